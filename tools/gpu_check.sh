#!/bin/bash
# Standard GPU-box check: GPU test tier, then a short bench (both logged under gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value', d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'],
          'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'),
          'clocks', d.get('clocks'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -n 3 gpurun_out/bench.err
