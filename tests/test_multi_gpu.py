"""The multi-GPU product path on real GPUs: runs tests/multi_gpu_check.py under torch.distributed.run with two ranks
(one per GPU, NCCL) when at least two GPUs are visible.  The index arithmetic of the same path is covered on the CPU by
tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_sharded_batch_assembles_one_result_on_rank_0():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count(": OK") == 3, res.stdout[-3000:]
