#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` + `--page source --csv` pair (read here, no GPU needed)."""
import csv
import sys


def main(raw, src):
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
            'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_alu.sum',
            'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'lts__t_bytes.sum']
    for i, h in enumerate(hdr):
        if h in want or ('pipe' in h and 'pct_of_peak_sustained_active' in h and h.startswith('sm__inst_executed_pipe')):
            print(f"{h:80s} {units[i]:16s} {vals[i]}")
    rows = list(csv.reader(open(src)))
    hdr, data = rows[1], rows[2:]
    ia, isamp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
    tot = sum(int(r[ia]) for r in data)
    tots = sum(int(r[isamp]) for r in data)
    print('warp instructions', tot, 'samples', tots)
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tr = {h: sum(int(r[hdr.index(h)]) for r in data) for h in reasons}
    print('stalls:', ', '.join(f"{k[6:]} {100 * v / tots:.1f}%" for k, v in sorted(tr.items(), key=lambda kv: -kv[1])[:9]))
    base = int(data[0][0], 16)
    start, acc_i, acc_s = 0, 0, 0
    for r in data:
        acc_i += int(r[ia])
        acc_s += int(r[isamp])
        if 'BAR.SYNC' in r[isrc] or 'EXIT' in r[isrc] or 'WARPSYNC.ALL' in r[isrc]:
            off = int(r[0], 16) - base
            if acc_i > 0.005 * tot or acc_s > 0.005 * tots:
                print(f"  {start:5x}-{off:5x}  inst {100 * acc_i / tot:5.1f}%  samples {100 * acc_s / tots:5.1f}%   ends with {r[isrc].strip()[:40]}")
            start, acc_i, acc_s = off + 16, 0, 0
    top = sorted(data, key=lambda r: -int(r[isamp]))[:14]
    for r in top:
        print(f"  {int(r[0], 16) - base:5x} s={100 * int(r[isamp]) / tots:5.2f}% n={int(r[ia]):9d} {r[isrc].strip()[:70]}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
