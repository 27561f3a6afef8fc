#!/usr/bin/env python3
"""Reads gpurun_out/trace_<tag>.npy (tools/trace_timeline.py) and prints, for the CTAs that ran on SM 0, the phase
timeline per chunk of the TWO-group detection kernel (4 filter + 6 test warps: the product kernel): when each group
could start, how long each phase took, who waited for whom.  (tools/trace_report.py is the three-group form used for
the v16 / v18 experiments.)
usage: trace_report_two_groups.py gpurun_out/trace_c3.npy [first chunk] [chunks to list]"""
import sys
import numpy as np

a = np.load(sys.argv[1])
c0 = int(sys.argv[2]) if len(sys.argv) > 2 else 40
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 12
NF = 4  # filter warps 0..3, test warps 4..9
ctas = [k for k in range(a.shape[0]) if (a[k] != 0).any()]
print("traced CTAs:", ctas)
for k in ctas:
    t = a[k].astype(np.float64)
    t[t == 0] = np.nan
    nch = int(np.sum(~np.isnan(t[0, :, 0])))
    print(f"\n=== CTA {k}: {nch} chunks traced")
    base = np.nanmin(t)
    t -= base
    F = t[:NF]      # [warp][chunk][slot]
    T = t[NF:10]
    sl = slice(8, max(9, nch - 8))  # steady state
    def m(x): return float(np.nanmean(x))
    period = m(np.diff(np.nanmax(T[:, :, 8], axis=0)[sl]))
    print(f"period per chunk {period:.0f} cycles")
    # filter group
    f_ready, f_s1, f_done = F[:, :, 0], F[:, :, 9], F[:, :, 1]
    prev_done = np.concatenate([np.full((NF, 1), np.nan), f_done[:, :-1]], axis=1)
    print(f"F: wait tile {m((f_ready - prev_done)[:, sl]):.0f}  stage1 {m((f_s1 - f_ready)[:, sl]):.0f}  "
          f"stage2+push {m((f_done - f_s1)[:, sl]):.0f}  | spread of A end across the 4 warps "
          f"{m((np.nanmax(f_done, axis=0) - np.nanmin(f_done, axis=0))[sl]):.0f}")
    # slots: 4 queue ready, 5 phase B done, 6 past the group barrier (lane 0 of the last warp: and the tile request),
    # 7 NMS pass done, 8 past the second barrier (lane 0 of the last warp: and the run records)
    q, b, req, nms, end = T[:, :, 4], T[:, :, 5], T[:, :, 6], T[:, :, 7], T[:, :, 8]
    bar1 = req
    prev_end = np.concatenate([np.full((T.shape[0], 1), np.nan), end[:, :-1]], axis=1)
    print(f"T: wait queue {m((q - prev_end)[:, sl]):.0f}  B {m((b - q)[:, sl]):.0f}  barrier (+ tile request on the last warp) "
          f"{m((req - b)[:, sl]):.0f}  NMS {m((nms - req)[:, sl]):.0f}  barrier (+ run records on the last warp) {m((end - nms)[:, sl]):.0f}")
    print(f"   last test warp alone: barrier + request {m((req - b)[-1, sl]):.0f}  NMS {m((nms - req)[-1, sl]):.0f}  barrier + records {m((end - nms)[-1, sl]):.0f}")
    # hand-offs: last filter arrival -> first test start; B end (bar1) -> tile ready
    lastA = np.nanmax(f_done, axis=0)
    print(f"hand-off last A end -> T start of B: {m((np.nanmin(q, axis=0) - lastA)[sl]):.0f}  "
          f"(negative = queue was ready before the test group asked)")
    bar1_all = np.nanmax(bar1, axis=0)
    tile2 = np.nanmin(f_ready, axis=0)
    if nch > 4:
        lat = (tile2[2:] - bar1_all[:-2])
        print(f"request (bar1 of chunk k) -> first filter warp starts chunk k+2: {m(lat[sl]):.0f}")
    print("chunk:  F ready(min) A end(max) | T start(min) B end(max=bar1) NMS end(max) chunk end(max)   [cycles from previous chunk end]")
    for c in range(c0, min(c0 + nl, nch)):
        ref = np.nanmax(end[:, c - 1]) if c > 0 else 0.0
        print(f"{c:4d}:  {np.nanmin(f_ready[:, c]) - ref:8.0f} {np.nanmax(f_done[:, c]) - ref:8.0f} | {np.nanmin(q[:, c]) - ref:8.0f} "
              f"{np.nanmax(bar1[:, c]) - ref:8.0f} {np.nanmax(nms[:, c]) - ref:8.0f} {np.nanmax(end[:, c]) - ref:8.0f}")
