#!/usr/bin/env python3
"""Reads gpurun_out/trace_<tag>.npy (tools/trace_timeline.py) and prints, for the CTAs that ran on SM 0, the phase
timeline per chunk of the three warp groups of the detection kernel (filter / test / emit).
usage: trace_report.py trace.npy [first chunk] [chunks to list] [test warps] [emit warps]"""
import sys
import numpy as np

a = np.load(sys.argv[1])
c0 = int(sys.argv[2]) if len(sys.argv) > 2 else 40
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 10
NT = int(sys.argv[4]) if len(sys.argv) > 4 else 4
NE = int(sys.argv[5]) if len(sys.argv) > 5 else 2
NF = 4
ctas = [k for k in range(a.shape[0]) if (a[k] != 0).any()]
print("traced CTAs:", ctas)
m = lambda x: float(np.nanmean(x))
for k in ctas:
    t = a[k].astype(np.float64)
    t[t == 0] = np.nan
    nch = int(np.sum(~np.isnan(t[0, :, 0])))
    t -= np.nanmin(t)
    F, T, E = t[:NF], t[NF:NF + NT], t[NF + NT:NF + NT + NE]
    sl = slice(8, max(9, nch - 8))
    print(f"\n=== CTA {k}: {nch} chunks traced; period per chunk {m(np.diff(np.nanmax(E[:, :, 11], axis=0)[sl])):.0f} cycles")
    f_ready, f_s1, f_bar, f_done = F[:, :, 0], F[:, :, 9], F[:, :, 10], F[:, :, 1]
    prev = np.concatenate([np.full((NF, 1), np.nan), f_done[:, :-1]], axis=1)
    print(f"filter: wait tile {m((f_ready - prev)[:, sl]):.0f}  stage 1 {m((f_s1 - f_ready)[:, sl]):.0f}  barrier {m((f_bar - f_s1)[:, sl]):.0f}"
          f"  stage 2 + push {m((f_done - f_bar)[:, sl]):.0f}  | spread of the warps' finish {m((np.nanmax(f_done, 0) - np.nanmin(f_done, 0))[sl]):.0f}")
    t_go, t_b, t_end = T[:, :, 4], T[:, :, 5], T[:, :, 6]
    prev = np.concatenate([np.full((NT, 1), np.nan), t_end[:, :-1]], axis=1)
    print(f"test:   wait (queue, plane) {m((t_go - prev)[:, sl]):.0f}  phase B {m((t_b - t_go)[:, sl]):.0f}  hand-over (+ tile request by the last warp) {m((t_end - t_b)[:, sl]):.0f}"
          f"  | spread of the warps' finish {m((np.nanmax(t_b, 0) - np.nanmin(t_b, 0))[sl]):.0f}")
    e_go, e_nms, e_end = E[:, :, 7], E[:, :, 8], E[:, :, 11]
    prev = np.concatenate([np.full((NE, 1), np.nan), e_end[:, :-1]], axis=1)
    print(f"emit:   wait list {m((e_go - prev)[:, sl]):.0f}  NMS + staging {m((e_nms - e_go)[:, sl]):.0f}  records, wipe, release {m((e_end - e_nms)[:, sl]):.0f}")
    print(f"hand-offs: last filter arrival -> first test warp starts {m((np.nanmin(t_go, 0) - np.nanmax(f_done, 0))[sl]):.0f}   "
          f"last test warp done -> first emit warp starts {m((np.nanmin(e_go, 0) - np.nanmax(t_b, 0))[sl]):.0f}")
    if nch > 6:
        print(f"tile request (last test warp done with chunk k) -> first filter warp starts chunk k+2: "
              f"{m((np.nanmin(f_ready, 0)[2:] - np.nanmax(t_b, 0)[:-2])[sl]):.0f}")
    print("chunk:  filter start(min) end(max) | test start(min) end(max) | emit start(min) end(max)   [cycles from the first event of the listing]")
    ref = np.nanmin(t[:, c0, :]) if c0 < nch else 0.0
    for c in range(c0, min(c0 + nl, nch)):
        print(f"{c:4d}:  {np.nanmin(f_ready[:, c]) - ref:8.0f} {np.nanmax(f_done[:, c]) - ref:8.0f} | {np.nanmin(t_go[:, c]) - ref:8.0f} "
              f"{np.nanmax(t_b[:, c]) - ref:8.0f} | {np.nanmin(e_go[:, c]) - ref:8.0f} {np.nanmax(e_end[:, c]) - ref:8.0f}")
