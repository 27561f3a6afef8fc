#!/usr/bin/env python3
"""Per-batch time line of the sharded path (ShardedDetector.trace): when, relative to the first batch's start, each
rank's detection started / ended (current stream) and its all-gather started / ended and its push ended (exchange
stream).  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/shard_timeline.py --frames 512
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf  # noqa: E402
from feature_detector_fast_b200 import sharding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=512, help="frames of the whole batch (sharded over the ranks)")
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--idle-sm-stride", type=int, default=37, help="0 = the detection kernel uses every SM")
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
det = fdf.Detector(local)
lo, hi = sharding.frame_shard(a.frames, rank, world)
frames = det.synth_frames(hi - lo, 3840, 2160, seed=20240, first_frame=lo)
cfg = fdf.Config(20, 9, fdf.NonMaximalSuppression.MaxThreshold)
sd = sharding.ShardedDetector(det, a.frames, cap_total=a.frames * 30000, cap_local=(hi - lo) * 30000,
                              idle_sm_stride=a.idle_sm_stride)
for _ in range(6):
    sd.detect(frames, cfg)
sd.fence()
torch.cuda.synchronize()
dist.barrier()
sd.trace = []
for _ in range(a.steps):
    sd.detect(frames, cfg)
sd.fence()
torch.cuda.synchronize()
t0 = sd.trace[0][0]
rows = [[t0.elapsed_time(e) for e in tr] for tr in sd.trace]
out = [None] * world
dist.all_gather_object(out, rows)
if rank == 0:
    print(f"{world} ranks, {a.frames} frames per batch ({hi - lo} per rank), idle SM stride {a.idle_sm_stride}; "
          "ms from the rank's first batch start")
    for r in (0, 1, world - 1):
        print(f"rank {r}:  batch  detect start   detect end   gather start   gather end   push end   | step")
        prev = None
        for k, row in enumerate(out[r][4:20], start=4):
            step = row[0] - prev if prev is not None else float('nan')
            prev = row[0]
            print(f"        {k:4d}   {row[0]:10.3f}   {row[1]:10.3f}   {row[2]:10.3f}   {row[3]:10.3f}   {row[4]:9.3f}   | {step:6.3f}")
sd.close()
dist.destroy_process_group()
