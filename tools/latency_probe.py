#!/usr/bin/env python3
"""Single-frame latency of the host entry point (fdf_detect) and of the device entry point (fdf_detect_device + sync)
on one 1920x1080 frame: where the 0.2 ms of `bench.py --criterion` go."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf  # noqa: E402

det = fdf.Detector(0)
frames = det.synth_frames(1, 1920, 1080, seed=20240, kind=0)
host = frames[0].cpu().numpy()
cfg = fdf.Config(16, 9, fdf.NonMaximalSuppression.MaxThreshold)
pts = torch.empty((200000, 2), dtype=torch.int32, device="cuda")
offs = torch.empty(2, dtype=torch.int64, device="cuda")


def timeit(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def dev_sync():
    det.detect_device(frames, cfg, points=pts, offsets=offs)
    torch.cuda.synchronize()


def dev_async():
    det.detect_device(frames, cfg, points=pts, offsets=offs)


print("fdf_detect (pageable host image)   %.1f us" % timeit(lambda: det.detect_array(host, cfg)))
pinned = torch.empty((1080, 1920), dtype=torch.uint8, pin_memory=True)
pinned.copy_(frames[0].cpu())
pinned_np = pinned.numpy()
print("fdf_detect (pinned host image)     %.1f us" % timeit(lambda: det.detect_array(pinned_np, cfg)))
for nms in (0, 2):
    c2 = fdf.Config(16, 9, fdf.NonMaximalSuppression(nms))
    print("fdf_detect (pinned, nms %d)         %.1f us" % (nms, timeit(lambda: det.detect_array(pinned_np, c2))))
for sr in (64, 48, 32):
    det.set_tuning(strip_rows=sr)
    print("fdf_detect (pinned, strip rows %d)  %.1f us" % (sr, timeit(lambda: det.detect_array(pinned_np, cfg))))
det.set_tuning()
print("fdf_detect_device + synchronize    %.1f us" % timeit(dev_sync))
print("fdf_detect_device, back to back    %.1f us per call (enqueue-bound)" % timeit(dev_async))
det.set_timing(8)
for _ in range(8):
    dev_sync()
ms = [det.get_timing(i) for i in range(8)]
print("kernels: detect %.1f us, scan %.1f us, gather %.1f us" % tuple(1e3 * sum(m[k] for m in ms) / 8 for k in range(3)))
