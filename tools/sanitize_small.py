#!/usr/bin/env python3
"""A small, fast exercise of every kernel path (a few seconds), meant to be run under
   compute-sanitizer --tool memcheck|synccheck python tools/sanitize_small.py
where that tool is available (it is closed on the pool this round was developed on, so it has only been run plain).
Shapes are chosen to hit: several chunks per row, several strips, the last-chunk validity table, all three NMS modes,
every count, the dense fallback (uniform noise), the RGB path and a small batch.  No oracle: results are only printed."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf  # noqa: E402

det = fdf.Detector(0)
rng = np.random.default_rng(1)
total = 0
for (w, h, kind) in [(700, 150, 0), (333, 77, 0), (520, 80, 1), (64, 40, 1)]:
    frames = det.synth_frames(2, (w + 15) // 16 * 16, h, seed=3, kind=kind)[:, :, :w]
    for nms in (0, 1, 2):
        for n in (9, 12, 16):
            t = 3 if kind == 1 else 16
            pts, offs = det.detect_device(frames, fdf.Config(t, n, fdf.NonMaximalSuppression(nms)))
            torch.cuda.synchronize()
            total += int(offs[-1])
            assert det.device_flags() == 0
img = rng.integers(0, 256, (90, 301), dtype=np.uint8)
total += len(det.detect_array(img, fdf.Config(20, 9, fdf.NonMaximalSuppression.MaxThreshold)))
rgb = rng.integers(0, 256, (50, 123, 3), dtype=np.uint8)
total += len(det.detect_rgb8_array(rgb, fdf.Config(20, 9, fdf.NonMaximalSuppression.SumAbsolute)))
batch = rng.integers(0, 256, (3, 60, 250), dtype=np.uint8)
total += len(det.detect_batch(batch, fdf.Config(40, 9, fdf.NonMaximalSuppression.Off))[0])
print("sanitize_small: keypoints", total)
