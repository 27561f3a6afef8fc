#!/usr/bin/env python3
"""GPU-box diagnostic: GPU vs oracle on a handful of cases with verbose mismatch output (not a test)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf  # noqa: E402
import oracle  # noqa: E402


def main():
    det = fdf.Detector(0)
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "shipped_300x200.npz"))["grey"]
    cases = [("shipped", g)]
    cases.append(("synth640x360", oracle.synth_frame(640, 360, 7, 0, 0, 4)))
    cases.append(("synth1080p", oracle.synth_frame(1920, 1080, 1234, 0, 0, 4)))
    cases.append(("noise300x100", oracle.synth_frame(300, 100, 1, 0, 1)))
    bad = 0
    for name, img in cases:
        for nms in (0, 1, 2):
            for n in (9, 12):
                cfg = fdf.Config(16, n, fdf.NonMaximalSuppression(nms))
                t0 = time.perf_counter()
                got = det.detect_array(img, cfg)
                dt = time.perf_counter() - t0
                want = oracle.port_detect(img, 16, n, nms)
                ok = got.shape == want.shape and np.array_equal(got, want)
                flags = det.device_flags()
                print(f"{name:14s} nms={nms} n={n}: gpu {len(got):6d} oracle {len(want):6d} "
                      f"{'OK ' if ok else 'MISMATCH'} flags={flags} {dt * 1e3:.2f} ms", flush=True)
                if not ok:
                    bad += 1
                    gs, ws = set(map(tuple, got.tolist())), set(map(tuple, want.tolist()))
                    print("   missing:", sorted(ws - gs)[:10], " extra:", sorted(gs - ws)[:10],
                          " same set:", gs == ws, flush=True)
    print("diag mismatches:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
