#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference checkout (run in the build container only).

Reads  /root/reference/media/*.png  (the reference's own shipped input and golden renders) and
writes small fixtures that travel to the GPU box, where /root/reference does not exist:

  shipped_300x200.npz
      grey            (200, 300) u8  -- media/Screenshot315_torch_grey.png (r == g == b)
      rust_off        (309, 2)  u32  -- red pixels of media/with_rust_threshold_16_consecutive_9.png
      rust_nonmax     (131, 2)  u32  -- red pixels of ..._consecutive_9.png_nonmax.png
      opencv_off      (309, 2)  u32  -- red pixels of media/with_opencv_threshold_16_type_9_16.png
      opencv_nonmax   (131, 2)  u32  -- red pixels of ..._type_9_16_nonmax.png
    main.rs:74-77 paints exactly one pure-red pixel per keypoint (draw_plus_sized(.., RED, 1)), so the
    red pixels ARE the reference's keypoint sets for (t=16, n=9, Off) and (t=16, n=9, MaxThreshold);
    listed row-major, which is the reference's output order (fast_simd.rs:550, 596-613).

  oracle_derived.json
      keypoint counts and tests/compare.rs-format SipHash-1-3 values for the five configurations of
      tests/compare.rs:66-114 plus (16,12,Off) and (16,16,Off) on the shipped image, and the
      known-answer vector of fast_simd.rs:919-937.  The first two are pinned by the renders above;
      the others are produced by the scalar oracle (they rest on the restatement only) and guard
      against regressions of the oracle itself.
"""
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
MEDIA = "/root/reference/media"


def red_points(path, grey):
    rgb = np.array(Image.open(path).convert("RGB"))
    red = (rgb[..., 0] == 255) & (rgb[..., 1] == 0) & (rgb[..., 2] == 0)
    # every other pixel must be the untouched grey input, otherwise the render is not of this image
    same = (rgb[..., 0] == grey) & (rgb[..., 1] == grey) & (rgb[..., 2] == grey)
    assert (red | same).all(), path
    ys, xs = np.nonzero(red)  # np.nonzero is row-major: y ascending then x ascending
    return np.stack([xs, ys], axis=1).astype(np.uint32)


def main():
    import oracle

    rgb = np.array(Image.open(os.path.join(MEDIA, "Screenshot315_torch_grey.png")).convert("RGB"))
    assert (rgb[..., 0] == rgb[..., 1]).all() and (rgb[..., 1] == rgb[..., 2]).all()
    grey = np.ascontiguousarray(rgb[..., 0])
    fix = dict(
        grey=grey,
        rust_off=red_points(os.path.join(MEDIA, "with_rust_threshold_16_consecutive_9.png"), grey),
        rust_nonmax=red_points(os.path.join(MEDIA, "with_rust_threshold_16_consecutive_9.png_nonmax.png"), grey),
        opencv_off=red_points(os.path.join(MEDIA, "with_opencv_threshold_16_type_9_16.png"), grey),
        opencv_nonmax=red_points(os.path.join(MEDIA, "with_opencv_threshold_16_type_9_16_nonmax.png"), grey),
    )
    for k, v in fix.items():
        print(k, v.shape)
    np.savez_compressed(os.path.join(HERE, "shipped_300x200.npz"), **fix)

    derived = {"image": "media/Screenshot315_torch_grey.png", "configs": []}
    for t, n, nms in [(16, 9, 0), (16, 9, 1), (16, 9, 2), (16, 12, 2), (32, 12, 2), (16, 12, 0), (16, 16, 0)]:
        pts = oracle.detect(grey, t, n, nms)
        derived["configs"].append(
            {"threshold": t, "count": n, "nms": nms, "keypoints": int(len(pts)),
             "siphash13": "0x%016x" % oracle.hash_points(pts)})
    ring = [37, 37, 39, 39, 37, 42, 43, 16, 14, 13, 15, 16, 15, 38, 37, 38]
    derived["kat"] = {"centre": 17, "ring": ring, "threshold": 16, "count": 9,
                      "score_max_threshold": oracle.score_max_threshold_px(17, ring, 9)}
    import struct
    derived["rgb8_siphash13"] = "0x%016x" % oracle.siphash13(struct.pack("<Q", rgb.size) + rgb.tobytes())
    with open(os.path.join(HERE, "oracle_derived.json"), "w") as f:
        json.dump(derived, f, indent=1)
    print(json.dumps(derived, indent=1))


if __name__ == "__main__":
    main()
