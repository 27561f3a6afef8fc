"""CPU oracle for the FAST-n detection path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the CPU baseline.
The product (``feature_detector_fast_b200``) never imports it.

``oracle.detect``        scalar C restatement of /root/reference/src/opencv_compat.rs:79-306
``oracle.port_detect``   AVX2 C++ port of /root/reference/src/fast_simd.rs:115-824 (timed baseline)
"""
from .oracle import (  # noqa: F401
    NMS_MAX_THRESHOLD,
    NMS_OFF,
    NMS_SUM_ABSOLUTE,
    build,
    circle,
    consecutive,
    detect,
    hash_points,
    is_keypoint,
    kat_random_max_threshold,
    kat_random_sum_abs,
    port_detect,
    port_detect_many,
    port_score_max_threshold_px,
    port_score_sum_abs_px,
    score_max_threshold_px,
    score_sum_abs_px,
    siphash13,
    synth_frame,
)
