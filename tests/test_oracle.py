"""Pins the CPU oracle (and its AVX2 port) to the reference's own golden vectors.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, same_points

CONFIGS = [(16, 9, 0), (16, 9, 1), (16, 9, 2), (16, 12, 2), (32, 12, 2)]  # tests/compare.rs:66-114


def test_circle_matches_reference_table(oracle_mod):
    # fast_simd.rs:79-98 / opencv_compat.rs:42-61
    want = [(0, -3), (1, -3), (2, -2), (3, -1), (3, 0), (3, 1), (2, 2), (1, 3), (0, 3), (-1, 3), (-2, 2), (-3, 1),
            (-3, 0), (-3, -1), (-2, -2), (-1, -3)]
    assert [tuple(r) for r in oracle_mod.circle().tolist()] == want


def test_known_answer_vector_scores_20(oracle_mod):
    # fast_simd.rs:919-937 and :965-1021: centre 17, this ring, t=16, n=9 -> keypoint, MaxThreshold score 20
    ring = [37, 37, 39, 39, 37, 42, 43, 16, 14, 13, 15, 16, 15, 38, 37, 38]
    assert oracle_mod.score_max_threshold_px(17, ring, 9) == 20
    assert oracle_mod.port_score_max_threshold_px(17, ring, 9) == 20
    img = np.zeros((128, 128), np.uint8)  # create_sample_image, fast_simd.rs:866-881
    img[64, 64] = 17
    for (dx, dy), v in zip(oracle_mod.circle().tolist(), ring):
        img[64 + dy, 64 + dx] = v
    assert oracle_mod.is_keypoint(img, 64, 64, 16, 9)
    for fn in (oracle_mod.detect, oracle_mod.port_detect):
        pts = fn(img, 16, 9, 0)
        assert [64, 64] in pts.tolist()
    pts, scores = oracle_mod.detect(img, 16, 9, 1, return_scores=True)
    assert dict(zip(map(tuple, pts.tolist()), scores.tolist())).get((64, 64), 20) == 20


def test_consecutive_vectors(oracle_mod):
    # opencv_compat.rs:327-345
    c = oracle_mod.consecutive
    assert c([0, 0, 0, 1], 3) is False
    assert c([1, 0, 0, 1], 3) is False
    assert c([1, 0, 1, 1], 2) is True
    assert c([0, 1, 1, 1], 3) is True
    assert c([1, 0, 1, 1], 3) is True
    assert c([1, 1, 0, 1], 3) is True
    assert c([1, 1, 1, 0], 3) is True
    assert c([1, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1], 3) is False
    assert c([1, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 1, 1, 1], 4) is True


def test_shipped_golden_renders(oracle_mod, golden):
    """media/with_rust_*.png and media/with_opencv_*.png: 309 (Off) and 131 (MaxThreshold) keypoints."""
    grey = golden["grey"]
    assert grey.shape == (200, 300)
    assert len(golden["rust_off"]) == 309 and len(golden["rust_nonmax"]) == 131
    assert same_points(golden["rust_off"], golden["opencv_off"])
    assert same_points(golden["rust_nonmax"], golden["opencv_nonmax"])
    for fn in (oracle_mod.detect, oracle_mod.port_detect):
        assert same_points(fn(grey, 16, 9, 0), golden["rust_off"])
        assert same_points(fn(grey, 16, 9, 1), golden["rust_nonmax"])


def test_derived_counts_and_hashes(oracle_mod, golden):
    derived = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_derived.json")))
    for cfg in derived["configs"]:
        pts = oracle_mod.detect(golden["grey"], cfg["threshold"], cfg["count"], cfg["nms"])
        assert len(pts) == cfg["keypoints"]
        assert "0x%016x" % oracle_mod.hash_points(pts) == cfg["siphash13"]
    counts = {(c["threshold"], c["count"], c["nms"]): c["keypoints"] for c in derived["configs"]}
    assert counts[(16, 9, 0)] == 309 and counts[(16, 9, 1)] == 131 and counts[(16, 9, 2)] == 135


def test_siphash_reference_vectors(oracle_mod):
    # SipHash-1-3 with key (0, 0) == Rust's DefaultHasher; empty input value is well known
    assert oracle_mod.siphash13(b"") == 0xD1FBA762150C532C  # std DefaultHasher::new().finish()


def test_port_equals_scalar_on_compare_rs_configs(oracle_mod, golden):
    # tests/compare.rs:39-64: fast_simd::detector == opencv_compat::detector, order included
    for t, n, nms in CONFIGS + [(16, 12, 0), (16, 16, 0), (0, 9, 1), (255, 9, 0)]:
        assert same_points(oracle_mod.port_detect(golden["grey"], t, n, nms), oracle_mod.detect(golden["grey"], t, n, nms))


@pytest.mark.parametrize("n", range(9, 17))
def test_port_equals_scalar_n_sweep_on_synthetic(oracle_mod, n):
    img = oracle_mod.synth_frame(333, 150, seed=11, frame=n, kind=0, amp=6)
    for nms in (0, 1, 2):
        assert same_points(oracle_mod.port_detect(img, 16, n, nms), oracle_mod.detect(img, 16, n, nms))


def test_port_equals_scalar_random_sizes(oracle_mod):
    rng = np.random.default_rng(5)
    for trial in range(30):
        w, h = int(rng.integers(7, 90)), int(rng.integers(7, 60))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        t, n, nms = int(rng.choice([0, 3, 16, 40, 128, 250])), int(rng.integers(9, 17)), trial % 3
        assert same_points(oracle_mod.port_detect(img, t, n, nms), oracle_mod.detect(img, t, n, nms))


def test_randomised_score_equalities(oracle_mod):
    # fast_simd.rs:939-945 (20000 seeds, n = 9; here every n) and :1198-1236 (SAD, 10 M there, 2 M here)
    for n in range(9, 17):
        assert oracle_mod.kat_random_max_threshold(20000, n) == 0
    assert oracle_mod.kat_random_sum_abs(2_000_000) == 0


def test_degenerate_sizes_and_invalid_count(oracle_mod):
    for w, h in [(6, 6), (6, 50), (50, 6), (3, 3)]:
        assert len(oracle_mod.detect(np.full((h, w), 9, np.uint8), 1, 9, 0)) == 0
    img = np.zeros((20, 20), np.uint8)
    for bad in (8, 17, 0):
        with pytest.raises(ValueError):
            oracle_mod.detect(img, 16, bad, 0)
        with pytest.raises(ValueError):
            oracle_mod.port_detect(img, 16, bad, 0)
    # NMS never emits rows 3 and h-4 (opencv_compat.rs:238-240): an 8-row image has no emit rows
    assert len(oracle_mod.detect(np.random.default_rng(1).integers(0, 256, (8, 64), dtype=np.uint8), 5, 9, 1)) == 0


def test_nms_rows_3_and_h_minus_4_never_emitted(oracle_mod):
    img = np.random.default_rng(2).integers(0, 256, (40, 80), dtype=np.uint8)
    off = oracle_mod.detect(img, 10, 9, 0)
    assert (off[:, 1] == 3).any() and (off[:, 1] == 36).any()
    for nms in (1, 2):
        pts = oracle_mod.detect(img, 10, 9, nms)
        assert len(pts) and not (pts[:, 1] == 3).any() and not (pts[:, 1] == 36).any()


def test_opencv_cross_check(oracle_mod, golden):
    """cv2 FAST TYPE_9_16 agrees with the reference semantics (README.md:7): Off exactly; NMS after dropping
    rows 3 and h-4 from cv2's output, for t >= 1."""
    cv2 = pytest.importorskip("cv2")
    imgs = [golden["grey"], oracle_mod.synth_frame(257, 131, 3, 0, 0, 5)]
    for img in imgs:
        h = img.shape[0]
        for t in (7, 16, 40):
            fast = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=False,
                                                  type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kp = [(int(k.pt[0]), int(k.pt[1])) for k in fast.detect(img)]
            assert kp == [tuple(p) for p in oracle_mod.detect(img, t, 9, 0).tolist()]
            fast = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True,
                                                  type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kp = [(int(k.pt[0]), int(k.pt[1])) for k in fast.detect(img)]
            kp = [p for p in kp if p[1] != 3 and p[1] != h - 4]
            assert kp == [tuple(p) for p in oracle_mod.detect(img, t, 9, 1).tolist()]


def test_synthetic_scene_density_is_realistic(oracle_mod):
    # the generator is tuned to the reference's published 1080p keypoint count (README.md:58-59: 23184 at t16 n9)
    img = oracle_mod.synth_frame(1920, 1080, 1234, 0, 0, 4)
    n_off = len(oracle_mod.port_detect(img, 16, 9, 0))
    assert 15000 < n_off < 35000
