"""The detection kernel's per-thread phase bodies (csrc/fdf_strip.cuh + fdf_core.cuh -- the code the GPU
executes) run thread by thread on the CPU against the oracle.  Catches tiling / halo / validity / NMS-row and
SWAR-arithmetic errors without a GPU.  CPU only."""
import numpy as np
import pytest

from conftest import same_points


def test_device_arithmetic_against_oracle(emulator):
    # ring masks, arc test (n = 9..16), both scores, and "the filter never rejects a keypoint"
    assert emulator.core_check(2_000_000, 42) == 0


@pytest.mark.parametrize("sr", [32, 64])
def test_shipped_image_all_reference_configs(emulator, oracle_mod, golden, sr):
    grey = golden["grey"]
    assert same_points(emulator(grey, 16, 9, 0, sr), golden["rust_off"])
    assert same_points(emulator(grey, 16, 9, 1, sr), golden["rust_nonmax"])
    for t, n, nms in [(16, 9, 2), (16, 12, 2), (32, 12, 2), (16, 12, 0), (16, 16, 0), (0, 9, 1), (255, 9, 0),
                      (130, 10, 1), (127, 9, 2), (128, 9, 2)]:
        assert same_points(emulator(grey, t, n, nms, sr), oracle_mod.detect(grey, t, n, nms)), (t, n, nms)


@pytest.mark.parametrize("n", range(9, 17))
def test_count_sweep(emulator, oracle_mod, n):
    img = oracle_mod.synth_frame(500, 90, seed=21, frame=n, kind=0, amp=5)
    for nms in (0, 1, 2):
        assert same_points(emulator(img, 16, n, nms, 32), oracle_mod.detect(img, 16, n, nms))


def test_widths_around_chunk_and_word_boundaries(emulator, oracle_mod):
    # chunk = 240 output columns, tile = 256, bit-plane word = 32 columns
    for w in (7, 8, 9, 31, 32, 33, 239, 240, 241, 243, 244, 246, 247, 248, 255, 256, 257, 479, 480, 481, 487):
        img = oracle_mod.synth_frame(w, 41, seed=w, frame=0, kind=1)
        for nms in (0, 1):
            assert same_points(emulator(img, 30, 9, nms, 32), oracle_mod.detect(img, 30, 9, nms)), w


def test_heights_around_strip_boundaries(emulator, oracle_mod):
    # strips emit 32/64 rows (Off) or 30/62 rows (NMS) starting at row 3/4
    for h in (7, 8, 9, 10, 17, 23, 33, 34, 35, 36, 37, 38, 39, 40, 64, 65, 66, 67, 68, 69, 70, 71, 72, 73, 74, 130, 131, 132):
        img = oracle_mod.synth_frame(70, h, seed=h, frame=1, kind=1)
        for nms in (0, 2):
            for sr in (32, 64):
                assert same_points(emulator(img, 25, 9, nms, sr), oracle_mod.detect(img, 25, 9, nms)), (h, nms, sr)


def test_random_images(emulator, oracle_mod):
    rng = np.random.default_rng(0)
    for trial in range(40):
        w, h = int(rng.integers(7, 600)), int(rng.integers(7, 120))
        kind = trial % 3
        if kind == 0:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 1:
            img = oracle_mod.synth_frame(w, h, trial, 0, 0, 4)
        else:  # binary blobs: many exact ties for the NMS
            img = (rng.integers(0, 2, (h, w)) * rng.integers(20, 255)).astype(np.uint8)
        t = int(rng.choice([0, 1, 7, 16, 20, 60, 127, 128, 200]))
        n, nms, sr = int(rng.integers(9, 17)), trial % 3, (64 if trial % 2 else 32)
        assert same_points(emulator(img, t, n, nms, sr), oracle_mod.detect(img, t, n, nms)), (w, h, t, n, nms, sr)


def test_wide_image_many_chunks(emulator, oracle_mod):
    # more than 15 chunks per row: the score-plane tag sequence restarts (plane cleared) inside a strip
    img = oracle_mod.synth_frame(3840, 40, seed=77, frame=0, kind=0, amp=4)
    for nms in (1, 2):
        assert same_points(emulator(img, 16, 9, nms, 32), oracle_mod.detect(img, 16, 9, nms))
    img = oracle_mod.synth_frame(4100, 24, seed=78, frame=0, kind=1)
    assert same_points(emulator(img, 40, 9, 1, 64), oracle_mod.detect(img, 40, 9, 1))


def test_dense_content_takes_the_fallback_paths(emulator, oracle_mod):
    # uniform noise at a low threshold: > 1024 candidates per chunk (every pixel tested straight from the tile, then
    # the emit warps scan the plane); must be exercised and still match the oracle bit for bit
    img = oracle_mod.synth_frame(520, 80, seed=5, frame=0, kind=1)
    for nms in (0, 1, 2):
        got, fb = emulator(img, 3, 9, nms, 32, want_fallbacks=True)
        assert same_points(got, oracle_mod.detect(img, 3, 9, nms))
        assert fb[0] > 0 and (nms == 0 or fb[1] > 0), fb
    got, fb = emulator(oracle_mod.synth_frame(520, 80, 6, 0, 0, 4), 16, 9, 1, 32, want_fallbacks=True)
    assert fb == [0, 0]  # realistic content never leaves the fast path


def test_saturated_and_flat_images(emulator, oracle_mod):
    for v in (0, 255, 17):
        img = np.full((50, 300), v, np.uint8)
        assert len(emulator(img, 0, 9, 1, 32)) == 0
    img = np.zeros((60, 520), np.uint8)
    img[::2, ::2] = 255  # every pixel at the extremes: exercises the >= 128 difference paths
    for t in (0, 100, 127, 128, 254, 255):
        for nms in (0, 1, 2):
            assert same_points(emulator(img, t, 9, nms, 32), oracle_mod.detect(img, t, 9, nms)), (t, nms)
