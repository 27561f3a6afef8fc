"""Parity of the CUDA path against the oracle and the reference's golden vectors, through the C ABI.

Bit-exact (integer work): keypoint sets AND order must be identical.  Needs a B200: `-m gpu`."""
import numpy as np
import pytest

from conftest import same_points

pytestmark = pytest.mark.gpu


def _cfg(t, n, nms):
    import feature_detector_fast_b200 as fdf

    return fdf.Config(threshold=t, count=n, non_maximal_supression=fdf.NonMaximalSuppression(nms))


def test_library_loaded_is_the_in_tree_cuda_build(detector):
    import feature_detector_fast_b200 as fdf

    assert fdf.LIB_PATH.endswith("feature_detector_fast_b200/lib/libfdf_cuda.so")
    assert detector.kernel_launches == 0 or detector.kernel_launches > 0


def test_shipped_golden_renders(detector, golden):
    """tests/compare.rs configs on media/Screenshot315_torch_grey.png against the shipped renders."""
    before = detector.kernel_launches
    assert same_points(detector.detect_array(golden["grey"], _cfg(16, 9, 0)), golden["rust_off"])
    assert same_points(detector.detect_array(golden["grey"], _cfg(16, 9, 1)), golden["rust_nonmax"])
    # the CUDA kernels really ran: detect + gather per call (one small image: the gather kernel scans the strip
    # counts itself; batches launch detect, scan, gather)
    assert detector.kernel_launches == before + 4
    assert detector.device_flags() == 0


def test_compare_rs_configs_against_oracle(detector, oracle_mod, golden):
    # tests/compare.rs:66-114 (+ the Off variants and the threshold extremes)
    for t, n, nms in [(16, 9, 0), (16, 9, 1), (16, 9, 2), (16, 12, 2), (32, 12, 2), (16, 12, 0), (16, 16, 0),
                      (0, 9, 1), (0, 9, 2), (255, 9, 0), (130, 10, 1), (127, 9, 2), (128, 9, 2)]:
        got = detector.detect_array(golden["grey"], _cfg(t, n, nms))
        assert same_points(got, oracle_mod.detect(golden["grey"], t, n, nms)), (t, n, nms)


def test_public_detect_returns_points(golden):
    import feature_detector_fast_b200 as fdf

    cfg = fdf.Config(16, 9, fdf.NonMaximalSuppression.MaxThreshold)
    pts = fdf.detect(golden["grey"], cfg)
    assert pts == [fdf.Point(int(x), int(y)) for x, y in golden["rust_nonmax"]]
    assert cfg.detect(golden["grey"]) == pts


def test_known_answer_vector(detector, oracle_mod):
    # fast_simd.rs:950-1022: 128x128 black image, centre (64,64)=17 with the hand ring -> keypoint
    ring = [37, 37, 39, 39, 37, 42, 43, 16, 14, 13, 15, 16, 15, 38, 37, 38]
    img = np.zeros((128, 128), np.uint8)
    img[64, 64] = 17
    for (dx, dy), v in zip(oracle_mod.circle().tolist(), ring):
        img[64 + dy, 64 + dx] = v
    got = detector.detect_array(img, _cfg(16, 9, 0))
    assert [64, 64] in got.tolist()
    assert same_points(got, oracle_mod.detect(img, 16, 9, 0))


@pytest.mark.parametrize("n", range(9, 17))
def test_count_sweep_1080p(detector, oracle_mod, n):
    """BASELINE config 4: n = 9..16 on synthetic 1080p noise+edge frames, all three modes."""
    img = oracle_mod.synth_frame(1920, 1080, seed=4, frame=n, kind=0, amp=6)
    for nms in (0, 1, 2):
        got = detector.detect_array(img, _cfg(16, n, nms))
        assert same_points(got, oracle_mod.port_detect(img, 16, n, nms)), (n, nms)
    # the scalar oracle itself on one mode (the port is validated against it in the CPU tier)
    assert same_points(detector.detect_array(img, _cfg(16, n, 1)), oracle_mod.detect(img, 16, n, 1))


def test_1080p_three_modes(detector, oracle_mod):
    """BASELINE configs 1-3 (t=16, n=9, Off / MaxThreshold / SumAbsolute) on a synthetic 1080p frame."""
    img = oracle_mod.synth_frame(1920, 1080, seed=1234, frame=0, kind=0, amp=4)
    for nms in (0, 1, 2):
        assert same_points(detector.detect_array(img, _cfg(16, 9, nms)), oracle_mod.detect(img, 16, 9, nms))


def test_widths_heights_and_tiny_images(detector, oracle_mod):
    for w, h in [(7, 7), (8, 8), (9, 9), (16, 9), (6, 30), (30, 6), (31, 17), (239, 20), (240, 21), (241, 22),
                 (247, 33), (248, 34), (255, 35), (256, 36), (257, 37), (300, 38), (479, 39), (481, 40),
                 (487, 64), (70, 65), (70, 66), (70, 67), (70, 68), (70, 69), (1000, 70)]:
        img = oracle_mod.synth_frame(w, h, seed=w * 1000 + h, frame=0, kind=1)
        for nms in (0, 1, 2):
            got = detector.detect_array(img, _cfg(30, 9, nms))
            assert same_points(got, oracle_mod.detect(img, 30, 9, nms)), (w, h, nms)


def test_pitch_larger_than_width(detector, oracle_mod):
    big = oracle_mod.synth_frame(400, 100, seed=9, frame=0, kind=0, amp=8)
    view = big[:, 13:313]  # 300 wide, pitch 400, misaligned start
    assert same_points(detector.detect_array(view, _cfg(16, 9, 1)),
                       oracle_mod.detect(np.ascontiguousarray(view), 16, 9, 1))


def test_uniform_noise_dense_keypoints(detector, oracle_mod):
    img = oracle_mod.synth_frame(640, 480, seed=1, frame=0, kind=1)
    for t, nms in [(16, 0), (16, 1), (16, 2), (0, 0), (60, 2)]:
        assert same_points(detector.detect_array(img, _cfg(t, 9, nms)), oracle_mod.port_detect(img, t, 9, nms))


def test_exact_ties_binary_image(detector, oracle_mod):
    rng = np.random.default_rng(3)
    img = (rng.integers(0, 2, (200, 520)) * 200).astype(np.uint8)
    for nms in (0, 1, 2):
        assert same_points(detector.detect_array(img, _cfg(50, 9, nms)), oracle_mod.detect(img, 50, 9, nms))


def test_invalid_count_panics_like_the_reference(detector):
    import feature_detector_fast_b200 as fdf

    img = np.zeros((32, 32), np.uint8)
    for bad in (8, 17, 0, 255):  # fast_simd.rs:302-305 and :797-801
        with pytest.raises(fdf.FdfPanic):
            detector.detect_array(img, _cfg(16, bad, 0))


def test_capacity_overflow_reports_needed_size(detector, oracle_mod):
    import feature_detector_fast_b200 as fdf

    img = oracle_mod.synth_frame(320, 200, seed=2, frame=0, kind=1)
    want = oracle_mod.detect(img, 16, 9, 0)
    assert len(want) > 100
    with pytest.raises(fdf.FdfError) as ei:
        detector.detect_array(img, _cfg(16, 9, 0), cap=100)
    assert ei.value.status == 4


def test_capacity_equal_to_the_exact_count_and_one_less(detector, oracle_mod):
    """ADVICE r1 (staging overflow): the unordered staging area is sized from the caller's capacity, so the tightest
    legal capacity (== the number of keypoints) is the case where a silent overflow would drop points.  It must give
    the full list with no device flag; one less must be FDF_ERR_CAPACITY carrying the true count -- for sparse and
    for dense (noise: dense path, hundreds of keypoints per chunk) content, every NMS mode."""
    import feature_detector_fast_b200 as fdf

    for kind, t in ((0, 16), (1, 3)):
        frames = np.stack([oracle_mod.synth_frame(1296, 300, seed=31, frame=f, kind=kind, amp=5) for f in range(3)])
        for nms in (0, 1, 2):
            want = [oracle_mod.port_detect(frames[f], t, 9, nms) for f in range(3)]
            k = sum(len(x) for x in want)
            assert k > 50
            pts, offs = detector.detect_batch(frames, _cfg(t, 9, nms), cap=k)
            assert offs[-1] == k and detector.device_flags() == 0
            for f in range(3):
                assert same_points(pts[int(offs[f]):int(offs[f + 1])], want[f])
            with pytest.raises(fdf.FdfError) as ei:
                detector.detect_batch(frames, _cfg(t, 9, nms), cap=k - 1)
            assert ei.value.status == 4 and str(k) in str(ei.value)
            with pytest.raises(fdf.FdfError) as ei:
                detector.detect_batch(frames, _cfg(t, 9, nms), cap=1)
            assert ei.value.status == 4 and str(k) in str(ei.value)
            one = detector.detect_array(frames[1], _cfg(t, 9, nms), cap=len(want[1]))
            assert same_points(one, want[1])


def test_batch_csr_output(detector, oracle_mod):
    frames = np.stack([oracle_mod.synth_frame(400, 130, seed=77, frame=f, kind=0, amp=5) for f in range(9)])
    for nms in (0, 1, 2):
        pts, offs = detector.detect_batch(frames, _cfg(20, 9, nms))
        assert offs[0] == 0 and len(offs) == 10 and offs[-1] == len(pts)
        for f in range(9):
            assert same_points(pts[int(offs[f]):int(offs[f + 1])], oracle_mod.detect(frames[f], 20, 9, nms)), (f, nms)


def test_gpu_generator_equals_oracle_generator(detector, oracle_mod):
    for kind, amp in [(0, 4), (0, 9), (1, 0)]:
        dev = detector.synth_frames(3, 333, 77, seed=5, first_frame=2, kind=kind, amp=amp)
        host = dev.cpu().numpy()
        for f in range(3):
            assert np.array_equal(host[f], oracle_mod.synth_frame(333, 77, 5, 2 + f, kind, amp))


def test_device_resident_path(detector, oracle_mod):
    import torch

    frames = detector.synth_frames(6, 1920, 1080, seed=31, first_frame=0, kind=0, amp=4)
    pts, offs = detector.detect_device(frames, _cfg(20, 9, 1))
    torch.cuda.synchronize()
    assert detector.device_flags() == 0
    offs_h = offs.cpu().numpy()
    pts_h = pts[: int(offs_h[-1])].cpu().numpy().astype(np.uint32)
    host = frames.cpu().numpy()
    for f in range(6):
        assert same_points(pts_h[offs_h[f]:offs_h[f + 1]], oracle_mod.port_detect(host[f], 20, 9, 1)), f


def test_misaligned_device_buffer_is_rejected(detector):
    import torch

    import feature_detector_fast_b200 as fdf

    frames = torch.zeros((2, 50, 301), dtype=torch.uint8, device="cuda")  # pitch 301: not a multiple of 16
    with pytest.raises(fdf.FdfError) as ei:
        detector.detect_device(frames, _cfg(16, 9, 0))
    assert ei.value.status == 3


def test_4k_batch_config5_sample(detector, oracle_mod):
    """BASELINE config 5 geometry (3840x2160, t=20, n=9, MaxThreshold) on a 24-frame resident batch:
    every frame's keypoint count and compare.rs-format hash equal the AVX2 port's; three frames are also
    compared point by point."""
    import torch

    n = 24
    frames = detector.synth_frames(n, 3840, 2160, seed=2024, first_frame=0, kind=0, amp=4)
    pts, offs = detector.detect_device(frames, _cfg(20, 9, 1))
    torch.cuda.synchronize()
    assert detector.device_flags() == 0
    offs_h = offs.cpu().numpy()
    pts_h = pts[: int(offs_h[-1])].cpu().numpy().astype(np.uint32)
    host = frames.cpu().numpy()
    counts, hashes = oracle_mod.port_detect_many(host, 20, 9, 1, n_threads=8)
    assert (np.diff(offs_h) == counts).all()
    for f in range(n):
        assert oracle_mod.hash_points(pts_h[offs_h[f]:offs_h[f + 1]]) == int(hashes[f]), f
    for f in (0, 11, 23):
        assert same_points(pts_h[offs_h[f]:offs_h[f + 1]], oracle_mod.port_detect(host[f], 20, 9, 1))
    # size-independent properties: row-major order inside every frame, centres inside the NMS emit range
    for f in range(n):
        p = pts_h[offs_h[f]:offs_h[f + 1]].astype(np.int64)
        key = p[:, 1] * 3840 + p[:, 0]
        assert (np.diff(key) > 0).all()
        assert p[:, 0].min() >= 3 and p[:, 0].max() < 3840 - 3 and p[:, 1].min() >= 4 and p[:, 1].max() <= 2160 - 5


def test_idempotent_and_order_independent_of_batching(detector, oracle_mod):
    """Detecting a frame alone, or as part of a batch, gives the same list (the look-back only adds offsets)."""
    frames = np.stack([oracle_mod.synth_frame(960, 540, seed=8, frame=f, kind=0, amp=4) for f in range(5)])
    pts, offs = detector.detect_batch(frames, _cfg(16, 9, 2))
    for f in range(5):
        alone = detector.detect_array(frames[f], _cfg(16, 9, 2))
        assert same_points(pts[int(offs[f]):int(offs[f + 1])], alone)
        assert same_points(alone, detector.detect_array(frames[f], _cfg(16, 9, 2)))


def test_cpp_binding_and_cli_regenerate_the_shipped_renders(golden, tmp_path):
    """include/fdf.hpp (the C++ mirror of lib.rs) through tools/fdf_cli.cpp (the counterpart of main.rs):
    the output image must be the shipped golden render (grey input + one pure-red pixel per keypoint,
    main.rs:74-77) and the text file "x y" per keypoint in order (main.rs:4-15)."""
    import subprocess

    import __graft_entry__ as entry

    cli = entry.build_cli()
    grey = golden["grey"]
    h, w = grey.shape
    pgm = tmp_path / "in.pgm"
    pgm.write_bytes(b"P5\n%d %d\n255\n" % (w, h) + grey.tobytes())
    for mode, want in (("off", golden["rust_off"]), ("max_threshold", golden["rust_nonmax"])):
        out = tmp_path / f"out_{mode}.ppm"
        res = subprocess.run([cli, str(pgm), str(out), "16", "9", mode], capture_output=True, text=True, timeout=120)
        assert res.returncode == 0, res.stderr
        assert f"found {len(want)} keypoints" in res.stdout
        raw = out.read_bytes()
        header = b"P6\n%d %d\n255\n" % (w, h)
        assert raw.startswith(header)
        rgb = np.frombuffer(raw[len(header):], np.uint8).reshape(h, w, 3)
        render = np.repeat(grey[:, :, None], 3, axis=2).copy()
        render[want[:, 1], want[:, 0]] = (255, 0, 0)
        assert np.array_equal(rgb, render)
        lines = (tmp_path / f"out_{mode}.txt").read_text().split("\n")[:-1]
        assert lines == [f"{x} {y}" for x, y in want.tolist()]
    # the reference panics for count < 9 (exit status 101 is what a Rust panic gives)
    res = subprocess.run([cli, str(pgm), str(tmp_path / "x.ppm"), "16", "8", "off"], capture_output=True, text=True,
                         timeout=120)
    assert res.returncode == 101 and "needs to exceed 9" in res.stderr


@pytest.mark.parametrize("sr", [32, 48, 64])
def test_forced_strip_heights(detector, oracle_mod, sr):
    """The launcher picks 32- or 64-row strips by batch size; every compiled strip height must give the same list."""
    detector.set_tuning(strip_rows=sr)
    img = oracle_mod.synth_frame(1920, 1080, seed=77, frame=sr, kind=0, amp=5)
    for nms in (0, 1, 2):
        assert same_points(detector.detect_array(img, _cfg(16, 9, nms)), oracle_mod.port_detect(img, 16, 9, nms)), nms
    # heights around the strip boundaries of this strip height
    for h in (sr - 1, sr, sr + 1, sr + 5, sr + 6, sr + 7, 2 * sr + 3):
        small = oracle_mod.synth_frame(300, h, seed=h, frame=1, kind=1)
        for nms in (0, 1):
            assert same_points(detector.detect_array(small, _cfg(25, 9, nms)), oracle_mod.detect(small, 25, 9, nms)), (h, nms)
    detector.set_tuning()


@pytest.mark.parametrize("sr", [32, 64])
def test_dense_and_sparse_chunks_in_one_strip(detector, oracle_mod, sr):
    """Noise next to flat and scene content: the same CTA alternates between the candidate-queue path and the
    dense path (queue overflow), in all three modes, and the staging blocks are reused across both."""
    detector.set_tuning(strip_rows=sr)
    scene = oracle_mod.synth_frame(1500, 200, seed=5, frame=0, kind=0, amp=4)
    noise = oracle_mod.synth_frame(1500, 200, seed=6, frame=0, kind=1)
    img = scene.copy()
    img[:, 300:700] = noise[:, 300:700]
    img[:, 900:1000] = 128
    img[:, 1200:1500] = noise[:, 1200:1500]
    for t, nms in [(5, 0), (5, 1), (5, 2), (16, 1)]:
        assert same_points(detector.detect_array(img, _cfg(t, 9, nms)), oracle_mod.port_detect(img, t, 9, nms)), (t, nms)
    assert detector.device_flags() == 0
    detector.set_tuning()


def test_noise_batch_takes_the_fallback_everywhere(detector, oracle_mod):
    """A resident batch of noise frames (64-row strips, every chunk overflows the queue): counts and hashes per frame."""
    import torch

    n = 40
    frames = detector.synth_frames(n, 1280, 720, seed=31, first_frame=0, kind=1, amp=0)
    for nms in (0, 1):
        pts, offs = detector.detect_device(frames, _cfg(12, 9, nms))
        torch.cuda.synchronize()
        assert detector.device_flags() == 0
        offs_h = offs.cpu().numpy()
        pts_h = pts[: int(offs_h[-1])].cpu().numpy().astype(np.uint32)
        counts, hashes = oracle_mod.port_detect_many(frames.cpu().numpy(), 12, 9, nms, n_threads=8)
        assert (np.diff(offs_h) == counts).all()
        for f in range(n):
            assert oracle_mod.hash_points(pts_h[offs_h[f]:offs_h[f + 1]]) == int(hashes[f]), (nms, f)


def test_host_batch_is_pipelined_in_sub_batches(detector, oracle_mod):
    """fdf_detect_batch splits large batches into sub-batches (copy of the next one overlaps the kernels of the
    current one); the packed CSR output must not depend on the split, including when the capacity runs out."""
    import feature_detector_fast_b200 as fdf

    frames = np.stack([oracle_mod.synth_frame(1280, 720, seed=12, frame=f, kind=0, amp=4) for f in range(7)])
    detector.set_tuning(sub_batch_mb=4096)
    pts1, offs1 = detector.detect_batch(frames, _cfg(16, 9, 1))
    detector.set_tuning(sub_batch_mb=2)  # 2 frames per sub-batch -> 4 sub-batches
    pts4, offs4 = detector.detect_batch(frames, _cfg(16, 9, 1))
    assert (np.asarray(offs1) == np.asarray(offs4)).all()
    assert same_points(pts1, pts4)
    for f in (0, 3, 6):
        assert same_points(pts4[int(offs4[f]):int(offs4[f + 1])], oracle_mod.port_detect(frames[f], 16, 9, 1))
    total = int(offs4[-1])
    small = np.zeros((total // 2, 2), np.uint32)
    with pytest.raises(fdf.FdfError) as ei:
        detector.detect_batch(frames, _cfg(16, 9, 1), out=small)
    assert ei.value.status == 4  # FDF_ERR_CAPACITY
    detector.set_tuning()


# ---- the step in front of the path: RGB8 -> luma8 (main.rs:53-58), fdf_rgb8_to_luma8_device / fdf_detect_rgb8 ----
def _luma_restated(rgb: np.ndarray) -> np.ndarray:
    """image 0.24.6 `to_luma8`: (2126 r + 7152 g + 722 b) / 10000 in u32, truncating (restated; the crate is not vendored)."""
    r, g, b = (rgb[..., i].astype(np.uint32) for i in range(3))
    return ((2126 * r + 7152 * g + 722 * b) // 10000).astype(np.uint8)


def test_rgb8_to_luma8_device_matches_the_restated_weights(detector):
    import torch

    rng = np.random.default_rng(11)
    for (f, h, w) in [(1, 9, 7), (2, 33, 61), (1, 40, 128), (3, 17, 250)]:
        rgb = rng.integers(0, 256, (f, h, w, 3), dtype=np.uint8)
        rgb[0, 0, 0] = (255, 255, 255)
        rgb[0, 0, 1] = (0, 0, 0)
        got = detector.rgb8_to_luma8_device(torch.from_numpy(rgb).cuda())
        assert got.stride(1) % 16 == 0
        assert np.array_equal(got.cpu().numpy(), _luma_restated(rgb)), (f, h, w)
        # util.rs:5-41: Rgb8ToLuma16View pixel = r + g + b (u16), to_grey = that / 3 as u8
        got3 = detector.rgb8_to_luma8_device(torch.from_numpy(rgb).cuda(), sum3=True)
        want3 = (rgb.astype(np.uint16).sum(axis=3) // 3).astype(np.uint8)
        assert np.array_equal(got3.cpu().numpy(), want3), (f, h, w)
    # unaligned source rows (a view into a wider buffer): the byte path of the kernel
    wide = torch.from_numpy(rng.integers(0, 256, (1, 20, 71, 3), dtype=np.uint8)).cuda()
    view = wide[:, :, 1:66, :]
    got = detector.rgb8_to_luma8_device(view.contiguous())
    assert np.array_equal(got.cpu().numpy(), _luma_restated(view.cpu().numpy()))


def test_detect_rgb8_is_detect_of_the_luma_image(detector, golden, oracle_mod):
    import feature_detector_fast_b200 as fdf

    grey = golden["grey"]
    # r == g == b: the conversion is the identity (what the shipped PNG exercises), so the golden lists must come out
    rgb = np.repeat(grey[:, :, None], 3, axis=2)
    assert same_points(detector.detect_rgb8_array(rgb, fdf.Config(16, 9, fdf.NonMaximalSuppression.Off)), golden["rust_off"])
    assert same_points(detector.detect_rgb8_array(rgb, fdf.Config(16, 9, fdf.NonMaximalSuppression.MaxThreshold)),
                       golden["rust_nonmax"])
    # a colour image: same as the oracle on the restated luma
    rng = np.random.default_rng(5)
    base = oracle_mod.synth_frame(333, 97, 3, 0, 0, 4).astype(np.int16)
    col = np.stack([np.clip(base + rng.integers(-30, 30, base.shape), 0, 255) for _ in range(3)], axis=2).astype(np.uint8)
    luma = _luma_restated(col)
    for nms in (0, 1, 2):
        cfg = fdf.Config(12, 9, fdf.NonMaximalSuppression(nms))
        assert same_points(detector.detect_rgb8_array(col, cfg), oracle_mod.detect(luma, 12, 9, nms)), nms
    with pytest.raises(fdf.FdfPanic):
        detector.detect_rgb8_array(col, fdf.Config(12, 8, fdf.NonMaximalSuppression.Off))


def test_config5_full_size_properties(detector):
    """BASELINE config 5 at its FULL size (512 resident 3840x2160 frames, t=20, n=9, MaxThreshold; 4.2 GB of pixels,
    ~9.5 M keypoints): too big for the oracle, so size-independent properties are checked on the device --
    (1) CSR offsets are non-decreasing and end at the number of points written; (2) every frame's list is strictly
    increasing in (y, x), i.e. row-major and duplicate-free, with centres inside the NMS emit range; (3) detection is
    a pure function of the frame: the 256 frames that a second batch (generator frames 256..767) shares with the
    first one (0..511) give identical lists at different positions of the batch; (4) a second run is bit-identical."""
    import torch

    F, W, H = 512, 3840, 2160
    cfg = _cfg(20, 9, 1)
    a = detector.synth_frames(F, W, H, seed=99, first_frame=0, kind=0, amp=4)
    pts_a, offs_a = detector.detect_device(a, cfg, points=torch.empty((F * 60000, 2), dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    assert detector.device_flags() == 0
    offs_a = offs_a.clone()
    total = int(offs_a[-1])
    assert 0 < total <= pts_a.shape[0]
    assert int(offs_a[0]) == 0 and bool((offs_a[1:] >= offs_a[:-1]).all())
    p = pts_a[:total].to(torch.int64)
    frame_of = torch.searchsorted(offs_a[1:].contiguous(), torch.arange(total, device="cuda"), right=True)
    key = (frame_of * H + p[:, 1]) * W + p[:, 0]
    assert bool((key[1:] > key[:-1]).all())  # row-major inside each frame, frames in order, no duplicates
    assert int(p[:, 0].min()) >= 3 and int(p[:, 0].max()) < W - 3 and int(p[:, 1].min()) >= 4 and int(p[:, 1].max()) <= H - 5
    first = pts_a[:total].clone()
    del a, p, key, frame_of
    # (3) the same frames at other batch positions
    b = detector.synth_frames(F, W, H, seed=99, first_frame=256, kind=0, amp=4)
    pts_b, offs_b = detector.detect_device(b, cfg, points=torch.empty((F * 60000, 2), dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    assert detector.device_flags() == 0
    lo_a, hi_a = int(offs_a[256]), int(offs_a[512])
    hi_b = int(offs_b[256])
    assert hi_a - lo_a == hi_b
    assert bool((offs_a[256:] - offs_a[256] == offs_b[:257]).all())
    assert torch.equal(first[lo_a:hi_a], pts_b[:hi_b])
    # (4) idempotence
    pts_c, offs_c = detector.detect_device(b, cfg, points=torch.empty((F * 60000, 2), dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    assert torch.equal(offs_b, offs_c) and torch.equal(pts_b[: int(offs_b[-1])], pts_c[: int(offs_c[-1])])


def test_very_wide_images_take_the_multi_round_gather(detector, oracle_mod):
    """Widths beyond ~8000 pixels: more level-2 bitmap words than gather threads (several rounds of the block prefix
    sum in fdf_gather_kernel), many chunks per strip (tag wrap, > 32 run records per strip)."""
    for (w, h, kind, t) in [(9000, 41, 0, 16), (16001, 23, 0, 16), (12345, 70, 1, 60)]:
        img = oracle_mod.synth_frame(w, h, seed=w, frame=0, kind=kind, amp=5)
        for nms in (0, 1, 2):
            assert same_points(detector.detect_array(img, _cfg(t, 9, nms)), oracle_mod.port_detect(img, t, 9, nms)), (w, h, nms)


def test_randomised_shapes_contents_and_configs(detector, oracle_mod):
    """Seeded fuzz: 500 random (width, height, content, threshold, count, mode) cases against the oracle's AVX2 port
    (itself pinned to the scalar restatement in the CPU tier); every 10th case also against the scalar oracle."""
    rng = np.random.default_rng(20261018)
    for case in range(500):
        w, h = int(rng.integers(7, 900)), int(rng.integers(7, 260))
        style = case % 5
        if style == 0:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)                       # uniform noise (dense fallback)
        elif style == 1:
            img = oracle_mod.synth_frame(w, h, case, 0, 0, int(rng.integers(0, 12)))  # scene
        elif style == 2:
            img = (rng.integers(0, 2, (h, w)) * rng.integers(1, 256)).astype(np.uint8)  # two levels: exact score ties
        elif style == 3:
            img = np.clip(rng.normal(128, rng.integers(1, 60), (h, w)), 0, 255).astype(np.uint8)  # gaussian noise
        else:
            img = oracle_mod.synth_frame(w, h, case, 1, 0, 4)
            img[:: int(rng.integers(2, 9)), :] = int(rng.integers(0, 256))            # scene with flat rows
        t = int(rng.choice([0, 1, 2, 5, 10, 16, 20, 31, 64, 100, 127, 128, 129, 200, 254, 255]))
        n, nms = int(rng.integers(9, 17)), int(rng.integers(0, 3))
        got = detector.detect_array(img, _cfg(t, n, nms))
        assert same_points(got, oracle_mod.port_detect(img, t, n, nms)), (case, w, h, style, t, n, nms)
        if case % 10 == 0:
            assert same_points(got, oracle_mod.detect(img, t, n, nms)), (case, w, h, style, t, n, nms)
    assert detector.device_flags() == 0


def test_shard_push_assembles_the_batch_result_on_one_gpu(detector, oracle_mod):
    """The multi-GPU exchange step (fdf_shard_push, include/fdf.h) with the ranks simulated on ONE GPU: the batch is cut
    into sharding.frame_shard blocks (uneven, and one rank without frames), every "rank" detects its block into a local
    buffer with local offsets, the blocks are laid out as the all-gather would, and every rank's push writes into the
    same result buffer.  The assembled points and global offsets must equal one detection of the whole batch.
    (tests/test_multi_gpu.py does the same with real ranks, NCCL and peer memory when two GPUs are visible.)"""
    import ctypes as C

    import torch

    import feature_detector_fast_b200 as fdf
    from feature_detector_fast_b200 import sharding

    lib, ctx = detector._lib, detector._ctx
    for n_frames, world in ((7, 3), (2, 3), (5, 1)):
        frames = detector.synth_frames(n_frames, 656, 210, seed=91, kind=0)
        for nms in (0, 1):
            cfg = _cfg(18, 9, nms)
            want_pts, want_offs = detector.detect_device(frames, cfg)
            torch.cuda.synchronize()
            total = int(want_offs[-1])
            block = sharding.shard_block(n_frames, world)
            all_offs = torch.zeros(world * block, dtype=torch.int64, device="cuda")
            local = []
            for r in range(world):
                lo, hi = sharding.frame_shard(n_frames, r, world)
                pts = torch.zeros((max(1, total), 2), dtype=torch.int32, device="cuda")
                if hi > lo:
                    _, offs = detector.detect_device(frames[lo:hi], cfg, points=pts,
                                                     offsets=all_offs[r * block: r * block + hi - lo + 1])
                local.append(pts)
            result = torch.full((total + 3, 2), -1, dtype=torch.int32, device="cuda")
            for r in range(world):
                goffs = torch.zeros(n_frames + 1, dtype=torch.int64, device="cuda")
                st = lib.fdf_shard_push(ctx, all_offs.data_ptr(), block, world, r, n_frames, local[r].data_ptr(),
                                        result.data_ptr(), total, goffs.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
                assert st == 0, lib.fdf_last_error(ctx)
                torch.cuda.synchronize()
                assert torch.equal(goffs, want_offs), (n_frames, world, r)
            assert torch.equal(result[:total], want_pts[:total]), (n_frames, world, nms)
            assert bool((result[total:] == -1).all())  # nothing written past the batch


def test_pipe_streams_images_first_in_first_out(detector, oracle_mod):
    """fdf_pipe_* (streaming host API): images of different sizes, contents and configs kept in flight,
    collected in submission order, each list identical to what fdf_detect returns for the same image; submit on a full
    pipe is FDF_ERR_BUSY, collect on an empty one an argument error, an image beyond the pipe's size is refused; pinned
    and pageable sources; an image too small to hold a keypoint; a result larger than the first asynchronous copy."""
    import torch

    import feature_detector_fast_b200 as fdf

    shapes = [(640, 360, 0, 16), (333, 77, 0, 12), (6, 40, 0, 16), (512, 300, 1, 3), (640, 360, 0, 20), (97, 211, 0, 9),
              (640, 352, 0, 16)]
    imgs = [oracle_mod.synth_frame(w, h, seed=5, frame=i, kind=k, amp=5) for i, (w, h, k, _) in enumerate(shapes)]
    cfgs = [_cfg(t, 9 + i % 3, i % 3) for i, (_, _, _, t) in enumerate(shapes)]
    wide = np.zeros((77, 400), np.uint8)  # the second image is a strided view (pitch 400 > width 333)
    wide[:, :333] = imgs[1]
    imgs[1] = wide[:, :333]
    assert imgs[1].strides[0] == 400
    pinned = torch.empty((352, 640), dtype=torch.uint8, pin_memory=True)  # the last image lives in pinned memory
    pinned.copy_(torch.from_numpy(imgs[-1]))
    imgs[-1] = pinned.numpy()
    want = [detector.detect_array(a, c) for a, c in zip(imgs, cfgs)]
    assert len(want[3]) > 4096 + 1024  # the noise image needs the second, synchronous copy on its first appearance
    pipe = detector.pipe(depth=3, max_w=640, max_h=360, cap=max(len(x) for x in want))
    try:
        with pytest.raises(fdf.FdfError) as ei:
            pipe.collect()
        assert ei.value.status == 3
        with pytest.raises(fdf.FdfError) as ei:
            pipe.submit(np.zeros((361, 640), np.uint8), cfgs[0])
        assert ei.value.status == 3 and pipe.in_flight == 0
        with pytest.raises(fdf.FdfPanic):
            pipe.submit(imgs[0], _cfg(16, 8, 0))
        got = []
        for rounds in range(2):  # twice: every slot is reused
            for a, c in zip(imgs, cfgs):
                if pipe.in_flight == pipe.depth:
                    with pytest.raises(fdf.FdfError) as ei:
                        pipe.submit(a, c)
                    assert ei.value.status == 8  # FDF_ERR_BUSY, nothing enqueued
                    assert pipe.in_flight == pipe.depth
                    got.append(pipe.collect())
                pipe.submit(a, c)
            while pipe.in_flight:
                got.append(pipe.collect())
        assert len(got) == 2 * len(imgs)
        for i, g in enumerate(got):
            assert same_points(g, want[i % len(imgs)]), i
        assert detector.device_flags() == 0
        # a pipe created for fewer keypoints than an image has: FDF_ERR_CAPACITY with the number found, image retired
        small = detector.pipe(depth=2, max_w=640, max_h=360, cap=100)
        small.submit(imgs[0], cfgs[0])
        with pytest.raises(fdf.FdfError) as ei:
            small.collect()
        assert ei.value.status == 4 and str(len(want[0])) in str(ei.value) and small.in_flight == 0
        small.close()
    finally:
        pipe.close()


def test_idle_sms_do_not_change_the_result(detector, oracle_mod):
    """fdf_set_idle_sms (the detection kernel leaves every n-th SM to the exchange kernels of a sharded batch): the CTAs
    that land there return before taking a ticket, the others do all the strips -- same points, same order, no flag."""
    import torch

    frames = detector.synth_frames(6, 3840, 2160, seed=404, kind=0)
    cfg = _cfg(20, 9, 1)
    want_pts, want_offs = detector.detect_device(frames, cfg)
    torch.cuda.synchronize()
    k = int(want_offs[-1])
    try:
        for stride in (24, 2, 148):
            detector.set_idle_sms(stride)
            pts, offs = detector.detect_device(frames, cfg)
            torch.cuda.synchronize()
            assert torch.equal(offs, want_offs) and torch.equal(pts[:k], want_pts[:k]), stride
            assert detector.device_flags() == 0
    finally:
        detector.set_idle_sms(0)


def test_item_parts_do_not_change_the_result(detector, oracle_mod):
    """fdf_set_item_parts: the detection kernel's work items are whole strips or 2 / 4 / 8 equal chunk ranges of a strip
    (automatic for small batches).  Same points, same order, same offsets for every split -- sparse and dense content,
    every NMS mode, widths whose chunk count does and does not divide, one image and a batch."""
    import torch

    cases = [(detector.synth_frames(3, 3840, 2160, seed=77, kind=0), 20),    # 16 chunks per strip
             (detector.synth_frames(1, 1920, 1080, seed=78, kind=0), 16),    # 8 chunks, one image
             (detector.synth_frames(2, 1936, 200, seed=79, kind=1), 3),      # noise: dense path
             (detector.synth_frames(2, 1456, 300, seed=80, kind=0)[:, :, :1450], 12)]  # 7 chunks: cannot be split evenly
    try:
        for frames, t in cases:
            for nms in (0, 1, 2):
                cfg = _cfg(t, 9, nms)
                detector.set_item_parts(1)
                want_pts, want_offs = detector.detect_device(frames, cfg)
                torch.cuda.synchronize()
                k = int(want_offs[-1])
                assert k > 0
                for parts in (2, 4, 8, 0):
                    detector.set_item_parts(parts)
                    pts, offs = detector.detect_device(frames, cfg)
                    torch.cuda.synchronize()
                    assert torch.equal(offs, want_offs), (tuple(frames.shape), nms, parts)
                    assert torch.equal(pts[:k], want_pts[:k]), (tuple(frames.shape), nms, parts)
                    assert detector.device_flags() == 0
        # and against the CPU oracle for one split case
        detector.set_item_parts(4)
        f = cases[1][0]
        pts, offs = detector.detect_device(f, _cfg(16, 9, 1))
        torch.cuda.synchronize()
        assert same_points(pts[: int(offs[1])].cpu().numpy(), oracle_mod.port_detect(f[0].cpu().numpy(), 16, 9, 1))
    finally:
        detector.set_item_parts(0)
