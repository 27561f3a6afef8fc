#!/usr/bin/env python3
"""bench.py -- throughput of the FAST-n detection path on B200, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4]): a batch of synthetic 3840x2160 grey frames, threshold 20, count 9,
max-threshold NMS.  One "step" = one pass of the detection path over this rank's resident batch
(512 frames per GPU; frames are independent units, so ranks share nothing on the data path and the headline is
weak-scaled: N GPUs process N x 512 frames per step).  For N > 1 a step goes through the product's sharded path
(feature_detector_fast_b200.sharding.ShardedDetector): detection per rank, ONE NCCL all-gather of the ranks' local CSR
offsets, and every rank's emission kernel writes its points straight into rank 0's result buffer over NVLink -- one
batch result.  `strong` repeats the measurement for the literal config 5: ONE 512-frame batch sharded over the N GPUs.

  value        Mpix/s, whole job, frames already resident in HBM when the timed region starts
  e2e          same metric through the C-ABI call `fdf_detect_batch` on HOST (pinned) buffers: host->device
               copy of the frames and device->host copy of offsets + keypoints inside the timed region
  roofline     dominant kernel (fdf_detect_kernel): algorithmic bytes (W*H per frame read once + 8 B per
               keypoint + 8 B per frame offset) / CUDA-event duration of that launch (events recorded by the
               library on the launching stream during the timed steps), against the measured HBM copy
               bandwidth in MEASURED_PEAKS.json
  cpu_baseline AVX2 port of the reference's fast_simd.rs (oracle/fdf_avx2_port.cpp) on the host cores,
               bounded sample of the same frames (rank 0, N = 1 only); reported, not the target.  It also re-checks the
               GPU result: per-frame hashes (tests/compare.rs:5-20 format) of the sampled frames must be equal.
  by_config    (rank 0, N = 1) the other BASELINE configs on resident synthetic 1080p frames: configs 1-3 (t16 n9, the
               three NMS modes), config 4 (n = 9..16 x three modes), the uniform-noise stress frame, and the crate's
               criterion triple (benches/benchmark.rs: one frame per call through fdf_detect, mean and 95 % CI)

`--impl reference` times that CPU port alone (the reference crate is Rust and cannot be built in this
image, so the port stands in for it), with all host threads, on bounded samples of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 3840, 2160
THRESHOLD, COUNT, NMS = 20, 9, 1
FRAMES_PER_GPU = 512
SEED = 20240
METRIC = "Mpixels/sec (FAST-n detection; also 1080p frames/sec and HBM GB/s vs peak)"
WORKLOAD = "configs[4]: 512 synthetic 3840x2160 frames per GPU, t=20, n=9, max-threshold NMS"


def workload_config(frames_per_gpu):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "frames_per_gpu": frames_per_gpu, "width": W, "height": H, "threshold": THRESHOLD,
            "count": COUNT, "nms": "max_threshold",
            "l2": "inputs larger than L2 (4.2 GB per GPU per step vs 126 MB)"}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).

    nvidia-smi takes a few hundred ms to start, so it is launched early (`start`, then `wait_ready`) and only
    the samples whose timestamps fall inside [mark_begin, mark_end] are used."""

    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_ready(self, timeout=5.0):
        if self.proc is None:
            return
        end = time.time() + timeout
        while time.time() < end:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.02)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            time.sleep(0.05)
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, reasons, smax, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.t0 is not None and not (self.t0 - 0.02 <= ts <= self.t1 + 0.02):
                    continue
                sm.append(float(r[2]))
                smax = float(r[3])
                power.append(float(r[4]))
            except ValueError:
                continue
            for name, v in zip(names, r[6:10]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power) if power else None)
        return out


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: pin this rank to the CPUs that are local to its GPU (sysfs local_cpulist of the GPU's PCI
    function) BEFORE the pinned host buffers are allocated, so that their pages sit on the GPU's NUMA node and the
    end-to-end copies do not cross the socket interconnect.  Returns the cpu list it bound to, or None."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "%d cpus local to GPU %d" % (len(cpus), local_rank)
    except Exception:
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def criterion_triple(run, label):
    """benches/benchmark.rs restated: the crate's three criterion functions -- simd_t16_c_9_off /
    _max_threshold / _sum_abs, one 1920x1080 frame per iteration -- timed criterion-style (warm-up, then samples; mean
    and a 95 % confidence interval of the mean).  `run(nms)` performs one iteration and returns the keypoint count."""
    import math

    names = {0: "simd_t16_c_9_off", 1: "simd_t16_c_9_max_threshold", 2: "simd_t16_c_9_sum_abs"}
    out = {}
    for nms, name in names.items():
        t_end = time.perf_counter() + 0.3  # warm-up
        found = run(nms)
        while time.perf_counter() < t_end:
            found = run(nms)
        samples = []
        t_end = time.perf_counter() + 1.0
        while len(samples) < 100 or (time.perf_counter() < t_end and len(samples) < 2000):
            t0 = time.perf_counter()
            run(nms)
            samples.append((time.perf_counter() - t0) * 1e3)
        mean = sum(samples) / len(samples)
        sd = math.sqrt(sum((x - mean) ** 2 for x in samples) / (len(samples) - 1))
        half = 1.96 * sd / math.sqrt(len(samples))
        out[name] = {"ci95_ms": [round(mean - half, 4), round(mean, 4), round(mean + half, 4)], "samples": len(samples),
                     "keypoints": int(found), "mpix_per_s": round(1920 * 1080 / mean / 1e3, 1)}
    out["protocol"] = ("benches/benchmark.rs restated: one synthetic 1920x1080 scene frame per call (the reference's "
                       "screenshot is not shipped), t=16 n=9, warm-up then >= 100 samples; " + label)
    return out


def synth_frames_host(oracle, n, w, h, first=0, threads=1):
    """n synthetic frames (F, H, W) generated on the host cores, with slack after the last frame (the AVX2 port's gathers
    over-read by <= 3 bytes)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor

    pad = np.zeros(n * h * w + 64, np.uint8)
    frames = pad[: n * h * w].reshape(n, h, w)

    def fill(f):
        frames[f] = oracle.synth_frame(w, h, SEED, first + f, 0, 4)

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(fill, range(n)))
    return frames


def time_cpu_port(frames_np, n_threads, min_seconds, oracle, want_hashes=False):
    """AVX2 port over `frames_np` (F, H, W) with n_threads workers, repeated for >= min_seconds."""
    f = frames_np.shape[0]
    reps, t0 = 0, time.perf_counter()
    counts = hashes = None
    while True:
        counts, hashes = oracle.port_detect_many(frames_np, THRESHOLD, COUNT, NMS, n_threads=n_threads,
                                                 want_hashes=want_hashes)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return (reps * f * W * H) / dt / 1e6, dt, reps, counts, hashes


def run_reference_arm(args, rank):
    """CPU port of the reference path, all host threads, the full workload per step (512 frames)."""
    if rank != 0:
        return 0
    import numpy as np

    import oracle

    oracle.build()
    cores = host_threads()
    n = args.frames
    frames = synth_frames_host(oracle, n, W, H, threads=cores)
    for _ in range(args.warmup):
        oracle.port_detect_many(frames, THRESHOLD, COUNT, NMS, n_threads=cores, want_hashes=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.port_detect_many(frames, THRESHOLD, COUNT, NMS, n_threads=cores, want_hashes=False)
    dt = time.perf_counter() - t0
    mpix = args.steps * n * W * H / dt / 1e6
    sample = f"the workload's {n} 3840x2160 frames per step, one frame per thread at a time, {cores} threads"
    frame1080 = synth_frames_host(oracle, 1, 1920, 1080)[0]
    crit = criterion_triple(lambda nms: len(oracle.port_detect(frame1080, 16, 9, nms)),
                            "CPU port of fast_simd.rs on one thread of this host")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mpix, 2), "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(n),
        "note": "CPU port of fast_simd.rs (the Rust reference cannot be built here: no cargo/rustc); the CPU arm is one "
                "host whatever --gpus says",
        "cpu_baseline": {"value": round(mpix, 2), "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(mpix, 2), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps_1080p_equiv": round(mpix * 1e6 / (1920 * 1080), 1),
        "criterion": crit,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_criterion(args, rank):
    """`--criterion`: only the benches/benchmark.rs triple, one JSON line per arm."""
    if rank != 0:
        return 0
    import oracle  # (synthetic frame generator; the CPU arm also times the port -- bench.py's CPU-baseline role)

    oracle.build()
    frame = synth_frames_host(oracle, 1, 1920, 1080)[0]
    if args.impl == "b200":
        import feature_detector_fast_b200 as fdf

        det = fdf.Detector(0)
        crit = criterion_triple(lambda nms: len(det.detect_array(frame, fdf.Config(16, 9, fdf.NonMaximalSuppression(nms)))),
                                "fdf_detect through the C ABI, host buffers, copies included")
    else:
        crit = criterion_triple(lambda nms: len(oracle.port_detect(frame, 16, 9, nms)),
                                "CPU port of fast_simd.rs on one thread of this host")
    print(json.dumps({"criterion": crit, "impl": args.impl}), flush=True)
    return 0


def bench_by_config(det, fdf, torch, peak):
    """The other BASELINE configs, device-resident, on synthetic 1080p frames (rank 0, N = 1).  Every entry: Mpix/s of
    the whole step (three kernels), the detection kernel's milliseconds and its fraction of the HBM roofline."""
    w, h, f = 1920, 1080, 256
    frames = det.synth_frames(f, w, h, seed=SEED + 1, first_frame=0, kind=0, amp=4)
    points = torch.empty((f * 60000, 2), dtype=torch.int32, device=frames.device)
    offsets = torch.empty(f + 1, dtype=torch.int64, device=frames.device)
    names = {0: "off", 1: "max_threshold", 2: "sum_absolute"}

    def one(frames, t, n, nms, steps=5):
        cfg = fdf.Config(t, n, fdf.NonMaximalSuppression(nms))
        nf = frames.shape[0]
        for _ in range(2):
            det.detect_device(frames, cfg, points=points, offsets=offsets)
        torch.cuda.synchronize()
        found = int(offsets[nf].item())
        if found > points.shape[0] or det.device_flags() != 0:
            raise SystemExit("bench.py: by_config run invalid (capacity or device flags)")
        det.set_timing(steps)
        for _ in range(steps):
            det.detect_device(frames, cfg, points=points, offsets=offsets)
        torch.cuda.synchronize()
        ms = [det.get_timing(i) for i in range(steps)]
        det.set_timing(0)
        k = sum(m[0] for m in ms) / steps
        step = sum(sum(m) for m in ms) / steps
        algo = nf * w * h + 8 * found + 8 * (nf + 1)
        return {"mpix_per_s": round(nf * w * h / step / 1e3, 1), "detect_kernel_ms": round(k, 4), "step_ms": round(step, 4),
                "roofline_frac": round(algo / (k * 1e-3) / 1e9 / peak, 4), "keypoints_per_frame": round(found / nf, 1)}

    out = {"frames": f"{f} resident synthetic 1920x1080 scene frames (seed {SEED + 1}); the reference's 1080p screenshot is "
                     "not shipped, its published counts 23184 / 7646 / 8307 are not reproducible"}
    out["configs_1_to_3_t16_n9"] = {names[nms]: one(frames, 16, 9, nms) for nms in (0, 1, 2)}
    sweep = {}
    for n in range(9, 17):
        sweep[f"n{n}"] = {names[nms]: one(frames[:64], 16, n, nms, steps=3) for nms in (0, 1, 2)}
    out["config_4_count_sweep_t16"] = {"frames": 64, "by_count": sweep}
    noise = det.synth_frames(32, w, h, seed=SEED + 2, first_frame=0, kind=1, amp=0)
    del points
    points = torch.empty((32 * w * h // 3, 2), dtype=torch.int32, device=frames.device)
    out["uniform_noise_stress_t20_n9_max_threshold"] = one(noise, 20, 9, 1, steps=3)
    return out


def streaming_pipe(det, fdf, frame_pageable, frame_pinned, depth=4, n=3000):
    """fdf_pipe_* (the streaming form of `detect`): 1080p frames kept `depth` in flight, frames per second
    through the C ABI with host buffers, copies included -- the criterion workload without the per-call round trip."""
    out = {"depth": depth, "frames": n, "protocol": "one synthetic 1920x1080 scene frame submitted over and over, t=16 n=9, "
           "`depth` images in flight, collect-then-submit loop; host image -> keypoints in host memory"}
    for name, frame in (("pageable_input", frame_pageable), ("pinned_input", frame_pinned)):
        for nms in (0, 1, 2):
            cfg = fdf.Config(16, 9, fdf.NonMaximalSuppression(nms))
            pipe = det.pipe(depth, 1920, 1080)
            for _ in range(200):  # warm-up, and the first-copy size settles
                if pipe.in_flight == depth:
                    pipe.collect()
                pipe.submit(frame, cfg)
            t0 = time.perf_counter()
            k = 0
            for _ in range(n):
                if pipe.in_flight == depth:
                    k = len(pipe.collect())
                pipe.submit(frame, cfg)
            while pipe.in_flight:
                k = len(pipe.collect())
            dt = time.perf_counter() - t0
            pipe.close()
            out.setdefault(name, {})[("off", "max_threshold", "sum_absolute")[nms]] = {
                "frames_per_s": round(n / dt, 1), "us_per_frame": round(dt / n * 1e6, 2),
                "mpix_per_s": round(n * 1920 * 1080 / dt / 1e6, 1), "keypoints": k}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--criterion", action="store_true",
                    help="benches/benchmark.rs protocol (three 1080p single-frame functions) instead of the batch workload")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU (default: the named config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-by-config", action="store_true")
    ap.add_argument("--strong-frames", type=int, default=FRAMES_PER_GPU,
                    help="total frames of the strong-scaling batch (default: the 512 of config 5)")
    ap.add_argument("--idle-sm-stride", type=int, default=37,
                    help="N > 1: the detection kernel leaves every n-th SM to the exchange kernels (0 = none)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.criterion:
        return run_criterion(args, rank)
    if args.impl == "reference":
        return run_reference_arm(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist

    import feature_detector_fast_b200 as fdf
    from feature_detector_fast_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    det = fdf.Detector(local_rank)
    cfg = fdf.Config(THRESHOLD, COUNT, fdf.NonMaximalSuppression.MaxThreshold)
    F = args.frames
    n_total = F * world
    per_frame_cap = 100000  # ~2.4x the expected keypoints of this workload (about 18.6k per 4K frame); checked below
    frames = det.synth_frames(F, W, H, seed=SEED, first_frame=rank * F, kind=0, amp=4)
    cap = F * per_frame_cap
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # One step of the product path.  N = 1: fdf_detect_device.  N > 1: the sharded path (one batch result on rank 0).
    sd = None
    if world == 1:
        points = torch.empty((cap, 2), dtype=torch.int32, device=dev)
        offsets = torch.empty(F + 1, dtype=torch.int64, device=dev)

        def step():
            det.detect_device(frames, cfg, points=points, offsets=offsets)
    else:
        sd = sharding.ShardedDetector(det, n_total, cap_total=n_total * per_frame_cap, cap_local=cap,
                                      idle_sm_stride=args.idle_sm_stride)

        def step():
            sd.detect(frames, cfg)

    def timed_steps(step_fn, steps, fence=None):
        """exactly `steps` steps between barriers; CUDA events on the launching stream; max over ranks"""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            step_fn()
        if fence is not None:
            fence()  # (rank 0 may read the other ranks' points only after this; it is part of the timed region)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt[0])
        return ms

    for _ in range(args.warmup):
        step()
    if sd is not None:
        sd.fence()
    torch.cuda.synchronize()
    if world == 1:
        found = int(offsets[-1].item())
        local_found = found
    else:
        goffs = sd.global_offsets
        found = int(goffs[-1].item())
        lo, hi = sharding.frame_shard(n_total, rank, world)
        local_found = int((goffs[hi] - goffs[lo]).item())
    if local_found > cap or det.device_flags() != 0:
        raise SystemExit(f"bench.py: invalid run (found {local_found} > cap {cap} or device flags set)")

    # ---- device-resident timing: exactly K steps, CUDA events, max over ranks --------------------
    sampler = ClockSampler(local_rank)
    det.set_timing(args.steps)  # CUDA events around each launch, recorded by the library on the launching stream
    launches0 = det.kernel_launches
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    barrier()
    sampler.mark_begin()
    total_ms = timed_steps(step, args.steps, fence=sd.fence if sd is not None else None)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = det.kernel_launches - launches0
    per_launch = [det.get_timing(i) for i in range(args.steps)]
    det.set_timing(0)
    kern_ms = sum(t[0] for t in per_launch) / args.steps        # the dominant kernel: fdf_detect_kernel
    scan_ms = sum(t[1] for t in per_launch) / args.steps
    gather_ms = sum(t[2] for t in per_launch) / args.steps       # (0 in the sharded path: its emission launch is not timed)
    if world > 1:
        tt = torch.tensor([kern_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        kern_ms = float(tt[0])
    ms_per_step = total_ms / args.steps
    mpix = n_total * W * H / (ms_per_step * 1e-3) / 1e6

    peak, peak_src = measured_hbm_peak()
    algo_bytes = F * W * H + 8 * local_found + 8 * (F + 1)
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": committed_traffic(), "peak_source": peak_src,
                "kernel": "fdf_detect_kernel<MaxThreshold,64>", "kernel_ms": round(kern_ms, 4),
                "other_kernels_ms": {"fdf_scan_kernel": round(scan_ms, 4), "fdf_gather_kernel": round(gather_ms, 4)},
                "algorithmic_bytes_per_launch": algo_bytes,
                "read_only_frac": round(F * W * H / (kern_ms * 1e-3) / 1e9 / peak, 4),
                "whole_step_frac": round(algo_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4)}

    # ---- N > 1: rank 0 verifies the one batch result the timed steps produced ---------------------------
    multi_check = None
    if world > 1:
        ok, detail = True, ""
        blocks = sd._all.clone()
        want_offs, bases = sharding.global_offsets_from_blocks(blocks, n_total, world)
        if not torch.equal(want_offs, sd.global_offsets):
            ok, detail = False, "global offsets differ from the scan of the ranks' counts"
        if rank == 0:
            import oracle

            oracle.build()
            offs_h = sd.global_offsets.cpu().numpy()
            checked = []
            for r in range(world):
                lo_r, hi_r = sharding.frame_shard(n_total, r, world)
                for f in sorted({lo_r, hi_r - 1}):
                    want = oracle.port_detect(oracle.synth_frame(W, H, SEED, f, 0, 4), THRESHOLD, COUNT, NMS)
                    got = sd.points[int(offs_h[f]):int(offs_h[f + 1])].cpu().numpy().astype(np.uint32)
                    same = got.shape == want.shape and oracle.hash_points(got) == oracle.hash_points(want)
                    checked.append(f)
                    if not same:
                        ok, detail = False, f"frame {f} (rank {r}) differs from the CPU port"
            multi_check = {"frames_hash_checked_on_rank0": checked}
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            raise SystemExit(f"bench.py: multi-GPU result check failed on some rank ({detail})")
        if rank == 0:
            multi_check.update({"global_offsets_equal_scan_of_all_ranks_counts": True, "hashes_equal_cpu_port": True,
                                "keypoints_in_batch_result": found})

    # ---- strong scaling: the literal config 5, ONE 512-frame batch sharded over the N GPUs -----------------
    strong = None
    strong_total = args.strong_frames
    if world == 1:
        if F == strong_total:
            strong = {"frames_total": strong_total, "frames_per_gpu": F, "value": round(mpix, 1), "unit": "Mpix/s",
                      "ms_per_step": round(ms_per_step, 4), "steps": args.steps, "note": "same measurement as `value` at N = 1"}
    elif strong_total % world == 0 and strong_total // world <= F:
        fs = strong_total // world
        s_frames = frames[:fs]  # (frame content does not matter for the timing; the check above covered correctness)
        sds = sharding.ShardedDetector(det, strong_total, cap_total=strong_total * per_frame_cap,
                                       cap_local=fs * per_frame_cap, idle_sm_stride=args.idle_sm_stride)
        s_steps = max(20, args.steps)
        for _ in range(3):
            sds.detect(s_frames, cfg)
        sds.fence()
        t_host = time.perf_counter()
        for _ in range(s_steps):
            sds.detect(s_frames, cfg)
        host_us = (time.perf_counter() - t_host) / s_steps * 1e6  # host time to enqueue one step (the GPU lags behind)
        sds.fence()
        torch.cuda.synchronize()
        s_ms = timed_steps(lambda: sds.detect(s_frames, cfg), s_steps, fence=sds.fence)
        strong = {"frames_total": strong_total, "frames_per_gpu": fs, "unit": "Mpix/s", "steps": s_steps,
                  "host_enqueue_us_per_step": round(host_us, 1),
                  "value": round(strong_total * W * H / (s_ms / s_steps * 1e-3) / 1e6, 1),
                  "ms_per_step": round(s_ms / s_steps, 4),
                  "idle_sm_stride": args.idle_sm_stride,
                  "note": "one 512-frame batch, looped; every step ends in one batch result on rank 0"}
        sds.close()

    # ---- end to end through the C ABI with host (pinned) buffers ------------------------------------
    e2e = None
    if not args.no_e2e:
        h_frames = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
        h_frames.copy_(frames)
        h_points = torch.empty((cap, 2), dtype=torch.int32, pin_memory=True)
        h_offsets = torch.empty(F + 1, dtype=torch.int64, pin_memory=True)
        torch.cuda.synchronize()
        e2e_steps = max(2, min(args.steps, 5))

        def e2e_step():
            det.detect_batch_pinned(h_frames.data_ptr(), F, W, H, cfg, h_points.data_ptr(), cap, h_offsets.data_ptr())
            return int(h_offsets[F])  # the device->host read of the step's result

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            k = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
        assert k == local_found, "end-to-end path found a different number of keypoints"
        e2e = {"value": round(n_total * W * H * e2e_steps / dt / 1e6, 1), "unit": "Mpix/s",
               "h2d_bytes_per_step": F * W * H, "d2h_bytes_per_step": 8 * (F + 1) + 8 * local_found + 4,
               "steps": e2e_steps, "ms_per_step": round(dt / e2e_steps * 1e3, 3),
               "api": "fdf_detect_batch (C ABI, pinned host buffers)", "host_placement": numa or "default"}
        del h_frames, h_points

    # ---- CPU baseline beside it (rank 0, N = 1 only), which also re-checks the GPU result ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle

        oracle.build()
        cores = host_threads()
        n_s = min(F, max(8, 2 * cores))
        sample = frames[:n_s].cpu().numpy()
        pad = np.zeros(sample.size + 64, np.uint8)
        pad[: sample.size] = sample.reshape(-1)
        sample = pad[: sample.size].reshape(n_s, H, W)
        one_mpix, _, _, _, _ = time_cpu_port(sample[:4], 1, 4.0, oracle)
        all_mpix, dt_all, reps, counts, hashes = time_cpu_port(sample, cores, 8.0, oracle, want_hashes=True)
        offs_h = offsets.cpu().numpy()
        pts_h = points[: int(offs_h[n_s])].cpu().numpy().astype(np.uint32)
        for f in range(n_s):
            if oracle.hash_points(pts_h[offs_h[f]:offs_h[f + 1]]) != int(hashes[f]) or offs_h[f + 1] - offs_h[f] != counts[f]:
                raise SystemExit(f"bench.py: GPU keypoints of frame {f} differ from the CPU port (hash of the ordered list)")
        cpu = {"value": round(all_mpix, 1), "unit": "Mpix/s", "cores": cores, "kind": "port",
               "sample": f"{n_s} of the batch's frames x {reps} passes, one frame per thread ({dt_all:.1f} s); "
                         f"single thread on 4 frames: {one_mpix:.1f} Mpix/s",
               "single_thread_value": round(one_mpix, 1), "hashes_match_gpu": True}

    # ---- the other BASELINE configs (rank 0, N = 1 only) ---------------------------------------------
    by_config = None
    if rank == 0 and world == 1 and not args.no_by_config:
        del points
        torch.cuda.empty_cache()
        by_config = bench_by_config(det, fdf, torch, peak)
        import oracle

        oracle.build()
        frame1080 = synth_frames_host(oracle, 1, 1920, 1080)[0]
        by_config["criterion"] = criterion_triple(
            lambda nms: len(det.detect_array(frame1080, fdf.Config(16, 9, fdf.NonMaximalSuppression(nms)))),
            "fdf_detect through the C ABI, PAGEABLE host image (what image::GrayImage's Vec<u8> is), copies included")
        pinned = torch.empty((1080, 1920), dtype=torch.uint8, pin_memory=True)
        pinned.copy_(torch.from_numpy(frame1080))
        pinned_np = pinned.numpy()
        by_config["criterion_pinned_input"] = criterion_triple(
            lambda nms: len(det.detect_array(pinned_np, fdf.Config(16, 9, fdf.NonMaximalSuppression(nms)))),
            "fdf_detect through the C ABI, host image in PINNED memory (cudaHostRegister'ed by the caller), copies included")

        by_config["streaming_pipe_1080p"] = streaming_pipe(det, fdf, frame1080, pinned_np)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(mpix, 1), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(F),
            "details": {"keypoints_per_step_rank0": local_found, "keypoints_per_step_all_ranks": found,
                        "collective": ("one all_gather of the ranks' local CSR offsets (NCCL) + one push kernel per rank "
                                       "that copies its points into rank 0's result over NVLink, both on an exchange "
                                       "stream; the detection kernel leaves every %d-th SM to them" % args.idle_sm_stride)
                        if world > 1
                        else "none (1 GPU)"},
            "fps_1080p_equiv": round(mpix * 1e6 / (1920 * 1080), 1),
            "fps_4k": round(mpix * 1e6 / (W * H), 1),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "strong": strong, "multi_gpu_check": multi_check,
            "by_config": by_config, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if sd is not None:
        sd.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    det.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
