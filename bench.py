#!/usr/bin/env python3
"""bench.py -- throughput of the FAST-n detection path on B200, one JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4]): a batch of synthetic 3840x2160 grey frames, threshold 20, count 9,
max-threshold NMS.  One "step" = one pass of the detection path over this rank's resident batch
(512 frames per GPU; frames are independent units, so ranks share nothing on the data path and the job is
weak-scaled: N GPUs process N x 512 frames per step) followed, for N > 1, by the path's only exchange
step: one NCCL all-gather of per-frame keypoint counts -> global CSR offsets.

  value        Mpix/s, whole job, frames already resident in HBM when the timed region starts
  e2e          same metric through the C-ABI call `fdf_detect_batch` on HOST (pinned) buffers: host->device
               copy of the frames and device->host copy of offsets + keypoints inside the timed region
  roofline     dominant kernel (fdf_detect_kernel): algorithmic bytes (W*H per frame read once + 8 B per
               keypoint + 8 B per frame offset) / CUDA-event duration of that launch (events recorded by the
               library on the launching stream during the timed steps), against the measured HBM copy
               bandwidth in MEASURED_PEAKS.json
  cpu_baseline AVX2 port of the reference's fast_simd.rs (oracle/fdf_avx2_port.cpp) on the host cores,
               bounded sample of the same frames (rank 0, N = 1 only); reported, not the target

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


def time_cpu_port(frames_np, n_threads, min_seconds, oracle):
    """AVX2 port over `frames_np` (F, H, W) with n_threads workers, repeated for >= min_seconds."""
    f = frames_np.shape[0]
    reps, t0 = 0, time.perf_counter()
    counts = None
    while True:
        counts, _ = oracle.port_detect_many(frames_np, THRESHOLD, COUNT, NMS, n_threads=n_threads, want_hashes=False)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return (reps * f * W * H) / dt / 1e6, dt, reps, counts


def run_reference_arm(args, rank):
    """CPU port of the reference path, all host threads, bounded samples of the same workload."""
    if rank != 0:
        return 0
    import numpy as np

    import oracle

    oracle.build()
    cores = host_threads()
    n = max(cores, 8)
    frames = np.zeros((n, H, W), np.uint8)
    for f in range(n):
        frames[f] = oracle.synth_frame(W, H, SEED, f, 0, 4)
    pad = np.zeros(frames.size + 64, np.uint8)
    pad[: frames.size] = frames.reshape(-1)
    frames = pad[: frames.size].reshape(n, H, W)
    for _ in range(args.warmup):
        oracle.port_detect_many(frames, THRESHOLD, COUNT, NMS, n_threads=cores, want_hashes=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.port_detect_many(frames, THRESHOLD, COUNT, NMS, n_threads=cores, want_hashes=False)
    dt = time.perf_counter() - t0
    mpix = args.steps * n * W * H / dt / 1e6
    sample = f"{n} of the workload's 3840x2160 frames per step, one frame per thread, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mpix, 2), "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": n, "width": W, "height": H, "threshold": THRESHOLD,
                   "count": COUNT, "nms": "max_threshold",
                   "note": "CPU port of fast_simd.rs (the Rust reference cannot be built here: no cargo/rustc)"},
        "cpu_baseline": {"value": round(mpix, 2), "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(mpix, 2), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps_1080p_equiv": round(mpix * 1e6 / (1920 * 1080), 1),
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_criterion(args, rank):
    """benches/benchmark.rs restated: the crate's three criterion functions -- simd_t16_c_9_off /
    _max_threshold / _sum_abs, one 1920x1080 frame per iteration -- timed criterion-style (warm-up, then samples;
    mean and a 95 % confidence interval of the mean) for one arm: `--impl b200` = one fdf_detect call per iteration
    through the C ABI with HOST buffers (copies included), `--impl reference` = the CPU port on one thread.
    Prints one JSON line per function.  The reference's published numbers (README.md:55-65, i7-4770TE, its private
    1080p screenshot): 5.34 / 8.71 / 7.23 ms."""
    if rank != 0:
        return 0
    import math

    import numpy as np

    import oracle  # (synthetic frame generator; the CPU arm also times the port -- bench.py's CPU-baseline role)

    oracle.build()
    w, h = 1920, 1080
    frame = oracle.synth_frame(w, h, SEED, 0, 0, 4)
    pad = np.zeros(frame.size + 64, np.uint8)
    pad[: frame.size] = frame.reshape(-1)
    frame = pad[: frame.size].reshape(h, w)
    names = {0: "simd_t16_c_9_off", 1: "simd_t16_c_9_max_threshold", 2: "simd_t16_c_9_sum_abs"}
    if args.impl == "b200":
        import feature_detector_fast_b200 as fdf

        det = fdf.Detector(0)
        run = lambda nms: len(det.detect_array(frame, fdf.Config(16, 9, fdf.NonMaximalSuppression(nms))))
    else:
        run = lambda nms: len(oracle.port_detect(frame, 16, 9, nms))
    for nms, name in names.items():
        t_end = time.perf_counter() + 1.0  # warm-up
        while time.perf_counter() < t_end:
            found = run(nms)
        samples = []
        t_end = time.perf_counter() + 4.0
        while len(samples) < 100 or (time.perf_counter() < t_end and len(samples) < 2000):
            t0 = time.perf_counter()
            run(nms)
            samples.append((time.perf_counter() - t0) * 1e3)
        mean = sum(samples) / len(samples)
        sd = math.sqrt(sum((x - mean) ** 2 for x in samples) / (len(samples) - 1))
        half = 1.96 * sd / math.sqrt(len(samples))
        print(json.dumps({"criterion": name, "impl": args.impl, "unit": "ms per 1920x1080 frame",
                          "ci95": [round(mean - half, 4), round(mean, 4), round(mean + half, 4)],
                          "samples": len(samples), "keypoints": found, "mpix_per_s": round(w * h / mean / 1e3, 1),
                          "data": "synthetic 1080p scene frame (the reference's screenshot is not shipped)",
                          "api": "fdf_detect, host buffers" if args.impl == "b200" else "CPU port of fast_simd.rs, 1 thread"}),
              flush=True)
    return 0


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
    frames = det.synth_frames(F, W, H, seed=SEED, first_frame=rank * F, kind=0, amp=4)
    cap = F * 100000  # ~2.4x the expected keypoints of this workload (about 42k per 4K frame); checked below
    points = torch.empty((cap, 2), dtype=torch.int32, device=dev)
    offsets = torch.empty(F + 1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step():
        det.detect_device(frames, cfg, points=points, offsets=offsets)
        if world > 1:
            counts = sharding.counts_from_offsets(offsets)
            return sharding.global_offsets(sharding.gather_frame_counts(counts, n_total))
        return offsets

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    found = int(offsets[-1].item())
    if found > cap or det.device_flags() != 0:
        raise SystemExit(f"bench.py: invalid run (found {found} > cap {cap} or device flags set)")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: exactly K steps, CUDA events, max over ranks --------------------
    sampler = ClockSampler(local_rank)
    det.set_timing(args.steps)  # CUDA events around each launch, recorded by the library on the launching stream
    launches0 = det.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    barrier()
    sampler.mark_begin()
    ev0.record()
    for i in range(args.steps):
        det.detect_device(frames, cfg, points=points, offsets=offsets)
        if world > 1:
            counts = sharding.counts_from_offsets(offsets)
            sharding.global_offsets(sharding.gather_frame_counts(counts, n_total))
    ev1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = det.kernel_launches - launches0
    total_ms = ev0.elapsed_time(ev1)
    per_launch = [det.get_timing(i) for i in range(args.steps)]
    det.set_timing(0)
    kern_ms = sum(t[0] for t in per_launch) / args.steps        # the dominant kernel: fdf_detect_kernel
    scan_ms = sum(t[1] for t in per_launch) / args.steps
    gather_ms = sum(t[2] for t in per_launch) / args.steps
    if world > 1:
        tt = torch.tensor([total_ms, kern_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, kern_ms = float(tt[0]), float(tt[1])
    ms_per_step = total_ms / args.steps
    mpix = n_total * W * H / (ms_per_step * 1e-3) / 1e6

    peak, peak_src = measured_hbm_peak()
    algo_bytes = F * W * H + 8 * found + 8 * (F + 1)
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": committed_traffic(), "peak_source": peak_src,
                "kernel": "fdf_detect_kernel<MaxThreshold,64>", "kernel_ms": round(kern_ms, 4),
                "other_kernels_ms": {"fdf_scan_kernel": round(scan_ms, 4), "fdf_gather_kernel": round(gather_ms, 4)},
                "algorithmic_bytes_per_launch": algo_bytes,
                "read_only_frac": round(F * W * H / (kern_ms * 1e-3) / 1e9 / peak, 4)}

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
        assert k == found, "end-to-end path found a different number of keypoints"
        e2e = {"value": round(n_total * W * H * e2e_steps / dt / 1e6, 1), "unit": "Mpix/s",
               "h2d_bytes_per_step": F * W * H, "d2h_bytes_per_step": 8 * (F + 1) + 8 * found + 4,
               "steps": e2e_steps, "ms_per_step": round(dt / e2e_steps * 1e3, 3),
               "api": "fdf_detect_batch (C ABI, pinned host buffers)", "host_placement": numa or "default"}
        del h_frames, h_points

    # ---- CPU baseline beside it (rank 0, N = 1 only), which also re-checks the GPU counts ---------------
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
        one_mpix, _, _, _ = time_cpu_port(sample[:4], 1, 4.0, oracle)
        all_mpix, dt_all, reps, counts = time_cpu_port(sample, cores, 8.0, oracle)
        gpu_counts = (offsets[1:n_s + 1] - offsets[:n_s]).cpu().numpy()
        if not (gpu_counts == counts).all():
            raise SystemExit("bench.py: GPU keypoint counts differ from the CPU port on the sampled frames")
        cpu = {"value": round(all_mpix, 1), "unit": "Mpix/s", "cores": cores, "kind": "port",
               "sample": f"{n_s} of the batch's frames x {reps} passes, one frame per thread ({dt_all:.1f} s); "
                         f"single thread on 4 frames: {one_mpix:.1f} Mpix/s",
               "single_thread_value": round(one_mpix, 1), "counts_match_gpu": True}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(mpix, 1), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": F, "width": W, "height": H, "threshold": THRESHOLD,
                       "count": COUNT, "nms": "max_threshold", "keypoints_per_step_rank0": found,
                       "l2": "inputs larger than L2 (4.2 GB per GPU per step vs 126 MB)",
                       "collective": "all_gather of per-frame counts (NCCL)" if world > 1 else "none (1 GPU)"},
            "fps_1080p_equiv": round(mpix * 1e6 / (1920 * 1080), 1),
            "fps_4k": round(mpix * 1e6 / (W * H), 1),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    det.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
