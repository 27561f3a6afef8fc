#!/usr/bin/env python3
"""Raw host-to-device copy bandwidth per GPU when 1, 2, 4, 8 GPUs of the box copy at the same time: separates what the
BOX can deliver (PCIe switches / root ports / host memory) from what the library's end-to-end path achieves.

One process per GPU (spawned here, no torch.distributed): each pins itself to the CPUs local to its GPU, allocates a
pinned buffer of the benchmark's input size (4.25 GB by default), waits on a file barrier and runs cudaMemcpyAsync
host->device in a loop; reports GB/s per GPU (min / mean / max over the ranks) for every N, plus `nvidia-smi topo -m`
and the PCIe tree.  usage: python tools/h2d_probe.py [--gb 4.25] [--reps 4] [--ns 1,2,4,8]"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time


def worker(rank, n, nbytes, reps, t_start, q):
    import torch

    torch.cuda.set_device(rank)
    try:  # the CPUs local to this GPU (same rule as bench.py)
        pr = torch.cuda.get_device_properties(rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    while time.time() < t_start:  # all ranks start together
        time.sleep(0.0005)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    q.put((rank, nbytes * reps / (ev0.elapsed_time(ev1) * 1e-3) / 1e9))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.25)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--ns", default="1,2,4,8")
    a = ap.parse_args()
    import torch

    have = torch.cuda.device_count()
    out = {"bytes_per_copy": int(a.gb * 1e9), "reps": a.reps, "gpus_visible": have, "by_n": {}}
    ctx = mp.get_context("spawn")
    for n in [int(x) for x in a.ns.split(",")]:
        if n > have:
            continue
        q = ctx.Queue()
        t_start = time.time() + 20.0 + 2.0 * n  # generous: pinned allocation of 4 GB takes seconds
        procs = [ctx.Process(target=worker, args=(r, n, int(a.gb * 1e9), a.reps, t_start, q)) for r in range(n)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=600) for _ in procs)
        for p in procs:
            p.join(timeout=60)
        g = [x[1] for x in res]
        out["by_n"][str(n)] = {"gbs_per_gpu": [round(x, 2) for x in g], "min": round(min(g), 2),
                               "mean": round(sum(g) / n, 2), "total": round(sum(g), 1)}
        print(f"N={n}: H2D GB/s per GPU min {min(g):.1f} mean {sum(g) / n:.1f} max {max(g):.1f}  total {sum(g):.1f}", flush=True)
    for name, cmd in (("topo", ["nvidia-smi", "topo", "-m"]), ("lspci_tree", ["lspci", "-tv"]), ("numa", ["numactl", "-H"]),
                      ("lscpu", ["lscpu"])):
        try:
            out[name] = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout[-6000:]
        except Exception as e:
            out[name] = f"unavailable: {e}"
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/h2d_probe.json", "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
