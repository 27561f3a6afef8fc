#!/usr/bin/env python3
"""torchrun check of the multi-GPU path on real GPUs: frames sharded contiguously over the ranks, detection per
rank, ONE NCCL all-gather of per-frame counts -> global CSR offsets; compared with the CPU port's counts."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf  # noqa: E402
from feature_detector_fast_b200 import sharding  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_frames, w, h = 13, 1920, 1080  # uneven shards on purpose
    lo, hi = sharding.frame_shard(n_frames, rank, world)
    det = fdf.Detector(local)
    cfg = fdf.Config(20, 9, fdf.NonMaximalSuppression.MaxThreshold)
    frames = det.synth_frames(hi - lo, w, h, seed=99, first_frame=lo)
    pts, offs = det.detect_device(frames, cfg)
    counts = sharding.counts_from_offsets(offs)
    g = sharding.gather_frame_counts(counts, n_frames)
    goffs = sharding.global_offsets(g)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        import oracle

        want = [len(oracle.port_detect(oracle.synth_frame(w, h, 99, f, 0, 4), 20, 9, 1)) for f in range(n_frames)]
        ok = g.cpu().tolist() == want and goffs.cpu().tolist() == np.concatenate([[0], np.cumsum(want)]).tolist()
        print("multi_gpu_check world", world, "counts", g.cpu().tolist(), "OK" if ok else "MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    det.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
