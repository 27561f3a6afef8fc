#!/usr/bin/env python3
"""torchrun check of the multi-GPU product path on real GPUs (test code: tests/test_multi_gpu.py runs it under
`torch.distributed.run --nproc-per-node 2` when two GPUs are visible):

    frames sharded contiguously over the ranks (sharding.frame_shard) -> detection per rank -> ONE NCCL all-gather of
    the ranks' local CSR offsets -> every rank's emission kernel writes its points straight into rank 0's result
    buffer (peer-mapped over NVLink) and the global CSR offsets.

Rank 0 then checks the assembled batch result against the CPU port frame by frame (counts, offsets, points).
Uneven shards (13 frames) and a rank without frames (1 frame, 2+ ranks) are both exercised.  Exit status 0 = OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf  # noqa: E402
from feature_detector_fast_b200 import sharding  # noqa: E402


def check(det, rank, world, n_frames, w, h, nms, seed):
    import oracle

    lo, hi = sharding.frame_shard(n_frames, rank, world)
    cfg = fdf.Config(20, 9, fdf.NonMaximalSuppression(nms))
    dev = torch.device("cuda", det.device)
    frames = (det.synth_frames(hi - lo, w, h, seed=seed, first_frame=lo) if hi > lo
              else torch.empty((0, h, w), dtype=torch.uint8, device=dev))
    sd = sharding.ShardedDetector(det, n_frames, cap_total=n_frames * 60000)
    ok = True
    for rep in range(2):  # twice: the buffers are reused
        points, goffs = sd.detect(frames, cfg)
        sd.fence()
        torch.cuda.synchronize()
        if det.device_flags() != 0:
            print(f"rank {rank}: device flags set", flush=True)
            ok = False
        # every rank holds the same global offsets
        ref = goffs.clone()
        dist.broadcast(ref, src=0)
        if not torch.equal(ref, goffs):
            print(f"rank {rank}: global offsets differ from rank 0's", flush=True)
            ok = False
        if rank == 0:
            offs = goffs.cpu().numpy()
            pts = points[: int(offs[-1])].cpu().numpy().astype(np.uint32)
            for f in range(n_frames):
                want = oracle.port_detect(oracle.synth_frame(w, h, seed, f, 0, 4), 20, 9, nms)
                got = pts[offs[f]:offs[f + 1]]
                if got.shape != want.shape or not np.array_equal(got, want):
                    print(f"frame {f}: {len(got)} points, CPU port has {len(want)}", flush=True)
                    ok = False
    sd.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        import oracle

        oracle.build()
    dist.barrier()
    det = fdf.Detector(local)
    ok = True
    for n_frames, w, h, nms in [(13, 1920, 1080, 1), (1, 640, 360, 0), (8, 1280, 720, 2)]:
        good = check(det, rank, world, n_frames, w, h, nms, seed=99)
        if rank == 0:
            print(f"multi_gpu_check world {world}: {n_frames} frames {w}x{h} nms {nms}:", "OK" if good else "MISMATCH", flush=True)
        ok = ok and good
    dist.barrier()
    dist.destroy_process_group()
    det.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
