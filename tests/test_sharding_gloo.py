"""world_size-2 `gloo` test of the multi-GPU host logic: contiguous frame shards + one all-gather of the ranks' local
CSR offsets reproduce the single-process CSR offsets and give every rank its base position in the batch result.  The per-rank detector here is the CPU oracle port (it is only the
stand-in producing counts; the GPU tier exercises the same code with NCCL).  CPU only."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    import torch
    import torch.distributed as dist

    import oracle
    from feature_detector_fast_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.frame_shard(n_frames, rank, world)
    counts = [len(oracle.port_detect(oracle.synth_frame(160, 90, 99, f, 0, 4), 16, 9, 1)) for f in range(lo, hi)]
    local = torch.zeros(hi - lo + 1, dtype=torch.int64)
    local[1:] = torch.cumsum(torch.tensor(counts, dtype=torch.int64), 0) if counts else local[1:]
    blocks = sharding.gather_offset_blocks(local, n_frames)
    offs, bases = sharding.global_offsets_from_blocks(blocks, n_frames, world)
    q.put((rank, sharding.counts_from_offsets(offs).tolist(), offs.tolist(), bases.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [8, 7, 1])
def test_two_rank_count_gather(n_frames):
    import torch.multiprocessing as mp

    import oracle

    oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [len(oracle.port_detect(oracle.synth_frame(160, 90, 99, f, 0, 4), 16, 9, 1)) for f in range(n_frames)]
    want_offs = np.concatenate([[0], np.cumsum(want)]).tolist()
    for _, counts, offs, bases in results:
        assert counts == want
        assert offs == want_offs
        assert bases == [want_offs[(r * n_frames) // 2] for r in range(2)]


def test_frame_shard_partitions():
    from feature_detector_fast_b200.sharding import frame_shard

    for n in (0, 1, 7, 8, 512, 513):
        for world in (1, 2, 3, 4, 8):
            blocks = [frame_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert [frame_shard(512, r, 8) for r in range(8)][3] == (192, 256)


def test_shard_block_and_device_rule_agree():
    """sharding.frame_shard is the rule csrc/fdf_kernels.cuh::shard_lo implements: (r * total) / n."""
    import re

    from feature_detector_fast_b200 import sharding

    src = open(os.path.join(os.path.dirname(__file__), "..", "feature_detector_fast_b200", "csrc", "fdf_kernels.cuh")).read()
    assert re.search(r"shard_lo\(uint32_t total, uint32_t r, uint32_t n\)\s*\{\s*return \(uint32_t\)\(\(\(unsigned long long\)r \* total\) / n\);", src)
    assert sharding.shard_block(13, 2) == 8 and sharding.shard_block(512, 8) == 65 and sharding.shard_block(1, 2) == 2
