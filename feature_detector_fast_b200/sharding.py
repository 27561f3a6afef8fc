"""Frame-level data parallelism: one process per GPU, frames sharded contiguously, one collective.

The detection path has no cross-frame dependency (a frame is never split across GPUs), so the only
exchange step is ONE all-gather of the ranks' local CSR offsets, from which every rank derives the
global offsets of the batch result and each rank learns where its points start in it.
The points themselves are not sent by a second collective: one kernel per rank copies them to their final position in
rank 0's result buffer, which the other ranks map over NVLink (CUDA IPC, `fdf_shared_alloc` / `fdf_shared_open`); the
collective and that copy run on a second stream and overlap the detection of the next batch.  `torch.distributed` is plumbing: NCCL on the GPUs (the all-gather,
the 64-byte handle broadcast, the closing barrier), gloo in the CPU tests of the index arithmetic.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple


def frame_shard(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of frames owned by `rank`: [rank*F/G, (rank+1)*F/G)  (csrc/fdf_kernels.cuh: shard_lo)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def shard_block(n_frames: int, world: int) -> int:
    """Entries per rank in the all-gather buffer: the largest shard's frames + 1 (its local CSR offsets)."""
    return max(frame_shard(n_frames, r, world)[1] - frame_shard(n_frames, r, world)[0] for r in range(world)) + 1


def counts_from_offsets(offsets):
    """Per-frame keypoint counts from a CSR offsets tensor (F + 1,)."""
    return offsets[1:] - offsets[:-1]


def global_offsets_from_blocks(all_offsets, n_frames: int, world: int):
    """What `fdf_detect_shard_finish` computes on the device, restated on tensors (any device; the CPU tests and the
    GPU checks compare against it): all_offsets is (world * block,) = every rank's local offsets block after the
    all-gather.  Returns (global CSR offsets (F + 1,), per-rank base positions (world,))."""
    import torch

    block = all_offsets.numel() // world
    out = torch.zeros(n_frames + 1, dtype=torch.int64, device=all_offsets.device)
    bases = torch.zeros(world, dtype=torch.int64, device=all_offsets.device)
    base = 0
    for r in range(world):
        lo, hi = frame_shard(n_frames, r, world)
        blk = all_offsets[r * block: r * block + (hi - lo) + 1].to(torch.int64)
        out[lo:hi] = blk[: hi - lo] + base
        bases[r] = base
        base = base + int(blk[hi - lo])
    out[n_frames] = base
    return out, bases


def gather_offset_blocks(local_offsets, n_frames: int, group=None):
    """The path's one collective, stand-alone (used by the gloo tests; ShardedDetector keeps its buffers resident):
    all-gathers every rank's local offsets, each padded to shard_block entries, into a (world * block,) tensor."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = frame_shard(n_frames, rank, world)
    if local_offsets.numel() != hi - lo + 1:
        raise ValueError(f"rank {rank} owns {hi - lo} frames but passed {local_offsets.numel()} offsets")
    block = shard_block(n_frames, world)
    recv = torch.zeros(world * block, dtype=torch.int64, device=local_offsets.device)
    send = recv[rank * block: (rank + 1) * block]
    send[: hi - lo + 1] = local_offsets.to(torch.int64)
    dist.all_gather_into_tensor(recv, send.clone(), group=group)
    return recv


class ShardedDetector:
    """One batch, all GPUs of the box, one result (north star: "NCCL is used only to gather per-GPU keypoint counts and
    offsets into one batch result over NVLink").

    Every rank constructs it with the same arguments (collective).  Rank 0 owns the result buffer of `cap_total` points;
    the other ranks map it.  `detect(local_frames, config)` enqueues

      on torch's current stream   detection, offset scan and ordered emission of the local frames into a local buffer
                                  (the three launches of the single-GPU path: nothing of the exchange is on the
                                  critical path of the next batch)
      on the exchange stream      all_gather_into_tensor of the ranks' local offsets (the one collective), then ONE
                                  kernel that copies the rank's points to their final position in rank 0's buffer over
                                  NVLink (coalesced 16-byte stores) and writes the global CSR offsets.  (Rank 0's own
                                  emission already lands in the result: it copies nothing.)

    and returns (points, global_offsets): `points` is the (cap_total, 2) int32 view of the result on rank 0 and None
    elsewhere; global_offsets is an int64 (F + 1,) device tensor on every rank.  Local buffers rotate between calls,
    so the exchange of batch k overlaps the detection of the following batches.  `fence()` makes both valid for work enqueued
    afterwards on the current stream (also the other ranks' points: it ends in an all-reduce).
    """

    def __init__(self, detector, n_frames: int, cap_total: int, cap_local: Optional[int] = None, group=None,
                 depth: int = 4, idle_sm_stride: int = 37):
        import torch
        import torch.distributed as dist

        self.det, self.depth = detector, max(2, int(depth))
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        # The detection kernel leaves every idle_sm_stride-th SM (4 of 148 for 37) to the exchange kernels: otherwise
        # every all-gather and push waits for a whole persistent detection launch to drain (fdf_set_idle_sms).
        self._idle_sm_stride = int(idle_sm_stride) if self.world > 1 else 0
        detector.set_idle_sms(self._idle_sm_stride)
        detector._sharded_users = getattr(detector, "_sharded_users", 0) + 1  # (the last one to close gives the SMs back)
        # The per-batch all-gather gets a communicator of its own on a HIGH-PRIORITY stream: the detection kernel is
        # persistent and fills every SM, so a normal-priority NCCL kernel that becomes ready a few microseconds after
        # the next batch's detection was launched waits for that whole launch.  With priority its (few) CTAs are
        # placed first whenever CTA slots free up.  Same for the exchange stream the push kernel runs on.
        self._own_group = False
        if dist.get_backend(group) == "nccl":
            ranks = dist.get_process_group_ranks(group if group is not None else dist.group.WORLD)
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            group = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
            self._own_group = True
        self.group = group
        self.n_frames, self.cap_total = int(n_frames), int(cap_total)
        self.lo, self.hi = frame_shard(self.n_frames, self.rank, self.world)
        self.cap_local = int(cap_local) if cap_local is not None else self.cap_total
        self.block = shard_block(self.n_frames, self.world)
        self.device = torch.device("cuda", detector.device)
        lib, ctx = detector._lib, detector._ctx
        self._all = torch.zeros(self.world * self.block, dtype=torch.int64, device=self.device)
        self.global_offsets = torch.zeros(self.n_frames + 1, dtype=torch.int64, device=self.device)
        self._fence = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._xstream = torch.cuda.Stream(device=self.device, priority=-1)
        self._turn = 0
        self.trace = None  # set to [] to record (batch start, detection done, all-gather start / end, push end) events
        # per buffer: the rank's block of the all-gather (its local offsets) and, except on rank 0, its local points
        # (`depth` of each: the all-gather is a collective, so the ranks' exchanges run in lock step; a few batches of
        # slack keep a rank that is ahead from waiting for the slowest one)
        self._mine = [torch.zeros(self.block, dtype=torch.int64, device=self.device) for _ in range(self.depth)]
        self._local = [None] * self.depth
        if self.rank != 0:
            self._local = [torch.empty((max(1, self.cap_local), 2), dtype=torch.int32, device=self.device)
                           for _ in range(self.depth)]
        self._done = [None] * self.depth  # exchange-stream events: buffer i may be written again
        # rank 0 allocates the result and broadcasts its IPC handle; the others map it (peer access over NVLink)
        handle = torch.zeros(64, dtype=torch.uint8, device=self.device)
        self._ptr = C.c_void_p()
        if self.rank == 0:
            hbuf = (C.c_uint8 * 64)()
            st = lib.fdf_shared_alloc(ctx, max(1, self.cap_total) * 8, C.byref(self._ptr), hbuf)
            if st != 0:
                raise RuntimeError("fdf_shared_alloc: " + lib.fdf_last_error(ctx).decode())
            handle.copy_(torch.tensor(list(hbuf), dtype=torch.uint8))
        dist.broadcast(handle, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if self.rank != 0:
            hbuf = (C.c_uint8 * 64)(*handle.cpu().tolist())
            st = lib.fdf_shared_open(ctx, hbuf, C.byref(self._ptr))
            if st != 0:
                raise RuntimeError("fdf_shared_open: " + lib.fdf_last_error(ctx).decode())
        self.points = None
        if self.rank == 0:
            self.points = _tensor_from_pointer(self._ptr.value, (max(1, self.cap_total), 2), self.device)
        dist.barrier(group=group)

    def detect(self, local_frames, config):
        import torch
        import torch.distributed as dist

        from .api import _raise

        det, lib = self.det, self.det._lib
        f, h, w = local_frames.shape
        if f != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns frames [{self.lo}, {self.hi}) but got {f} frames")
        i = self._turn
        self._turn = (i + 1) % self.depth
        main = torch.cuda.current_stream(self.device)
        if self._done[i] is not None:
            main.wait_event(self._done[i])  # the exchange that last used buffer i (`depth` batches ago) is over
        mine = self._mine[i]
        tr = None
        if self.trace is not None:
            tr = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            self.trace.append(tr)
            tr[0].record(main)
        # rank 0 emits straight into the result (its points start at 0); the others into a local buffer
        out_ptr = self._ptr.value if self.rank == 0 else self._local[i].data_ptr()
        out_cap = self.cap_total if self.rank == 0 else self.cap_local
        if f > 0:
            st = lib.fdf_detect_device(det._ctx, local_frames.data_ptr(), f, w, h, local_frames.stride(1),
                                       local_frames.stride(0), int(config.threshold), int(config.count),
                                       int(config.non_maximal_supression), out_ptr, out_cap, mine.data_ptr(),
                                       main.cuda_stream)
            if st != 0:
                _raise(lib, det._ctx, st)
        else:
            mine.zero_()
        ready = torch.cuda.Event()
        ready.record(main)
        if tr:
            tr[1].record(main)
        xs = self._xstream
        xs.wait_event(ready)
        with torch.cuda.stream(xs):
            if tr:
                tr[2].record(xs)
            dist.all_gather_into_tensor(self._all, mine, group=self.group)
            if tr:
                tr[3].record(xs)
            st = lib.fdf_shard_push(det._ctx, self._all.data_ptr(), self.block, self.world, self.rank, self.n_frames,
                                    None if self.rank == 0 else self._local[i].data_ptr(),
                                    None if self.rank == 0 else self._ptr.value, self.cap_total,
                                    self.global_offsets.data_ptr(), xs.cuda_stream)
            if st != 0:
                _raise(lib, det._ctx, st)
            self._done[i] = torch.cuda.Event()
            self._done[i].record(xs)
            if tr:
                tr[4].record(xs)
        return self.points, self.global_offsets

    def fence(self) -> None:
        """Orders every rank's exchange (offsets, points in rank 0's buffer) before whatever is enqueued next on the
        current stream."""
        import torch
        import torch.distributed as dist

        main = torch.cuda.current_stream(self.device)
        for ev in self._done:
            if ev is not None:
                main.wait_event(ev)
        dist.all_reduce(self._fence, group=self.group)

    def close(self) -> None:
        import torch
        import torch.distributed as dist

        if getattr(self, "_ptr", None) is not None and self._ptr.value:
            torch.cuda.synchronize(self.device)
            if self.rank != 0:
                self.det._lib.fdf_shared_close(self.det._ctx, self._ptr)
            dist.barrier(group=self.group)  # nobody has the buffer mapped any more
            if self.rank == 0:
                self.points = None
                self.det._lib.fdf_shared_close(self.det._ctx, self._ptr)
            self._ptr = C.c_void_p()
            if self._own_group:
                dist.destroy_process_group(self.group)
                self._own_group = False
            self.det._sharded_users = max(0, getattr(self.det, "_sharded_users", 1) - 1)
            if self.det._sharded_users == 0:
                self.det.set_idle_sms(0)


class _RawCudaBuffer:
    """Minimal __cuda_array_interface__ carrier so that torch can view library-owned device memory."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


def _tensor_from_pointer(ptr: int, shape, device):
    import torch

    with torch.cuda.device(device):
        return torch.as_tensor(_RawCudaBuffer(ptr, shape, "<i4"), device=device)
