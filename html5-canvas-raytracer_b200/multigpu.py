"""Multi-GPU host layer: one process per GPU (torchrun), scene replicated, frame partitioned.

  * row stripes (BASELINE config 3): stripes of `stripe_rows` rows are dealt round-robin to the ranks
    (stripe i -> rank i % world).  Contiguous bands would be badly imbalanced — the horizon splits the
    demo into cheap sky rows and expensive floor/sphere rows.  Every rank renders its
    rows PACKED into a [padded_rows, W, 4] buffer, one `all_gather_into_tensor` (NCCL over NVLink)
    assembles [world, padded_rows, W, 4] on every rank, and a small kernel de-interleaves it into
    the row-major framebuffer.  The gather is the only collective and exists only when world > 1.
  * frame batches (config 5): frame f -> rank f % world; the rendered frames are all-gathered.

torch is used for process-group plumbing and device buffers only; all rendering goes through the
C ABI.  The partition arithmetic is pure Python so that it can be tested with gloo on CPU.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

from . import _abi as A
from .scene import make_opts

DEFAULT_STRIPE_ROWS = 4     # 2160 rows / 4 = 540 stripes: per-rank row counts differ by < 1.5 % at 8 ranks


# ---- partition arithmetic (mirrors RowMap in csrc/rtb_device.cuh; tested against it) ----------------

def stripe_rows_of(height: int, stripe_rows: int, rank: int, world: int) -> List[int]:
    if world <= 1:
        return list(range(height))
    return [y for y in range(height) if (y // stripe_rows) % world == rank]


def padded_rows(height: int, stripe_rows: int, world: int) -> int:
    return max(len(stripe_rows_of(height, stripe_rows, r, world)) for r in range(max(world, 1)))


def unstripe_numpy(gathered: np.ndarray, height: int, stripe_rows: int, world: int) -> np.ndarray:
    """[world, padded_rows, W, 4] -> [H, W, 4]; the CPU twin of rtb_unstripe_device."""
    out = np.empty((height,) + gathered.shape[2:], gathered.dtype)
    for r in range(world):
        ys = stripe_rows_of(height, stripe_rows, r, world)
        out[ys] = gathered[r, :len(ys)]
    return out


def frames_of(nframes: int, rank: int, world: int) -> List[int]:
    return list(range(rank, nframes, max(world, 1)))


# ---- process group ------------------------------------------------------------------------------------

def init_distributed(backend: str | None = None):
    """Join the torchrun process group (env: RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR/PORT).
    Returns (rank, local_rank, world).  world == 1 without torchrun: no group is created."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


class StripedRenderer:
    """Row-stripe sharded redraw(): every rank ends up with the full framebuffer on its GPU.

    Gather layout per rank: [padded_rows + 1][W][4] — the packed stripe rows plus one trailer row whose
    first word is the rank's status (non-zero: its ray queues overflowed, the frame must be repeated).
    Shipping the status inside the payload gives every rank the same verdict from the one collective."""

    def __init__(self, renderer, width: int, height: int, rank: int, world: int,
                 stripe_rows: int = DEFAULT_STRIPE_ROWS):
        import torch

        self.r, self.W, self.H = renderer, width, height
        self.rank, self.world, self.S = rank, world, stripe_rows
        self.lib = A.load()
        self.n_local = self.lib.rtb_stripe_local_rows(height, stripe_rows, rank, world)
        self.padded = self.lib.rtb_stripe_padded_rows(height, stripe_rows, world)
        self.stride = self.padded + 1
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.frame = torch.empty((height, width, 4), dtype=torch.uint8, device=dev)
        if world > 1:
            # two sets: frame i+1 renders while frame i is still being gathered (redraw_sequence)
            self.local = [torch.zeros((self.stride, width, 4), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.gathered = [torch.empty((world, self.stride, width, 4), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.flags_host = torch.zeros((world,), dtype=torch.int32).pin_memory()

    # -- building blocks ---------------------------------------------------------------------------
    def _enqueue_render(self, camera, slot, stats, kw):
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        o = make_opts(stripe_rows=self.S, stripe_rank=self.rank, stripe_world=self.world, out_packed=1, **kw)
        loc = self.local[slot]
        self.r.redraw_device(camera, o, loc.data_ptr(), stream, stats)
        A.check(self.lib.rtb_copy_status_device(self.r._h, loc[self.padded].data_ptr(), A.stream_arg(stream)))

    def _enqueue_gather(self, slot):
        import torch.distributed as dist

        # concatenated-along-dim-0 view: the form both NCCL and gloo accept
        return dist.all_gather_into_tensor(self.gathered[slot].view(-1, self.W, 4), self.local[slot], async_op=True)

    def _enqueue_unstripe(self, slot):
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        A.check(self.lib.rtb_unstripe_device(self.gathered[slot].data_ptr(), self.frame.data_ptr(), self.W, self.H,
                                             self.S, self.world, self.stride, A.stream_arg(stream)))
        flags = self.gathered[slot][:, self.padded, 0, :].contiguous().view(torch.int32).view(-1)
        self.flags_host.copy_(flags, non_blocking=True)

    def _verdict(self) -> bool:
        """Host-side: wait for the current stream, collect this rank's deferred status, read all flags."""
        import torch

        torch.cuda.current_stream().synchronize()
        self.r.finish()                                   # clears the library's pending state (already complete)
        return not bool(self.flags_host.any())

    # -- one frame (latency mode) ----------------------------------------------------------------------
    def redraw(self, camera, stats=None, **kw):
        """Render this rank's stripes, all-gather, de-interleave.  Returns the device framebuffer.
        Everything is only ENQUEUED; the host waits once.  If any rank's ray queues overflowed, every
        rank sees that in the gathered status words and repeats the frame (the library has shrunk the
        offending rank's bands, deterministically)."""
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        if self.world <= 1:
            while True:
                self.r.redraw_device(camera, make_opts(**kw), self.frame.data_ptr(), stream, stats)
                if self.r.finish():
                    return self.frame
        while True:
            self._enqueue_render(camera, 0, stats, kw)
            self._enqueue_gather(0).wait()
            self._enqueue_unstripe(0)
            if self._verdict():
                return self.frame

    # -- a sequence of frames (throughput mode) ----------------------------------------------------------
    def redraw_sequence(self, cameras, on_frame=None, **kw):
        """Software-pipelined: frame i+1 is rendered while frame i is in the collective.  `on_frame(i,
        frame)` is called (device tensor, valid until the next call) once frame i is assembled.  Returns
        the number of frames that had to be repeated because of queue overflows."""
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        n = len(cameras)
        if self.world <= 1:
            repeats = 0
            for i, c in enumerate(cameras):
                while True:
                    self.r.redraw_device(c, make_opts(**kw), self.frame.data_ptr(), stream, None)
                    if self.r.finish():
                        break
                    repeats += 1
                if on_frame:
                    on_frame(i, self.frame)
            return repeats
        repeats = 0
        i = 0
        work = None
        while i < n or work is not None:
            # enqueue the next render first so that it overlaps the collective of the previous frame
            nxt = None
            if i < n:
                slot = i & 1
                try:
                    self._enqueue_render(cameras[i], slot, None, kw)
                    nxt = (i, slot)
                except A.RtbError as e:
                    # enqueueing collected the verdict of the frame in flight: THIS rank overflowed on it.
                    # Nothing was enqueued; fall through to that frame's status words so that every rank
                    # restarts at the same frame.
                    if e.code != A.RTB_ERR_OVERFLOW or work is None:
                        raise
            if work is not None:
                (j, jslot), w = work
                w.wait()
                self._enqueue_unstripe(jslot)
                ev = torch.cuda.Event()
                ev.record()
                ev.synchronize()                          # host waits for frame j only; frame i keeps the GPU busy
                if bool(self.flags_host.any()):           # same words on every rank: same decision everywhere
                    repeats += 1
                    torch.cuda.current_stream().synchronize()
                    self.r.finish()
                    i, work = j, None                     # restart the pipeline at frame j
                    continue
                if on_frame:
                    on_frame(j, self.frame)
            if nxt is not None:
                work = (nxt, self._enqueue_gather(nxt[1]))
                i += 1
            else:
                work = None
        torch.cuda.current_stream().synchronize()
        self.r.finish()
        return repeats


class FrameBatchRenderer:
    """Frame-batch sharding for sequences: frame f is rendered by rank f % world; each step of
    `world` frames is assembled with one all-gather."""

    def __init__(self, renderer, width: int, height: int, rank: int, world: int):
        import torch

        self.r, self.W, self.H, self.rank, self.world = renderer, width, height, rank, world
        dev = torch.device("cuda", torch.cuda.current_device())
        self.mine = torch.zeros((height, width, 4), dtype=torch.uint8, device=dev)
        self.step = torch.empty((max(world, 1), height, width, 4), dtype=torch.uint8, device=dev)

    def redraw_step(self, cameras: Sequence, first: int, stats=None, **kw):
        """Frames first .. first+world-1 (cameras beyond the end are skipped).  Returns [world,H,W,4]."""
        import torch
        import torch.distributed as dist

        stream = torch.cuda.current_stream().cuda_stream
        f = first + self.rank
        if f < len(cameras):
            self.r.redraw_device(cameras[f], make_opts(**kw), self.mine.data_ptr(), stream, stats)
        if self.world <= 1:
            self.step[0].copy_(self.mine)
        else:
            dist.all_gather_into_tensor(self.step.view(-1, self.W, 4), self.mine)
        return self.step
