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

DEFAULT_STRIPE_ROWS = 4
MAX_IN_FLIGHT = 4            # = RTB_MAX_IN_FLIGHT of the library
DEFAULT_IN_FLIGHT = 3        # frames redraw_sequence keeps in flight (measured: 1.67 / 1.41 / 1.28 / 1.25 ms per
                             # 1/8-frame for 1 / 2 / 3 / 4 on one B200)     # 2160 rows / 4 = 540 stripes: per-rank row counts differ by < 1.5 % at 8 ranks


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
        # MAX_IN_FLIGHT sets of everything: redraw_sequence keeps several frames in flight
        K = MAX_IN_FLIGHT
        self.frames = [torch.empty((height, width, 4), dtype=torch.uint8, device=dev) for _ in range(K)]
        self.frame = self.frames[0]
        self.flags_host = [torch.zeros((max(world, 1),), dtype=torch.int32).pin_memory() for _ in range(K)]
        self.status_dev = torch.zeros((K, 4), dtype=torch.int32, device=dev)      # world == 1: status words
        self._streams = None
        if world > 1:
            self.local = [torch.zeros((self.stride, width, 4), dtype=torch.uint8, device=dev) for _ in range(K)]
            self.gathered = [torch.empty((world, self.stride, width, 4), dtype=torch.uint8, device=dev) for _ in range(K)]

    # -- building blocks (everything is enqueued on the CURRENT torch stream) --------------------------
    def _enqueue_render(self, camera, slot, stats, kw):
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        if self.world <= 1:
            self.r.redraw_device(camera, make_opts(**kw), self.frames[slot].data_ptr(), stream, stats)
            A.check(self.lib.rtb_copy_status_device(self.r._h, self.status_dev[slot].data_ptr(), A.stream_arg(stream)))
            return
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
        A.check(self.lib.rtb_unstripe_device(self.gathered[slot].data_ptr(), self.frames[slot].data_ptr(), self.W,
                                             self.H, self.S, self.world, self.stride, A.stream_arg(stream)))
        flags = self.gathered[slot][:, self.padded, 0, :].contiguous().view(torch.int32).view(-1)
        self.flags_host[slot].copy_(flags, non_blocking=True)

    def _enqueue_frame(self, camera, slot, kw):
        """Render (+ gather + de-interleave) of one frame into self.frames[slot]; status words to the host."""
        self._enqueue_render(camera, slot, None, kw)
        if self.world > 1:
            self._enqueue_gather(slot).wait()             # stream-level wait, the host does not block
            self._enqueue_unstripe(slot)
        else:
            self.flags_host[slot].copy_(self.status_dev[slot, :1], non_blocking=True)

    # -- one frame (latency mode) ----------------------------------------------------------------------
    def redraw(self, camera, stats=None, **kw):
        """Render this rank's stripes, all-gather, de-interleave.  Returns the device framebuffer.
        Everything is only ENQUEUED; the host waits once.  If any rank's ray queues overflowed, every
        rank sees that in the gathered status words and repeats the frame (the library has shrunk the
        offending rank's bands, deterministically)."""
        import torch

        while True:
            if stats is not None:
                self._enqueue_render(camera, 0, stats, kw)
                if self.world > 1:
                    self._enqueue_gather(0).wait()
                    self._enqueue_unstripe(0)
                else:
                    self.flags_host[0].copy_(self.status_dev[0, :1], non_blocking=True)
            else:
                self._enqueue_frame(camera, 0, kw)
            torch.cuda.current_stream().synchronize()
            self.r.finish()                               # clears the library's pending state (already complete)
            if not bool(self.flags_host[0].any()):
                self.frame = self.frames[0]
                return self.frame

    # -- a sequence of frames (throughput mode) ----------------------------------------------------------
    def redraw_sequence(self, cameras, on_frame=None, frames_in_flight: int = DEFAULT_IN_FLIGHT, **kw):
        """Software-pipelined: up to `frames_in_flight` (1..4) frames are enqueued before the host waits
        for the oldest one, each on its own stream with its own library workspace.  Later frames render
        while frame i is in the collective, and the kernels of the frames interleave on the GPU: the
        short per-level launches of one fill the tails of the others.  `on_frame(i, frame)` is called
        (device tensor, valid for the next frames_in_flight - 1 calls) once frame i is assembled on this
        rank.  Returns the number of frames that had to be repeated because of queue overflows (every
        rank reads the same gathered status words, so every rank repeats the same frames)."""
        import torch

        depth = max(1, min(MAX_IN_FLIGHT, int(frames_in_flight)))
        n = len(cameras)
        main = torch.cuda.current_stream()
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.dev) for _ in range(MAX_IN_FLIGHT)]
        done = [torch.cuda.Event() for _ in range(MAX_IN_FLIGHT)]
        for st in self._streams:
            st.wait_stream(main)
        self.r.set_frames_in_flight(depth)
        repeats = 0
        i = 0
        inflight = []
        try:
            while i < n or inflight:
                while i < n and len(inflight) < depth:
                    slot = i % depth
                    with torch.cuda.stream(self._streams[slot]):
                        self._enqueue_frame(cameras[i], slot, kw)
                        done[slot].record()
                    inflight.append(i)
                    i += 1
                j = inflight.pop(0)
                jslot = j % depth
                done[jslot].synchronize()                 # the host waits for frame j only
                if bool(self.flags_host[jslot].any()):    # same words on every rank: same decision everywhere
                    repeats += 1
                    for st in self._streams:
                        st.synchronize()
                    self.r.finish()                       # collects every frame in flight; bands were shrunk
                    i, inflight = j, []                   # restart the pipeline at frame j
                    continue
                self.frame = self.frames[jslot]
                if on_frame:
                    with torch.cuda.stream(self._streams[jslot]):
                        on_frame(j, self.frame)
        finally:
            for st in self._streams:
                st.synchronize()
            self.r.finish()
            self.r.set_frames_in_flight(1)
            for st in self._streams:
                main.wait_stream(st)
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
