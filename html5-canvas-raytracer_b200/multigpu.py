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
    """Row-stripe sharded redraw(): every rank ends up with the full framebuffer on its GPU."""

    def __init__(self, renderer, width: int, height: int, rank: int, world: int,
                 stripe_rows: int = DEFAULT_STRIPE_ROWS):
        import torch

        self.r, self.W, self.H = renderer, width, height
        self.rank, self.world, self.S = rank, world, stripe_rows
        self.lib = A.load()
        self.n_local = self.lib.rtb_stripe_local_rows(height, stripe_rows, rank, world)
        self.padded = self.lib.rtb_stripe_padded_rows(height, stripe_rows, world)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.frame = torch.empty((height, width, 4), dtype=torch.uint8, device=dev)
        if world > 1:
            self.local = torch.zeros((self.padded, width, 4), dtype=torch.uint8, device=dev)
            self.gathered = torch.empty((world, self.padded, width, 4), dtype=torch.uint8, device=dev)

    def redraw(self, camera, stats=None, **kw):
        """Render this rank's stripes, all-gather, de-interleave.  Returns the device framebuffer."""
        import torch
        import torch.distributed as dist

        stream = torch.cuda.current_stream().cuda_stream
        # Everything below is only ENQUEUED (render, collective, de-interleave); the host then waits once
        # in finish().  If a rank's ray queues overflowed, every rank repeats the frame (the collective
        # must stay matched), which the deterministic band shrink makes succeed.
        while True:
            if self.world <= 1:
                self.r.redraw_device(camera, make_opts(**kw), self.frame.data_ptr(), stream, stats)
            else:
                o = make_opts(stripe_rows=self.S, stripe_rank=self.rank, stripe_world=self.world, out_packed=1, **kw)
                self.r.redraw_device(camera, o, self.local.data_ptr(), stream, stats)
                # concatenated-along-dim-0 view: the form both NCCL and gloo accept
                dist.all_gather_into_tensor(self.gathered.view(-1, self.W, 4), self.local)
                A.check(self.lib.rtb_unstripe_device(self.gathered.data_ptr(), self.frame.data_ptr(), self.W,
                                                     self.H, self.S, self.world, stream))
            ok = self.r.finish()
            if self.world > 1:
                flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=self.frame.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                ok = int(flag.item()) == 0
            if ok:
                return self.frame


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
