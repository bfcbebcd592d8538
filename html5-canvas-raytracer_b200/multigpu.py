"""Multi-GPU host layer: one process per GPU (torchrun), scene replicated, frame partitioned.

  * row stripes (BASELINE config 3): stripes of `stripe_rows` rows are dealt round-robin to the ranks
    (stripe i -> rank i % world).  Contiguous bands would be badly imbalanced — the horizon splits the
    demo into cheap sky rows and expensive floor/sphere rows.  Every rank renders its
    rows PACKED into a [padded_rows, W, 4] buffer; the exchange is either the library's scatter kernel
    over NVLink peer memory (rows stored straight into every rank's framebuffer) or one
    `all_gather_into_tensor` (NCCL) plus a de-interleave kernel.  It exists only when world > 1.
    Sequences (redraw_sequence) are rendered in GROUPS of frames per wavefront pass (rtb_render_device_group),
    several passes in flight per GPU; with SharedHostFrames every rank also copies its own rows of every frame
    into one ring of host framebuffers in shared, page-locked memory (the box's PCIe links in parallel).
  * frame batches (config 5): frame f -> rank f % world; the rendered frames are all-gathered.

torch is used for process-group plumbing and device buffers only; all rendering goes through the
C ABI.  The partition arithmetic is pure Python so that it can be tested with gloo on CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from . import _abi as A
from .scene import make_opts

DEFAULT_STRIPE_ROWS = 1      # measured on the 4K / 1024-sphere frame at 8 ranks: the per-rank node counts spread by
                             # 0.24 % with single rows, 0.6 % with 2, 2 % with 4, 5 % with 8 rows per stripe, and
                             # the slowest rank's frame time falls from 1.660 ms (4) to 1.617 ms (1)
MAX_IN_FLIGHT = 8            # = RTB_MAX_IN_FLIGHT of the library: wavefront passes that may overlap on one GPU
MAX_GROUP = A.RTB_MAX_GROUP  # frames one wavefront pass renders together (rtb_render_device_group)
DEFAULT_IN_FLIGHT = 4        # passes redraw_sequence keeps in flight on one GPU (measured, 4K frames on one B200:
                             # 9.58 ms per frame with one in flight, 9.16 with four)
DEFAULT_IN_FLIGHT_MULTI = 6  # ... with world > 1, where a pass is a GROUP of frames (default_group): measured with 20-frame
                             # 4K sequences, ms per frame, (group, passes in flight): 2 GPUs (2,3) 4.07, (2,6) 3.92, (1,6) 3.98;
                             # 4 GPUs (4,3) 2.07, (4,6) 2.05, (2,6) 2.03; 8 GPUs (8,3) 1.058, (4,6) 1.068 [end to end 1.34 vs
                             # 1.23: smaller groups complete earlier, their host copies overlap the next passes], (2,8) 1.078
MAX_REPEATS = 64             # frames re-rendered because of ray-queue overflows before giving up (as rtb_render_batch)


def default_in_flight(world: int) -> int:
    return DEFAULT_IN_FLIGHT if world <= 1 else DEFAULT_IN_FLIGHT_MULTI


def default_group(world: int) -> int:
    """Frames per wavefront pass in sequence mode.  With the rows of a frame dealt over `world` GPUs, `world`
    frames per pass give every kernel launch the size it has on a single GPU; half of that (from 4 GPUs on) with
    six passes in flight measured as fast or faster and delivers frames earlier (see DEFAULT_IN_FLIGHT_MULTI)."""
    if world <= 2:
        return max(1, world)
    return max(1, min(world // 2, MAX_GROUP))


def split_groups(n: int, group: int) -> List[int]:
    """Sizes of the frame groups of an n-frame sequence: as few groups of at most `group` frames as possible,
    evened out (20 frames, group 8 -> 7, 7, 6)."""
    if n <= 0:
        return []
    k = -(-n // max(1, group))
    base, rem = divmod(n, k)
    return [base + 1] * rem + [base] * (k - rem)


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


class _DeviceBytes:
    """A raw device allocation of the library, viewable as a torch tensor (__cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr="|u1"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


# control block of the peer exchange, uint32 words; one row of RTB_MAX_PEERS words per frame slot
_CTL_READY, _CTL_STATUS, _CTL_ACK, _CTL_TIMEOUT, _CTL_WORDS = 0, 128, 256, 384, 512
# > 0: one-warp wait kernel that gives up after that many ms (a dead peer becomes an error, not a hang);
# 0: the stream itself waits (cuStreamWaitValue32: no kernel, no SM slot, no timeout).  Same speed at 8 GPUs.
_PEER_TIMEOUT_MS = int(os.environ.get("RTB_PEER_TIMEOUT_MS", "20000"))


class SharedHostFrames:
    """A ring of full host framebuffers ([n, H, W, 4] uint8) in POSIX shared memory that EVERY rank's process
    has mapped and page-locked.  In a striped sequence each rank then copies its own rows of every frame straight
    from its GPU into the frames (rtb_stripe_rows_to_host): the N GPUs of a box deliver a frame over their N PCIe
    links at once, instead of funnelling 33 MB per 4K frame through one GPU's link after the NVLink exchange.
    The consumer (any of the processes, or a host process that maps the same segment) reads `frames`.
    Collective: every rank constructs it and every rank calls close()."""

    def __init__(self, n: int, height: int, width: int, rank: int, world: int):
        import atexit
        import torch.distributed as dist
        from multiprocessing import resource_tracker, shared_memory

        self.n, self.H, self.W, self.rank, self.world = n, height, width, rank, world
        self.frame_bytes = height * width * 4
        self.shm = None
        size = max(1, n * self.frame_bytes)
        name, err = [None], None
        if rank == 0:
            try:
                self.shm = shared_memory.SharedMemory(create=True, size=size)
                name[0] = self.shm.name
                atexit.register(self._unlink_at_exit)     # a crash before close() must not leave the segment behind
            except Exception as e:                        # e.g. a /dev/shm too small for the ring
                err = f"shared memory: {e}"
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        self.lib = A.load()
        if name[0] is not None and rank != 0:
            try:
                self.shm = shared_memory.SharedMemory(name=name[0], create=False)
                try:                                      # attaching is not owning: only rank 0 unlinks
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:
                    pass
            except Exception as e:
                err = f"shared memory: {e}"
        registered = False
        if self.shm is not None and err is None:
            self.frames = np.ndarray((n, height, width, 4), dtype=np.uint8, buffer=self.shm.buf)
            self.base = self.frames.ctypes.data
            rc = self.lib.rtb_host_register(C.c_void_p(self.base), size)
            registered = rc == 0
            if not registered:
                err = "cudaHostRegister: " + self.lib.rtb_last_error().decode("utf-8", "replace")
        # every rank takes the same decision
        oks = [err is None and name[0] is not None]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, (oks[0], err))
            oks = [g[0] for g in gathered]
            err = next((g[1] for g in gathered if g[1]), err)
        if not all(oks):
            if registered:
                self.lib.rtb_host_unregister(C.c_void_p(self.base))
            self.frames = None
            if self.shm is not None:
                self.shm.close()
                if rank == 0:
                    self.shm.unlink()
                self.shm = None
            raise RuntimeError(f"shared host frames unavailable ({err or 'a rank could not map the segment'})")

    def _unlink_at_exit(self):
        try:
            if self.shm is not None and self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass

    def ptr(self, i: int) -> int:
        return self.base + (i % self.n) * self.frame_bytes

    def close(self):
        import torch.distributed as dist

        if getattr(self, "shm", None) is None:
            return
        self.lib.rtb_host_unregister(C.c_void_p(self.base))
        self.frames = None
        if self.world > 1:
            dist.barrier()
        shm, self.shm = self.shm, None
        shm.close()
        if self.rank == 0:
            shm.unlink()


class StripedRenderer:
    """Row-stripe sharded redraw(): every rank ends up with the full framebuffer on its GPU.

    exchange = "p2p" (default on GPUs): each rank stores its packed stripe rows straight into every rank's
    framebuffer over NVLink with one kernel of the library (rtb_peer_scatter_group: all-gather + de-interleave in
    one pass, no collective library in the data path, nothing spins while holding SMs); sequence numbers
    and per-rank status words travel in a small control block.  Buffers are shared with CUDA IPC handles
    exchanged once over torch.distributed.
    exchange = "nccl": `all_gather_into_tensor` of [group][padded_rows + 1][W][4] per rank — the packed stripe
    rows of every frame of the group plus one trailer row each, whose first word (frame 0) is the rank's
    status — followed by rtb_unstripe_device per frame.
    Either way a non-zero status word (that rank's ray queues overflowed) reaches every rank with the
    frames, so all ranks take the same repeat decision without a second exchange.
    "auto" tries p2p and falls back to nccl when the ranks cannot map each other's memory.

    Sequences (redraw_sequence) are rendered in GROUPS of `group` frames per wavefront pass
    (rtb_render_device_group; default = world), `slots` passes in flight."""

    def __init__(self, renderer, width: int, height: int, rank: int, world: int,
                 stripe_rows: int = DEFAULT_STRIPE_ROWS, exchange: str = "auto", group: int = 0, slots: int = 0):
        import torch

        self.r, self.W, self.H = renderer, width, height
        self.rank, self.world, self.S = rank, world, stripe_rows
        self.lib = A.load()
        self.n_local = self.lib.rtb_stripe_local_rows(height, stripe_rows, rank, world)
        self.padded = self.lib.rtb_stripe_padded_rows(height, stripe_rows, world)
        self.stride = self.padded + 1
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.G = G = max(1, min(MAX_GROUP, int(group) or default_group(world)))
        # `slots` sets of everything: redraw_sequence keeps several passes in flight
        self.K = K = max(1, min(MAX_IN_FLIGHT, int(slots) or (MAX_IN_FLIGHT if G == 1 else max(3, default_in_flight(world)))))
        self.flags_host = [torch.zeros((max(world, 1) + 1,), dtype=torch.int32).pin_memory() for _ in range(K)]
        self.status_dev = torch.zeros((K, 4), dtype=torch.int32, device=dev)      # world == 1: status words
        self._streams = None
        self.exchange = "none"
        self.exchange_note = ""
        self._peer = None
        if world > 1:
            self.local = [torch.zeros((G, self.stride, width, 4), dtype=torch.uint8, device=dev) for _ in range(K)]
            if exchange not in ("auto", "p2p", "nccl"):
                raise ValueError(f"exchange must be auto, p2p or nccl (got {exchange!r})")
            if exchange in ("auto", "p2p"):
                err = self._setup_p2p()
                if err is None:
                    self.exchange = "p2p"
                elif exchange == "p2p":
                    raise RuntimeError(f"peer-memory exchange unavailable: {err}")
                else:
                    self.exchange_note = f"p2p unavailable ({err})"
            if self.exchange != "p2p":
                self.exchange = "nccl"
                self.gathered = [torch.empty((world, G, self.stride, width, 4), dtype=torch.uint8, device=dev)
                                 for _ in range(K)]
        if self.exchange != "p2p":
            self.frames = [torch.empty((G, height, width, 4), dtype=torch.uint8, device=dev) for _ in range(K)]
        self.frame = self.frames[0][0]

    # -- peer-memory exchange: set-up ----------------------------------------------------------------------
    def _setup_p2p(self):
        """Allocate this rank's framebuffers + control block in the library, swap IPC handles, map the
        peers'.  Returns None on success, else a one-line reason — the same on every rank."""
        import torch
        import torch.distributed as dist

        K, G, world = self.K, self.G, self.world
        if world > 16 or self.W % 4 != 0:
            return "needs world <= 16 and a width that is a multiple of 4"
        frame_bytes = self.H * self.W * 4
        own = {"frames": C.c_void_p(), "ctl": C.c_void_p()}
        handles = {k: (C.c_uint8 * 64)() for k in own}
        err = None
        try:
            A.check(self.lib.rtb_peer_alloc(K * G * frame_bytes, C.byref(own["frames"]), handles["frames"]))
            A.check(self.lib.rtb_peer_alloc(_CTL_WORDS * 4, C.byref(own["ctl"]), handles["ctl"]))
        except A.RtbError as e:
            err = str(e)
        mine = None if err else (bytes(handles["frames"]), bytes(handles["ctl"]))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        if any(h is None for h in everyone):
            err = err or "a peer could not allocate its buffers"
        bases = {"frames": [0] * world, "ctl": [0] * world}
        opened = []
        if err is None:
            for p in range(world):
                for k, key in enumerate(("frames", "ctl")):
                    if p == self.rank:
                        bases[key][p] = own[key].value
                        continue
                    ptr = C.c_void_p()
                    h = (C.c_uint8 * 64).from_buffer_copy(everyone[p][k])
                    rc = self.lib.rtb_peer_open(h, C.byref(ptr))
                    if rc != 0:
                        err = f"rank {self.rank} cannot map rank {p}: {self.lib.rtb_last_error().decode()}"
                        break
                    opened.append(ptr.value)
                    bases[key][p] = ptr.value
                if err:
                    break
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            for ptr in opened:
                self.lib.rtb_peer_close(C.c_void_p(ptr))
            dist.barrier()
            for key in own:
                if own[key].value:
                    self.lib.rtb_peer_free(own[key])
            return err or "a peer could not map this rank's buffers"
        VP = C.c_void_p * world
        self._peer = {
            "own": own, "opened": opened, "frame_bytes": frame_bytes,
            # per slot: every rank's first framebuffer of that slot (the group's frames follow contiguously)
            "frames": [VP(*[bases["frames"][p] + k * G * frame_bytes for p in range(world)]) for k in range(K)],
            "ctl": VP(*bases["ctl"]),
            "ctl_base": own["ctl"].value,
            "seq": 0, "slot_seq": [0] * K, "unacked": [0] * K,
        }
        self.frames = [torch.as_tensor(_DeviceBytes(own["frames"].value + k * G * frame_bytes, (G, self.H, self.W, 4)),
                                       device=self.dev) for k in range(K)]
        self._ctl = torch.as_tensor(_DeviceBytes(own["ctl"].value, (_CTL_WORDS,), "<i4"), device=self.dev)
        return None

    def close(self):
        """Unmap the peers' buffers and free this rank's (collective: every rank calls it)."""
        import torch
        import torch.distributed as dist

        if self._peer is None:
            return
        torch.cuda.synchronize()
        dist.barrier()
        for ptr in self._peer["opened"]:
            self.lib.rtb_peer_close(C.c_void_p(ptr))
        dist.barrier()
        for key in ("frames", "ctl"):
            self.lib.rtb_peer_free(self._peer["own"][key])
        self._peer = None
        self.frames = []

    # -- building blocks (everything is enqueued on the CURRENT torch stream) --------------------------
    def _enqueue_render(self, cams, slot, stats, kw):
        """One wavefront pass for the frames `cams` (a group) into slot `slot`."""
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        n = len(cams)
        if self.world <= 1:
            outs = [self.frames[slot][b].data_ptr() for b in range(n)]
            self.r.redraw_device_group(cams, make_opts(**kw), outs, stream, stats)
            A.check(self.lib.rtb_copy_status_device(self.r._h, self.status_dev[slot].data_ptr(), A.stream_arg(stream)))
            return
        o = make_opts(stripe_rows=self.S, stripe_rank=self.rank, stripe_world=self.world, out_packed=1, **kw)
        loc = self.local[slot]
        self.r.redraw_device_group(cams, o, [loc[b].data_ptr() for b in range(n)], stream, stats)
        A.check(self.lib.rtb_copy_status_device(self.r._h, loc[0, self.padded].data_ptr(), A.stream_arg(stream)))

    def _enqueue_gather(self, slot):
        import torch.distributed as dist

        # concatenated-along-dim-0 view: the form both NCCL and gloo accept
        return dist.all_gather_into_tensor(self.gathered[slot].view(-1, self.W, 4), self.local[slot].view(-1, self.W, 4),
                                           async_op=True)

    def _enqueue_unstripe(self, slot, n):
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        for b in range(n):
            # frame b of every rank sits at [rank][b]; rank stride = G * stride rows
            A.check(self.lib.rtb_unstripe_device(self.gathered[slot][0, b].data_ptr(), self.frames[slot][b].data_ptr(),
                                                 self.W, self.H, self.S, self.world, self.G * self.stride,
                                                 A.stream_arg(stream)))
        flags = self.gathered[slot][:, 0, self.padded, 0, :].contiguous().view(torch.int32).view(-1)
        self.flags_host[slot][:self.world].copy_(flags, non_blocking=True)

    def _enqueue_release(self, slot):
        """p2p: tell every peer that this rank no longer reads the framebuffers of `slot` (its previous group)."""
        import torch

        P = self._peer
        if P["unacked"][slot]:
            st = A.stream_arg(torch.cuda.current_stream().cuda_stream)
            A.check(self.lib.rtb_peer_signal(P["ctl"], self.world, self.rank, _CTL_ACK + 16 * slot, P["unacked"][slot], st))
            P["unacked"][slot] = 0

    def _enqueue_scatter(self, slot, n, host=None, first_frame=0):
        """p2p: wait until every peer has released the framebuffers of `slot`, store this rank's rows of the
        group's n frames into all of them, wait until every rank's rows have arrived here.  With `host`, the
        rank's rows also leave for the shared host frames right behind the scatter kernel (PCIe, copy engine),
        while the stream waits for the peers."""
        import torch

        P = self._peer
        st = A.stream_arg(torch.cuda.current_stream().cuda_stream)
        P["seq"] += 1
        seq, prev = P["seq"], P["slot_seq"][slot]
        ctl = P["ctl_base"]
        tflag = C.c_void_p(ctl + 4 * _CTL_TIMEOUT)
        if prev:
            A.check(self.lib.rtb_peer_wait(C.c_void_p(ctl + 4 * (_CTL_ACK + 16 * slot)), self.world, prev,
                                           _PEER_TIMEOUT_MS, tflag, st))
        loc = self.local[slot]
        A.check(self.lib.rtb_peer_scatter_group(C.c_void_p(loc.data_ptr()), self.stride * self.W * 4, n,
                                                P["frames"][slot], P["frame_bytes"], P["ctl"], self.W, self.H,
                                                self.S, self.rank, self.world,
                                                C.c_void_p(loc[0, self.padded].data_ptr()),
                                                _CTL_READY + 16 * slot, _CTL_STATUS + 16 * slot, seq, st))
        if host is not None:
            self._enqueue_rows_to_host(slot, n, host, first_frame)
        A.check(self.lib.rtb_peer_wait(C.c_void_p(ctl + 4 * (_CTL_READY + 16 * slot)), self.world, seq,
                                       _PEER_TIMEOUT_MS, tflag, st))
        P["slot_seq"][slot] = seq
        P["unacked"][slot] = seq
        s0 = _CTL_STATUS + 16 * slot
        self.flags_host[slot][:self.world].copy_(self._ctl[s0:s0 + self.world], non_blocking=True)
        self.flags_host[slot][self.world:].copy_(self._ctl[_CTL_TIMEOUT:_CTL_TIMEOUT + 1], non_blocking=True)

    def _mark(self, tr, name):
        """Tracing (tools/seq_trace.py): a timed CUDA event on the current stream + the host clock."""
        if tr is not None:
            import time
            import torch
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            tr[name] = (ev, time.perf_counter())

    def _enqueue_rows_to_host(self, slot, n, host, first_frame):
        """This rank's rows of the group's n frames -> the shared host frames first_frame .. (one strided D2H per
        frame, on the current stream, straight from the packed stripe buffers / world 1: the framebuffers)."""
        import torch

        st = A.stream_arg(torch.cuda.current_stream().cuda_stream)
        for b in range(n):
            src = self.frames[slot][b] if self.world <= 1 else self.local[slot][b]
            A.check(self.lib.rtb_stripe_rows_to_host(C.c_void_p(src.data_ptr()), C.c_void_p(host.ptr(first_frame + b)),
                                                     self.W, self.H, self.S, self.rank, max(self.world, 1), st))

    def _enqueue_group(self, cams, slot, kw, stats=None, tr=None, host=None, first_frame=0):
        """Render + exchange of one group of frames into self.frames[slot][0..n); status words to the host."""
        n = len(cams)
        if n > self.G:
            raise ValueError(f"group of {n} frames exceeds this renderer's group size {self.G}")
        self._mark(tr, "begin")
        if self.exchange == "p2p":
            self._enqueue_release(slot)
        self._enqueue_render(cams, slot, stats, kw)
        self._mark(tr, "rendered")
        if host is not None and self.exchange == "nccl":
            self._enqueue_rows_to_host(slot, n, host, first_frame)
        if self.exchange == "p2p":
            self._enqueue_scatter(slot, n, host, first_frame)
        elif self.exchange == "nccl":
            self._enqueue_gather(slot).wait()             # stream-level wait, the host does not block
            self._enqueue_unstripe(slot, n)
        else:
            if host is not None:
                self._enqueue_rows_to_host(slot, n, host, first_frame)
            self.flags_host[slot][:1].copy_(self.status_dev[slot, :1], non_blocking=True)

    def _frame_ok(self, slot) -> bool:
        """Host side, after the group's stream work has completed: did every rank render it validly?"""
        f = self.flags_host[slot]
        if self.exchange == "p2p" and int(f[self.world]) != 0:
            raise RuntimeError("peer exchange timed out: a rank stopped publishing frames")
        return not bool(f[:max(self.world, 1)].any())

    @staticmethod
    def _overflow_error():
        return A.RtbError(A.RTB_ERR_OVERFLOW, f"a frame still overflows its ray queues after {MAX_REPEATS} attempts "
                                              "(one-row bands do not fit the queue capacity)")

    # -- one frame (latency mode) ----------------------------------------------------------------------
    def redraw(self, camera, stats=None, **kw):
        """Render this rank's stripes, exchange, return the device framebuffer (valid until the next call).
        Everything is only ENQUEUED; the host waits once.  If any rank's ray queues overflowed, every
        rank sees that in the status words and repeats the frame (the library has shrunk the offending
        rank's bands, deterministically) - at most MAX_REPEATS times, then RtbError(RTB_ERR_OVERFLOW)."""
        return self.redraw_group([camera], stats, **kw)[0]

    def redraw_group(self, cameras, stats=None, **kw):
        """The same for a group of up to `group` frames in ONE wavefront pass; returns the [n, H, W, 4] device view."""
        import torch

        for _ in range(MAX_REPEATS):
            self._enqueue_group(list(cameras), 0, kw, stats)
            torch.cuda.current_stream().synchronize()
            self.r.finish()                               # clears the library's pending state (already complete)
            if self._frame_ok(0):
                self.frame = self.frames[0][0]
                return self.frames[0][:len(cameras)]
        raise self._overflow_error()

    # -- a sequence of frames (throughput mode) ----------------------------------------------------------
    def redraw_sequence(self, cameras, on_frame=None, frames_in_flight: int = 0, group: int = 0, trace=None,
                        host_frames=None, **kw):
        """Software-pipelined: the sequence is cut into groups of up to `group` frames (0 = this renderer's group
        size; one wavefront pass each) and up to `frames_in_flight` passes (0 = default_in_flight(world)) are
        enqueued before the host waits for the oldest one, each on its own stream with its own library workspace.
        Later groups render while group i is in the exchange, and the kernels of the passes interleave on the
        GPU.  `on_frame(i, frame)` is called (device tensor, valid until `frames_in_flight - 1` further groups
        have been issued) once frame i is assembled on this rank.  Returns the number of groups that had to be
        repeated because of queue overflows (every rank reads the same status words, so every rank repeats the
        same groups); more than MAX_REPEATS raise RtbError(RTB_ERR_OVERFLOW).
        `host_frames` (a SharedHostFrames of at least len(cameras) frames, constructed by every rank): every rank
        also copies ITS rows of frame i into host_frames.frames[i]; when the call returns on every rank (the caller
        puts a barrier behind it), the frames are complete in host memory."""
        import torch

        if host_frames is not None and host_frames.n < len(cameras):
            raise ValueError("host_frames holds fewer frames than the sequence")

        depth = max(1, min(self.K, int(frames_in_flight) or default_in_flight(self.world)))
        sizes = split_groups(len(cameras), max(1, min(self.G, int(group) or self.G)))
        starts = [sum(sizes[:i]) for i in range(len(sizes))]
        n = len(sizes)
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
                        tr = None
                        if trace is not None:
                            tr = {"group": i, "frames": sizes[i]}
                            trace.append(tr)
                        self._enqueue_group(cameras[starts[i]:starts[i] + sizes[i]], slot, kw, tr=tr,
                                            host=host_frames, first_frame=starts[i])
                        self._mark(tr, "done")
                        done[slot].record()
                    inflight.append(i)
                    i += 1
                j = inflight.pop(0)
                jslot = j % depth
                done[jslot].synchronize()                 # the host waits for group j only
                if trace is not None:
                    import time
                    trace[j]["host_done"] = time.perf_counter()
                if not self._frame_ok(jslot):             # same words on every rank: same decision everywhere
                    repeats += 1
                    if repeats > MAX_REPEATS:
                        raise self._overflow_error()
                    for st in self._streams:
                        st.synchronize()
                    self.r.finish()                       # collects every pass in flight; bands were shrunk
                    i, inflight = j, []                   # restart the pipeline at group j
                    continue
                self.frame = self.frames[jslot][sizes[j] - 1]
                if on_frame:
                    with torch.cuda.stream(self._streams[jslot]):
                        for b in range(sizes[j]):
                            on_frame(starts[j] + b, self.frames[jslot][b])
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

    def redraw_step(self, cameras: Sequence, first: int, stats=None, objects_for_frame=None, **kw):
        """Frames first .. first+world-1 (cameras beyond the end are skipped).  Returns [world,H,W,4].
        objects_for_frame(f) -> the frame's own (re-sorted) object list: config C5 re-sorts the scene
        for every camera position (scenes.orbit_frames); without it every frame uses self.r's scene."""
        import torch
        import torch.distributed as dist

        from .render import Renderer

        stream = torch.cuda.current_stream().cuda_stream
        f = first + self.rank

        def render_valid(r):
            # An asynchronous wavefront frame may come back invalid (ray queues overflowed; the library has
            # then shrunk its bands).  Collect the verdict and re-render HERE, before the collective: every
            # rank enters the all-gather exactly once, with a valid frame, whatever happened on the others.
            for _ in range(MAX_REPEATS):
                r.redraw_device(cameras[f], make_opts(**kw), self.mine.data_ptr(), stream, stats)
                torch.cuda.current_stream().synchronize()
                if r.finish():
                    return
            raise StripedRenderer._overflow_error()

        if f < len(cameras):
            if objects_for_frame is not None:
                with Renderer(objects_for_frame(f), device=torch.cuda.current_device()) as r:
                    render_valid(r)
            else:
                render_valid(self.r)
        if self.world <= 1:
            self.step[0].copy_(self.mine)
        else:
            dist.all_gather_into_tensor(self.step.view(-1, self.W, 4), self.mine)
        return self.step
