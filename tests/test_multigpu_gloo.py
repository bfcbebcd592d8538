"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over gloo: stripe partition, packed
per-rank buffers, all_gather layout and de-interleave must reassemble the single-process frame byte
for byte.  The per-rank 'renderer' here is the CPU oracle restricted to the rank's stripes (tests may
use the oracle; on a GPU box the same layer drives rtb_render_device + NCCL)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, S, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    rt = importlib.import_module("html5-canvas-raytracer_b200")
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    from oracle import oracle
    r, lr, w = mg.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    sd = rt.SceneDesc(rt.scenes.demo_objects())
    cam = rt.scenes.demo_camera(W, H)
    full = np.zeros((H, W, 4), np.uint8)
    oracle.render(sd, cam, rt.make_opts(segs=4, stripe_rows=S, stripe_rank=rank, stripe_world=world), out=full, threads=1)
    ys = mg.stripe_rows_of(H, S, rank, world)
    padded = mg.padded_rows(H, S, world)
    local = torch.zeros((padded, W, 4), dtype=torch.uint8)
    local[:len(ys)] = torch.from_numpy(full[ys])          # the packed layout rtb_render_device(out_packed=1) writes
    gathered = torch.empty((world, padded, W, 4), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered.view(-1, W, 4), local)   # same call shape as StripedRenderer.redraw
    frame = mg.unstripe_numpy(gathered.numpy(), H, S, world)
    # frame batches: frame f -> rank f % world
    mine = mg.frames_of(7, rank, world)
    counts = torch.tensor([len(mine)])
    dist.all_reduce(counts)
    q.put((rank, frame, int(counts.item()), mine))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,S,H", [(2, 4, 30), (3, 16, 50)])
def test_striped_gather_reassembles_the_frame(rt, oracle, world, S, H):
    import torch.multiprocessing as mp
    W = 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = rt.SceneDesc(rt.scenes.demo_objects())
    ref, _ = oracle.render(sd, rt.scenes.demo_camera(W, H), rt.make_opts(segs=4))
    seen = set()
    for rank, frame, total, mine in results:
        assert np.array_equal(frame, ref), f"rank {rank}"
        assert total == 7
        seen.update(mine)
    assert seen == set(range(7))


def test_single_process_degenerates_to_no_collective(rt):
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    assert mg.stripe_rows_of(10, 4, 0, 1) == list(range(10))
    assert mg.padded_rows(10, 4, 1) == 10
    g = np.arange(10 * 3 * 4, dtype=np.uint8).reshape(1, 10, 3, 4)
    assert np.array_equal(mg.unstripe_numpy(g, 10, 4, 1), g[0])
    os.environ.pop("WORLD_SIZE", None)
    assert mg.init_distributed()[2] == 1


def test_split_groups_partitions_a_sequence(rt):
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    assert mg.split_groups(20, 8) == [7, 7, 6] and mg.split_groups(20, 1) == [1] * 20
    assert mg.split_groups(16, 8) == [8, 8] and mg.split_groups(3, 8) == [3] and mg.split_groups(0, 8) == []
    for n in range(1, 60):
        for g in range(1, 9):
            s = mg.split_groups(n, g)
            assert sum(s) == n and max(s) <= g and max(s) - min(s) <= 1 and len(s) == -(-n // g)
    assert [mg.default_group(w) for w in (1, 2, 3, 4, 8, 16, 32)] == [1, 2, 1, 2, 4, 8, mg.MAX_GROUP]


def _ring_worker(rank, world, port, q):
    """SharedHostFrames over gloo WITHOUT a GPU: the segment is created and mapped by every rank, page-locking it
    fails (no device - there is no CPU fallback in the library), and every rank must reach the SAME verdict and
    leave nothing behind in /dev/shm."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    mg.init_distributed(backend="gloo")
    before = set(os.listdir("/dev/shm")) if os.path.isdir("/dev/shm") else set()
    try:
        ring = mg.SharedHostFrames(3, 8, 16, rank, world)
        verdict = "ok"
        ring.close()
    except RuntimeError as e:
        verdict = str(e)
    dist.barrier()
    after = set(os.listdir("/dev/shm")) if os.path.isdir("/dev/shm") else set()
    q.put((rank, verdict, sorted(after - before)))
    dist.barrier()
    dist.destroy_process_group()


def test_shared_host_frames_fail_collectively_without_a_gpu(rt):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("checks the no-device verdict")
    except ImportError:
        pytest.skip("torch missing")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, verdict, leaked in results:
        assert "unavailable" in verdict and "cudaHostRegister" in verdict, (rank, verdict)
        assert leaked == [], leaked
