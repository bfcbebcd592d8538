"""Multi-GPU correctness check (run under torchrun, >= 2 GPUs): StripedRenderer single frames and the
pipelined sequence must equal the single-GPU render byte for byte — also when ONE rank's ray queues
overflow (forced with a tiny RTB_WF_CAP_NODES on the last rank) and the frame has to be repeated."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
rt = importlib.import_module("html5-canvas-raytracer_b200")
mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
os.environ["NCCL_DEBUG"] = "WARN"
rank, local_rank, world = mg.init_distributed()
torch.cuda.set_device(local_rank)
W, H = 640, 362
XCH = os.environ.get("RTB_EXCHANGE", "auto")
objs = rt.scenes.synth_objects(256, 11)
demo = rt.scenes.demo_objects()
cams = rt.scenes.orbit_cameras(W, H, nframes=12)[:9]
kw = dict(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)
with rt.Renderer(objs, device=local_rank) as r:
    refs = [r.redraw(c, **kw) for c in cams]                       # single-GPU frames (every rank computes them)
    sr = mg.StripedRenderer(r, W, H, rank, world, exchange=XCH)
    ok = True
    for ci, (c, ref) in enumerate(zip(cams, refs)):
        f = sr.redraw(c, **kw).cpu().numpy()
        same = bool(np.array_equal(f, ref))
        if not same:
            bad = np.argwhere((f != ref).any(axis=2))
            print(f"rank {rank}: single frame {ci} differs in {len(bad)} px, rows {sorted(set(bad[:,0].tolist()))[:12]} flags {[f.tolist() for f in sr.flags_host]}", flush=True)
        ok &= same
    got = {}
    rep = sr.redraw_sequence(cams, on_frame=lambda i, fr: got.__setitem__(i, fr.cpu().numpy().copy()), **kw)
    seq_ok = [bool(np.array_equal(got[i], refs[i])) for i in range(len(cams))]
    if not all(seq_ok) or rep:
        print(f"rank {rank}: sequence ok={seq_ok} repeats={rep}", flush=True)
        for i in range(len(cams)):
            if not seq_ok[i]:
                bad = np.argwhere((got[i] != refs[i]).any(axis=2))
                rows = sorted(set(bad[:, 0].tolist()))
                owners = sorted(set((y // sr.S) % world for y in rows))
                same_as = [k for k in range(len(cams)) if np.array_equal(got[i], refs[k])]
                print(f"rank {rank}:   frame {i}: {len(bad)} px differ, {len(rows)} rows, owner ranks {owners}, equals ref of frame {same_as}", flush=True)
    ok &= all(seq_ok) and rep == 0
    # host delivery: every rank copies its own rows into a shared, page-locked host ring
    ring = mg.SharedHostFrames(len(cams), H, W, rank, world)
    rep_h = sr.redraw_sequence(cams, host_frames=ring, **kw)
    torch.cuda.synchronize()
    dist.barrier()
    host_ok = [bool(np.array_equal(ring.frames[i], refs[i])) for i in range(len(cams))]
    if not all(host_ok) or rep_h:
        print(f"rank {rank}: shared host frames ok={host_ok} repeats={rep_h}", flush=True)
    ok &= all(host_ok) and rep_h == 0
    ring.close()
    # force an overflow on the LAST rank only: the verdict travels in the gather payload
    if rank == world - 1:                                  # bands sized for 1.1 nodes per ray: the first attempt overflows,
        os.environ["RTB_WF_CAP_NODES"] = str(W * 9 * mg.default_group(world))   # halved bands (2.2, 4.4 ...) fit
        os.environ["RTB_WF_NODES_PER_RAY"] = "1.1"
    with rt.Renderer(demo, device=local_rank) as rd:       # demo scene: glass/mirror rows overflow tiny queues
        saved = {k: os.environ.pop(k, None) for k in ("RTB_WF_CAP_NODES", "RTB_WF_NODES_PER_RAY")}
        drefs = [rd.redraw(c, **kw) for c in cams]
        os.environ.update({k: v for k, v in saved.items() if v})
    with rt.Renderer(demo, device=local_rank) as r2:
        sr2 = mg.StripedRenderer(r2, W, H, rank, world, exchange=XCH)
        f = sr2.redraw(cams[1], **kw).cpu().numpy()
        s2 = bool(np.array_equal(f, drefs[1]))
        if not s2:
            print(f"rank {rank}: forced-overflow single frame differs", flush=True)
        ok &= s2
        sr2.close()
    with rt.Renderer(demo, device=local_rank) as r3:
        sr3 = mg.StripedRenderer(r3, W, H, rank, world, exchange=XCH)
        got2 = {}
        rep2 = sr3.redraw_sequence(cams, on_frame=lambda i, fr: got2.__setitem__(i, fr.cpu().numpy().copy()), **kw)
        s3 = [bool(np.array_equal(got2[i], drefs[i])) for i in range(len(cams))]
        if not all(s3) or rep2 == 0:
            print(f"rank {rank}: forced-overflow sequence ok={s3} repeats={rep2}", flush=True)
        ok &= all(s3) and rep2 > 0
        sr3.close()
    os.environ.pop("RTB_WF_CAP_NODES", None)
    os.environ.pop("RTB_WF_NODES_PER_RAY", None)
    sr.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"mgpu_check world={world} exchange={sr.exchange} {sr.exchange_note}: {'OK' if int(t.item()) else 'MISMATCH'} (sequence repeats without/with forced overflow: {rep}/{rep2})")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
