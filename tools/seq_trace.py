"""Timeline of a striped frame sequence (run under torchrun): per group and rank, when its render began and
ended on the GPU, when its exchange completed, and when the host enqueued / collected it.
usage: torchrun ... tools/seq_trace.py [frames] [passes in flight] [group]  -> gpurun_out/seq_trace_rank<r>.json"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rt = importlib.import_module("html5-canvas-raytracer_b200")
mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
os.environ["NCCL_DEBUG"] = "WARN"
rank, local_rank, world = mg.init_distributed()
torch.cuda.set_device(local_rank)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
fif = int(sys.argv[2]) if len(sys.argv) > 2 else 0
grp = int(sys.argv[3]) if len(sys.argv) > 3 else 0
W, H = 3840, 2160
cam = rt.scenes.demo_camera(W, H)
kw = dict(segs=8, kernel=rt.RTB_KERNEL_AUTO)
out = []
with rt.Renderer(rt.scenes.synth_objects(1024, 1024), device=local_rank) as r:
    sr = mg.StripedRenderer(r, W, H, rank, world, group=grp, slots=max(fif or mg.default_in_flight(world), 3) if world > 1 else 0)
    sr.redraw_sequence([cam] * (4 * sr.G), frames_in_flight=fif, **kw)
    for rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        trace = []
        sr.redraw_sequence([cam] * N, frames_in_flight=fif, trace=trace, **kw)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        rec = {"rep": rep, "total_ms_gpu": ev0.elapsed_time(ev1), "total_ms_host": (t1 - t0) * 1e3, "groups": []}
        for tr in trace:
            g = {"group": tr["group"], "frames": tr["frames"]}
            for k in ("begin", "rendered", "done"):
                g[k + "_gpu"] = round(ev0.elapsed_time(tr[k][0]), 3)
                g[k + "_host"] = round((tr[k][1] - t0) * 1e3, 3)
            g["host_done"] = round((tr["host_done"] - t0) * 1e3, 3)
            rec["groups"].append(g)
        out.append(rec)
    sr.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"seq_trace_rank{rank}.json"), "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
