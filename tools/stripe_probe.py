"""Per-rank cost of the striped render WITHOUT the collective: render the stripes of rank r of `world`
on one GPU (what each of N GPUs executes).  usage: stripe_probe.py [world ...]"""
import ctypes as C, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rt = importlib.import_module("html5-canvas-raytracer_b200")
A = rt._abi
objs = rt.scenes.synth_objects(1024, 1024)
cam = rt.scenes.demo_camera(3840, 2160)
with rt.Renderer(objs) as r:
    dev = torch.empty((2160, 3840, 4), dtype=torch.uint8, device="cuda")
    for world in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
        res = []
        for rank in sorted({0, world // 2, world - 1}):
            o = rt.make_opts(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT, stripe_rows=int(os.environ.get("STRIPE_ROWS", "4")), stripe_rank=rank,
                             stripe_world=world, out_packed=1)
            for _ in range(3):
                r.redraw_device(cam, o, dev.data_ptr())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                r.redraw_device(cam, o, dev.data_ptr())
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 100
            st = A.RtbStats()
            r.redraw_device(cam, o, dev.data_ptr(), None, st)
            res.append((rank, wall, st.kernel_ms, st.ms_trace_primary, st.ms_trace_secondary, st.ms_shadow, st.ms_compose, st.nodes))
        for x in res:
            print(f"world {world} rank {x[0]}: wall {x[1]:.3f} ms/frame  kernels {x[2]:.3f} (primary {x[3]:.3f} secondary {x[4]:.3f} shadow {x[5]:.3f} compose {x[6]:.3f}) nodes {x[7]}")
