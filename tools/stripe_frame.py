"""Render a few frames of rank `rank` of `world` (C3 scene, packed 4-row stripes) — the command ncu wraps
to list the per-level kernel durations of ONE rank's share.  usage: stripe_frame.py [world] [rank] [frames]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rt = importlib.import_module("html5-canvas-raytracer_b200")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
objs = rt.scenes.synth_objects(1024, 1024)
cam = rt.scenes.demo_camera(3840, 2160)
with rt.Renderer(objs) as r:
    dev = torch.empty((2160, 3840, 4), dtype=torch.uint8, device="cuda")
    o = rt.make_opts(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT, stripe_rows=4, stripe_rank=rank,
                     stripe_world=world, out_packed=1)
    for _ in range(frames):
        r.redraw_device(cam, o, dev.data_ptr())
    torch.cuda.synchronize()
    print("ok")
