"""A few tiny wavefront frames (fixed per-launch costs dominate) — the command ncu wraps."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rt = importlib.import_module("html5-canvas-raytracer_b200")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 36)
objs = rt.scenes.synth_objects(1024, 1024)
cam = rt.scenes.demo_camera(W, H)
with rt.Renderer(objs) as r:
    dev = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    o = rt.make_opts(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)
    for _ in range(3):
        r.redraw_device(cam, o, dev.data_ptr())
    torch.cuda.synchronize()
    print("ok")
