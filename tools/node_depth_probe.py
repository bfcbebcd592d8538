"""Where do fp32 and fp64 ray trees start to differ?  Cumulative node / shadow-ray counts by depth."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rt = importlib.import_module("html5-canvas-raytracer_b200")
objs = rt.scenes.synth_objects(1024, 1024)
cam = rt.scenes.demo_camera(960, 540)
with rt.Renderer(objs) as r:
    prev = (0, 0)
    for segs in range(1, 9):
        _, s64 = r.redraw(cam, stats=True, segs=segs, precision=rt.RTB_PRECISION_FP64)
        _, s32 = r.redraw(cam, stats=True, segs=segs, kernel=rt.RTB_KERNEL_WAVEFRONT)
        _, p32 = r.redraw(cam, stats=True, segs=segs, kernel=rt.RTB_KERNEL_PIXEL)
        print(f"segs {segs}: nodes fp64 {s64.nodes} wf32 {s32.nodes} ({s32.nodes - s64.nodes:+d}) px32 {p32.nodes} ({p32.nodes - s64.nodes:+d}) | level nodes fp64 {s64.nodes - prev[0]} wf32 {s32.nodes - prev[1]}")
        prev = (s64.nodes, s32.nodes)
