"""Render a few frames of a BASELINE config through the C ABI — the command ncu wraps.
usage: python tools/profile_frame.py [config] [frames]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rt = importlib.import_module("html5-canvas-raytracer_b200")

name = sys.argv[1] if len(sys.argv) > 1 else "c3_synth1024_4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
scene, W, H, segs, g = rt.scenes.config_workload(name)
objs = rt.scenes.demo_objects() if scene == "demo" else rt.scenes.synth_objects(1024, 1024)
cam = rt.scenes.demo_camera(W, H)
with rt.Renderer(objs) as r:
    for f in range(frames):
        img, st = r.redraw(cam, stats=True, segs=segs, spp_grid=g, kernel=rt.RTB_KERNEL_WAVEFRONT)
        print(f"frame {f}: {st.kernel_ms:.3f} ms kernels, {st.launches} launches, primary {st.ms_trace_primary:.3f} "
              f"secondary {st.ms_trace_secondary:.3f} shadow {st.ms_shadow:.3f} compose {st.ms_compose:.3f}")
