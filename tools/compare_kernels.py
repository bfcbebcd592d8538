"""fp32 per-pixel kernel vs fp32 wavefront on a BASELINE config (kernel_ms from CUDA events)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
rt = importlib.import_module("html5-canvas-raytracer_b200")
for name in sys.argv[1:] or ["c2_demo_1080p_segs5"]:
    scene, W, H, segs, g = rt.scenes.config_workload(name)
    objs = rt.scenes.demo_objects() if scene == "demo" else rt.scenes.synth_objects(1024, 1024)
    cam = rt.scenes.demo_camera(W, H)
    with rt.Renderer(objs) as r:
        imgs = {}
        for kname, k in (("pixel", rt.RTB_KERNEL_PIXEL), ("wavefront", rt.RTB_KERNEL_WAVEFRONT)):
            ms = []
            for _ in range(4):
                img, st = r.redraw(cam, stats=True, segs=segs, spp_grid=g, kernel=k)
                ms.append(st.kernel_ms)
            imgs[kname] = img
            print(f"{name} {kname}: kernel {min(ms[1:]):.3f} ms, launches {st.launches}, bands {st.bands}")
        d = np.abs(imgs["pixel"].astype(int) - imgs["wavefront"].astype(int))[..., :3].max(axis=2)
        print(f"  pixel vs wavefront: identical {float((d == 0).mean()):.5f}, within2 {float((d <= 2).mean()):.5f}")
