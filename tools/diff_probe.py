"""Classify fp32-vs-fp64 pixel differences on C3 (fp64 kernel == oracle bit for bit)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
rt = importlib.import_module("html5-canvas-raytracer_b200")
objs = rt.scenes.synth_objects(1024, 1024)
W, H = 1920, 1080
cam = rt.scenes.demo_camera(W, H)
with rt.Renderer(objs) as r:
    ref = r.redraw(cam, segs=8, precision=rt.RTB_PRECISION_FP64).astype(np.int32)
    for name, kw in (("wavefront", dict(kernel=rt.RTB_KERNEL_WAVEFRONT)), ("pixel32", dict(kernel=rt.RTB_KERNEL_PIXEL))):
        img = r.redraw(cam, segs=8, **kw).astype(np.int32)
        d = np.abs(img - ref)[..., :3].max(axis=2)
        bad = d > 2
        print(f"{name}: exact {float((d == 0).mean()):.5f} within1 {float((d <= 1).mean()):.5f} within2 {float((d <= 2).mean()):.5f}")
        hist = [(lo, hi, int(((d > lo) & (d <= hi)).sum())) for lo, hi in ((0, 1), (1, 2), (2, 5), (5, 20), (20, 100), (100, 255))]
        print("   diff histogram (lo,hi]:", hist, "total px", d.size)
        ys, xs = np.nonzero(bad)
        # is the reference pixel a pure checker colour (yellow 255,255,0 / magenta 255,0,255 scaled)? -> floor
        rp = ref[ys, xs, :3]
        ip = img[ys, xs, :3]
        floorish = ((rp[:, 0] > 0) & ((rp[:, 1] == 0) | (rp[:, 2] == 0)) & ((ip[:, 1] == 0) | (ip[:, 2] == 0)))
        swapped = floorish & (((rp[:, 1] == 0) & (ip[:, 2] == 0) & (ip[:, 1] > 0)) | ((rp[:, 2] == 0) & (ip[:, 1] == 0) & (ip[:, 2] > 0)))
        print(f"   >2 LSB: {bad.sum()} px; checker-colour swaps (yellow<->magenta): {int(swapped.sum())}; rows with most: {np.bincount(ys, minlength=H).argsort()[-5:][::-1].tolist()}")
        # neighbourhood: is the bad pixel on an edge of the reference image (silhouette)?
        gx = np.abs(np.diff(ref[..., :3], axis=1)).max(axis=2); gy = np.abs(np.diff(ref[..., :3], axis=0)).max(axis=2)
        edge = np.zeros((H, W), bool); edge[:, :-1] |= gx > 24; edge[:, 1:] |= gx > 24; edge[:-1] |= gy > 24; edge[1:] |= gy > 24
        print(f"   of those, on a reference-image edge (silhouette/checker border): {int((bad & edge).sum())}; interior: {int((bad & ~edge).sum())}")
