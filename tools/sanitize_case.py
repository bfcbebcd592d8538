"""A small case that walks every kernel of the split and fused pipelines, frame groups, stripes and the group
scatter - the command compute-sanitizer wraps (one tool per call, see B200_PROFILING.md).
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
rt = importlib.import_module("html5-canvas-raytracer_b200")
W, H = 96, 54
cams = rt.scenes.orbit_cameras(W, H, 3)
lights = [[5, 10, 5], [5, 10, 0], [-6, 8, 4]]
with rt.Renderer(rt.scenes.synth_objects(96, 3), lights=lights) as r:
    ref = None
    for bits in ("7", "0", "5"):
        os.environ["RTB_WF_SPLIT"] = bits
        dev = torch.zeros((3, H, W, 4), dtype=torch.uint8, device="cuda")
        r.redraw_device_group(cams, rt.make_opts(segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT, spp_grid=2,
                                                star_mode=rt.RTB_STARS_HASH), [dev[i].data_ptr() for i in range(3)])
        assert r.finish()
        torch.cuda.synchronize()
        a = dev.cpu().numpy()
        assert ref is None or np.array_equal(a, ref), bits
        ref = a
        # striped + packed, tiny queues (several bands)
        os.environ["RTB_WF_CAP_NODES"] = str(W * 2 * 3 * 4 * 6)
        rows = rt._abi.load().rtb_stripe_local_rows(H, 1, 1, 3)
        loc = torch.zeros((2, rows, W, 4), dtype=torch.uint8, device="cuda")
        st = rt._abi.RtbStats()
        r.redraw_device_group(cams[:2], rt.make_opts(segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT, stripe_rows=1, stripe_rank=1,
                                                     stripe_world=3, out_packed=1), [loc[i].data_ptr() for i in range(2)], stats=st)
        del os.environ["RTB_WF_CAP_NODES"]
        print(f"split {bits}: group ok, striped group {st.bands} bands, {st.nodes} nodes, {st.launches} launches", flush=True)
    img = r.redraw(cams[0], precision=rt.RTB_PRECISION_FP64, segs=6)
    img32 = r.redraw(cams[0], kernel=rt.RTB_KERNEL_PIXEL, segs=6)
print("sanitize case ok")
