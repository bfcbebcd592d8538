"""Throughput of a frame sequence with 1 vs 2 frames in flight (rtb_set_frames_in_flight), for the whole
frame (world 1) and for ONE rank's stripes of an N-GPU frame (no collective: what each GPU executes).
usage: inflight_probe.py [world ...]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rt = importlib.import_module("html5-canvas-raytracer_b200")
objs = rt.scenes.synth_objects(1024, 1024)
cam = rt.scenes.demo_camera(3840, 2160)
N = 40
with rt.Renderer(objs) as r:
    dev = [torch.empty((2160, 3840, 4), dtype=torch.uint8, device="cuda") for _ in range(8)]
    streams = [torch.cuda.Stream() for _ in range(8)]
    for world in [int(a) for a in sys.argv[1:]] or [1, 8]:
        o = rt.make_opts(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT, stripe_rows=4, stripe_rank=0,
                         stripe_world=world, out_packed=1)
        for depth in (1, 2, 3, 4, 6, 8):
            r.set_frames_in_flight(depth)
            for rep in range(2):                          # first repetition warms up
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(N):
                    s = streams[i % depth]
                    r.redraw_device(cam, o, dev[i % depth].data_ptr(), s.cuda_stream)
                host = (time.perf_counter() - t0) / N * 1e3
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / N * 1e3
            ok = r.finish()
            print(f"world {world} frames in flight {depth}: {dt:.3f} ms/frame, host enqueue {host:.3f} ms/frame (ok={ok})")
    r.set_frames_in_flight(1)
