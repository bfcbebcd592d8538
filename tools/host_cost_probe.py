"""Host-side cost of enqueueing one wavefront frame: a frame so small that the GPU is never the
bottleneck, several frames in flight.  usage: host_cost_probe.py"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rt = importlib.import_module("html5-canvas-raytracer_b200")
objs = rt.scenes.synth_objects(1024, 1024)
for (W, H) in ((64, 36), (640, 360)):
    cam = rt.scenes.demo_camera(W, H)
    N = 200
    with rt.Renderer(objs) as r:
        dev = [torch.empty((H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(8)]
        streams = [torch.cuda.Stream() for _ in range(8)]
        o = rt.make_opts(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)
        for depth in (1, 4):
            r.set_frames_in_flight(depth)
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(N):
                    r.redraw_device(cam, o, dev[i % depth].data_ptr(), streams[i % depth].cuda_stream)
                host = (time.perf_counter() - t0) / N * 1e3
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / N * 1e3
            r.finish()
            # pure host cost: enqueue `depth` frames into empty slots right after a full synchronise
            burst = []
            for rep in range(5):
                torch.cuda.synchronize(); r.finish()
                t0 = time.perf_counter()
                for i in range(depth):
                    r.redraw_device(cam, o, dev[i % depth].data_ptr(), streams[i % depth].cuda_stream)
                burst.append((time.perf_counter() - t0) / depth * 1e3)
            torch.cuda.synchronize(); r.finish()
            print(f"   host cost of one enqueue (non-blocking): {min(burst):.3f} ms")
            print(f"{W}x{H} frames in flight {depth}: {dt:.3f} ms/frame, host loop {host:.3f} ms/frame")
        r.set_frames_in_flight(1)
