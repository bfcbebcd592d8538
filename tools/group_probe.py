"""What ONE rank of an N-GPU row-striped sequence executes (no exchange): frame groups of G frames per wavefront
pass, D passes in flight.  usage: group_probe.py [world] [frames]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rt = importlib.import_module("html5-canvas-raytracer_b200")
mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
objs = rt.scenes.synth_objects(1024, 1024)
W, H = 3840, 2160
cam = rt.scenes.demo_camera(W, H)
rows = (H + world - 1) // world
sampler = None
if os.environ.get("PROBE_SAMPLER"):                      # does the NVML clock sampler of bench.py disturb the enqueue thread?
    import bench
    sampler = bench.ClockSampler(0, period=float(os.environ["PROBE_SAMPLER"]))
    sampler.start()
with rt.Renderer(objs) as r:
    dev = torch.empty((4, 8, rows, W, 4), dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(4)]
    o = rt.make_opts(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT, stripe_rows=1, stripe_rank=0, stripe_world=world, out_packed=1)
    for G, D in ((1, 8), (world, 1), (world, 2), (world, 3), (max(1, world // 2), 2), (max(1, world // 2), 4)):
        if G > 8:
            continue
        sizes = mg.split_groups(N, G)
        r.set_frames_in_flight(min(D, 8))
        for rep in range(3):                              # first repetitions warm up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i, n in enumerate(sizes):
                s = i % D
                with torch.cuda.stream(streams[s % 4]):
                    r.redraw_device_group([cam] * n, o, [dev[s % 4, b].data_ptr() for b in range(n)],
                                          streams[s % 4].cuda_stream)
            host = (time.perf_counter() - t0) / N * 1e3
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / N * 1e3
        ok = r.finish()
        print(f"world {world}: {N} frames, groups of {G} ({len(sizes)} passes), {D} in flight: {dt:.3f} ms/frame "
              f"(ideal {8.9 / world:.3f}), host enqueue {host:.3f} ms/frame ok={ok}", flush=True)
    r.set_frames_in_flight(1)
if sampler:
    print("sampler:", sampler.stop())
