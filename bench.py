#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

  metric   : Mrays/s (and frames/s) on the 4K / 1024-sphere synthetic scene (configs[2] — the
             configuration the metric is quoted on), fp32 wavefront kernels, segs = 8.
  step     : one full frame of the hot path.
  rays     : tree nodes + shadow rays of the REFERENCE control flow, counted once by the instrumented
             CPU oracle for this exact scene/resolution (tests/golden/counters_*.json); the GPU is
             credited with the oracle's count, never its own.
  value    : whole-job Mrays/s with the scene resident in HBM and the framebuffer left on the device.
  e2e      : same metric through the host API, host buffers on both sides, all copies inside the timed
             region: the scene description handed over from host memory (rtb_scene_create: H2D), a sequence
             of frames rendered by ONE rtb_render_batch call (N > 1: redraw_sequence), every frame copied
             into pinned host memory (D2H).  e2e.single_frame = the same for one frame per call.
  roofline : FP32 CUDA-core compute (not HBM, not tensor): algorithmic flops = 15 x T, T = the
             reference's intersectSphere evaluations for the frame (oracle counter).
  N > 1    : torchrun, one rank per GPU; interleaved single-row stripes; exchange by the library's scatter
             kernel over NVLink peer memory (--exchange nccl: one NCCL all-gather per frame).
             Fixed total work => "scaling": "strong".

`--impl reference` times the CPU float64 restatement of main.js (oracle/, all host threads) on a
bounded row sample of the same frame — the reference itself is browser JavaScript and no JS engine
exists in this image.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG = "c3_synth1024_4k"
FLOPS_PER_TEST = 15            # miss path of intersectSphere, main.js:422-425
CPU_SAMPLE_ROWS = 36           # evenly spaced rows


def load_counters(name):
    with open(os.path.join(ROOT, "tests", "golden", f"counters_{name}.json")) as f:
        return json.load(f)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clocks + throttle reasons of the job's GPUs under load (NVML), polled by ONE thread on rank 0.
    NOT inside the timed region: at N = 8 every NVML query stalls the peer-memory exchange for milliseconds
    (measured, 20 frames: 23.5 ms without polling, 27.1 ms with every rank polling its GPU, 29.0 ms with rank
    0 polling all eight - profiles/r02_notes.md).  The clock the timed region ran at is measured on the
    device instead (rtb_clock_probe before and after it: SM cycles / ns); this thread covers the phases right
    after it - the per-class pass and the end-to-end passes, GPUs busy throughout - plus one sample taken
    immediately before the region."""

    def __init__(self, indices=(0,), period=0.02):
        super().__init__(daemon=True)
        self.indices, self.period = list(indices), period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.in_region = 0
        self.region = False
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h[0], pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            for h in self.h:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    if self.region:
                        self.in_region += 1
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            self._halt.wait(self.period)

    def sample_once(self):
        if not self.ok:
            return
        nv = self.nv
        for h in self.h:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                                  (0x4, "sw_power_cap"), (0x80, "hw_power_brake")):
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_min_mhz": (s[0] if s else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "samples_in_timed_region": self.in_region, "gpus": self.indices,
                "window": "timed region + per-class pass + end-to-end passes (GPUs busy throughout)"}


def build_objects(rt, scene):
    return rt.scenes.demo_objects() if scene == "demo" else rt.scenes.synth_objects(1024, 1024)


def host_threads():
    """Host threads this process may use.  Counted from the affinity mask, not from OMP_NUM_THREADS: torchrun
    exports OMP_NUM_THREADS=1 to every rank, which would silently turn the CPU arm into a 1-thread run."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_config(scene, W, H, segs, g, cnt):
    """`config` of the JSON line: the WORKLOAD only, identical for the GPU arm and the CPU reference arm
    (what each arm runs it with is in `engine`)."""
    return {"workload": CONFIG,
            "scene": "synth1024 (numpy default_rng(1024), reference sort)" if scene != "demo" else "demo (main.js:107-163)",
            "width": W, "height": H, "segs": segs, "spp": g * g, "n_spheres": cnt["n_spheres"],
            "rays_per_frame": cnt["rays"],
            "l2": "no flush between frames: a frame's working set (ray/node queues, ~2.7 GB at 4K) exceeds the 126 MB L2"}


def cpu_sample(rt, desc, cam, segs, g, threads=0, rows=CPU_SAMPLE_ROWS):
    """Oracle (C port of main.js) on `rows` evenly spaced rows of the frame.  Returns (Mrays/s, dict)."""
    from oracle import oracle
    oracle.build()
    if threads <= 0:
        threads = host_threads()
    step = max(1, cam.height // rows)
    t0 = time.perf_counter()
    _, cnt = oracle.render(desc, cam, rt.make_opts(segs=segs, spp_grid=g), row_step=step, threads=threads)
    dt = time.perf_counter() - t0
    nrows = len(range(0, cam.height, step))
    cores = threads
    return cnt["rays"] / dt / 1e6, {
        "seconds": round(dt, 3), "rays": cnt["rays"], "sphere_tests": cnt["sphere_tests"], "cores": cores,
        "sample": f"{nrows} evenly spaced rows (every {step}th) of the {cam.width}x{cam.height} frame, "
                  f"{cnt['rays']} rays; extrapolated {dt * cam.height / nrows:.1f} s/frame"}


def run_reference(args):
    """CPU reference arm: rank 0 only; all host threads; each step = the bounded row sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rt = importlib.import_module("html5-canvas-raytracer_b200")
    scene, W, H, segs, g = rt.scenes.config_workload(CONFIG)
    desc = rt.SceneDesc(build_objects(rt, scene))
    cam = rt.scenes.demo_camera(W, H)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(rt, desc, cam, segs, g, rows=4)
    vals, info = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, info = cpu_sample(rt, desc, cam, segs, g)
        vals.append(v)
    total = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": round(value, 4), "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1000 * total / args.steps, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(scene, W, H, segs, g, load_counters(CONFIG)),
        "engine": {"what": "CPU float64 restatement of main.js (oracle/rt_oracle.c), OpenMP over rows; the reference "
                           "is browser JavaScript and this image has no JS engine",
                   "threads": info["cores"]},
        "cpu_baseline": {"value": round(value, 4), "unit": "Mrays/s", "cores": info["cores"], "kind": "port",
                         "sample": info["sample"]},
        "e2e": {"value": round(value, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def scene_h2d_bytes(desc):
    n = desc.c.n_spheres
    tex = sum(int(t.width) * int(t.height) * 4 for t in desc.textures)
    nl = desc.c.n_lights
    # geometry (f64 + f32), materials (f64 + f32 device structs), texture table + texels, light tables,
    # primary table — what rtb_scene_create + the first render upload
    return n * (32 + 16) + n * (176 + 96) + len(desc.textures) * 16 + tex + (nl + 1) * n * 16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU framebuffer exchange: peer-memory scatter kernel or NCCL all-gather")
    ap.add_argument("--frames-in-flight", type=int, default=0,
                    help="wavefront passes the sequence keeps in flight (0 = the host layer's default)")
    ap.add_argument("--group", type=int, default=0,
                    help="frames per wavefront pass in sequence mode (0 = the host layer's default: multigpu.default_group)")
    ap.add_argument("--config", default=CONFIG,
                    choices=["c3_synth1024_4k", "c2_demo_1080p_segs5", "c4_demo_4k_16spp"],
                    help="BASELINE.json config; the default (configs[2]) is the one the metric is quoted on")
    args = ap.parse_args()
    globals()["CONFIG"] = args.config
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                       # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
        return

    import ctypes as C

    import numpy as np
    import torch

    rt = importlib.import_module("html5-canvas-raytracer_b200")
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["NCCL_DEBUG"] = "WARN"       # keep NCCL's version banner off stdout: ONE JSON line
    rank, local_rank, world = mg.init_distributed()
    torch.cuda.set_device(local_rank)
    dist = torch.distributed if world > 1 else None

    scene, W, H, segs, g = rt.scenes.config_workload(CONFIG)
    cnt = load_counters(CONFIG)
    rays_per_frame, T = cnt["rays"], cnt["sphere_tests"]
    objs = build_objects(rt, scene)
    cam = rt.scenes.demo_camera(W, H)
    renderer = rt.Renderer(objs, device=local_rank)
    fif = args.frames_in_flight or mg.default_in_flight(world)
    sr = mg.StripedRenderer(renderer, W, H, rank, world, exchange=args.exchange, group=args.group,
                            slots=max(fif, 3) if world > 1 else 0)
    G = sr.G                                  # frames per wavefront pass (sequence mode)
    # own exchange kernels per GROUP (signal, 2 waits, scatter | one unstripe per frame)
    xl = 0 if world <= 1 else (4 if sr.exchange == "p2p" else G)
    A = rt._abi
    # the product's own choice (RTB_KERNEL_AUTO): wavefront for the 1024-sphere scene, fused per-pixel
    # kernel for the 14-sphere demo configs
    use_wf = cnt["n_spheres"] > 32
    opts_kw = dict(segs=segs, spp_grid=g, precision=rt.RTB_PRECISION_FP32, kernel=rt.RTB_KERNEL_AUTO)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    for _ in range(args.warmup):
        sr.redraw(cam, **opts_kw)
    # the sequence path too (its streams and one workspace per pass in flight are created on first use)
    sr.redraw_sequence([cam] * max(args.warmup, (fif + 1) * G), frames_in_flight=fif, **opts_kw)
    sr.redraw_sequence([cam] * args.steps, frames_in_flight=fif, **opts_kw)      # and once in the timed shape
    barrier()
    mode = os.environ.get("RTB_BENCH_SAMPLER", "device")     # device (default) | rank0 | all (NVML inside the region) | off
    sampler = None
    if mode == "all" or (mode in ("rank0", "device") and rank == 0):
        sampler = ClockSampler([local_rank] if mode == "all" else list(range(world)))
        if mode == "device":
            sampler.sample_once()                               # one NVML reading just before the region
            barrier()
        else:
            sampler.start()
            sampler.region = True
    elif mode == "device":
        barrier()
    n_sm = rt.device_info()["sm_count"]
    probe = torch.zeros((2, 2 * n_sm), dtype=torch.int64, device="cuda")
    pstream = A.stream_arg(torch.cuda.current_stream().cuda_stream)
    A.check(A.load().rtb_clock_probe(C.c_void_p(probe[0].data_ptr()), n_sm, pstream))
    st_acc = {"p": 0.0, "s": 0.0, "sh": 0.0, "c": 0.0, "k": 0.0, "launches": 0, "nodes": 0, "shadow": 0}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    # K frames as a sequence (StripedRenderer.redraw_sequence): groups of G frames per wavefront pass, `fif`
    # passes in flight, so the render of one group overlaps the exchange of the previous one; every frame is
    # fully assembled on every rank inside the region.
    repeats = sr.redraw_sequence([cam] * args.steps, frames_in_flight=fif, **opts_kw)   # no instrumentation inside
    ev1.record()
    A.check(A.load().rtb_clock_probe(C.c_void_p(probe[1].data_ptr()), n_sm, pstream))
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    if sampler:
        sampler.region = False
        if mode == "device":
            sampler.start()                                     # from here on: per-class pass + end-to-end passes
    # SM clock of the timed region, measured on the device: cycles / ns per SM between the two probes
    pr = probe.cpu().numpy().reshape(2, n_sm, 2)
    okm = (pr[0, :, 1] > 0) & (pr[1, :, 1] > pr[0, :, 1])
    dev_mhz = float(np.median((pr[1, okm, 0] - pr[0, okm, 0]) / (pr[1, okm, 1] - pr[0, okm, 1]) * 1e3)) if okm.any() else 0.0
    tm = torch.tensor([dev_mhz if dev_mhz > 0 else 1e9], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MIN)
    dev_mhz_min = float(tm.item())
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = rays_per_frame / (ms_step * 1e-3) / 1e6

    # per-kernel-class durations (CUDA events inside the library, on the launching stream): measured
    # live, right after the timed region, over the same number of frames
    # (one group of G frames per instrumented pass, the launch shape of the timed region; normalised per frame)
    n_passes = len(mg.split_groups(args.steps, G))
    pass_launches = 0
    for _ in range(n_passes):
        st = A.RtbStats()
        sr.redraw_group([cam] * G, stats=st, **opts_kw)
        st_acc["p"] += st.ms_trace_primary; st_acc["s"] += st.ms_trace_secondary
        st_acc["sh"] += st.ms_shadow; st_acc["c"] += st.ms_compose; st_acc["k"] += st.kernel_ms
        pass_launches = st.launches + xl
        st_acc["launches"] += 2 * pass_launches                                # timed + instrumented pass
        st_acc["nodes"], st_acc["shadow"] = int(st.nodes) // G, int(st.shadow_rays) // G
    inst_frames = n_passes * G                                                 # frames the sums above cover
    barrier()

    # ---- end to end through the host API (rank-local D2H on rank 0; scene handed over per step) ----
    host = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    e2e_launches = 0

    lib = A.load()
    desc = renderer.desc                    # the scene description in HOST memory, marshalled once
    cam_c = cam.to_c()
    e2e_opts = rt.make_opts(**opts_kw)
    h2d_bytes = 0

    def e2e_step():
        """rtb_scene_create (H2D: the whole scene description) -> render -> frame in pinned host memory."""
        nonlocal e2e_launches, h2d_bytes
        h = C.c_void_p()
        A.check(lib.rtb_scene_create(C.byref(desc.c), C.byref(h)))
        h2d_bytes = int(lib.rtb_scene_h2d_bytes(h)) + 2 * 16 * ((cnt["n_spheres"] + 15) // 16 * 16)   # + primary tables
        st = A.RtbStats()
        try:
            if world <= 1:
                A.check(lib.rtb_render(h, C.byref(cam_c), C.byref(e2e_opts), host.data_ptr(), C.byref(st)))   # D2H inside
                e2e_launches += st.launches
            else:
                old = sr.r._h
                sr.r._h = h
                try:
                    frame = sr.redraw(cam, stats=st, **opts_kw)
                finally:
                    sr.r._h = old
                e2e_launches += st.launches + xl
                if rank == 0:
                    host.copy_(frame, non_blocking=True)
                torch.cuda.synchronize()
        finally:
            lib.rtb_scene_destroy(h)

    # (1) one frame at a time: create -> render -> frame on the host -> destroy; the latency a caller of a
    #     single redraw() sees
    e2e_step()
    barrier()
    k2 = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(k2):
        e2e_step()
    barrier()
    single_ms = (time.perf_counter() - t0) * 1000 / k2

    # (2) a frame SEQUENCE through the host API, everything inside the timed region: the scene description
    #     handed over from host memory (rtb_scene_create: H2D), kb frames rendered with several in flight,
    #     every frame copied into pinned host memory (D2H under the kernels of the next frames), destroy.
    #     N = 1: ONE C-ABI call, rtb_render_batch.  N > 1: StripedRenderer.redraw_sequence, rank 0 reads back.
    kb = max(4, min(args.steps, 40))                    # 40 x 33 MB of pinned host memory at 4K
    host_seq = torch.empty((kb, H, W, 4), dtype=torch.uint8).pin_memory() if (rank == 0 and world <= 1) else None
    # N > 1: the frames land in ONE ring of host framebuffers in shared memory that every rank has mapped and
    # page-locked; each rank copies its own rows of every frame over its own PCIe link (SharedHostFrames)
    ring, ring_note = None, ""
    if world > 1:
        try:
            ring = mg.SharedHostFrames(kb, H, W, rank, world)       # collective: all ranks succeed or all raise
        except RuntimeError as e:                                  # fall back: rank 0 reads every assembled frame back
            ring_note = str(e)
            host_seq = torch.empty((kb, H, W, 4), dtype=torch.uint8).pin_memory() if rank == 0 else None
    cams_c = (A.RtbCamera * kb)(*([cam_c] * kb))

    def e2e_sequence():
        nonlocal e2e_launches
        h = C.c_void_p()
        A.check(lib.rtb_scene_create(C.byref(desc.c), C.byref(h)))
        try:
            if world <= 1:
                A.check(lib.rtb_render_batch(h, cams_c, kb, C.byref(e2e_opts), host_seq.data_ptr(), None))
            else:
                old = sr.r._h
                sr.r._h = h
                try:
                    if ring is not None:
                        sr.redraw_sequence([cam] * kb, frames_in_flight=fif, host_frames=ring, **opts_kw)
                    else:
                        def to_host(j, frame):
                            if rank == 0:
                                host_seq[j].copy_(frame, non_blocking=True)    # on the group's own stream
                        sr.redraw_sequence([cam] * kb, on_frame=to_host, frames_in_flight=fif, **opts_kw)
                finally:
                    sr.r._h = old
                torch.cuda.synchronize()
                dist.barrier()                           # every rank's rows of every frame are in host memory
            e2e_launches += pass_launches * len(mg.split_groups(kb, G))
        finally:
            lib.rtb_scene_destroy(h)

    e2e_sequence()
    barrier()
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        e2e_sequence()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000 / (reps * kb)
    te = torch.tensor([e2e_ms, single_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms, single_ms = float(te[0].item()), float(te[1].item())
    e2e_value = rays_per_frame / (e2e_ms * 1e-3) / 1e6
    if ring is not None:
        ring.close()

    clocks = sampler.stop() if sampler else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if mode == "device" and dev_mhz_min < 1e8:
        clocks["nvml_sm_mhz"] = clocks.get("sm_mhz")
        clocks["sm_mhz"] = round(dev_mhz_min, 1)
        clocks["how"] = ("sm_mhz: SM cycles / ns between two device probes bracketing the timed region (median over SMs, "
                         "MIN over ranks); reasons / nvml_sm_mhz: NVML on rank 0, one sample right before the region and "
                         "every 20 ms during the phases right after it (per-class pass, end-to-end passes) - NVML polling "
                         "inside a multi-GPU timed region stalls the peer exchange")
    sr.close()                               # collective: unmap the peers' buffers, free this rank's
    if rank != 0:
        renderer.close()
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---- roofline (FP32 compute) ------------------------------------------------------------------
    peaks, peak_src = measured_peaks()
    info = rt.device_info()
    sm_max_mhz = float(peaks.get("sm_max_mhz") or 1965.0)
    peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    achieved_tflops = FLOPS_PER_TEST * T / (ms_step * 1e-3) / 1e12
    K = inst_frames
    n_sub = W * H * g * g
    T_primary = n_sub * cnt["n_spheres"]
    T_secondary = (cnt["nodes"] - n_sub) * cnt["n_spheres"]
    T_shadow = T - cnt["nodes"] * cnt["n_spheres"]

    def cls(ms, tests):
        ms = ms / K
        return {"ms": round(ms, 4), "tests": int(tests / world),
                "tflops": round(FLOPS_PER_TEST * tests / world / (ms * 1e-3) / 1e12, 3) if ms > 0 else None}

    # kernel classes of the split pipeline (csrc/wavefront.cu): a sweep kernel + a hit / post kernel per level / light
    K_PRIMARY = "level 1: wf_raygen_kernel + wf_sweep_kernel<primary> + wf_hit_kernel<primary>"
    K_SECOND = "levels 2..segs: wf_sweep_kernel<general> + wf_hit_kernel<general>"
    K_LIGHTS = "lights: wf_ssweep_kernel + wf_spost_kernel"
    kernels = {K_PRIMARY: cls(st_acc["p"], T_primary),
               K_SECOND: cls(st_acc["s"], T_secondary),
               K_LIGHTS: cls(st_acc["sh"], T_shadow),
               "wf_compose_kernel + wf_resolve_kernel": {"ms": round(st_acc["c"] / K, 4)}}
    kernel_ms = (st_acc["p"] + st_acc["s"] + st_acc["sh"] + st_acc["c"]) / K
    dom = max((K_PRIMARY, K_SECOND, K_LIGHTS), key=lambda k: kernels[k]["ms"])
    if not use_wf:                      # one fused kernel per frame
        kernels = {"pixel_kernel<float>": cls(st_acc["k"], T)}
        kernel_ms = st_acc["k"] / K
        dom = "pixel_kernel<float>"
    # DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full capture
    traffic, traffic_note = None, "no ncu capture for this workload"
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_traffic.json")
    if use_wf and CONFIG == "c3_synth1024_4k" and os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj["dram_bytes_per_launch"]
        traffic_note = f"{tj['kernel']}, one launch of {tj['launch_us']} us; {tj['source']}"
    # What the DEVICE executes (for orientation, next to the credited reference flops): the shared-origin sweeps
    # (primary, shadow) spend 6 flop per test, the general sweep 16, and every shadow ray sweeps all spheres
    # (the reference stops at the first opaque occluder).  st_acc nodes / shadow rays are per frame, whole frame.
    exec_flops = None
    if use_wf and st_acc["nodes"] * world >= n_sub:
        ns = cnt["n_spheres"]
        exec_flops = (6.0 * n_sub + 16.0 * (st_acc["nodes"] * world - n_sub) + 6.0 * st_acc["shadow"] * world) * ns
    roofline = {
        "bound": "fp32", "achieved": round(achieved_tflops * (1 if world == 1 else 1), 3), "peak": round(peak_tflops * world, 3),
        "unit": "TFLOP/s", "frac": round(achieved_tflops / (peak_tflops * world), 4),
        "traffic": traffic, "traffic_note": traffic_note,
        "peak_source": f"{info['sm_count']} SMs x 128 FP32 lanes x 2 x sm_max_mhz {sm_max_mhz:.0f} ({peak_src})",
        "flops": f"{FLOPS_PER_TEST} x T, T = {T} reference intersectSphere evaluations per frame (oracle)",
        "executed": None if exec_flops is None else {
            "tflops": round(exec_flops / (ms_step * 1e-3) / 1e12, 3),
            "frac": round(exec_flops / (ms_step * 1e-3) / 1e12 / (peak_tflops * world), 4),
            "what": "FMA flops the sweeps execute per frame: 6 per test for primary and shadow rays (shared origin), 16 for "
                    "reflect / refract rays, every shadow ray against all spheres; the credited figure above uses the "
                    "reference's 15 per test, which is why it can exceed 1"},
        "dominant_kernel": dom, "dominant": kernels[dom],
        "dominant_frac": round(kernels[dom]["tflops"] / peak_tflops, 4) if kernels[dom].get("tflops") else None,
        "kernels": kernels, "kernel_ms_per_step": round(kernel_ms, 4),
        "hbm": {"framebuffer_bytes": W * H * 4, "gbs_framebuffer_only": round(W * H * 4 / (ms_step * 1e-3) / 1e9, 2),
                "peak_gbs": peaks.get("hbm_gbs")},
    }

    line = {
        "metric": "Mrays/s", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": round(1000.0 / ms_step, 3),
        "config": workload_config(scene, W, H, segs, g, cnt),
        "engine": {"kernel": "wavefront fp32" if use_wf else "per-pixel fp32 (AUTO, N <= 32)",
                   "partition": "single GPU" if world == 1 else (
                       f"interleaved {sr.S}-row stripes + " +
                       ("peer-memory scatter kernel over NVLink (rtb_peer_scatter_group)" if sr.exchange == "p2p"
                        else "NCCL all-gather + de-interleave kernel") +
                       (f"; {sr.exchange_note}" if sr.exchange_note else "")),
                   "frames_per_pass": G, "passes_in_flight": fif,
                   "groups": mg.split_groups(args.steps, G)},
        "roofline": roofline,
        "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "ms_per_step": round(e2e_ms, 4),
                "frames_per_s": round(1000.0 / e2e_ms, 3),
                "h2d_bytes_per_step": int(round(h2d_bytes / kb)), "d2h_bytes_per_step": W * H * 4,
                "frames_per_call": kb,
                "what": f"a sequence of {kb} frames through the host API, all inside the timed region: rtb_scene_create "
                        "(the whole scene description from host memory, H2D once per sequence: "
                        f"{h2d_bytes} B) + " +
                        ("ONE rtb_render_batch call (every frame D2H into pinned host memory, under the kernels of the "
                         "next frames)" if world == 1 else
                         ("StripedRenderer.redraw_sequence into a ring of host framebuffers in shared, page-locked memory: "
                          "every rank copies its own rows of every frame over its own PCIe link (rtb_stripe_rows_to_host), "
                          "barrier at the end" if not ring_note else
                          f"StripedRenderer.redraw_sequence, rank 0 copies every assembled frame to pinned host memory ({ring_note})")
                         ) + " + rtb_scene_destroy",
                "single_frame": {"ms": round(single_ms, 4), "value": round(rays_per_frame / (single_ms * 1e-3) / 1e6, 2),
                                 "what": "one frame per call: rtb_scene_create + rtb_render into a pinned host buffer "
                                         "+ rtb_scene_destroy (nothing overlaps)"}},
        "gpu_launches": int(st_acc["launches"] + e2e_launches),
        "clocks": clocks,
        "device": info,
        "device_counts": {"nodes": st_acc["nodes"], "shadow_rays_executed": st_acc["shadow"],
                          "oracle_nodes": cnt["nodes"], "oracle_shadow_rays": cnt["shadow_rays"]},
    }
    if not args.no_cpu_baseline and world == 1:
        v, ci = cpu_sample(rt, renderer.desc, cam, segs, g)
        line["cpu_baseline"] = {"value": round(v, 4), "unit": "Mrays/s", "cores": ci["cores"], "kind": "port",
                                "sample": ci["sample"]}
    renderer.close()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
