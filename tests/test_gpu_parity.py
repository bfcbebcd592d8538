"""GPU parity: every CUDA path, called through the C ABI (ctypes -> csrc/librtb200.so), against the
CPU float64 oracle on the same seeded scenes.  Gates are the north_star's:
  fp64 validation kernel : +-1 LSB per channel on >= 99.99 % of pixels, PSNR >= 50 dB
  fp32 kernels           : +-2 LSB on >= 99.9 % of pixels, PSNR >= 40 dB
plus size-independent properties (row ranges, stripes, bands and frame batches must not change a byte).
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import FP32_PSNR_DB, FP32_WITHIN2, FP64_PSNR_DB, FP64_WITHIN1, lsb_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def demo(rt):
    objs = rt.scenes.demo_objects()
    r = rt.Renderer(objs)
    yield r
    r.close()


@pytest.fixture(scope="module")
def synth(rt):
    objs = rt.scenes.synth_objects(n_total=160, seed=7)
    r = rt.Renderer(objs)
    yield r
    r.close()


def _oracle(oracle, rt, renderer, cam, **kw):
    img, cnt = oracle.render(renderer.desc, cam, rt.make_opts(**kw))
    return img, cnt


@pytest.mark.parametrize("size", [(320, 180), (641, 359)])
def test_fp64_pixel_kernel_demo(rt, oracle, demo, size):
    cam = rt.scenes.demo_camera(*size)
    ref, cnt = _oracle(oracle, rt, demo, cam, segs=8)
    img, st = demo.redraw(cam, stats=True, segs=8, precision=rt.RTB_PRECISION_FP64)
    s = lsb_stats(img, ref)
    print("fp64 demo", size, s)
    assert s["alpha_ok"]
    assert s["within1"] >= FP64_WITHIN1 and s["psnr"] >= FP64_PSNR_DB
    # same arithmetic for every hit/miss decision => identical ray trees
    assert st.nodes == cnt["nodes"] and st.shadow_rays == cnt["shadow_rays"]
    assert st.sphere_tests == cnt["sphere_tests"]


def test_fp64_pixel_kernel_segs5(rt, oracle, demo):
    cam = rt.scenes.demo_camera(480, 270)
    ref, cnt = _oracle(oracle, rt, demo, cam, segs=5)
    img, st = demo.redraw(cam, stats=True, segs=5, precision=rt.RTB_PRECISION_FP64)
    s = lsb_stats(img, ref)
    assert s["within1"] >= FP64_WITHIN1 and s["psnr"] >= FP64_PSNR_DB
    assert st.nodes == cnt["nodes"]


def test_fp64_synth(rt, oracle, synth):
    cam = rt.scenes.demo_camera(384, 216)
    ref, cnt = _oracle(oracle, rt, synth, cam, segs=8)
    img, st = synth.redraw(cam, stats=True, segs=8, precision=rt.RTB_PRECISION_FP64)
    s = lsb_stats(img, ref)
    print("fp64 synth", s)
    assert s["within1"] >= FP64_WITHIN1 and s["psnr"] >= FP64_PSNR_DB
    assert st.nodes == cnt["nodes"] and st.shadow_rays == cnt["shadow_rays"]


@pytest.mark.parametrize("kernel", ["pixel", "wavefront"])
@pytest.mark.parametrize("size", [(320, 180), (640, 360)])
def test_fp32_demo(rt, oracle, demo, kernel, size):
    cam = rt.scenes.demo_camera(*size)
    ref, cnt = _oracle(oracle, rt, demo, cam, segs=8)
    k = rt.RTB_KERNEL_PIXEL if kernel == "pixel" else rt.RTB_KERNEL_WAVEFRONT
    img, st = demo.redraw(cam, stats=True, segs=8, precision=rt.RTB_PRECISION_FP32, kernel=k)
    s = lsb_stats(img, ref)
    print("fp32 demo", kernel, size, s, "nodes", st.nodes, cnt["nodes"])
    assert s["alpha_ok"]
    assert s["within2"] >= FP32_WITHIN2 and s["psnr"] >= FP32_PSNR_DB
    assert abs(int(st.nodes) - cnt["nodes"]) <= 0.002 * cnt["nodes"]


@pytest.mark.parametrize("kernel", ["pixel", "wavefront"])
def test_fp32_synth(rt, oracle, synth, kernel):
    cam = rt.scenes.demo_camera(384, 216)
    ref, cnt = _oracle(oracle, rt, synth, cam, segs=8)
    k = rt.RTB_KERNEL_PIXEL if kernel == "pixel" else rt.RTB_KERNEL_WAVEFRONT
    img, st = synth.redraw(cam, stats=True, segs=8, kernel=k)
    s = lsb_stats(img, ref)
    print("fp32 synth", kernel, s, "nodes", st.nodes, cnt["nodes"], "shadow", st.shadow_rays, cnt["shadow_rays"])
    assert s["within2"] >= FP32_WITHIN2 and s["psnr"] >= FP32_PSNR_DB


def test_wavefront_segs5_and_spp(rt, oracle, demo):
    cam = rt.scenes.demo_camera(200, 120)
    ref, _ = _oracle(oracle, rt, demo, cam, segs=5, spp_grid=4)
    img = demo.redraw(cam, segs=5, spp_grid=4, kernel=rt.RTB_KERNEL_WAVEFRONT)
    s = lsb_stats(img, ref)
    print("wavefront 16spp", s)
    assert s["within2"] >= FP32_WITHIN2
    img64 = demo.redraw(cam, segs=5, spp_grid=4, precision=rt.RTB_PRECISION_FP64)
    s64 = lsb_stats(img64, ref)
    assert s64["within1"] >= FP64_WITHIN1


def test_camera_basis_mode(rt, oracle, demo):
    cams = rt.scenes.orbit_cameras(240, 135, nframes=8)
    # frame 0 of the orbit equals the reference camera: both modes give the same bytes
    ref0, _ = _oracle(oracle, rt, demo, rt.scenes.demo_camera(240, 135), segs=8)
    img0 = demo.redraw(cams[0], segs=8, precision=rt.RTB_PRECISION_FP64)
    assert lsb_stats(img0, ref0)["within1"] >= FP64_WITHIN1
    ref3, _ = _oracle(oracle, rt, demo, cams[3], segs=8)
    for prec, gate, key in ((rt.RTB_PRECISION_FP64, FP64_WITHIN1, "within1"), (rt.RTB_PRECISION_FP32, FP32_WITHIN2, "within2")):
        img3 = demo.redraw(cams[3], segs=8, precision=prec)
        assert lsb_stats(img3, ref3)[key] >= gate


def test_row_range_and_stripes_do_not_change_bytes(rt, demo):
    cam = rt.scenes.demo_camera(256, 150)
    for prec, kern in ((rt.RTB_PRECISION_FP64, rt.RTB_KERNEL_PIXEL), (rt.RTB_PRECISION_FP32, rt.RTB_KERNEL_WAVEFRONT)):
        full = demo.redraw(cam, precision=prec, kernel=kern)
        # progressive hand-back: rows [40,90) only; the rest of the caller's buffer is untouched
        part = np.full_like(full, 7)
        demo.redraw(cam, out=part, precision=prec, kernel=kern, y0=40, y1=90)
        assert np.array_equal(part[40:90], full[40:90])
        assert (part[:40] == 7).all() and (part[90:] == 7).all()
        # interleaved stripes over 3 "ranks" reassemble the identical frame (ragged: 150 = 9*16+6)
        asm = np.zeros_like(full)
        for r in range(3):
            demo.redraw(cam, out=asm, precision=prec, kernel=kern, stripe_rows=16, stripe_rank=r, stripe_world=3)
        assert np.array_equal(asm, full)
        # packed layout = the owned rows in ascending y
        lib = rt._abi.load()
        for r in range(3):
            nloc = lib.rtb_stripe_local_rows(150, 16, r, 3)
            buf = np.zeros((150, 256, 4), np.uint8)
            demo.redraw(cam, out=buf, precision=prec, kernel=kern, stripe_rows=16, stripe_rank=r,
                        stripe_world=3, out_packed=1)
            ys = [y for y in range(150) if (y // 16) % 3 == r]
            assert nloc == len(ys)
            assert np.array_equal(buf[:nloc], full[ys])


def test_wavefront_bands_and_overflow_retry(rt, demo):
    cam = rt.scenes.demo_camera(320, 200)
    full, st1 = demo.redraw(cam, stats=True, kernel=rt.RTB_KERNEL_WAVEFRONT)
    r2 = rt.Renderer(rt.scenes.demo_objects())
    os.environ["RTB_WF_CAP_NODES"] = str(320 * 12)   # tiny queues: forces many bands and retries
    try:
        banded, st2 = r2.redraw(cam, stats=True, kernel=rt.RTB_KERNEL_WAVEFRONT)
    finally:
        del os.environ["RTB_WF_CAP_NODES"]
        r2.close()
    assert st2.bands > st1.bands
    assert np.array_equal(banded, full)
    assert st2.nodes == st1.nodes and st2.shadow_rays == st1.shadow_rays


def test_batch_equals_single_frames(rt, demo):
    cams = rt.scenes.orbit_cameras(160, 90, nframes=5)
    batch = demo.redraw_batch(cams, segs=8)
    for f, c in enumerate(cams):
        assert np.array_equal(batch[f], demo.redraw(c, segs=8))


def test_batch_keeps_frames_in_flight_and_survives_overflow(rt):
    """rtb_render_batch (wavefront path): four frames in flight, D2H under the next frames; with queues
    forced tiny it restarts at the overflowing frame and still returns every frame of the batch exactly."""
    objs = rt.scenes.synth_objects(1024, 1024)
    cams = rt.scenes.orbit_cameras(160, 90, 9)
    with rt.Renderer(objs) as r0:
        refs = [r0.redraw(c, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT) for c in cams]
        got = r0.redraw_batch(cams, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
        assert all(np.array_equal(got[i], refs[i]) for i in range(9))
        got, st = r0.redraw_batch(cams, stats=True, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)   # instrumented path
        assert all(np.array_equal(got[i], refs[i]) for i in range(9)) and st.nodes > 0
    os.environ["RTB_WF_CAP_NODES"] = str(160 * 24)
    try:
        with rt.Renderer(objs) as r:
            got = r.redraw_batch(cams, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
            assert all(np.array_equal(got[i], refs[i]) for i in range(9))
    finally:
        del os.environ["RTB_WF_CAP_NODES"]


def test_c5_orbit_frames_with_per_frame_sort(rt, oracle):
    """Config C5: each orbit frame rendered from its own re-sorted scene (scenes.orbit_frames), basis
    camera; the fp64 kernel stays bit-exact against the oracle, FrameBatchRenderer takes the per-frame
    scenes."""
    torch = pytest.importorskip("torch")
    mg = __import__("importlib").import_module("html5-canvas-raytracer_b200.multigpu")
    frames = list(rt.scenes.orbit_frames(96, 54, nframes=6))[:4]
    refs = []
    for objs, cam in frames:
        ref, cnt = oracle.render(rt.SceneDesc(objs), cam, rt.make_opts(segs=6))
        with rt.Renderer(objs) as r:
            img, st = r.redraw(cam, stats=True, segs=6, precision=rt.RTB_PRECISION_FP64)
        assert np.array_equal(img, ref) and st.nodes == cnt["nodes"]
        refs.append(ref)
    with rt.Renderer(frames[0][0]) as r0:
        fb = mg.FrameBatchRenderer(r0, 96, 54, 0, 1)
        for f in range(4):
            out = fb.redraw_step([c for _, c in frames], f, objects_for_frame=lambda i: frames[i][0],
                                 segs=6, precision=rt.RTB_PRECISION_FP64)
            torch.cuda.synchronize()
            assert np.array_equal(out[0].cpu().numpy(), refs[f])


def test_unstripe_device_roundtrip(rt, demo):
    torch = pytest.importorskip("torch")
    cam = rt.scenes.demo_camera(128, 70)
    full = demo.redraw(cam)
    lib = rt._abi.load()
    S, world = 8, 4
    padded = lib.rtb_stripe_padded_rows(70, S, world)
    gathered = torch.zeros((world, padded, 128, 4), dtype=torch.uint8, device="cuda")
    frame = torch.zeros((70, 128, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()           # torch's fills run on torch's stream, the library has its own
    for r in range(world):
        o = rt.make_opts(stripe_rows=S, stripe_rank=r, stripe_world=world, out_packed=1)
        demo.redraw_device(cam, o, gathered[r].data_ptr())
    rt._abi.check(lib.rtb_unstripe_device(gathered.data_ptr(), frame.data_ptr(), 128, 70, S, world, 0, None))
    torch.cuda.synchronize()           # device-wide: drains the library stream too
    assert np.array_equal(frame.cpu().numpy(), full)


@pytest.mark.parametrize("S", [4, 1])
def test_peer_scatter_assembles_every_ranks_frame(rt, demo, S):
    """rtb_peer_* on ONE device: three 'ranks' (all buffers local) render their packed stripes, each
    scatters them into every rank's framebuffer, sequence numbers and status words arrive in every
    control block, and rtb_peer_wait / rtb_peer_signal / the timeout path behave.  (Across real GPUs and
    processes: tools/mgpu_check.py under torchrun.)"""
    torch = pytest.importorskip("torch")
    import ctypes as C
    A = rt._abi
    lib = A.load()
    W, H, world = 128, 70, 3
    cam = rt.scenes.demo_camera(W, H)
    full = demo.redraw(cam)
    mg = __import__("importlib").import_module("html5-canvas-raytracer_b200.multigpu")
    frames, ctls = [], []
    for _ in range(world):
        for nbytes, lst in ((H * W * 4, frames), (256 * 4, ctls)):
            p, h = C.c_void_p(), (C.c_uint8 * 64)()
            A.check(lib.rtb_peer_alloc(nbytes, C.byref(p), h))
            assert any(h)                                   # an IPC handle was produced
            lst.append(p.value)
    VP = C.c_void_p * world
    pf, pc = VP(*frames), VP(*ctls)
    padded = lib.rtb_stripe_padded_rows(H, S, world)
    local = torch.zeros((world, padded + 1, W, 4), dtype=torch.uint8, device="cuda")
    local[1, padded, 0, 0] = 7                              # rank 1 reports a non-zero status word
    torch.cuda.synchronize()
    READY, STATUS, ACK, TMO = 0, 16, 32, 48
    try:
        for r in range(world):
            o = rt.make_opts(stripe_rows=S, stripe_rank=r, stripe_world=world, out_packed=1)
            demo.redraw_device(cam, o, local[r].data_ptr())
            A.check(lib.rtb_peer_scatter(C.c_void_p(local[r].data_ptr()), pf, pc, W, H, S, r, world,
                                         C.c_void_p(local[r, padded].data_ptr()), READY, STATUS, 5, None))
        for r in range(world):                              # every rank waits for all three sequence numbers
            A.check(lib.rtb_peer_wait(C.c_void_p(ctls[r] + 4 * READY), world, 5, 2000, C.c_void_p(ctls[r] + 4 * TMO), None))
            A.check(lib.rtb_peer_signal(pc, world, r, ACK, 5, None))
        torch.cuda.synchronize()
        for r in range(world):
            fr = torch.as_tensor(mg._DeviceBytes(frames[r], (H, W, 4)), device="cuda").cpu().numpy()
            assert np.array_equal(fr, full), r
            ctl = torch.as_tensor(mg._DeviceBytes(ctls[r], (256,), "<i4"), device="cuda").cpu().numpy()
            assert ctl[READY:READY + world].tolist() == [5, 5, 5]
            assert ctl[STATUS:STATUS + world].tolist() == [0, 7, 0]
            assert ctl[ACK:ACK + world].tolist() == [5, 5, 5]
            assert ctl[TMO] == 0
        # timeout 0 = the stream itself waits (cuStreamWaitValue32), no kernel: already satisfied here
        A.check(lib.rtb_peer_wait(C.c_void_p(ctls[2] + 4 * READY), world, 5, 0, C.c_void_p(ctls[2] + 4 * TMO), None))
        torch.cuda.synchronize()
        # a sequence number that never arrives: the wait gives up and raises the flag instead of hanging
        A.check(lib.rtb_peer_wait(C.c_void_p(ctls[0] + 4 * READY), world, 6, 50, C.c_void_p(ctls[0] + 4 * TMO), None))
        torch.cuda.synchronize()
        ctl = torch.as_tensor(mg._DeviceBytes(ctls[0], (256,), "<i4"), device="cuda").cpu().numpy()
        assert ctl[TMO] == 1
        # argument checks
        assert lib.rtb_peer_scatter(C.c_void_p(local[0].data_ptr()), pf, pc, W + 1, H, S, 0, world, None, 0, 16, 1, None) == A.RTB_ERR_UNSUPPORTED
        assert lib.rtb_peer_signal(pc, 17, 0, 0, 1, None) == A.RTB_ERR_INVALID
        assert lib.rtb_peer_wait(None, 1, 1, 1, None, None) == A.RTB_ERR_INVALID
    finally:
        for p in frames + ctls:
            lib.rtb_peer_free(C.c_void_p(p))


def test_star_hash_mode_matches_oracle(rt, oracle, demo):
    cam = rt.scenes.demo_camera(256, 144)
    kw = dict(segs=8, star_mode=rt.RTB_STARS_HASH, star_seed=12345)
    ref, cnt = _oracle(oracle, rt, demo, cam, **kw)
    assert (ref[..., :3].max(axis=2) > 0)[:40].any()     # some stars in the sky rows
    img64 = demo.redraw(cam, precision=rt.RTB_PRECISION_FP64, **kw)
    assert lsb_stats(img64, ref)["within1"] >= FP64_WITHIN1
    img32 = demo.redraw(cam, kernel=rt.RTB_KERNEL_WAVEFRONT, **kw)
    assert lsb_stats(img32, ref)["within2"] >= FP32_WITHIN2


def test_error_paths(rt, demo):
    cam = rt.scenes.demo_camera(64, 36)
    with pytest.raises(rt.RtbError) as e:
        demo.redraw(cam, segs=99)
    assert e.value.code == rt._abi.RTB_ERR_INVALID and "segs" in str(e.value)
    with pytest.raises(rt.RtbError) as e:
        demo.redraw(cam, precision=rt.RTB_PRECISION_FP64, kernel=rt.RTB_KERNEL_WAVEFRONT)
    assert e.value.code == rt._abi.RTB_ERR_UNSUPPORTED
    bad = rt.scenes.demo_objects()
    desc = rt.SceneDesc(bad)
    desc.c.abi_version = 99
    h = C.c_void_p()
    lib = rt._abi.load()
    assert lib.rtb_scene_create(C.byref(desc.c), C.byref(h)) == rt._abi.RTB_ERR_INVALID
    assert b"abi_version" in lib.rtb_last_error()


def test_empty_scene_is_red(rt, oracle):
    # main.js:231 — no object hit: [1,0,0]
    r = rt.Renderer([])
    cam = rt.scenes.demo_camera(32, 16)
    for prec in (rt.RTB_PRECISION_FP64, rt.RTB_PRECISION_FP32):
        img = r.redraw(cam, precision=prec)
        assert (img.reshape(-1, 4) == np.array([255, 0, 0, 255], np.uint8)).all()
    r.close()


def test_c3_full_size_sampled_rows(rt, oracle):
    """BASELINE config 3 at its full size (3840x2160, 1024 spheres, segs 8): the GPU renders the whole
    frame, the oracle every 54th row (40 rows) of it — the +-LSB gates must hold there too."""
    objs = rt.scenes.synth_objects(1024, 1024)
    cam = rt.scenes.demo_camera(3840, 2160)
    step = 54
    with rt.Renderer(objs) as r:
        ref, cnt = oracle.render(r.desc, cam, rt.make_opts(segs=8), row_step=step)
        rows = list(range(0, 2160, step))
        img32, st = r.redraw(cam, stats=True, segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)
        s32 = lsb_stats(img32[rows], ref[rows])
        print("C3 4K fp32 wavefront on", len(rows), "rows:", s32, "kernel ms", st.kernel_ms, "bands", st.bands)
        assert s32["within2"] >= FP32_WITHIN2 and s32["psnr"] >= FP32_PSNR_DB
        img64 = np.zeros_like(img32)
        for y in rows:   # fp64 validation kernel on exactly those rows (progressive row hand-back)
            r.redraw(cam, out=img64, segs=8, precision=rt.RTB_PRECISION_FP64, y0=y, y1=y + 1)
        s64 = lsb_stats(img64[rows], ref[rows])
        print("C3 4K fp64 on", len(rows), "rows:", s64)
        assert s64["within1"] >= FP64_WITHIN1 and s64["psnr"] >= FP64_PSNR_DB


def test_async_device_render_reports_overflow_through_finish(rt, demo):
    """rtb_render_device without stats only enqueues; a ray-queue overflow surfaces in rtb_finish(), the
    library shrinks its bands, and repeating the same call converges to the identical frame."""
    torch = pytest.importorskip("torch")
    cam = rt.scenes.demo_camera(320, 200)
    full = demo.redraw(cam, kernel=rt.RTB_KERNEL_WAVEFRONT)
    dev = torch.zeros((200, 320, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    o = rt.make_opts(kernel=rt.RTB_KERNEL_WAVEFRONT)
    r2 = rt.Renderer(rt.scenes.demo_objects())
    os.environ["RTB_WF_CAP_NODES"] = str(320 * 12)
    try:
        tries = 0
        while True:
            r2.redraw_device(cam, o, dev.data_ptr())
            tries += 1
            if r2.finish():
                break
            assert tries < 24
        assert tries > 1                                   # the first, optimistic band size did overflow
        torch.cuda.synchronize()
        assert np.array_equal(dev.cpu().numpy(), full)
        # an uncollected verdict is not lost: the next call on the scene reports it
        r3 = rt.Renderer(rt.scenes.demo_objects())
        r3.redraw_device(cam, o, dev.data_ptr())
        with pytest.raises(rt.RtbError) as e:
            r3.redraw_device(cam, o, dev.data_ptr())
        assert e.value.code == rt._abi.RTB_ERR_OVERFLOW
        r3.close()
    finally:
        del os.environ["RTB_WF_CAP_NODES"]
        r2.close()


def test_striped_renderer_single_rank_sequence(rt, synth):
    """StripedRenderer at world = 1 (no collective): single frames and the sequence mode return the bytes
    of a plain redraw; the multi-rank logic is covered by tests/test_multigpu_gloo.py and by the 2-GPU
    bench."""
    torch = pytest.importorskip("torch")
    mg = __import__("importlib").import_module("html5-canvas-raytracer_b200.multigpu")
    cam = rt.scenes.demo_camera(192, 108)
    ref = synth.redraw(cam, segs=6)
    sr = mg.StripedRenderer(synth, 192, 108, 0, 1)
    f = sr.redraw(cam, segs=6)
    torch.cuda.synchronize()
    assert np.array_equal(f.cpu().numpy(), ref)
    seen = []
    rep = sr.redraw_sequence([cam, cam, cam], on_frame=lambda i, fr: seen.append((i, fr.cpu().numpy().copy())), segs=6)
    assert rep == 0 and [i for i, _ in seen] == [0, 1, 2]
    assert all(np.array_equal(a, ref) for _, a in seen)
    # a moving camera with several frames in flight (one workspace, primary-table set and stream each):
    # every frame must be the frame a plain redraw of that camera gives, with one frame in flight too
    cams = rt.scenes.orbit_cameras(192, 108, 7)
    refs = [synth.redraw(c, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT) for c in cams]
    assert any(not np.array_equal(refs[0], x) for x in refs[1:])
    for depth in (4, 3, 2, 1):
        seen = []
        rep = sr.redraw_sequence(cams, on_frame=lambda i, fr: seen.append((i, fr.cpu().numpy().copy())),
                                 frames_in_flight=depth, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
        assert rep == 0 and [i for i, _ in seen] == list(range(7))
        assert all(np.array_equal(a, refs[i]) for i, a in seen), depth


def test_two_frames_in_flight_overflow_restarts_the_pipeline(rt):
    """Forced queue overflow while two frames are in flight: the sequence repeats the frame (smaller
    bands) and still delivers every frame byte-identical to an unconstrained render."""
    torch = pytest.importorskip("torch")
    import os
    mg = __import__("importlib").import_module("html5-canvas-raytracer_b200.multigpu")
    objs = rt.scenes.synth_objects(1024, 1024)
    cams = rt.scenes.orbit_cameras(160, 90, 4)
    with rt.Renderer(objs) as r0:
        refs = [r0.redraw(c, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT) for c in cams]
    os.environ["RTB_WF_CAP_NODES"] = str(160 * 24)
    try:
        with rt.Renderer(objs) as r:
            sr = mg.StripedRenderer(r, 160, 90, 0, 1)
            seen = {}
            rep = sr.redraw_sequence(cams, on_frame=lambda i, fr: seen.__setitem__(i, fr.cpu().numpy().copy()),
                                     segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
            assert rep >= 1
            assert sorted(seen) == [0, 1, 2, 3]
            assert all(np.array_equal(seen[i], refs[i]) for i in range(4))
    finally:
        del os.environ["RTB_WF_CAP_NODES"]


def test_star_stream_replay_matches_unmodified_main_js(rt, oracle, demo):
    """RTB_STARS_STREAM: the fp64 kernel consumes a supplied Math.random() stream in the reference's own
    call order (count pass -> raster-order scan -> render pass) and reproduces the framebuffer of the
    UNMODIFIED main.js run with that stream (tests/golden/mainjs_demo_stars_96x54.npz) byte for byte."""
    from oracle import gen_golden
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mainjs_demo_stars_96x54.npz")
    z = np.load(golden)
    gold, n = z["rgba"], int(z["n_draws"])
    stream = gen_golden.star_stream(n)
    cam = rt.scenes.demo_camera(96, 54)
    img64 = demo.redraw(cam, segs=8, precision=rt.RTB_PRECISION_FP64, star_mode=rt.RTB_STARS_STREAM, star_stream=stream)
    assert np.array_equal(img64, gold)
    # fp32 cannot replay a stream: one differing hit/miss shifts every later index -> refused
    with pytest.raises(rt.RtbError) as e:
        demo.redraw(cam, segs=8, precision=rt.RTB_PRECISION_FP32, star_mode=rt.RTB_STARS_STREAM, star_stream=stream)
    assert e.value.code == rt._abi.RTB_ERR_UNSUPPORTED
    # larger frame against the oracle fed with the same stream
    cam2 = rt.scenes.demo_camera(320, 180)
    big = gen_golden.star_stream(200000)
    ref, cnt = oracle.render(demo.desc, cam2, rt.make_opts(segs=8), threads=1, rand_stream=big)
    img = demo.redraw(cam2, segs=8, precision=rt.RTB_PRECISION_FP64, star_mode=rt.RTB_STARS_STREAM, star_stream=big)
    assert np.array_equal(img, ref)
    # a stream that is too short is refused, and so are partial frames / the wavefront path
    with pytest.raises(rt.RtbError) as e:
        demo.redraw(cam2, segs=8, precision=rt.RTB_PRECISION_FP64, star_mode=rt.RTB_STARS_STREAM, star_stream=big[:100])
    assert e.value.code == rt._abi.RTB_ERR_INVALID and "draws" in str(e.value)
    with pytest.raises(rt.RtbError) as e:
        demo.redraw(cam2, precision=rt.RTB_PRECISION_FP64, star_mode=rt.RTB_STARS_STREAM, star_stream=big, y0=0, y1=10)
    assert e.value.code == rt._abi.RTB_ERR_UNSUPPORTED


def test_gpu_kernels_against_the_unmodified_main_js_frames(rt, demo):
    """The committed framebuffers of the UNMODIFIED main.js (tests/golden/mainjs_demo_*.npz, oracle/gen_golden.py)
    are the reference itself: the fp64 validation kernel and the fp32 kernels are held to the north star's gates
    against THEM directly, not only against the C restatement (which reproduces them bit-exactly)."""
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    files = sorted(f for f in os.listdir(golden) if f.startswith("mainjs_demo_") and f.endswith(".npz") and "stars" not in f)
    assert len(files) >= 4
    for f in files:
        gold = np.load(os.path.join(golden, f))["rgba"]
        h, w = gold.shape[:2]
        cam = rt.scenes.demo_camera(w, h)
        s64 = lsb_stats(demo.redraw(cam, segs=8, precision=rt.RTB_PRECISION_FP64), gold)
        assert s64["within1"] >= FP64_WITHIN1 and s64["psnr"] >= FP64_PSNR_DB and s64["alpha_ok"], (f, s64)
        if w * h >= 160 * 90:       # the fp32 gate is a fraction of pixels: meaningful from ~10^4 pixels on
            for kernel in (rt.RTB_KERNEL_WAVEFRONT, rt.RTB_KERNEL_PIXEL):
                s32 = lsb_stats(demo.redraw(cam, segs=8, kernel=kernel), gold)
                assert s32["within2"] >= FP32_WITHIN2 and s32["psnr"] >= FP32_PSNR_DB, (f, kernel, s32)
        print(f, "fp64 exact", s64["exact"])
