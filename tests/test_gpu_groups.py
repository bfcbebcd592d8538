"""GPU tests of FRAME GROUPS: rtb_render_device_group renders up to RTB_MAX_GROUP frames of a sequence in one
pass of the wavefront pipeline (the multi-GPU throughput mode: each rank renders 1/world of `world` frames
per pass).  Size-independent property: grouping must not change a single byte of any frame - whatever the
group size, camera path (one table run or one per frame), stripe layout, band split or overflow retry."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth(rt):
    r = rt.Renderer(rt.scenes.synth_objects(160, 7))
    yield r
    r.close()


@pytest.fixture(scope="module")
def demo(rt):
    r = rt.Renderer(rt.scenes.demo_objects())
    yield r
    r.close()


def _group(rt, r, cams, H, W, rows=None, **kw):
    """Render `cams` as ONE group into device buffers; returns the list of host arrays."""
    import torch
    n = len(cams)
    rows = H if rows is None else rows
    dev = torch.full((n, rows, W, 4), 17, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    r.redraw_device_group(cams, rt.make_opts(**kw), [dev[i].data_ptr() for i in range(n)])
    assert r.finish()
    torch.cuda.synchronize()
    return [dev[i].cpu().numpy() for i in range(n)]


@pytest.mark.parametrize("n", [1, 2, 5, 8])
def test_group_of_moving_cameras_equals_single_frames(rt, synth, n):
    """Orbiting camera: every frame has its own origin -> one primary-table run per frame."""
    W, H = 192, 108
    cams = rt.scenes.orbit_cameras(W, H, 8)[:n]
    kw = dict(segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
    refs = [synth.redraw(c, **kw) for c in cams]
    got = _group(rt, synth, cams, H, W, **kw)
    for i in range(n):
        assert np.array_equal(got[i], refs[i]), (n, i)


def test_group_runs_of_shared_origins(rt, synth):
    """Origins A A B B B A: three runs (the third shares A's origin but is not adjacent to the first)."""
    W, H = 160, 90
    oc = rt.scenes.orbit_cameras(W, H, 12)
    a, b = oc[0], oc[5]
    cams = [a, a, b, b, b, a]
    kw = dict(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)
    ra, rb = synth.redraw(a, **kw), synth.redraw(b, **kw)
    assert not np.array_equal(ra, rb)
    got = _group(rt, synth, cams, H, W, **kw)
    for i, c in enumerate(cams):
        assert np.array_equal(got[i], ra if c is a else rb), i


def test_group_fixed_camera_stripes_spp_and_stars(rt, demo):
    """One run; packed 3-rank stripes (ragged), 4 spp, star hash: each frame of the group equals the frame
    a single striped call gives."""
    import torch
    W, H, world, S = 128, 70, 3, 4
    cam = rt.scenes.demo_camera(W, H)
    kw = dict(segs=8, spp_grid=2, kernel=rt.RTB_KERNEL_WAVEFRONT, star_mode=rt.RTB_STARS_HASH, star_seed=99)
    full = demo.redraw(cam, **kw)
    lib = rt._abi.load()
    for rank in range(world):
        rows = lib.rtb_stripe_local_rows(H, S, rank, world)
        ys = [y for y in range(H) if (y // S) % world == rank]
        got = _group(rt, demo, [cam] * 4, H, W, rows=rows, stripe_rows=S, stripe_rank=rank, stripe_world=world,
                     out_packed=1, **kw)
        for g in got:
            assert np.array_equal(g, full[ys]), rank


def test_group_bands_and_overflow_retry(rt, synth):
    """Tiny queues: the group is split into many bands (rows of ALL frames per band), possibly after the first,
    optimistic band size overflowed; the group is byte-identical either way."""
    W, H = 160, 90
    cams = rt.scenes.orbit_cameras(W, H, 5)
    kw = dict(segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
    refs = [synth.redraw(c, **kw) for c in cams]
    import torch
    dev = torch.zeros((5, H, W, 4), dtype=torch.uint8, device="cuda")
    os.environ["RTB_WF_CAP_NODES"] = str(160 * 5 * 8)
    try:
        with rt.Renderer(rt.scenes.synth_objects(160, 7)) as r:
            tries = 0
            while True:
                r.redraw_device_group(cams, rt.make_opts(**kw), [dev[i].data_ptr() for i in range(5)])
                tries += 1
                if r.finish():
                    break
                assert tries < 24
            torch.cuda.synchronize()
            for i in range(5):
                assert np.array_equal(dev[i].cpu().numpy(), refs[i]), i
            # synchronous (stats) path: retries inside the call, counters summed over the group
            st = rt._abi.RtbStats()
            r.redraw_device_group(cams, rt.make_opts(**kw), [dev[i].data_ptr() for i in range(5)], stats=st)
            assert st.bands > 1 and st.nodes >= 5 * W * H
            for i in range(5):
                assert np.array_equal(dev[i].cpu().numpy(), refs[i]), i
    finally:
        del os.environ["RTB_WF_CAP_NODES"]


def test_group_counters_are_the_sum_of_the_frames(rt, synth):
    W, H = 128, 72
    cams = rt.scenes.orbit_cameras(W, H, 3)
    kw = dict(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)
    single = []
    for c in cams:
        _, st = synth.redraw(c, stats=True, **kw)
        single.append(st)
    import torch
    dev = torch.zeros((3, H, W, 4), dtype=torch.uint8, device="cuda")
    st = rt._abi.RtbStats()
    synth.redraw_device_group(cams, rt.make_opts(**kw), [dev[i].data_ptr() for i in range(3)], stats=st)
    assert st.nodes == sum(s.nodes for s in single)
    assert st.shadow_rays == sum(s.shadow_rays for s in single)
    assert st.launches == single[0].launches             # one pass: the launch count of ONE frame


def test_group_pixel_kernel_and_argument_checks(rt, demo):
    """Small scenes (AUTO -> per-pixel kernel) render the frames of a group one after the other."""
    W, H = 96, 54
    cams = rt.scenes.orbit_cameras(W, H, 3)
    for prec in (rt.RTB_PRECISION_FP32, rt.RTB_PRECISION_FP64):
        refs = [demo.redraw(c, precision=prec) for c in cams]
        got = _group(rt, demo, cams, H, W, precision=prec)
        for i in range(3):
            assert np.array_equal(got[i], refs[i])
    import torch
    dev = torch.zeros((9, H, W, 4), dtype=torch.uint8, device="cuda")
    with pytest.raises(rt.RtbError) as e:
        demo.redraw_device_group([cams[0]] * 9, rt.make_opts(), [dev[i].data_ptr() for i in range(9)])
    assert e.value.code == rt._abi.RTB_ERR_INVALID
    other = rt.scenes.demo_camera(W, H + 2)
    with pytest.raises(rt.RtbError) as e:
        demo.redraw_device_group([cams[0], other], rt.make_opts(), [dev[0].data_ptr(), dev[1].data_ptr()])
    assert e.value.code == rt._abi.RTB_ERR_UNSUPPORTED


@pytest.mark.parametrize("group,depth", [(4, 2), (3, 3), (8, 1), (2, 4)])
def test_striped_renderer_sequences_in_groups(rt, synth, group, depth):
    """StripedRenderer (world 1) with frame groups: every frame of a moving-camera sequence is the frame a
    plain redraw gives, in order, for several group sizes / passes in flight."""
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    W, H = 192, 108
    cams = rt.scenes.orbit_cameras(W, H, 11)
    kw = dict(segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
    refs = [synth.redraw(c, **kw) for c in cams]
    sr = mg.StripedRenderer(synth, W, H, 0, 1, group=group, slots=4)
    seen = []
    rep = sr.redraw_sequence(cams, on_frame=lambda i, fr: seen.append((i, fr.cpu().numpy().copy())),
                             frames_in_flight=depth, **kw)
    assert rep == 0 and [i for i, _ in seen] == list(range(11))
    assert all(np.array_equal(a, refs[i]) for i, a in seen)
    g = sr.redraw_group(cams[:min(group, 3)], **kw)
    assert all(np.array_equal(g[i].cpu().numpy(), refs[i]) for i in range(min(group, 3)))


def test_grouped_sequence_overflow_restarts(rt):
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    objs = rt.scenes.synth_objects(1024, 1024)
    cams = rt.scenes.orbit_cameras(160, 90, 7)
    with rt.Renderer(objs) as r0:
        refs = [r0.redraw(c, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT) for c in cams]
    os.environ["RTB_WF_CAP_NODES"] = str(160 * 3 * 24)
    try:
        with rt.Renderer(objs) as r:
            sr = mg.StripedRenderer(r, 160, 90, 0, 1, group=3, slots=3)
            seen = {}
            rep = sr.redraw_sequence(cams, on_frame=lambda i, fr: seen.__setitem__(i, fr.cpu().numpy().copy()),
                                     frames_in_flight=2, segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
            assert rep >= 1 and sorted(seen) == list(range(7))
            assert all(np.array_equal(seen[i], refs[i]) for i in range(7))
    finally:
        del os.environ["RTB_WF_CAP_NODES"]


def test_peer_scatter_group_on_one_device(rt, demo):
    """rtb_peer_scatter_group with three 'ranks' on one device: the packed stripe rows of a 3-frame group
    (three different cameras) land in every rank's framebuffers, one status word / sequence number per group."""
    import torch
    A = rt._abi
    lib = A.load()
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    W, H, world, S, G = 128, 70, 3, 1, 3
    cams = rt.scenes.orbit_cameras(W, H, 3)
    fulls = [demo.redraw(c) for c in cams]
    frame_bytes = H * W * 4
    frames, ctls = [], []
    for _ in range(world):
        for nbytes, lst in ((G * frame_bytes, frames), (256 * 4, ctls)):
            p, h = C.c_void_p(), (C.c_uint8 * 64)()
            A.check(lib.rtb_peer_alloc(nbytes, C.byref(p), h))
            lst.append(p.value)
    VP = C.c_void_p * world
    pf, pc = VP(*frames), VP(*ctls)
    padded = lib.rtb_stripe_padded_rows(H, S, world)
    stride = padded + 1
    local = torch.zeros((world, G, stride, W, 4), dtype=torch.uint8, device="cuda")
    local[2, 0, padded, 0, 0] = 9                            # rank 2 reports a non-zero status word
    torch.cuda.synchronize()
    READY, STATUS = 0, 16
    try:
        for r in range(world):
            o = rt.make_opts(stripe_rows=S, stripe_rank=r, stripe_world=world, out_packed=1)
            demo.redraw_device_group(cams, o, [local[r, b].data_ptr() for b in range(G)])
            A.check(lib.rtb_peer_scatter_group(C.c_void_p(local[r].data_ptr()), stride * W * 4, G, pf, frame_bytes, pc,
                                               W, H, S, r, world, C.c_void_p(local[r, 0, padded].data_ptr()),
                                               READY, STATUS, 3, None))
        torch.cuda.synchronize()
        assert demo.finish()
        for r in range(world):
            fr = torch.as_tensor(mg._DeviceBytes(frames[r], (G, H, W, 4)), device="cuda").cpu().numpy()
            for b in range(G):
                assert np.array_equal(fr[b], fulls[b]), (r, b)
            ctl = torch.as_tensor(mg._DeviceBytes(ctls[r], (256,), "<i4"), device="cuda").cpu().numpy()
            assert ctl[READY:READY + world].tolist() == [3, 3, 3]
            assert ctl[STATUS:STATUS + world].tolist() == [0, 0, 9]
        assert lib.rtb_peer_scatter_group(C.c_void_p(local[0].data_ptr()), stride * W * 4 + 4, G, pf, frame_bytes, pc,
                                          W, H, S, 0, world, None, READY, STATUS, 4, None) == A.RTB_ERR_INVALID
    finally:
        for p in frames + ctls:
            lib.rtb_peer_free(C.c_void_p(p))


@pytest.mark.parametrize("scene", ["demo", "synth"])
def test_split_pipeline_equals_fused_pipeline(rt, demo, synth, scene):
    """The split pipeline (RTB_WF_SPLIT bits: sweep kernels + hit-processing kernels for the primary level /
    the deeper levels, occluder sweeps from ray records + post kernels for the lights) and the fused persistent
    kernels do the same arithmetic: byte-identical frames, identical device counters - whole frames, 4 spp with
    star hash, ragged packed stripes, a moving-camera frame group, and tiny queues (many bands)."""
    import torch
    r = demo if scene == "demo" else synth
    W, H = 256, 144
    cam = rt.scenes.demo_camera(W, H)
    cams = rt.scenes.orbit_cameras(W, H, 5)

    def both(fn):
        out = []
        for split in ("0", "7"):
            os.environ["RTB_WF_SPLIT"] = split
            try:
                out.append(fn())
            finally:
                del os.environ["RTB_WF_SPLIT"]
        return out

    kw = dict(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)
    (a, sa), (b, sb) = both(lambda: r.redraw(cam, stats=True, **kw))
    assert np.array_equal(a, b)
    for bits in ("1", "2", "4", "5"):                      # per stage: primary / deeper levels / lights
        os.environ["RTB_WF_SPLIT"] = bits
        try:
            assert np.array_equal(r.redraw(cam, **kw), a), bits
        finally:
            del os.environ["RTB_WF_SPLIT"]
    assert (sa.nodes, sa.shadow_rays, sa.sphere_tests) == (sb.nodes, sb.shadow_rays, sb.sphere_tests)
    a, b = both(lambda: r.redraw(cam, spp_grid=2, star_mode=rt.RTB_STARS_HASH, star_seed=5, **kw))
    assert np.array_equal(a, b)
    a, b = both(lambda: r.redraw(cam, stripe_rows=3, stripe_rank=1, stripe_world=4, out_packed=1, **kw))
    assert np.array_equal(a, b)
    a, b = both(lambda: _group(rt, r, cams, H, W, **kw))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    os.environ["RTB_WF_CAP_NODES"] = str(W * 5 * 10)
    try:
        with rt.Renderer(rt.scenes.demo_objects() if scene == "demo" else rt.scenes.synth_objects(160, 7)) as r2:
            def banded():
                st = rt._abi.RtbStats()
                dev = torch.zeros((5, H, W, 4), dtype=torch.uint8, device="cuda")
                r2.redraw_device_group(cams, rt.make_opts(**kw), [dev[i].data_ptr() for i in range(5)], stats=st)
                return dev.cpu().numpy(), st.bands
            (c, nb), (d, _) = both(banded)
            assert nb > 1 and np.array_equal(c, d)
            assert all(np.array_equal(c[i], a[i]) for i in range(5))
    finally:
        del os.environ["RTB_WF_CAP_NODES"]


@pytest.mark.parametrize("S,H", [(1, 70), (4, 70), (4, 64), (8, 21)])
def test_stripe_rows_to_host_assembles_the_frame(rt, demo, S, H):
    """rtb_stripe_rows_to_host: every 'rank' copies its packed stripe rows from the device to their natural
    rows of ONE host frame (page-locked with rtb_host_register) - complete stripes as one strided 2-D copy, a
    ragged last stripe on its own; together the ranks' copies are the whole frame."""
    import torch
    A = rt._abi
    lib = A.load()
    W, world = 96, 3
    cam = rt.scenes.demo_camera(W, H)
    full = demo.redraw(cam)
    host = np.zeros((H, W, 4), np.uint8)
    A.check(lib.rtb_host_register(C.c_void_p(host.ctypes.data), host.nbytes))
    try:
        for rank in range(world):
            rows = lib.rtb_stripe_local_rows(H, S, rank, world)
            dev = torch.zeros((max(rows, 1), W, 4), dtype=torch.uint8, device="cuda")
            if rows:
                demo.redraw_device(cam, rt.make_opts(stripe_rows=S, stripe_rank=rank, stripe_world=world, out_packed=1),
                                   dev.data_ptr())
            A.check(lib.rtb_stripe_rows_to_host(C.c_void_p(dev.data_ptr()), C.c_void_p(host.ctypes.data), W, H, S,
                                                rank, world, None))
            torch.cuda.synchronize()
        assert demo.finish()
        assert np.array_equal(host, full)
    finally:
        lib.rtb_host_unregister(C.c_void_p(host.ctypes.data))
    assert lib.rtb_stripe_rows_to_host(None, C.c_void_p(host.ctypes.data), W, H, S, 0, world, None) == A.RTB_ERR_INVALID


def test_sequence_into_shared_host_frames(rt, synth):
    """redraw_sequence(host_frames=...): every frame of the sequence lands in the shared, page-locked host ring."""
    import torch
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    W, H = 160, 90
    cams = rt.scenes.orbit_cameras(W, H, 7)
    kw = dict(segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
    refs = [synth.redraw(c, **kw) for c in cams]
    sr = mg.StripedRenderer(synth, W, H, 0, 1, group=3, slots=3)
    ring = mg.SharedHostFrames(7, H, W, 0, 1)
    try:
        rep = sr.redraw_sequence(cams, frames_in_flight=2, host_frames=ring, **kw)
        torch.cuda.synchronize()
        assert rep == 0
        for i in range(7):
            assert np.array_equal(ring.frames[i], refs[i]), i
    finally:
        ring.close()


def test_frame_batch_renderer_collects_the_overflow_verdict(rt, synth):
    """FrameBatchRenderer on the wavefront path (> 32 spheres): an asynchronous frame whose ray queues overflow is
    re-rendered BEFORE the all-gather (advisor finding of round 1: the invalid frame used to be gathered silently,
    or the ranks desynchronised)."""
    import torch
    mg = importlib.import_module("html5-canvas-raytracer_b200.multigpu")
    W, H = 160, 90
    cams = rt.scenes.orbit_cameras(W, H, 3)
    kw = dict(segs=6, kernel=rt.RTB_KERNEL_WAVEFRONT)
    refs = [synth.redraw(c, **kw) for c in cams]
    os.environ["RTB_WF_CAP_NODES"] = str(W * 40)
    os.environ["RTB_WF_NODES_PER_RAY"] = "1.05"            # bands sized too optimistically: the first attempt overflows
    try:
        with rt.Renderer(rt.scenes.synth_objects(160, 7)) as r:
            fb = mg.FrameBatchRenderer(r, W, H, 0, 1)
            for f in range(3):
                out = fb.redraw_step(cams, f, **kw)
                torch.cuda.synchronize()
                assert np.array_equal(out[0].cpu().numpy(), refs[f]), f
            assert r._lib.rtb_finish(r._h) == 0            # nothing left uncollected
    finally:
        del os.environ["RTB_WF_CAP_NODES"]
        del os.environ["RTB_WF_NODES_PER_RAY"]


def test_redraw_batch_validates_the_output_buffer(rt, demo):
    cams = rt.scenes.orbit_cameras(64, 36, 3)
    with pytest.raises(ValueError):
        demo.redraw_batch(cams, out=np.zeros((2, 36, 64, 4), np.uint8))           # too small for 3 frames
    with pytest.raises(ValueError):
        demo.redraw_batch(cams, out=np.zeros((3, 36, 64, 4), np.float32))         # wrong dtype
    with pytest.raises(ValueError):
        demo.redraw_batch(cams, out=np.zeros((3, 36, 128, 4), np.uint8)[:, :, ::2])   # not contiguous
    with pytest.raises(ValueError):
        demo.redraw_batch([cams[0], rt.scenes.demo_camera(64, 40)])               # frames of different sizes
    out = demo.redraw_batch(cams)
    assert out.shape == (3, 36, 64, 4) and all(np.array_equal(out[i], demo.redraw(cams[i])) for i in range(3))


@pytest.mark.parametrize("lights", [
    [],                                                              # no light: ambient / reflection only
    [[5, 10, 5]],                                                    # one
    [[5, 10, 5], [5, 10, 0], [-6, 8, 4]],                            # three: the second lq record (lights 2, 3)
    [[5, 10, 5], [5, 10, 0], [-6, 8, 4], [0, 12, -6]],               # four = RTB_MAX_LIGHTS (main.js:283 lists 2 + 2)
])
def test_light_counts_against_the_oracle(rt, oracle, lights):
    """0, 1, 3 and 4 lights: the light_intensity carry (main.js:284, declared once outside the loop) runs through
    every facing light in order, through refractive occluders too (glass, bubble).  fp64 kernel bit-exact, fp32
    kernels inside the gate, split and fused pipelines byte-identical."""
    from conftest import FP32_WITHIN2, lsb_stats
    W, H = 256, 144
    cam = rt.scenes.demo_camera(W, H)
    with rt.Renderer(rt.scenes.demo_objects(), lights=lights, light_intensity=50.0) as r:
        ref, cnt = oracle.render(r.desc, cam, rt.make_opts(segs=8))
        img64, st = r.redraw(cam, stats=True, segs=8, precision=rt.RTB_PRECISION_FP64)
        assert np.array_equal(img64, ref) and st.shadow_rays == cnt["shadow_rays"]
        frames = {}
        for bits in ("7", "0"):
            os.environ["RTB_WF_SPLIT"] = bits
            try:
                frames[bits] = r.redraw(cam, segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)
            finally:
                del os.environ["RTB_WF_SPLIT"]
        assert np.array_equal(frames["7"], frames["0"])
        assert lsb_stats(frames["7"], ref)["within2"] >= FP32_WITHIN2
        assert lsb_stats(r.redraw(cam, segs=8, kernel=rt.RTB_KERNEL_PIXEL), ref)["within2"] >= FP32_WITHIN2
