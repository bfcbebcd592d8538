"""The drop-in boundary without a GPU: the built C-ABI library loads, exports every symbol that
include/rtb200.h declares, its struct layouts equal the ctypes mirror, the pure-host entry points
behave, and computing entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtb200.h")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def lib(rt):
    if not os.path.exists(rt._abi.LIB_PATH):
        builder = __import__("importlib").import_module("html5-canvas-raytracer_b200.build")
        builder.build_library()
    return rt._abi.load()


def test_every_declared_symbol_is_exported(rt, lib):
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", src)))
    assert declared and sorted(rt._abi.EXPORTS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", rt._abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rtb_[a-z0-9_]+)", out))
    assert set(declared) <= exported, set(declared) - exported
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rtb_abi_version() == rt._abi.RTB_ABI_VERSION
    assert b"sm_100a" in lib.rtb_version()


def test_struct_layout_matches_header(rt, tmp_path):
    """sizeof/offsetof from a C compile of the public header vs the ctypes mirror."""
    A = rt._abi
    structs = {"RtbMaterial": A.RtbMaterial, "RtbSphere": A.RtbSphere, "RtbTexture": A.RtbTexture,
               "RtbSceneDesc": A.RtbSceneDesc, "RtbCamera": A.RtbCamera, "RtbRenderOpts": A.RtbRenderOpts,
               "RtbStats": A.RtbStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for name, st in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f, _ in st._fields_:
            lines.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines.append("return 0;}")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c11", str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for name, st in structs.items():
        assert int(got[name]) == C.sizeof(st), name
        for f, _ in st._fields_:
            assert int(got[f"{name}.{f}"]) == getattr(st, f).offset, f"{name}.{f}"


def test_stripe_arithmetic_matches_python_twin(rt, lib):
    mg = __import__("importlib").import_module("html5-canvas-raytracer_b200.multigpu")
    for H in (1, 7, 16, 70, 150, 1080, 2160):
        for S in (1, 4, 16):
            for world in (1, 2, 3, 4, 8):
                tot = 0
                for r in range(world):
                    n = lib.rtb_stripe_local_rows(H, S, r, world)
                    assert n == len(mg.stripe_rows_of(H, S, r, world))
                    tot += n
                assert tot == H
                assert lib.rtb_stripe_padded_rows(H, S, world) == mg.padded_rows(H, S, world)


def test_default_opts_are_the_reference_literals(rt, lib):
    o = rt._abi.RtbRenderOpts()
    lib.rtb_default_opts(C.byref(o))
    assert (o.segs, o.spp_grid, o.precision, o.kernel, o.star_mode) == (8, 1, rt.RTB_PRECISION_FP32, rt.RTB_KERNEL_AUTO, rt.RTB_STARS_OFF)
    assert (o.y0, o.y1, o.stripe_world, o.out_packed) == (0, 0, 0, 0)


def test_scene_flattening(rt):
    objs = rt.scenes.demo_objects()
    d = rt.SceneDesc(objs)
    assert d.c.n_spheres == 14 and d.c.n_textures == 3 and d.c.n_lights == 2
    assert list(d.c.lights[0]) == [5.0, 10.0, 5.0] and list(d.c.lights[1]) == [5.0, 10.0, 0.0]
    assert d.c.light_intensity == 50.0 and d.c.eps == 0.001
    home = d.c.spheres[12]
    assert home.mtl.sampler_kind == rt._abi.RTB_SAMPLER_CHECKER
    assert list(home.mtl.checker_freq) == [5000.0, 2500.0]
    assert [list(c) for c in home.mtl.checker_color] == [[1, 1, 0], [1, 0, 1]]
    sky = d.c.spheres[13]
    assert (sky.mtl.star_threshold, sky.mtl.star_gain) == (0.001, 1000.0)
    units = {d.c.spheres[i].mtl.texture_unit for i in (1, 5, 11)}
    assert units == {0, 1, 2}
    with pytest.raises(ValueError):
        rt.createMaterial([1, 1], [0] * 5, 1, 1)
    t = rt.createTexture()
    m = rt.createMaterial([1, 1, 1], [1, 0, 0, 0, 0], 0, 1)
    m.sampler = rt.textureSampler(t)
    with pytest.raises(ValueError):                       # texture never loaded (main.js:167-178 would poll forever)
        rt.SceneDesc([rt.createSphere([0, 0, 0], 1, m)])


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback(rt, lib):
    assert lib.rtb_init(-1) == rt._abi.RTB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.rtb_last_error()
    with pytest.raises(rt.RtbError) as e:
        rt.Renderer(rt.scenes.demo_objects())
    assert e.value.code == rt._abi.RTB_ERR_NO_DEVICE
    # nothing initialised: computing entry points refuse
    h = C.c_void_p()
    d = rt.SceneDesc(rt.scenes.demo_objects())
    assert lib.rtb_scene_create(C.byref(d.c), C.byref(h)) == rt._abi.RTB_ERR_NOT_INIT


def test_product_never_imports_the_oracle():
    """The oracle is a checker: nothing under the product package (or its C sources) may reference it."""
    pkg = os.path.join(ROOT, "html5-canvas-raytracer_b200")
    for dp, _, files in os.walk(pkg):
        if "_obj" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".cc")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "rt_oracle" not in txt, f


def test_host_side_look_at_and_sort_match_the_reference_shaped_host_layer(rt, lib):
    """rtb_look_at (main.js:92-105) and rtb_sort_spheres (main.js:159-163) are pure host code: bit-identical to the
    Python mirror of main.js's lookAt / projection / objects.sort, which the golden frames of the unmodified
    main.js pin (default camera: axisX = (-1,0,0), axisY = (0,1,0), axisZ = (0,0,-1))."""
    A = rt._abi
    out = A.RtbCamera()
    assert lib.rtb_look_at(A.D3(0, 1.5, 10), A.D3(0, 1.5, 0), A.D3(0, 1, 0), 1280, 720, A.RTB_CAMERA_REFERENCE, C.byref(out)) == 0
    assert list(out.axisX) == [-1.0, 0.0, 0.0] and list(out.axisY) == [0.0, 1.0, 0.0] and list(out.axisZ) == [0.0, 0.0, -1.0]
    ref = rt.scenes.demo_camera(1280, 720).to_c()
    for f in ("projW", "projH", "projD"):
        assert getattr(out, f) == getattr(ref, f)
    for cam in rt.scenes.orbit_cameras(640, 360, 24):      # rotated cameras
        c = cam.to_c()
        tgt = A.D3(0, 1.5, 0)
        assert lib.rtb_look_at(c.origin, tgt, A.D3(0, 1, 0), 640, 360, A.RTB_CAMERA_BASIS, C.byref(out)) == 0
        for f in ("origin", "axisX", "axisY", "axisZ"):
            assert list(getattr(out, f)) == list(getattr(c, f)), f
        assert out.camera_mode == A.RTB_CAMERA_BASIS
    assert lib.rtb_look_at(A.D3(0, 0, 0), A.D3(0, 0, 1), A.D3(0, 1, 0), 0, 10, 0, C.byref(out)) == A.RTB_ERR_INVALID
    # sort: the demo scene in SOURCE order and the 1024-sphere scene shuffled, against sortObjects
    import random
    for objs, origin in ((rt.scenes.demo_objects(), (0.0, 1.5, 10.0)), (rt.scenes.synth_objects(1024, 1024), (3.0, 2.0, -7.0))):
        shuffled = list(objs)
        random.Random(5).shuffle(shuffled)
        want = list(shuffled)
        rt.sortObjects(want, list(origin))
        desc = rt.SceneDesc(shuffled)
        assert lib.rtb_sort_spheres(desc.c.spheres, len(shuffled), A.D3(*origin)) == 0
        got = [(tuple(desc.c.spheres[i].origin), desc.c.spheres[i].r2) for i in range(len(shuffled))]
        assert got == [(tuple(float(c) for c in o.origin), o.r2) for o in want]
    assert lib.rtb_sort_spheres(None, 3, A.D3(0, 0, 0)) == A.RTB_ERR_INVALID


def test_bench_reference_arm_prints_the_contract_line_with_all_host_threads():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): ONE JSON line with the contract's
    keys, the workload `config` the GPU arm prints too, and every host thread at work even under torchrun's
    OMP_NUM_THREADS=1 (round 1: the N > 1 reference runs silently used one thread)."""
    import json
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["config"]["workload"] == "c3_synth1024_4k" and d["config"]["n_spheres"] == 1024
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # every other rank of a torchrun launch exits 0 without work and without output
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == ""


def _png_bytes(w, h, ctype, raw_rows, extra=b""):
    import struct
    import zlib

    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xffffffff)
    z = zlib.compress(b"".join(raw_rows))
    half = len(z) // 2                                         # two IDAT chunks: the stream spans chunk borders
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) + extra + \
        chunk(b"IDAT", z[:half]) + chunk(b"IDAT", z[half:]) + chunk(b"IEND", b""), chunk


def _c_decode(lib, data):
    import ctypes as C
    w, h = C.c_int32(), C.c_int32()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    rc = lib.rtb_decode_png(buf, len(data), C.byref(w), C.byref(h), None, 0)
    if rc != 0:
        return rc, None
    out = np.zeros((h.value, w.value, 4), np.uint8)
    rc = lib.rtb_decode_png(buf, len(data), C.byref(w), C.byref(h), out.ctypes.data, out.nbytes)
    return rc, out


def test_c_abi_png_decoder_matches_the_python_loadTexture_path(rt, lib):
    """rtb_decode_png (loadTexture for C / N-API hosts, main.js:375-395) against the package's png.py on every
    colour type and filter type, on the reference's own textures when they are at hand, and on malformed input."""
    rng = np.random.default_rng(5)
    w, h = 13, 11
    for ctype, ch in ((0, 1), (2, 3), (3, 1), (4, 2), (6, 4)):
        img = rng.integers(0, 4 if ctype == 3 else 256, (h, w * ch), dtype=np.uint8).astype(np.int32)
        rows, prev = [], np.zeros(w * ch, np.int32)
        for y in range(h):
            line, ft = img[y], y % 5
            out = np.zeros_like(line)
            for i in range(len(line)):
                a = line[i - ch] if i >= ch else 0
                b = prev[i]
                c = prev[i - ch] if i >= ch else 0
                pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                p = [0, a, b, (a + b) >> 1, a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)][ft]
                out[i] = (line[i] - p) & 255
            rows.append(bytes([ft]) + bytes(out.astype(np.uint8)))
            prev = line
        extra = b""
        if ctype == 3:
            _, chunk = _png_bytes(1, 1, 3, [b"\0\0"])
            extra = chunk(b"PLTE", rng.integers(0, 256, 12, dtype=np.uint8).tobytes()) + chunk(b"tRNS", bytes([7, 200]))
        data, _ = _png_bytes(w, h, ctype, rows, extra)
        rc, got = _c_decode(lib, data)
        assert rc == 0, (ctype, lib.rtb_last_error())
        ww, hh, want = rt.png.decode_rgba8(data)
        assert (ww, hh) == (w, h) and np.array_equal(got, want), ctype
    # the reference's textures (present in this container only): the C decoder yields the committed texels
    earth, mars = rt.scenes.load_planet_textures()
    for name, tex in (("earth", earth), ("mars", mars)):
        p = os.path.join("/root/reference", name + ".png")
        if os.path.exists(p):
            rc, got = _c_decode(lib, open(p, "rb").read())
            assert rc == 0 and np.array_equal(got.reshape(-1), np.asarray(tex.texels).reshape(-1)), name
    # malformed input: loud errors, no crash
    import ctypes as C
    assert _c_decode(lib, b"not a png at all")[0] != 0
    assert _c_decode(lib, data[:40])[0] != 0                               # truncated before IDAT ends
    bad = bytearray(data)
    bad[24] = 16                                                           # 16-bit depth: unsupported
    assert _c_decode(lib, bytes(bad))[0] != 0
    w_, h_ = C.c_int32(), C.c_int32()
    small = np.zeros(8, np.uint8)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    assert lib.rtb_decode_png(buf, len(data), C.byref(w_), C.byref(h_), small.ctypes.data, small.nbytes) != 0
    junk, _ = _png_bytes(w, h, 6, [b"\x09" + bytes(w * 4)] * h)            # filter type 9
    assert _c_decode(lib, junk)[0] != 0
    short, _ = _png_bytes(w, h, 6, [b"\x00" + bytes(w * 4)] * (h - 1))     # one scanline missing
    assert _c_decode(lib, short)[0] != 0
