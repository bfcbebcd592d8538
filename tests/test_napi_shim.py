"""csrc/napi_shim.cc EXECUTED against a mock of the Node-API surface it uses (tests/napi_mock/mock_node.cc): this
image has no Node, so the addon could otherwise only be compiled.  The mock is not Node - it checks the shim's own
logic: walking the scene / camera / opts objects main.js builds, Uint8ClampedArray and plain-Array texels, the
error paths (JS exceptions, never a C++ exception or an abort), and the promise-returning renderAsync (work on a
worker thread, completion on the main thread, the framebuffer kept alive by a reference)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_SRC = os.path.join(ROOT, "tests", "napi_mock", "mock_node.cc")
MOCK_LIB = os.path.join(ROOT, "tests", "napi_mock", "_build", "libmocknode.so")
ADDON = os.path.join(ROOT, "html5-canvas-raytracer_b200", "csrc", "rtb200.node")
UINT8, UINT8_CLAMPED = 1, 2     # napi_typedarray_type


@pytest.fixture(scope="module")
def node(rt):
    builder = __import__("importlib").import_module("html5-canvas-raytracer_b200.build")
    builder.build_library()
    builder.build_napi()
    if not os.path.exists(MOCK_LIB) or os.path.getmtime(MOCK_LIB) < os.path.getmtime(MOCK_SRC):
        os.makedirs(os.path.dirname(MOCK_LIB), exist_ok=True)
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", MOCK_SRC, "-o", MOCK_LIB, "-ldl"], check=True)
    m = C.CDLL(MOCK_LIB, mode=C.RTLD_GLOBAL)      # napi_* become global: rtb200.node binds to them on dlopen
    vp = C.c_void_p
    for name, res, args in (
            ("mock_env_new", vp, []), ("mock_undefined", vp, [vp]), ("mock_object", vp, []), ("mock_array", vp, []),
            ("mock_number", vp, [C.c_double]), ("mock_string", vp, [C.c_char_p]),
            ("mock_typedarray", vp, [C.c_int, vp, C.c_size_t]), ("mock_set", None, [vp, C.c_char_p, vp]),
            ("mock_push", None, [vp, vp]), ("mock_get", vp, [vp, vp, C.c_char_p]), ("mock_kind", C.c_int, [vp]),
            ("mock_number_value", C.c_double, [vp]), ("mock_string_value", C.c_char_p, [vp]),
            ("mock_promise_state", C.c_int, [vp]), ("mock_promise_value", vp, [vp]), ("mock_live_refs", C.c_int, [vp]),
            ("mock_exception", C.c_char_p, [vp]), ("mock_clear_exception", None, [vp]),
            ("mock_load_addon", vp, [vp, C.c_char_p]), ("mock_call", vp, [vp, vp, C.c_int, C.POINTER(vp)]),
            ("mock_run_loop", C.c_int, [vp]), ("mock_typed_data", vp, [vp]), ("mock_typed_len", C.c_size_t, [vp]),
            ("mock_typed_type", C.c_int, [vp])):
        f = getattr(m, name)
        f.restype, f.argtypes = res, args
    env = m.mock_env_new()
    exports = m.mock_load_addon(env, ADDON.encode())
    assert exports, m.mock_exception(env)
    return m, env, exports


def js(m, x):
    """Python -> mock JS value: numbers, strings, lists -> Array, dicts -> Object, uint8 ndarrays -> Uint8ClampedArray."""
    if isinstance(x, np.ndarray):
        return m.mock_typedarray(UINT8_CLAMPED, x.ctypes.data, x.size)
    if isinstance(x, dict):
        o = m.mock_object()
        for k, v in x.items():
            m.mock_set(o, k.encode(), js(m, v))
        return o
    if isinstance(x, (list, tuple)):
        a = m.mock_array()
        for v in x:
            m.mock_push(a, js(m, v))
        return a
    if isinstance(x, str):
        return m.mock_string(x.encode())
    if isinstance(x, int) and not isinstance(x, bool) and x > 2 ** 60:
        return C.c_void_p(x)                                   # already a mock value
    return m.mock_number(float(x))


def scene_and_camera(rt, W, H, keep, plain_array_texels=False):
    """The arguments main.js would hand over, built from the flattened demo scene."""
    desc = rt.SceneDesc(rt.scenes.demo_objects())
    keep.append(desc)
    objs = []
    for i in range(desc.c.n_spheres):
        s = desc.c.spheres[i]
        objs.append({"origin": list(s.origin), "r2": s.r2,
                     "mtl": {"color": list(s.mtl.color), "albedo": list(s.mtl.albedo),
                             "specular_exponent": s.mtl.specular_exponent, "refract_index": s.mtl.refract_index,
                             "sampler_kind": s.mtl.sampler_kind, "texture_unit": s.mtl.texture_unit}})
    texs = []
    for t in range(desc.c.n_textures):
        tx = desc.c.textures[t]
        arr = np.ctypeslib.as_array(tx.texels, shape=(tx.width * tx.height * 4,)).copy()
        keep.append(arr)
        # checkerTexture fills a plain JS Array (main.js:340, 358-366); loadTexture a Uint8ClampedArray (main.js:391)
        texs.append({"width": tx.width, "height": tx.height,
                     "texels": [int(v) for v in arr] if (plain_array_texels and arr.size <= 4096) else arr})
    cam = rt.scenes.demo_camera(W, H)
    camera = {"width": W, "height": H, "origin": cam.origin, "axisX": cam.axisX, "axisY": cam.axisY,
              "axisZ": cam.axisZ, "projW": cam.projW, "projH": cam.projH, "projD": cam.projD}
    return {"objects": objs, "textures": texs}, camera


def call(m, env, exports, name, *args):
    fn = m.mock_get(env, exports, name.encode())
    argv = (C.c_void_p * len(args))(*args)
    r = m.mock_call(env, fn, len(args), argv)
    exc = m.mock_exception(env)
    m.mock_clear_exception(env)
    return r, (exc.decode() if exc else None)


def test_addon_registers_and_reports_errors_as_js_exceptions(rt, node):
    m, env, exports = node
    for name in ("render", "renderAsync", "decodeTexture", "version"):
        assert m.mock_kind(m.mock_get(env, exports, name.encode())) == 7          # a function
    v, exc = call(m, env, exports, "version")
    assert exc is None and b"sm_100a" in m.mock_string_value(v)
    keep = []
    scene, camera = scene_and_camera(rt, 64, 36, keep)
    out = np.zeros(64 * 36 * 4, np.uint8)
    r, exc = call(m, env, exports, "render", js(m, scene), js(m, camera))
    assert r is None and "4 arguments" in exc
    bad_cam = dict(camera); del bad_cam["projW"]
    r, exc = call(m, env, exports, "render", js(m, scene), js(m, bad_cam), js(m, {"segs": 8}), js(m, out))
    assert r is None and "projW" in exc
    r, exc = call(m, env, exports, "render", js(m, scene), js(m, camera), js(m, {"segs": 8}), js(m, out[:100]))
    assert r is None and "smaller" in exc
    r, exc = call(m, env, exports, "render", js(m, scene), js(m, camera), js(m, {"segs": 8}), js(m, [1, 2, 3]))
    assert r is None and "Uint8ClampedArray" in exc
    bad_scene = {"objects": scene["objects"], "textures": [{"width": 16, "height": 8, "texels": [0] * 10}]}
    r, exc = call(m, env, exports, "render", js(m, bad_scene), js(m, camera), js(m, {"segs": 8}), js(m, out))
    assert r is None and "texels" in exc
    r, exc = call(m, env, exports, "renderAsync", js(m, scene), js(m, bad_cam), js(m, {}), js(m, out))
    assert r is None and "projW" in exc and m.mock_live_refs(env) == 0             # nothing leaked on the error path


def test_addon_decodes_png_textures_like_loadTexture(rt, node):
    """decodeTexture(pngBytes): the {width, height, texels, loaded} object main.js's loadTexture leaves behind
    (main.js:385-393), texels equal to the package's own PNG path; usable as a `textures` entry as it is."""
    import struct
    import zlib
    m, env, exports = node

    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xffffffff)
    rng = np.random.default_rng(9)
    w, h = 16, 8
    pal = rng.integers(0, 256, (5, 3), dtype=np.uint8)
    idx = rng.integers(0, 5, (h, w), dtype=np.uint8)
    raw = b"".join(b"\x00" + bytes(idx[y]) for y in range(h))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 3, 0, 0, 0)) + \
        chunk(b"PLTE", pal.tobytes()) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    src = np.frombuffer(png, np.uint8).copy()
    tex, exc = call(m, env, exports, "decodeTexture", m.mock_typedarray(UINT8, src.ctypes.data, src.size))
    assert exc is None and tex
    assert m.mock_number_value(m.mock_get(env, tex, b"width")) == w
    assert m.mock_number_value(m.mock_get(env, tex, b"height")) == h
    assert m.mock_number_value(m.mock_get(env, tex, b"loaded")) == 1
    tx = m.mock_get(env, tex, b"texels")
    assert m.mock_typed_type(tx) == UINT8_CLAMPED and m.mock_typed_len(tx) == w * h * 4
    got = np.ctypeslib.as_array(C.cast(m.mock_typed_data(tx), C.POINTER(C.c_uint8)), shape=(h, w, 4))
    _, _, want = rt.png.decode_rgba8(png)
    assert np.array_equal(got, want)
    # errors are JS exceptions
    r, exc = call(m, env, exports, "decodeTexture", js(m, [1, 2, 3]))
    assert r is None and "Uint8Array" in exc
    junk = np.arange(64, dtype=np.uint8)
    r, exc = call(m, env, exports, "decodeTexture", m.mock_typedarray(UINT8, junk.ctypes.data, junk.size))
    assert r is None and "PNG" in exc
    r, exc = call(m, env, exports, "decodeTexture")
    assert r is None and "1 argument" in exc


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure")
def test_addon_without_a_gpu_fails_loudly(rt, node):
    m, env, exports = node
    keep = []
    scene, camera = scene_and_camera(rt, 64, 36, keep)
    out = np.zeros(64 * 36 * 4, np.uint8)
    r, exc = call(m, env, exports, "render", js(m, scene), js(m, camera), js(m, {"segs": 8}), js(m, out))
    assert r is None and "no CPU fallback" in exc and not out.any()
    p, exc = call(m, env, exports, "renderAsync", js(m, scene), js(m, camera), js(m, {"segs": 8}), js(m, out))
    assert exc is None and m.mock_promise_state(p) == 0
    assert m.mock_run_loop(env) == 1
    assert m.mock_promise_state(p) == 2 and b"no CPU fallback" in m.mock_string_value(m.mock_promise_value(p))
    assert m.mock_live_refs(env) == 0


@pytest.mark.gpu
def test_addon_renders_what_the_c_abi_renders(rt, node):
    m, env, exports = node
    W, H = 128, 72
    keep = []
    scene, camera = scene_and_camera(rt, W, H, keep, plain_array_texels=True)
    with rt.Renderer(rt.scenes.demo_objects()) as r:
        want = r.redraw(rt.scenes.demo_camera(W, H), segs=8)
        want5 = r.redraw(rt.scenes.demo_camera(W, H), segs=5, precision=rt.RTB_PRECISION_FP64)
    out = np.zeros((H, W, 4), np.uint8)
    res, exc = call(m, env, exports, "render", js(m, scene), js(m, camera), js(m, {"segs": 8}), js(m, out))
    assert exc is None and np.array_equal(out, want)
    assert m.mock_number_value(m.mock_get(env, res, b"launches")) >= 1
    # progressive rows (main.js:200-201): only rows [10, 20) are written
    part = np.zeros((H, W, 4), np.uint8)
    res, exc = call(m, env, exports, "render", js(m, scene), js(m, camera), js(m, {"segs": 8, "y0": 10, "y1": 20}), js(m, part))
    assert exc is None and np.array_equal(part[10:20], want[10:20]) and not part[:10].any() and not part[20:].any()
    # promise-returning render: pending until the loop runs, then resolved with the stats; fp64, depth 5
    out2 = np.zeros((H, W, 4), np.uint8)
    p, exc = call(m, env, exports, "renderAsync", js(m, scene), js(m, camera), js(m, {"segs": 5, "precision": 1}), js(m, out2))
    assert exc is None and m.mock_kind(p) == 8 and m.mock_live_refs(env) == 1
    assert m.mock_run_loop(env) == 1
    assert m.mock_promise_state(p) == 1 and np.array_equal(out2, want5) and m.mock_live_refs(env) == 0
    assert m.mock_number_value(m.mock_get(env, m.mock_promise_value(p), b"nodes")) > W * H
    # a request the library refuses: the promise REJECTS with the library's message
    p, exc = call(m, env, exports, "renderAsync", js(m, scene), js(m, camera), js(m, {"segs": 99}), js(m, out2))
    assert exc is None and m.mock_run_loop(env) == 1
    assert m.mock_promise_state(p) == 2 and b"segs" in m.mock_string_value(m.mock_promise_value(p))
