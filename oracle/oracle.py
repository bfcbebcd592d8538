"""ctypes wrapper around oracle/_build/liboracle.so (the C float64 restatement of main.js).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("html5-canvas-raytracer_b200")
A = importlib.import_module("html5-canvas-raytracer_b200._abi")

LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


class OracleCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "nodes", "shadow_rays", "sphere_tests", "sampler_calls", "tir", "depth_cap_px",
        "max_nodes_px", "star_draws")]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n, _ in self._fields_}
        d["rays"] = d["nodes"] + d["shadow_rays"]   # SURVEY.md §8d ray definition
        return d


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rt_oracle.c")
    hdr = os.path.join(_ROOT, "include", "rtb200.h")
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.oracle_render.argtypes = [C.POINTER(A.RtbSceneDesc), C.POINTER(A.RtbCamera), C.POINTER(A.RtbRenderOpts),
                                    vp, vp, C.POINTER(OracleCounters), C.c_int, C.c_int, vp, C.c_uint64]
        L.oracle_render.restype = C.c_int
        L.oracle_trace_ray.argtypes = [C.POINTER(A.RtbSceneDesc), C.c_int, A.D3, A.D3, A.D3, C.POINTER(OracleCounters)]
        L.oracle_trace_ray.restype = C.c_int
        L.oracle_intersect_sphere.argtypes = [C.POINTER(A.RtbSceneDesc), C.c_int, A.D3, A.D3, vp]
        L.oracle_intersect_sphere.restype = C.c_double
        L.oracle_sample_texture.argtypes = [C.POINTER(A.RtbTexture), C.c_double, C.c_double, A.D3]
        L.oracle_sample_texture.restype = None
        L.oracle_clamped_u8.argtypes = [C.c_double]; L.oracle_clamped_u8.restype = C.c_uint8
        L.oracle_toint32.argtypes = [C.c_double]; L.oracle_toint32.restype = C.c_int32
        L.oracle_star_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.oracle_star_uniform.restype = C.c_double
        L.oracle_max_threads.argtypes = []; L.oracle_max_threads.restype = C.c_int
        L.oracle_set_libm_jitter.argtypes = [C.c_uint64, C.c_int64]; L.oracle_set_libm_jitter.restype = None
        _lib = L
    return _lib


def render(scene_desc, camera, opts, row_step=1, threads=0, want_rgb=False, rand_stream=None, out=None):
    """scene_desc: scene.SceneDesc; camera: scene.Camera or RtbCamera; opts: RtbRenderOpts.
    Returns (rgba8 H x W x 4, counters dict[, rgb float64 H x W x 3])."""
    cam = camera.to_c() if hasattr(camera, "to_c") else camera
    H, W = cam.height, cam.width
    if out is None:
        out = np.zeros((H, W, 4), np.uint8)
    rgb = np.zeros((H, W, 3), np.float64) if want_rgb else None
    cnt = OracleCounters()
    rs = None if rand_stream is None else np.ascontiguousarray(rand_stream, np.float64)
    rc = lib().oracle_render(C.byref(scene_desc.c), C.byref(cam), C.byref(opts), out.ctypes.data,
                             rgb.ctypes.data if want_rgb else None, C.byref(cnt), row_step, threads,
                             rs.ctypes.data if rs is not None else None, 0 if rs is None else rs.size)
    if rc != 0:
        raise ValueError("oracle_render: bad arguments")
    return (out, cnt.as_dict(), rgb) if want_rgb else (out, cnt.as_dict())


def trace_ray(scene_desc, segs, org, dir_):
    o = A.D3()
    cnt = OracleCounters()
    lib().oracle_trace_ray(C.byref(scene_desc.c), segs, A.D3(*org), A.D3(*dir_), o, C.byref(cnt))
    return list(o), cnt.as_dict()


def intersect_sphere(scene_desc, idx, org, dir_, ext=True):
    e = (C.c_double * 13)()
    t = lib().oracle_intersect_sphere(C.byref(scene_desc.c), idx, A.D3(*org), A.D3(*dir_),
                                      C.cast(e, C.c_void_p) if ext else None)
    return t, list(e)


def max_threads() -> int:
    return int(lib().oracle_max_threads())
