"""ctypes mirror of include/rtb200.h and the loader for csrc/librtb200.so.

The structs here are byte-for-byte the C structs of the public header (checked by
tests/test_abi.py against sizeof/offsetof values the library itself reports).
There is NO CPU fallback: `load()` raises when the CUDA library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

RTB_ABI_VERSION = 1
RTB_MAX_LIGHTS = 4
RTB_MAX_SEGS = 16
RTB_MAX_GROUP = 8

RTB_OK = 0
RTB_ERR_INVALID, RTB_ERR_NO_DEVICE, RTB_ERR_CUDA, RTB_ERR_OOM = -1, -2, -3, -4
RTB_ERR_NOT_INIT, RTB_ERR_OVERFLOW, RTB_ERR_UNSUPPORTED = -5, -6, -7

RTB_SAMPLER_COLOR, RTB_SAMPLER_TEXTURE, RTB_SAMPLER_CHECKER, RTB_SAMPLER_STARS = 0, 1, 2, 3
RTB_PRECISION_FP32, RTB_PRECISION_FP64 = 0, 1
RTB_KERNEL_AUTO, RTB_KERNEL_PIXEL, RTB_KERNEL_WAVEFRONT = 0, 1, 2
RTB_STARS_OFF, RTB_STARS_HASH, RTB_STARS_STREAM = 0, 1, 2
RTB_CAMERA_REFERENCE, RTB_CAMERA_BASIS = 0, 1

D3 = C.c_double * 3


class RtbMaterial(C.Structure):
    _fields_ = [
        ("color", D3),
        ("albedo", C.c_double * 5),
        ("specular_exponent", C.c_double),
        ("refract_index", C.c_double),
        ("sampler_kind", C.c_int32),
        ("texture_unit", C.c_int32),
        ("checker_freq", C.c_double * 2),
        ("checker_color", D3 * 2),
        ("star_threshold", C.c_double),
        ("star_gain", C.c_double),
    ]


class RtbSphere(C.Structure):
    _fields_ = [("origin", D3), ("r2", C.c_double), ("mtl", RtbMaterial)]


class RtbTexture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("texels", C.POINTER(C.c_uint8))]


class RtbSceneDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("n_spheres", C.c_int32),
        ("spheres", C.POINTER(RtbSphere)),
        ("n_textures", C.c_int32),
        ("textures", C.POINTER(RtbTexture)),
        ("n_lights", C.c_int32),
        ("lights", D3 * RTB_MAX_LIGHTS),
        ("light_intensity", C.c_double),
        ("eps", C.c_double),
    ]


class RtbCamera(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("origin", D3),
        ("axisX", D3),
        ("axisY", D3),
        ("axisZ", D3),
        ("projW", C.c_double),
        ("projH", C.c_double),
        ("projD", C.c_double),
        ("camera_mode", C.c_int32),
        ("_pad", C.c_int32),
    ]


class RtbRenderOpts(C.Structure):
    _fields_ = [
        ("segs", C.c_int32),
        ("spp_grid", C.c_int32),
        ("precision", C.c_int32),
        ("kernel", C.c_int32),
        ("star_mode", C.c_int32),
        ("y0", C.c_int32),
        ("y1", C.c_int32),
        ("stripe_rows", C.c_int32),
        ("stripe_rank", C.c_int32),
        ("stripe_world", C.c_int32),
        ("out_packed", C.c_int32),
        ("star_seed", C.c_uint64),
        ("star_stream", C.POINTER(C.c_double)),
        ("star_stream_len", C.c_uint64),
    ]


class RtbStats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("nodes", C.c_uint64),
        ("shadow_rays", C.c_uint64),
        ("sphere_tests", C.c_uint64),
        ("launches", C.c_uint32),
        ("bands", C.c_uint32),
        ("ms_trace_primary", C.c_double),
        ("ms_trace_secondary", C.c_double),
        ("ms_shadow", C.c_double),
        ("ms_compose", C.c_double),
    ]


# Every symbol include/rtb200.h declares; tests assert the built library exports all of them.
EXPORTS = [
    "rtb_init", "rtb_shutdown", "rtb_scene_create", "rtb_scene_destroy", "rtb_scene_h2d_bytes", "rtb_finish", "rtb_set_frames_in_flight", "rtb_copy_status_device",
    "rtb_peer_alloc", "rtb_peer_open", "rtb_peer_close", "rtb_peer_free", "rtb_peer_scatter", "rtb_peer_scatter_group", "rtb_peer_signal", "rtb_peer_wait", "rtb_render",
    "rtb_render_device", "rtb_render_device_group", "rtb_render_batch", "rtb_unstripe_device", "rtb_stripe_local_rows",
    "rtb_stripe_padded_rows", "rtb_default_opts", "rtb_last_error", "rtb_version",
    "rtb_abi_version", "rtb_device_info", "rtb_clock_probe", "rtb_stripe_rows_to_host", "rtb_host_register",
    "rtb_host_unregister", "rtb_look_at", "rtb_sort_spheres", "rtb_debug_table_pad", "rtb_debug_sweep_tables",
    "rtb_decode_png",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTB200_LIB: another build of the SAME library (kernel experiments, tools/build_variant.py); never a fallback
LIB_PATH = os.environ.get("RTB200_LIB") or os.path.join(_HERE, "csrc", "librtb200.so")
_lib = None


class RtbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rtb200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """dlopen csrc/librtb200.so and declare prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the render path.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    lib.rtb_init.argtypes = [C.c_int]; lib.rtb_init.restype = C.c_int
    lib.rtb_shutdown.argtypes = []; lib.rtb_shutdown.restype = None
    lib.rtb_scene_create.argtypes = [C.POINTER(RtbSceneDesc), C.POINTER(vp)]
    lib.rtb_scene_create.restype = C.c_int
    lib.rtb_scene_destroy.argtypes = [vp]; lib.rtb_scene_destroy.restype = None
    lib.rtb_scene_h2d_bytes.argtypes = [vp]; lib.rtb_scene_h2d_bytes.restype = C.c_uint64
    lib.rtb_finish.argtypes = [vp]; lib.rtb_finish.restype = C.c_int
    lib.rtb_set_frames_in_flight.argtypes = [C.c_int32]; lib.rtb_set_frames_in_flight.restype = C.c_int
    u8p = C.POINTER(C.c_uint8)
    vpp = C.POINTER(C.c_void_p)
    lib.rtb_peer_alloc.argtypes = [C.c_uint64, vpp, u8p]; lib.rtb_peer_alloc.restype = C.c_int
    lib.rtb_peer_open.argtypes = [u8p, vpp]; lib.rtb_peer_open.restype = C.c_int
    lib.rtb_peer_close.argtypes = [vp]; lib.rtb_peer_close.restype = C.c_int
    lib.rtb_peer_free.argtypes = [vp]; lib.rtb_peer_free.restype = C.c_int
    lib.rtb_peer_scatter.argtypes = [vp, vpp, vpp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp,
                                     C.c_uint32, C.c_uint32, C.c_uint32, vp]
    lib.rtb_peer_scatter.restype = C.c_int
    lib.rtb_peer_scatter_group.argtypes = [vp, C.c_uint64, C.c_int32, vpp, C.c_uint64, vpp, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.c_int32, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    lib.rtb_peer_scatter_group.restype = C.c_int
    lib.rtb_peer_signal.argtypes = [vpp, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, vp]; lib.rtb_peer_signal.restype = C.c_int
    lib.rtb_peer_wait.argtypes = [vp, C.c_int32, C.c_uint32, C.c_uint32, vp, vp]; lib.rtb_peer_wait.restype = C.c_int
    lib.rtb_copy_status_device.argtypes = [vp, vp, vp]; lib.rtb_copy_status_device.restype = C.c_int
    lib.rtb_render.argtypes = [vp, C.POINTER(RtbCamera), C.POINTER(RtbRenderOpts), vp, C.POINTER(RtbStats)]
    lib.rtb_render.restype = C.c_int
    lib.rtb_render_device.argtypes = [vp, C.POINTER(RtbCamera), C.POINTER(RtbRenderOpts), vp, vp, C.POINTER(RtbStats)]
    lib.rtb_render_device.restype = C.c_int
    lib.rtb_render_device_group.argtypes = [vp, C.POINTER(RtbCamera), C.c_int32, C.POINTER(RtbRenderOpts), vpp, vp,
                                            C.POINTER(RtbStats)]
    lib.rtb_render_device_group.restype = C.c_int
    lib.rtb_render_batch.argtypes = [vp, C.POINTER(RtbCamera), C.c_int32, C.POINTER(RtbRenderOpts), vp, C.POINTER(RtbStats)]
    lib.rtb_render_batch.restype = C.c_int
    lib.rtb_unstripe_device.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.rtb_unstripe_device.restype = C.c_int
    lib.rtb_stripe_local_rows.argtypes = [C.c_int32] * 4; lib.rtb_stripe_local_rows.restype = C.c_int32
    lib.rtb_stripe_padded_rows.argtypes = [C.c_int32] * 3; lib.rtb_stripe_padded_rows.restype = C.c_int32
    lib.rtb_default_opts.argtypes = [C.POINTER(RtbRenderOpts)]; lib.rtb_default_opts.restype = None
    lib.rtb_last_error.argtypes = []; lib.rtb_last_error.restype = C.c_char_p
    lib.rtb_version.argtypes = []; lib.rtb_version.restype = C.c_char_p
    lib.rtb_abi_version.argtypes = []; lib.rtb_abi_version.restype = C.c_uint32
    lib.rtb_device_info.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_char_p, C.c_int32]
    lib.rtb_device_info.restype = C.c_int
    lib.rtb_stripe_rows_to_host.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.rtb_stripe_rows_to_host.restype = C.c_int
    lib.rtb_host_register.argtypes = [vp, C.c_uint64]; lib.rtb_host_register.restype = C.c_int
    lib.rtb_host_unregister.argtypes = [vp]; lib.rtb_host_unregister.restype = C.c_int
    lib.rtb_look_at.argtypes = [D3, D3, D3, C.c_int32, C.c_int32, C.c_int32, C.POINTER(RtbCamera)]
    lib.rtb_look_at.restype = C.c_int
    lib.rtb_sort_spheres.argtypes = [C.POINTER(RtbSphere), C.c_int32, D3]; lib.rtb_sort_spheres.restype = C.c_int
    lib.rtb_decode_png.argtypes = [vp, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp, C.c_uint64]
    lib.rtb_decode_png.restype = C.c_int
    lib.rtb_debug_table_pad.argtypes = [C.c_int32]; lib.rtb_debug_table_pad.restype = C.c_int32
    lib.rtb_debug_sweep_tables.argtypes = [C.POINTER(RtbSceneDesc), D3, vp, vp, vp]; lib.rtb_debug_sweep_tables.restype = C.c_int
    lib.rtb_clock_probe.argtypes = [vp, C.c_int32, vp]; lib.rtb_clock_probe.restype = C.c_int
    _lib = lib
    return lib


def stream_arg(stream):
    """cudaStream_t argument of the device entry points.  None -> NULL = the library's own stream.  An
    integer handle is passed through, except 0: that is CUDA's legacy default stream (what
    torch.cuda.current_stream().cuda_stream returns outside a stream context) and must be named
    explicitly as cudaStreamLegacy (0x1), because NULL already means "library stream"."""
    if stream is None:
        return None
    return C.c_void_p(1 if int(stream) == 0 else int(stream))


def check(code: int) -> None:
    if code != RTB_OK:
        msg = load().rtb_last_error()
        raise RtbError(code, msg.decode("utf-8", "replace") if msg else "")
