"""Renderer — the host-side stand-in for main.js's `redraw()` (main.js:180-213): it owns the uploaded
scene and fills a `colorbuf.data`-shaped RGBA8 buffer through the C ABI (csrc/librtb200.so).

Every compute call goes to the CUDA library; if the library is not built, or no GPU is visible,
construction raises — there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi as A
from .scene import Camera, SceneDesc, Sphere, make_opts


class Renderer:
    def __init__(self, objects: Sequence[Sphere], lights=None, light_intensity=50.0, eps=0.001,
                 device: int = -1):
        self._lib = A.load()
        A.check(self._lib.rtb_init(int(device)))
        self.desc = SceneDesc(objects, lights=lights, light_intensity=light_intensity, eps=eps)
        self._h = C.c_void_p()
        A.check(self._lib.rtb_scene_create(C.byref(self.desc.c), C.byref(self._h)))

    # -- lifetime ----------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.rtb_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- rendering -----------------------------------------------------------------------------------
    @staticmethod
    def _cam(camera) -> A.RtbCamera:
        return camera.to_c() if isinstance(camera, Camera) else camera

    def redraw(self, camera, out: Optional[np.ndarray] = None, stats: bool = False, opts=None, **kw):
        """Fill (and return) an H x W x 4 uint8 array — the bytes of `colorbuf.data` after the last
        `spanish` tick (main.js:195-200).  Keyword options are those of scene.make_opts (segs=8,
        spp_grid=1, precision, kernel, star_mode, y0/y1, stripe_*, out_packed)."""
        cam = self._cam(camera)
        o = opts if opts is not None else make_opts(**kw)
        if out is None:
            out = np.zeros((cam.height, cam.width, 4), np.uint8)
        if out.dtype != np.uint8 or not out.flags["C_CONTIGUOUS"] or out.size < cam.width * cam.height * 4:
            raise ValueError("out must be a C-contiguous uint8 buffer of width*height*4 bytes")
        st = A.RtbStats()
        A.check(self._lib.rtb_render(self._h, C.byref(cam), C.byref(o), out.ctypes.data,
                                     C.byref(st) if stats else None))
        return (out, st) if stats else out

    def redraw_ptr(self, camera, opts, host_ptr: int, stats: Optional[A.RtbStats] = None) -> None:
        """Same, into a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        cam = self._cam(camera)
        A.check(self._lib.rtb_render(self._h, C.byref(cam), C.byref(opts), C.c_void_p(host_ptr),
                                     C.byref(stats) if stats is not None else None))

    def redraw_device(self, camera, opts, device_ptr: int, stream=None,
                      stats: Optional[A.RtbStats] = None) -> None:
        """Render into DEVICE memory (no D2H): `device_ptr` is a CUDA pointer on the library's device,
        `stream` a cudaStream_t handle (None = the library's own stream; 0 = the legacy default stream,
        e.g. torch.cuda.current_stream().cuda_stream)."""
        cam = self._cam(camera)
        A.check(self._lib.rtb_render_device(self._h, C.byref(cam), C.byref(opts), C.c_void_p(device_ptr),
                                            A.stream_arg(stream),
                                            C.byref(stats) if stats is not None else None))

    def redraw_device_group(self, cameras, opts, device_ptrs, stream=None,
                            stats: Optional[A.RtbStats] = None) -> None:
        """A GROUP of up to RTB_MAX_GROUP frames (same size and options, one camera and one device buffer each)
        in one pass of the wavefront pipeline (rtb_render_device_group)."""
        cams = [self._cam(c) for c in cameras]
        n = len(cams)
        if n != len(device_ptrs):
            raise ValueError("one device buffer per camera")
        arr = (A.RtbCamera * n)(*cams)
        ptrs = (C.c_void_p * n)(*[int(p) for p in device_ptrs])
        A.check(self._lib.rtb_render_device_group(self._h, arr, n, C.byref(opts), ptrs, A.stream_arg(stream),
                                                  C.byref(stats) if stats is not None else None))

    def finish(self) -> bool:
        """Collect the verdict of the last asynchronous redraw_device().  True: the frame is valid.
        False: its ray queues overflowed - the library has shrunk its bands; issue the same call again."""
        rc = self._lib.rtb_finish(self._h)
        if rc == A.RTB_ERR_OVERFLOW:
            return False
        A.check(rc)
        return True

    def set_frames_in_flight(self, n: int) -> None:
        """1 (default) or 2 asynchronous wavefront frames in flight (rtb_set_frames_in_flight): with 2,
        consecutive redraw_device() calls on two streams overlap on the GPU."""
        A.check(self._lib.rtb_set_frames_in_flight(int(n)))

    def redraw_batch(self, cameras, out: Optional[np.ndarray] = None, stats: bool = False, opts=None, **kw):
        """Frame batch (orbit animation, BASELINE config 5): F x H x W x 4 uint8."""
        cams = [self._cam(c) for c in cameras]
        n = len(cams)
        arr = (A.RtbCamera * max(n, 1))(*cams)
        o = opts if opts is not None else make_opts(**kw)
        if any(c.width != cams[0].width or c.height != cams[0].height for c in cams):
            raise ValueError("all frames of a batch have the same size")
        if out is None:
            out = np.zeros((n, cams[0].height, cams[0].width, 4), np.uint8) if n else np.zeros((0, 0, 0, 4), np.uint8)
        need = n * cams[0].width * cams[0].height * 4 if n else 0
        if out.dtype != np.uint8 or not out.flags["C_CONTIGUOUS"] or out.size < need:
            raise ValueError("out must be a C-contiguous uint8 buffer of nframes*width*height*4 bytes")
        st = A.RtbStats()
        A.check(self._lib.rtb_render_batch(self._h, arr, n, C.byref(o), out.ctypes.data,
                                           C.byref(st) if stats else None))
        return (out, st) if stats else out


def device_info():
    lib = A.load()
    A.check(lib.rtb_init(-1))
    sm, khz = C.c_int32(), C.c_int32()
    name = C.create_string_buffer(128)
    A.check(lib.rtb_device_info(C.byref(sm), C.byref(khz), name, 128))
    return {"sm_count": sm.value, "clock_khz": khz.value, "name": name.value.decode()}
