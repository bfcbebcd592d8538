"""Where does the end-to-end time of one C3 frame go?  (scene marshalling, create, render, D2H, destroy)"""
import ctypes as C
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

rt = importlib.import_module("html5-canvas-raytracer_b200")
A = rt._abi
lib = A.load()
A.check(lib.rtb_init(0))
objs = rt.scenes.synth_objects(1024, 1024)
cam = rt.scenes.demo_camera(3840, 2160).to_c()
host = torch.empty((2160, 3840, 4), dtype=torch.uint8).pin_memory()
opts = rt.make_opts(segs=8, kernel=rt.RTB_KERNEL_WAVEFRONT)


def t(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1000 / n


desc = rt.SceneDesc(objs)
print(f"python SceneDesc marshalling: {t(lambda: rt.SceneDesc(objs)):.3f} ms")
h = C.c_void_p()


def create_destroy():
    A.check(lib.rtb_scene_create(C.byref(desc.c), C.byref(h)))
    lib.rtb_scene_destroy(h)


print(f"rtb_scene_create + destroy: {t(create_destroy):.3f} ms")
A.check(lib.rtb_scene_create(C.byref(desc.c), C.byref(h)))
st = A.RtbStats()
print(f"rtb_render (host, pinned): {t(lambda: A.check(lib.rtb_render(h, C.byref(cam), C.byref(opts), host.data_ptr(), C.byref(st)))):.3f} ms"
      f"  [kernel {st.kernel_ms:.3f} d2h {st.d2h_ms:.3f}]")
print(f"rtb_render (host, pinned, no stats): {t(lambda: A.check(lib.rtb_render(h, C.byref(cam), C.byref(opts), host.data_ptr(), None))):.3f} ms")
dev = torch.empty((2160, 3840, 4), dtype=torch.uint8, device="cuda")
print(f"rtb_render_device: {t(lambda: A.check(lib.rtb_render_device(h, C.byref(cam), C.byref(opts), dev.data_ptr(), None, None))):.3f} ms")
pageable = torch.empty((2160, 3840, 4), dtype=torch.uint8)
print(f"rtb_render (host, pageable): {t(lambda: A.check(lib.rtb_render(h, C.byref(cam), C.byref(opts), pageable.data_ptr(), None))):.3f} ms")
lib.rtb_scene_destroy(h)


def full():
    A.check(lib.rtb_scene_create(C.byref(desc.c), C.byref(h)))
    A.check(lib.rtb_render(h, C.byref(cam), C.byref(opts), host.data_ptr(), None))
    lib.rtb_scene_destroy(h)


print(f"create + render + destroy: {t(full):.3f} ms")
