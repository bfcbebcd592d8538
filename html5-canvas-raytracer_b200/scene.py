"""Host-side mirror of the reference's scene/texture interface (main.js), same names and argument
meaning, so that code written against main.js's constructors builds the scene handed to the CUDA
library.  Nothing here renders: this module only produces the flat `RtbSceneDesc` / `RtbCamera`
the C ABI takes.

Reference items mirrored (all in /root/reference/main.js):
  vec3 helpers 20-75 (host-only subset), lookAt 92-100, projection 102-105, scene literal 107-124,
  sampler overrides 126-157, sort 159-163, createTexture 339-341, checkerTexture 353-373,
  loadTexture 375-395 (PNG decode instead of Image+canvas), createMaterial 397-406,
  createSphere 408-418.
JS closures cannot cross a C ABI, so `mtl.sampler` takes one of the sampler *descriptors* below
instead of a function; the default is the reference's default `hit.m.color` (main.js:404).
"""
from __future__ import annotations

import ctypes as C
import functools
import math
from typing import List, Optional, Sequence

import numpy as np

from . import _abi as A
from . import png as _png

# ---- vec3 helpers (main.js:20-75), host-side subset ------------------------------------------------


def init3D(x, y, z):
    return [x, y, z]


def dot3D(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def cross3D(a, b):  # main.js:53-55
    return [a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]]


def length3D(v):
    return math.sqrt(dot3D(v, v))


def scale3D(v, t):
    return [v[0] * t, v[1] * t, v[2] * t]


def normal3D(v):  # main.js:62-66 — multiply by the reciprocal, return v itself when l == 0
    l = length3D(v)
    if l == 0:
        return v
    return scale3D(v, 1 / l)


def between3D(a, b):
    return [b[0] - a[0], b[1] - a[1], b[2] - a[2]]


def distance3D(a, b):
    return length3D(between3D(a, b))


# ---- sampler descriptors (what `mtl.sampler = function (hit) {...}` becomes) ------------------------


class Sampler:
    kind = A.RTB_SAMPLER_COLOR


class colorSampler(Sampler):
    """default sampler `return hit.m.color` (main.js:404)"""
    kind = A.RTB_SAMPLER_COLOR


class textureSampler(Sampler):
    """`return sampleTexture(texture, hit.u, hit.v)` (main.js:143-145, 149-151, 155-157)"""
    kind = A.RTB_SAMPLER_TEXTURE

    def __init__(self, texture: "Texture"):
        self.texture = texture


class checkerSampler(Sampler):
    """home-sphere procedural checker on hit.n (main.js:126-133)"""
    kind = A.RTB_SAMPLER_CHECKER

    def __init__(self, freq_u=5000, freq_v=2500, color0=(1, 1, 0), color1=(1, 0, 1)):
        self.freq = (float(freq_u), float(freq_v))
        self.colors = (tuple(map(float, color0)), tuple(map(float, color1)))


class starSampler(Sampler):
    """skybox night stars driven by Math.random() (main.js:135-139)"""
    kind = A.RTB_SAMPLER_STARS

    def __init__(self, threshold=0.001, gain=1000):
        self.threshold = float(threshold)
        self.gain = float(gain)


# ---- reference constructors --------------------------------------------------------------------------


class Texture:  # createTexture() main.js:339-341
    def __init__(self):
        self.width = 0
        self.height = 0
        self.texels = np.zeros(0, np.uint8)  # RGBA8
        self.loaded = False


def createTexture() -> Texture:
    return Texture()


def checkerTexture(texture: Texture, width: int, height: int, color1, color2) -> None:
    """main.js:353-373.  The reference writes `255 * color` doubles into a plain Array and
    sampleTexture divides by 255 again; colours in {0,1} make that an exact byte."""
    t = np.empty((height, width, 4), np.float64)
    for y in range(height):
        for x in range(width):
            c = color2 if ((x & 1) ^ (y & 1)) else color1
            t[y, x, 0] = 255 * c[0]
            t[y, x, 1] = 255 * c[1]
            t[y, x, 2] = 255 * c[2]
            t[y, x, 3] = 255
    if not np.all(t == np.round(t)) or t.min() < 0 or t.max() > 255:
        raise ValueError("checkerTexture colours must map to exact bytes (RGBA8 texel storage)")
    texture.texels = t.astype(np.uint8).reshape(-1)
    texture.width = width
    texture.height = height
    texture.loaded = True


def loadTexture(texture: Texture, src) -> None:
    """main.js:375-395 without the browser: decode the PNG (path or bytes) straight to the RGBA8
    bytes `getImageData` would return (the reference PNGs are opaque palette images without
    gamma/ICC chunks, so no colour management or premultiplication applies)."""
    w, h, rgba = _png.decode_rgba8(src)
    texture.width, texture.height = w, h
    texture.texels = rgba.reshape(-1)
    texture.loaded = True


def textureFromRGBA(rgba: np.ndarray) -> Texture:
    """Pre-decoded H x W x 4 uint8 pixels (what `getImageData(...).data` holds, main.js:388-391)."""
    t = Texture()
    rgba = np.ascontiguousarray(rgba, np.uint8)
    t.height, t.width = int(rgba.shape[0]), int(rgba.shape[1])
    t.texels = rgba.reshape(-1)
    t.loaded = True
    return t


class Material:  # createMaterial main.js:397-406
    def __init__(self, color, albedo, se, ri):
        self.color = [float(c) for c in color]
        self.albedo = [float(a) for a in albedo]  # ambient,diffuse,specular,reflect,refract
        self.specular_exponent = float(se)
        self.refract_index = float(ri)
        self.sampler: Sampler = colorSampler()


def createMaterial(color, albedo, se, ri) -> Material:
    if len(color) != 3 or len(albedo) != 5:
        raise ValueError("createMaterial(color[3], albedo[5], se, ri)")
    return Material(color, albedo, se, ri)


class Sphere:  # createSphere main.js:408-418
    def __init__(self, o, r, m: Material):
        r = float(r)
        self.origin = [float(c) for c in o]
        self.r2 = r * r
        self.mtl = m
        self.surface_area = 4 * math.pi * self.r2


def createSphere(o, r, m: Material) -> Sphere:
    return Sphere(o, r, m)


def sortObjects(objects: List[Sphere], origin) -> None:
    """objects.sort by surface_area / distance3D(origin, centre), ascending (main.js:159-163).
    Array.prototype.sort is stable (V8 >= 7.0) and so is list.sort; ties keep insertion order."""

    def cmp(a, b):
        aa = a.surface_area / distance3D(origin, a.origin)
        bb = b.surface_area / distance3D(origin, b.origin)
        d = aa - bb
        return -1 if d < 0 else (1 if d > 0 else 0)

    objects.sort(key=functools.cmp_to_key(cmp))


# ---- camera (closure state of main(), main.js:85-105) ----------------------------------------------


class Camera:
    def __init__(self, width: int, height: int, camera_mode: int = A.RTB_CAMERA_REFERENCE):
        self.width, self.height = int(width), int(height)
        self.origin = init3D(0, 1.5, 10)   # main.js:85-88
        self.axisX = init3D(-1, 0, 0)
        self.axisY = init3D(0, 1, 0)
        self.axisZ = init3D(0, 0, -1)
        self.camera_mode = camera_mode
        projA = 60 * math.pi / 180          # main.js:102-105
        self.projW = self.width / 2
        self.projH = self.height / 2
        self.projD = self.projW / math.tan(projA / 2)

    def lookAt(self, org, tgt, up) -> "Camera":  # main.js:92-100
        self.origin = [float(c) for c in org]
        self.axisZ = between3D(org, tgt)
        self.axisX = cross3D(up, self.axisZ)
        self.axisY = cross3D(self.axisZ, self.axisX)
        self.axisX = normal3D(self.axisX)
        self.axisY = normal3D(self.axisY)
        self.axisZ = normal3D(self.axisZ)
        return self

    def to_c(self) -> A.RtbCamera:
        c = A.RtbCamera()
        c.width, c.height = self.width, self.height
        c.origin = A.D3(*self.origin)
        c.axisX = A.D3(*self.axisX)
        c.axisY = A.D3(*self.axisY)
        c.axisZ = A.D3(*self.axisZ)
        c.projW, c.projH, c.projD = self.projW, self.projH, self.projD
        c.camera_mode = self.camera_mode
        return c


# ---- flattening to the C ABI --------------------------------------------------------------------------

DEFAULT_LIGHTS = [[5.0, 10.0, 5.0], [5.0, 10.0, 0.0]]  # main.js:283
DEFAULT_LIGHT_INTENSITY = 50.0                           # main.js:284
DEFAULT_EPS = 0.001                                      # main.js:430-436, 445


class SceneDesc:
    """Owns the ctypes arrays behind an RtbSceneDesc (keeps them alive while C reads them)."""

    def __init__(self, objects: Sequence[Sphere], lights=None, light_intensity=DEFAULT_LIGHT_INTENSITY,
                 eps=DEFAULT_EPS):
        lights = DEFAULT_LIGHTS if lights is None else lights
        if len(lights) > A.RTB_MAX_LIGHTS:
            raise ValueError(f"at most {A.RTB_MAX_LIGHTS} lights")
        self.objects = list(objects)
        self.textures: List[Texture] = []
        n = len(self.objects)
        self._spheres = (A.RtbSphere * max(n, 1))()
        for i, o in enumerate(self.objects):
            s = self._spheres[i]
            s.origin = A.D3(*o.origin)
            s.r2 = o.r2
            m = s.mtl
            m.color = A.D3(*o.mtl.color)
            m.albedo = (C.c_double * 5)(*o.mtl.albedo)
            m.specular_exponent = o.mtl.specular_exponent
            m.refract_index = o.mtl.refract_index
            sm = o.mtl.sampler
            m.sampler_kind = sm.kind
            m.texture_unit = 0
            # reference defaults so that a zeroed field never means "frequency 0"
            m.checker_freq = (C.c_double * 2)(5000.0, 2500.0)
            m.checker_color[0] = A.D3(1, 1, 0)
            m.checker_color[1] = A.D3(1, 0, 1)
            m.star_threshold, m.star_gain = 0.001, 1000.0
            if isinstance(sm, textureSampler):
                if not sm.texture.loaded:
                    raise ValueError("texture not loaded (main.js:167-178 would poll forever)")
                if sm.texture not in self.textures:
                    self.textures.append(sm.texture)
                m.texture_unit = self.textures.index(sm.texture)
            elif isinstance(sm, checkerSampler):
                m.checker_freq = (C.c_double * 2)(*sm.freq)
                m.checker_color[0] = A.D3(*sm.colors[0])
                m.checker_color[1] = A.D3(*sm.colors[1])
            elif isinstance(sm, starSampler):
                m.star_threshold, m.star_gain = sm.threshold, sm.gain
        self._tex_bufs = [np.ascontiguousarray(t.texels, np.uint8) for t in self.textures]
        self._textures = (A.RtbTexture * max(len(self.textures), 1))()
        for i, t in enumerate(self.textures):
            if self._tex_bufs[i].size != t.width * t.height * 4:
                raise ValueError("texels must be width*height*4 RGBA8 bytes")
            self._textures[i].width = t.width
            self._textures[i].height = t.height
            self._textures[i].texels = self._tex_bufs[i].ctypes.data_as(C.POINTER(C.c_uint8))
        d = A.RtbSceneDesc()
        d.abi_version = A.RTB_ABI_VERSION
        d.n_spheres = n
        d.spheres = C.cast(self._spheres, C.POINTER(A.RtbSphere))
        d.n_textures = len(self.textures)
        d.textures = C.cast(self._textures, C.POINTER(A.RtbTexture))
        d.n_lights = len(lights)
        for i, l in enumerate(lights):
            d.lights[i] = A.D3(*[float(c) for c in l])
        d.light_intensity = float(light_intensity)
        d.eps = float(eps)
        self.c = d


def make_opts(segs=8, spp_grid=1, precision=A.RTB_PRECISION_FP32, kernel=A.RTB_KERNEL_AUTO,
              star_mode=A.RTB_STARS_OFF, y0=0, y1=0, stripe_rows=0, stripe_rank=0, stripe_world=0,
              star_seed=0, out_packed=0, star_stream=None) -> A.RtbRenderOpts:
    o = A.RtbRenderOpts()
    o.segs, o.spp_grid, o.precision, o.kernel, o.star_mode = segs, spp_grid, precision, kernel, star_mode
    o.y0, o.y1 = y0, y1
    o.stripe_rows, o.stripe_rank, o.stripe_world = stripe_rows, stripe_rank, stripe_world
    o.star_seed = star_seed
    o.out_packed = out_packed
    if star_stream is not None:                        # RTB_STARS_STREAM: Math.random() values, call order
        arr = np.ascontiguousarray(star_stream, np.float64)
        o._keep_stream = arr                           # the C side reads the caller's memory during the call
        o.star_stream = arr.ctypes.data_as(C.POINTER(C.c_double))
        o.star_stream_len = arr.size
    return o
