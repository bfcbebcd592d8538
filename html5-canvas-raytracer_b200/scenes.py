"""The workloads of BASELINE.json `configs`, built through the reference-shaped constructors of
scene.py: the 14-sphere demo scene of main.js:107-163, the seeded 1024-sphere synthetic scene
 and the orbit cameras of C5.
"""
from __future__ import annotations

import math
import os
from typing import List, Tuple

import numpy as np

from . import _abi as A
from .scene import (Camera, Sphere, checkerSampler, checkerTexture, createMaterial, createSphere,
                    createTexture, sortObjects, starSampler, textureFromRGBA, textureSampler)

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "earth_mars_rgba8.npz")


def load_planet_textures():
    """earth.png / mars.png of the reference, pre-decoded to RGBA8 (256x128 each) by
    tools/make_assets.py — the bytes `getImageData` yields at main.js:388-391."""
    z = np.load(_ASSETS)
    return textureFromRGBA(z["earth"]), textureFromRGBA(z["mars"])


def demo_objects(origin=(0.0, 1.5, 10.0), textures=None) -> List[Sphere]:
    """main.js:107-163: the scene literal, the five sampler overrides, then the sort."""
    S, M = createSphere, createMaterial
    objects = [
        S([0.0, -500.0, 0.0], 500, M([1.0, 1.0, 1.0], [0.0, 0.5, 0.8, 0.0, 0.0], 10, 1.0)),   # home
        S([0.0, 0.0, 0.0], 5000, M([0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0, 0.0], 0, 1.0)),      # skybox
        S([50.0, 20.0, -100.0], 4.0, M([1.0, 1.0, 1.0], [1.0, 0.0, 0.0, 0.0, 0.0], 0, 1.0)),  # earth
        S([-50.0, 20.0, -100.0], 2.0, M([1.0, 1.0, 1.0], [1.0, 0.0, 0.0, 0.0, 0.0], 0, 1.0)), # mars
        S([0.0, 0.25, 3.0], 0.25, M([1.0, 1.0, 1.0], [0.0, 1.0, 0.1, 0.0, 0.0], 10, 1.0)),    # matte
        S([-2.5, 0.5, 3.0], 0.5, M([1.0, 1.0, 1.0], [0.0, 0.5, 0.2, 0.0, 0.8], 50, 1.5)),     # glass
        S([1.5, 2.5, 0.0], 0.5, M([0.5, 0.5, 0.5], [0.0, 0.5, 1.0, 0.4, 0.0], 20, 1.0)),      # chrome
        S([1.0, 0.25, 3.0], 0.25, M([0.5, 0.5, 0.5], [0.0, 0.5, 1.0, 0.5, 0.0], 20, 1.0)),    # chrome
        S([2.5, 0.5, 3.0], 0.5, M([0.5, 0.5, 0.5], [0.0, 0.4, 0.5, 0.2, 0.8], 20, 1.0)),      # bubble
        S([0.0, 2.5, -2.0], 0.5, M([1.0, 1.0, 1.0], [0.0, 0.1, 0.5, 0.6, 0.0], 500, 1.0)),    # mirror
        S([-1.5, 2.5, 0.0], 0.5, M([1.0, 1.0, 1.0], [0.0, 0.8, 0.2, 0.1, 0.0], 50, 1.0)),     # metal
        S([-1.5, 1.0, 0.0], 1.0, M([1.0, 0.0, 0.0], [0.0, 0.8, 0.3, 0.5, 0.0], 50, 1.0)),     # ornament
        S([1.5, 1.0, 0.0], 1.0, M([0.0, 1.0, 0.0], [0.0, 0.8, 0.3, 0.5, 0.0], 50, 1.0)),      # ornament
        S([0.0, 1.0, -2.0], 1.0, M([0.0, 0.0, 1.0], [0.0, 0.8, 0.3, 0.5, 0.0], 50, 1.0)),     # ornament
    ]
    objects[0].mtl.sampler = checkerSampler(5000, 2500, (1, 1, 0), (1, 0, 1))   # main.js:126-133
    objects[1].mtl.sampler = starSampler(0.001, 1000)                          # main.js:135-139
    earth, mars = textures if textures is not None else load_planet_textures()
    objects[2].mtl.sampler = textureSampler(earth)                             # main.js:141-145
    objects[3].mtl.sampler = textureSampler(mars)                              # main.js:147-151
    unit2 = createTexture()
    checkerTexture(unit2, 16, 8, [0, 0, 0], [1, 1, 1])                         # main.js:153-154
    objects[4].mtl.sampler = textureSampler(unit2)                             # main.js:155-157
    sortObjects(objects, list(origin))                                         # main.js:159-163
    return objects


def demo_camera(width: int, height: int) -> Camera:
    """main.js:85-105: lookAt([0,1.5,10],[0,1.5,0],[0,1,0]), FOV 60 deg, canvas-sized projection."""
    return Camera(width, height).lookAt([0, 1.5, 10], [0, 1.5, 0], [0, 1, 0])


def synth_objects(n_total: int = 1024, seed: int = 1024, origin=(0.0, 1.5, 10.0), textures=None) -> List[Sphere]:
    """home + skybox (as in the demo; the sky sampler stays the star
    sampler, black under RTB_STARS_OFF) + (n_total - 2) random spheres from
    numpy.random.default_rng(seed), then the reference sort."""
    if n_total < 2:
        raise ValueError("n_total counts the home and skybox spheres")
    S, M = createSphere, createMaterial
    rng = np.random.default_rng(seed)
    earth, mars = textures if textures is not None else load_planet_textures()
    objects = [
        S([0.0, -500.0, 0.0], 500, M([1.0, 1.0, 1.0], [0.0, 0.5, 0.8, 0.0, 0.0], 10, 1.0)),
        S([0.0, 0.0, 0.0], 5000, M([0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0, 0.0], 0, 1.0)),
    ]
    objects[0].mtl.sampler = checkerSampler(5000, 2500, (1, 1, 0), (1, 0, 1))
    objects[1].mtl.sampler = starSampler(0.001, 1000)
    for _ in range(n_total - 2):
        r = float(rng.uniform(0.25, 1.0))
        x = float(rng.uniform(-30.0, 30.0))
        z = float(rng.uniform(-70.0, 4.0))
        floor_height = math.sqrt(500.0 ** 2 - x * x - z * z) - 500.0
        y = floor_height + r + float(rng.uniform(0.0, 6.0))
        kind = float(rng.uniform())
        color = [float(c) for c in rng.uniform(0.2, 1.0, 3)]
        refl = float(rng.uniform(0.2, 0.6))
        coin = float(rng.uniform())
        if kind < 0.40:      # diffuse
            m = M(color, [0.0, 0.8, 0.3, 0.0, 0.0], 50, 1.0)
        elif kind < 0.70:    # reflective
            m = M(color, [0.0, 0.5, 0.5, refl, 0.0], 20, 1.0)
        elif kind < 0.80:    # glass
            m = M(color, [0.0, 0.4, 0.3, 0.1, 0.8], 50, 1.5)
        else:                # textured
            m = M([1.0, 1.0, 1.0], [0.0, 0.8, 0.2, 0.0, 0.0], 20, 1.0)
            m.sampler = textureSampler(earth if coin < 0.5 else mars)
        objects.append(S([x, y, z], r, m))
    sortObjects(objects, list(origin))
    return objects


def orbit_cameras(width: int, height: int, nframes: int = 360) -> List[Camera]:
    """org = (10 sin t, 1.5, 10 cos t), t = 2 pi f / nframes, looking at
    (0,1.5,0).  Needs RTB_CAMERA_BASIS: the reference formula (main.js:187-191) does not rotate.
    Frame 0 equals the reference camera exactly (sin 0 = 0, cos 0 = 1)."""
    cams = []
    for f in range(nframes):
        th = 2 * math.pi * f / nframes
        c = Camera(width, height, camera_mode=A.RTB_CAMERA_BASIS)
        c.lookAt([10 * math.sin(th), 1.5, 10 * math.cos(th)], [0, 1.5, 0], [0, 1, 0])
        cams.append(c)
    return cams


def orbit_frames(width: int, height: int, nframes: int = 360, textures=None):
    """Config C5 with the host-side scene preparation the reference would do for each frame: the
    objects re-sorted with main.js's comparator (159-163) for THAT frame's camera origin — array order
    decides closest-hit ties, the self-shadow index and the occluder order.  Yields (objects, camera)."""
    if textures is None:
        textures = load_planet_textures()
    for cam in orbit_cameras(width, height, nframes):
        yield demo_objects(origin=tuple(cam.origin), textures=textures), cam


def config_workload(name: str) -> Tuple[str, int, int, int, int]:
    """(scene, width, height, segs, spp_grid) of the named BASELINE.json config."""
    table = {
        "c1_demo_720p": ("demo", 1280, 720, 8, 1),
        "c2_demo_1080p_segs5": ("demo", 1920, 1080, 5, 1),
        "c3_synth1024_4k": ("synth1024", 3840, 2160, 8, 1),
        "c4_demo_4k_16spp": ("demo", 3840, 2160, 8, 4),
        "c5_orbit_1080p": ("demo", 1920, 1080, 8, 1),
    }
    return table[name]
