"""html5-canvas-raytracer_b200 — B200-native replacement for the per-pixel hot path of
termuxinator/html5-canvas-raytracer's main.js, behind the reference-shaped host interface.

The directory name carries hyphens (it is the name the build contract fixes), so import it with
`importlib.import_module("html5-canvas-raytracer_b200")` or through the root-level alias `rtb200`.

Layout: csrc/ (CUDA kernels + C ABI, built to csrc/librtb200.so), scene.py (mirror of main.js's
constructors), scenes.py (BASELINE.json workloads), render.py (Renderer: the redraw() replacement),
multigpu.py (row-stripe / frame-batch sharding over torch.distributed + NCCL).

Importing the package never touches CUDA; constructing a Renderer does, and raises if the CUDA
library is not built or no GPU is visible (there is no CPU render path).
"""
from . import _abi
from ._abi import (RTB_CAMERA_BASIS, RTB_CAMERA_REFERENCE, RTB_KERNEL_AUTO, RTB_KERNEL_PIXEL,
                   RTB_KERNEL_WAVEFRONT, RTB_PRECISION_FP32, RTB_PRECISION_FP64, RTB_STARS_HASH,
                   RTB_STARS_OFF, RTB_STARS_STREAM, RtbError)
from .scene import (Camera, Material, SceneDesc, Sphere, Texture, checkerSampler, checkerTexture,
                    colorSampler, createMaterial, createSphere, createTexture, loadTexture, make_opts,
                    sortObjects, starSampler, textureFromRGBA, textureSampler)
from . import scenes
from .render import Renderer, device_info
