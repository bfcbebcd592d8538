"""Decode the reference's earth.png / mars.png ONCE (in the build container, where /root/reference
exists) into the RGBA8 bytes `getImageData` returns (main.js:388-391) and store them as a small
compressed asset.  /root/reference does not exist on the GPU box, so nothing reads it at run time.
Cross-checks the package's own PNG decoder against PIL."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
png = importlib.import_module("html5-canvas-raytracer_b200.png")

out = {}
for name in ("earth", "mars"):
    w, h, rgba = png.decode_rgba8(f"/root/reference/{name}.png")
    try:
        from PIL import Image
        ref = np.asarray(Image.open(f"/root/reference/{name}.png").convert("RGBA"))
        assert ref.shape == rgba.shape and np.array_equal(ref, rgba), "decoder disagrees with PIL"
    except ImportError:
        pass
    assert (w, h) == (256, 128) and rgba[..., 3].min() == 255
    out[name] = rgba
np.savez_compressed(os.path.join(ROOT, "html5-canvas-raytracer_b200", "assets", "earth_mars_rgba8.npz"), **out)
print("ok", {k: v.shape for k, v in out.items()})
