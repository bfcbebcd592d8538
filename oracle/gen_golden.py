"""gen_golden.py — execute the UNMODIFIED /root/reference/main.js under oracle/minijs with DOM shims
and store its framebuffer (`colorbuf.data`) as golden vectors under tests/golden/.

Run in the build container (where /root/reference exists):
    python oracle/gen_golden.py [WxH ...]          default: 64x36 96x54

What the shims provide is what SURVEY.md §8c specifies for the Node `vm` harness: `document`
(createElement('pre'|'canvas'), getElementById, body{appendChild, clientWidth, clientHeight, onload}),
`window`, `Image` (setting `src` queues `onload({target:{width,height}})`), a fake 2-D context
(createImageData -> a real clamped byte array; drawImage remembers the image; getImageData returns
the decoded RGBA8 of the PNG; putImageData captures; fillText captures the `build #741 (Nms)` string),
a synchronous FIFO for setTimeout, and `Math.random = () => 0.5` (policy P0: sky black).
main.js itself is read from the reference tree and is not modified or copied.
"""
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.minijs.interp import (UNDEF, Interpreter, JSObject, NativeFunction, Uint8ClampedArray,  # noqa: E402
                                  call, to_number, to_string)

REFERENCE = "/root/reference"


class Host(JSObject):
    """JSObject with optional setter hooks."""

    def __init__(self, props=None, setters=None):
        super().__init__(props)
        self.setters = setters or {}

    def js_set(self, k, v):
        self.props[k] = v
        h = self.setters.get(k)
        if h:
            h(v)


def nf(fn, name="native"):
    return NativeFunction(lambda this, args: fn(this, args), name)


def star_stream(n):
    """Deterministic stand-in for the engine's Math.random() stream used by the `stars` golden frame:
    value k = ((k * 2654435761) mod 2^32) / 2^32 * 0.004, i.e. uniform in [0, 0.004) so that a quarter of
    the draws fall under main.js:137's 0.001 threshold and light a star."""
    k = np.arange(n, dtype=np.uint64)
    return ((k * np.uint64(2654435761)) % np.uint64(4294967296)).astype(np.float64) / 4294967296.0 * 0.004


def run_main_js(width, height, main_js_path=os.path.join(REFERENCE, "main.js"), verbose=False, stars=False,
                module_only=False):
    """module_only: evaluate the module (its top-level function declarations become globals of the interpreter)
    and return the Interpreter WITHOUT firing body.onload - for harness-driven calls of the reference's own
    functions (gen_golden_rays.py)."""
    png = importlib.import_module("html5-canvas-raytracer_b200.png")
    draws = [0]

    def rnd():
        if not stars:
            return 0.5
        k = draws[0]
        draws[0] += 1
        return ((k * 2654435761) % 4294967296) / 4294967296.0 * 0.004
    it = Interpreter(random_fn=rnd)
    captured = {"put": 0, "text": None, "imagedata": None}

    def make_context(canvas):
        ctx = Host({"imageSmoothingEnabled": True})

        def create_image_data(this, args):
            w, h = int(to_number(args[0])), int(to_number(args[1]))
            d = JSObject({"data": Uint8ClampedArray(w * h * 4), "width": float(w), "height": float(h)})
            return d

        def put_image_data(this, args):
            captured["put"] += 1
            captured["imagedata"] = args[0]
            return UNDEF

        def draw_image(this, args):
            ctx.props["_image"] = args[0]
            return UNDEF

        def get_image_data(this, args):
            img = ctx.props["_image"]
            rgba = img.props["_rgba"]
            return JSObject({"data": Uint8ClampedArray(rgba.tobytes()), "width": float(rgba.shape[1]),
                             "height": float(rgba.shape[0])})

        def fill_text(this, args):
            captured["text"] = to_string(args[0])
            return UNDEF

        ctx.props.update({
            "createImageData": nf(create_image_data), "putImageData": nf(put_image_data),
            "drawImage": nf(draw_image), "getImageData": nf(get_image_data), "fillText": nf(fill_text)})
        return ctx

    def make_canvas():
        c = Host({"width": 300.0, "height": 150.0})
        ctx = make_context(c)
        c.props["getContext"] = nf(lambda this, args: ctx)
        return c

    main_canvas = make_canvas()
    body = Host({"clientWidth": float(width), "clientHeight": float(height), "onload": None,
                 "appendChild": nf(lambda this, args: UNDEF)})

    def create_element(this, args):
        tag = to_string(args[0])
        if tag == "canvas":
            return make_canvas()
        return Host({"innerHTML": ""})

    document = Host({
        "body": body,
        "createElement": nf(create_element),
        "getElementById": nf(lambda this, args: main_canvas if to_string(args[0]) == "canvasID" else None),
    })

    def image_ctor(args):
        img = Host({"onload": None})

        def set_src(v):
            path = os.path.join(REFERENCE, os.path.basename(to_string(v)))
            w, h, rgba = png.decode_rgba8(path)
            img.props["width"], img.props["height"] = float(w), float(h)
            img.props["_rgba"] = rgba
            # the browser fires onload asynchronously: queue it (FIFO) with e.target = the image
            it.timers.append((nf(lambda this, a: call(img.props["onload"], img, [JSObject({"target": img})])), []))
        img.setters["src"] = set_src
        return img

    g = it.globals.vars
    g["document"] = document
    g["window"] = Host({"onerror": None})
    g["Image"] = NativeFunction(lambda this, args: UNDEF, "Image", construct=image_ctor)

    src = open(main_js_path).read()
    t0 = time.time()
    it.run(src)                                     # module evaluation: IIFE hooks console.log, sets body.onload
    if module_only:
        return it
    onload = body.props["onload"]
    call(onload, body, [])                          # document.body.onload = main  (main.js:17)
    n = it.drain_timers()
    dt = time.time() - t0
    data = captured["imagedata"].props["data"].buf
    rgba = np.frombuffer(bytes(data), np.uint8).reshape(height, width, 4).copy()
    if verbose:
        print(f"main.js {width}x{height}: {n} timer callbacks, {captured['put']} putImageData calls, "
              f"overlay {captured['text']!r}, {dt:.1f} s")
    assert captured["put"] == height, "spanish() must hand back one row per tick (main.js:200)"
    if stars:
        return rgba, captured["text"], draws[0]
    return rgba, captured["text"]


def main():
    sizes = [a for a in sys.argv[1:] if "x" in a] or ["64x36", "96x54"]
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if "--stars" in sys.argv:
        for s in sizes:
            w, h = (int(v) for v in s.split("x"))
            rgba, text, n = run_main_js(w, h, verbose=True, stars=True)
            np.savez_compressed(os.path.join(out_dir, f"mainjs_demo_stars_{w}x{h}.npz"), rgba=rgba, n_draws=np.array(n),
                                source=np.array("unmodified main.js build 741 under oracle/minijs, Math.random = "
                                                "gen_golden.star_stream (k*2654435761 mod 2^32)/2^32*0.004"))
            print("wrote", f"tests/golden/mainjs_demo_stars_{w}x{h}.npz", rgba.shape, "draws", n)
        return
    for s in sizes:
        w, h = (int(v) for v in s.split("x"))
        rgba, text = run_main_js(w, h, verbose=True)
        np.savez_compressed(os.path.join(out_dir, f"mainjs_demo_{w}x{h}.npz"), rgba=rgba,
                            overlay=np.array(text or ""), source=np.array("unmodified main.js build 741 under oracle/minijs, Math.random = () => 0.5"))
        print("wrote", f"tests/golden/mainjs_demo_{w}x{h}.npz", rgba.shape)


if __name__ == "__main__":
    main()
