"""gen_golden_rays.py — golden vectors for NON-demo inputs, produced by the reference's OWN functions.

main.js hard-codes its scene, camera and recursion depth, so its framebuffer only pins the oracle on the demo
scene at segs = 8.  Its tracer, however, is a set of top-level functions (`intersectWorld`, `createSphere`,
`createMaterial`, `createTexture`, `checkerTexture`, `sampleTexture`), which become globals when the UNMODIFIED
module is evaluated under oracle/minijs.  This script builds a seeded synthetic scene with those constructors
(diffuse, mirror-like, glass with refract_index 1.5, checker-textured spheres inside a ground sphere and a sky
sphere) and calls the reference's `intersectWorld(segs, objects, org, dir)` for a few hundred rays at several
recursion depths - including rays that start inside glass (total internal reflection) - and stores the returned
colours (float64) with the inputs under tests/golden/mainjs_rays_synth.npz.  main.js is read from the
reference tree; nothing of it is copied.

Run in the build container (where /root/reference exists):  python oracle/gen_golden_rays.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import gen_golden  # noqa: E402
from oracle.minijs.interp import to_number  # noqa: E402

N_SPHERES, N_RAYS, SEED = 40, 360, 20261018
DEPTHS = (1, 2, 5, 8)


def make_scene(rng):
    """Rows: cx, cy, cz, r, colour(3), albedo(5), specular_exponent, refract_index, textured(0/1)."""
    rows = [[0, -500, 0, 500, 0.9, 0.8, 0.7, 0, .5, .8, 0, 0, 10, 1.0, 0],          # ground
            [0, 0, 0, 5000, 0.1, 0.1, 0.3, 1, 0, 0, 0, 0, 0, 1.0, 0]]               # sky (plain colour: no Math.random)
    kinds = rng.random(N_SPHERES - 2)
    for k in kinds:
        r = rng.uniform(0.3, 1.2)
        c = [rng.uniform(-6, 6), r + rng.uniform(0, 4), rng.uniform(-8, 4)]
        col = list(rng.uniform(0.2, 1.0, 3))
        if k < 0.35:
            rows.append(c + [r] + col + [0, .8, .3, 0, 0, 50, 1.0, 0])               # diffuse
        elif k < 0.6:
            rows.append(c + [r] + col + [0, .5, .5, rng.uniform(.2, .6), 0, 20, 1.0, 0])   # reflective
        elif k < 0.8:
            rows.append(c + [r] + col + [0, .4, .3, .1, .8, 50, 1.5, 0])             # glass: reflect + refract
        else:
            rows.append(c + [r] + [1, 1, 1] + [0, .8, .2, 0, 0, 20, 1.0, 1])         # checker-textured
    return np.array(rows, np.float64)


def make_rays(rng, scene):
    org, dirs = [], []
    for i in range(N_RAYS):
        if i % 6 == 5:                                   # start INSIDE a glass sphere: exits, TIR candidates
            g = scene[(scene[:, 11] > 0)]
            s = g[rng.integers(len(g))]
            o = s[:3] + rng.uniform(-0.3, 0.3, 3) * s[3]
        else:
            o = np.array([rng.uniform(-3, 3), rng.uniform(0.5, 4), rng.uniform(6, 12)])
        t = np.array([rng.uniform(-6, 6), rng.uniform(-1, 4), rng.uniform(-8, 2)])
        d = t - o
        d /= np.sqrt((d * d).sum())
        org.append(o)
        dirs.append(d)
    return np.array(org), np.array(dirs)


def js_num(x):
    return repr(float(x))


def main():
    rng = np.random.default_rng(SEED)
    scene = make_scene(rng)
    org, dirs = make_rays(rng, scene)
    it = gen_golden.run_main_js(64, 36, module_only=True)          # evaluates the unmodified module only
    lines = ["const tex = createTexture();", "checkerTexture(tex, 16, 8, [0,0,0], [1,1,1]);", "const objs = [];"]
    for s in scene:
        lines.append("objs.push(createSphere([%s,%s,%s], %s, createMaterial([%s,%s,%s], [%s,%s,%s,%s,%s], %s, %s)));" %
                     tuple(js_num(v) for v in s[:14]))
        if s[14]:
            lines.append("objs[objs.length-1].mtl.sampler = function (hit) { return sampleTexture(tex, hit.u, hit.v); };")
    lines.append("const out = [];")
    lines.append("const rays = [%s];" % ",".join("[%s]" % ",".join(js_num(v) for v in list(o) + list(d))
                                                    for o, d in zip(org, dirs)))
    lines.append("const depths = [%s];" % ",".join(str(d) for d in DEPTHS))
    lines.append("for (let k = 0; k < depths.length; k++) { for (let i = 0; i < rays.length; i++) { const r = rays[i]; "
                 "out.push(intersectWorld(depths[k], objs, [r[0],r[1],r[2]], [r[3],r[4],r[5]])); } }")
    it.run("\n".join(lines))
    out = it.globals.vars["out"]
    res = np.array([[to_number(c) for c in v.items] for v in out.items], np.float64).reshape(len(DEPTHS), N_RAYS, 3)
    path = os.path.join(ROOT, "tests", "golden", "mainjs_rays_synth.npz")
    np.savez_compressed(path, scene=scene, org=org, dir=dirs, depths=np.array(DEPTHS), rgb=res,
                        source=np.array("intersectWorld / createSphere / createMaterial / checkerTexture / sampleTexture of the "
                                        "unmodified main.js build 741, evaluated under oracle/minijs (gen_golden_rays.py)"))
    print("wrote", path, res.shape, "non-trivial colours:", int((res.max(axis=2) > 0).sum()))

    # sampleTexture (main.js:343-351: nearest texel, ceil(u*w) - 1 addressing, no wrap) and checkerTexture
    # (main.js:353-373) themselves: a 7 x 5 texture with distinct texels, (u, v) on a grid that includes 0, 1,
    # exact texel boundaries and values just beside them
    tw, th = 7, 5
    texels = rng.integers(0, 256, tw * th * 4)
    us = sorted(set([0.0, 1.0, 1e-12, 1 - 1e-12] + [i / tw for i in range(tw + 1)] + [i / tw + 1e-9 for i in range(tw)] +
                    list(rng.random(12))))
    vs = sorted(set([0.0, 1.0] + [i / th for i in range(th + 1)] + [i / th - 1e-9 for i in range(1, th + 1)] + list(rng.random(8))))
    snippet = ["const t2 = createTexture();", "t2.width = %d; t2.height = %d;" % (tw, th),
               "t2.texels = [%s];" % ",".join(str(int(v)) for v in texels), "const so = [];",
               "const us = [%s]; const vs = [%s];" % (",".join(js_num(u) for u in us), ",".join(js_num(v) for v in vs)),
               "for (let j = 0; j < vs.length; j++) { for (let i = 0; i < us.length; i++) { so.push(sampleTexture(t2, us[i], vs[j])); } }",
               "const ck = createTexture();", "checkerTexture(ck, 6, 4, [0,1,1], [1,0,1]);"]
    it.run("\n".join(snippet))
    so = it.globals.vars["so"]
    samp = np.array([[to_number(c) for c in v.items] for v in so.items], np.float64).reshape(len(vs), len(us), 3)
    ck = it.globals.vars["ck"]
    ck_tex = np.array([to_number(c) for c in ck.props["texels"].items], np.float64)
    path2 = os.path.join(ROOT, "tests", "golden", "mainjs_sampler_vectors.npz")
    np.savez_compressed(path2, tex_w=np.array(tw), tex_h=np.array(th), texels=texels.astype(np.uint8), us=np.array(us),
                        vs=np.array(vs), rgb=samp, checker_w=np.array(6), checker_h=np.array(4),
                        checker_c0=np.array([0.0, 1.0, 1.0]), checker_c1=np.array([1.0, 0.0, 1.0]), checker_texels=ck_tex,
                        source=np.array("sampleTexture / checkerTexture of the unmodified main.js build 741 under oracle/minijs"))
    print("wrote", path2, samp.shape, ck_tex.shape)


if __name__ == "__main__":
    main()
