"""The safety margins of the fp32 sweep tables, checked on the CPU against float64.

The wavefront kernels decide "candidate or not" for every (ray, sphere) pair with a three-FMA (shared origin) or
eight-FMA (general ray) fp32 test against packed tables built on the host (wf_make_table / wf_make_geom_table,
csrc/rtb_internal.h); only candidates reach the exact re-test.  A true hit that the fast test does not flag would be
lost silently, so the tables carry margins: "fp32 rounding can only ADD candidates".  This test takes the REAL
tables from the library (rtb_debug_sweep_tables, host only - no GPU needed), models the kernels' fp32 arithmetic in
numpy (each FMA = one rounding of the float64-exact product-sum, the order of csrc/wavefront.cu sweep_primary /
sweep_general / ray_setup) and checks millions of pairs - rays aimed at sphere limbs included - against the
float64 geometry of main.js:420-428."""
import ctypes as C

import numpy as np
import pytest

F32, F64 = np.float32, np.float64


def fma32(a, b, c):
    """fp32 fused multiply-add: the product of two floats is exact in float64; one rounding to float32."""
    return (a.astype(F64) * b.astype(F64) + c.astype(F64)).astype(F32)


def tables(rt, lib, desc, origin):
    n = desc.c.n_spheres
    npad = lib.rtb_debug_table_pad(n)
    tab = np.zeros(npad * 4, F32)
    exact = np.zeros(npad * 4, F32)
    gen = np.zeros(npad * 4, F32)
    rc = lib.rtb_debug_sweep_tables(C.byref(desc.c), rt._abi.D3(*origin), tab.ctypes.data, exact.ctypes.data,
                                    gen.ctypes.data)
    assert rc == 0

    def unpack(t):              # per pair: x0 x1 y0 y1 z0 z1 w0 w1 -> [npad, 4] = x y z w per sphere
        p = t.reshape(-1, 8)
        return np.stack([p[:, 0:2].ravel(), p[:, 2:4].ravel(), p[:, 4:6].ravel(), p[:, 6:8].ravel()], axis=1)
    return unpack(tab), exact.reshape(-1, 4), unpack(gen), n


def spheres_of(desc):
    n = desc.c.n_spheres
    c = np.array([list(desc.c.spheres[i].origin) for i in range(n)], F64)
    r2 = np.array([desc.c.spheres[i].r2 for i in range(n)], F64)
    return c, r2


def limb_rays(rng, org, c, r2, k):
    """Directions from `org` aimed just inside / outside the silhouette of random spheres (grazing rays)."""
    idx = rng.integers(0, len(c), k)
    L = c[idx] - org
    dist = np.sqrt((L * L).sum(1))
    ok = dist > np.sqrt(r2[idx]) * 1.01                    # origin outside the sphere
    L, dist, r = L[ok], dist[ok], np.sqrt(r2[idx][ok])
    w = L / dist[:, None]
    t = rng.normal(size=L.shape)
    t -= (t * w).sum(1)[:, None] * w
    t /= np.sqrt((t * t).sum(1))[:, None]
    ang = np.arcsin(np.minimum(r / dist, 1.0)) * (1 + rng.choice([-1, 1], len(r)) * 10.0 ** rng.uniform(-7, -2, len(r)))
    d = np.cos(ang)[:, None] * w + np.sin(ang)[:, None] * t
    return d / np.sqrt((d * d).sum(1))[:, None]


def line_hits(org, d, c, r2):
    """float64: does the LINE through org along d pass within r of C (main.js:422-425), and tca.  d is the
    fp32 direction the kernel holds, taken as exact and re-normalised in float64: d2 = |L|^2 - tca^2 presumes a
    unit vector, and a float32 one is off by up to 6e-8 - times |L|^2 = 2.5e7 for a ray leaving the sky sphere."""
    d = d / np.sqrt((d * d).sum(1))[:, None]
    L = c[None, :, :] - org[:, None, :]
    tca = (L * d[:, None, :]).sum(2)
    d2 = (L * L).sum(2) - tca * tca
    return d2, tca


SCENES = ["demo", "synth1024", "stress"]


def make_desc(rt, name):
    if name == "demo":
        return rt.SceneDesc(rt.scenes.demo_objects())
    if name == "synth1024":
        return rt.SceneDesc(rt.scenes.synth_objects(1024, 1024))
    rng = np.random.default_rng(7)
    objs = []
    for i in range(300):                                   # centres out to |C| ~ 4000, radii over five decades
        r = 10.0 ** rng.uniform(-2, 2.7)
        c = rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(0, 3.6)
        objs.append(rt.createSphere(list(c), r, rt.createMaterial([1, 1, 1], [0, .5, .5, .3, 0], 10, 1.0)))
    objs.append(rt.createSphere([0, 0, 0], 5000, rt.createMaterial([0, 0, 0], [1, 0, 0, 0, 0], 0, 1.0)))
    return rt.SceneDesc(objs)


@pytest.mark.parametrize("scene", SCENES)
def test_shared_origin_table_never_loses_a_hit(rt, scene):
    lib = rt._abi.load()
    desc = make_desc(rt, scene)
    c, r2 = spheres_of(desc)
    rng = np.random.default_rng(11)
    origins = [np.array([0.0, 1.5, 10.0]), np.array([5.0, 10.0, 5.0]), np.array([5.0, 10.0, 0.0])]   # camera, lights
    if scene == "stress":
        origins.append(np.array([1234.5, -321.0, 777.0]))
    eps = 0.001
    checked = hits = 0
    for org in origins:
        tab, exact, gen, n = tables(rt, lib, desc, org)
        d = np.concatenate([limb_rays(rng, org, c, r2, 1500),
                            (lambda v: v / np.sqrt((v * v).sum(1))[:, None])(rng.normal(size=(1500, 3)))])
        d32 = d.astype(F32)                                # the ray the kernel holds
        nd = -d32
        t = fma32(nd[:, None, 2], tab[None, :n, 2], tab[None, :n, 3])
        t = fma32(nd[:, None, 1], tab[None, :n, 1], t)
        t = fma32(nd[:, None, 0], tab[None, :n, 0], t)
        cand = np.signbit(t)
        d2, tca = line_hits(org[None, :].repeat(len(d), 0), d32.astype(F64), c, r2)
        thc = np.sqrt(np.maximum(r2[None, :] - d2, 0))
        true_hit = (d2 <= r2[None, :]) & (tca + thc >= eps)          # some root >= eps (main.js:429-439)
        lost = true_hit & ~cand
        assert not lost.any(), (scene, org, int(lost.sum()), np.argwhere(lost)[:5])
        # padding entries never test positive
        tp = fma32(nd[:, None, 2], tab[None, n:, 2], tab[None, n:, 3])
        assert not np.signbit(fma32(nd[:, None, 0], tab[None, n:, 0], fma32(nd[:, None, 1], tab[None, n:, 1], tp))).any()
        checked += true_hit.size
        hits += int(true_hit.sum())
        # the margin costs little: candidates that are not hits stay a small fraction of all pairs
        assert (cand & ~true_hit).sum() <= 0.02 * cand.size + 2 * len(d), (scene, int((cand & ~true_hit).sum()))
    assert checked > 1e5 and hits > 1000


@pytest.mark.parametrize("scene", SCENES)
def test_general_table_never_loses_a_hit(rt, scene):
    lib = rt._abi.load()
    desc = make_desc(rt, scene)
    c, r2 = spheres_of(desc)
    rng = np.random.default_rng(13)
    tab, exact, gen, n = tables(rt, lib, desc, [0.0, 1.5, 10.0])
    # ray origins: points ON random spheres (hit points of parent rays), as the kernel stores them: fp32
    k = 2500
    idx = rng.integers(0, n, k)
    nrm = rng.normal(size=(k, 3))
    nrm /= np.sqrt((nrm * nrm).sum(1))[:, None]
    org = (c[idx] + nrm * np.sqrt(r2[idx])[:, None]).astype(F32)
    o64 = org.astype(F64)
    d = np.empty((k, 3))
    d[: k // 2] = rng.normal(size=(k // 2, 3))
    d[: k // 2] /= np.sqrt((d[: k // 2] ** 2).sum(1))[:, None]
    for i in range(k // 2, k):                             # aimed at the limb of another sphere
        d[i] = limb_rays(rng, o64[i], c, r2, 8)[0]
    d32 = d.astype(F32)
    dx, dy, dz = d32[:, 0], d32[:, 1], d32[:, 2]
    # branchless orthonormal basis of ray_setup<false> (Duff et al. 2017), fp32
    sg = np.copysign(F32(1), dz)
    ka = (F32(-1) / (sg + dz)).astype(F32)
    kb = (dx * dy * ka).astype(F32)
    u = np.stack([(F32(1) + sg * dx * dx * ka), sg * kb, -sg * dx], 1).astype(F32)
    v = np.stack([kb, (sg + dy * dy * ka), -dy], 1).astype(F32)
    d2, tca = line_hits(o64, d32.astype(F64), c, r2)
    line_hit = d2 <= r2[None, :]
    assert line_hit.sum() > 2000
    # -(u.org), -(v.org): the kernel evaluates them in fp32 and the compiler is free to contract; take both ends -
    # one rounding of the exact dot product, and plain fp32 products and sums without any FMA (the most rounding)
    exact_dots = ((-(u.astype(F64) * o64).sum(1)).astype(F32), (-(v.astype(F64) * o64).sum(1)).astype(F32))
    plain_dots = (-((u[:, 0] * org[:, 0] + u[:, 1] * org[:, 1]) + u[:, 2] * org[:, 2]),
                  -((v[:, 0] * org[:, 0] + v[:, 1] * org[:, 1]) + v[:, 2] * org[:, 2]))
    for nku, nkv in (exact_dots, plain_dots):
        a = fma32(u[:, None, 2], gen[None, :n, 2], nku[:, None])
        a = fma32(u[:, None, 1], gen[None, :n, 1], a)
        a = fma32(u[:, None, 0], gen[None, :n, 0], a)
        b = fma32(v[:, None, 2], gen[None, :n, 2], nkv[:, None])
        b = fma32(v[:, None, 1], gen[None, :n, 1], b)
        b = fma32(v[:, None, 0], gen[None, :n, 0], b)
        e = fma32(a, a, gen[None, :n, 3].repeat(k, 0))
        e = fma32(b, b, e)
        cand = np.signbit(e)
        lost = line_hit & ~cand
        # the sphere a ray starts on is handled analytically by the slow path's `self` branch, but it must still
        # be a candidate (the ray's origin lies ON it): included in the check
        assert not lost.any(), (scene, int(lost.sum()), np.argwhere(lost)[:5])
        assert (cand & ~line_hit).sum() <= 0.02 * cand.size + 2 * k
