"""Literal, slow Python transliteration of main.js's hot path — the SECOND WITNESS.

TEST INFRASTRUCTURE ONLY.  It is deliberately written independently of rt_oracle.c, list-of-three
vectors and recursion exactly as main.js has them, so that a transcription slip in either shows up
as a byte difference at small sizes.  Python floats are IEEE binary64 like JS numbers; `math.atan2 /
asin / pow` come from the host libm (the same one rt_oracle.c links).

Reference lines are cited per function (/root/reference/main.js).
"""
import math

import numpy as np

Infinity = math.inf


# ---- vec3 helpers main.js:20-75 -------------------------------------------------------------------------
def oppose3D(v):
    return [-v[0], -v[1], -v[2]]


def scale3D(v, t):
    return [v[0] * t, v[1] * t, v[2] * t]


def project3D(o, v, t):
    return [o[0] + v[0] * t, o[1] + v[1] * t, o[2] + v[2] * t]


def dot3D(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def reflect3D(v, n):
    t = -(2 * dot3D(v, n))
    return project3D(v, n, t)


def mag3D(v):
    return dot3D(v, v)


def length3D(v):
    return math.sqrt(mag3D(v))


def normal3D(v):
    l = length3D(v)
    if l == 0:
        return v
    return scale3D(v, 1 / l)


def between3D(a, b):
    return [b[0] - a[0], b[1] - a[1], b[2] - a[2]]


def _jsmax(a, b):
    if a != a or b != b:
        return math.nan
    return a if a > b else b


def _jsmin(a, b):
    if a != a or b != b:
        return math.nan
    return a if a < b else b


def _toint32(d):
    if d != d or d in (math.inf, -math.inf):
        return 0
    m = math.fmod(math.trunc(d), 4294967296.0)
    if m < 0:
        m += 4294967296.0
    u = int(m)
    return u - (1 << 32) if u >= (1 << 31) else u


def _clamped_u8(v):  # Uint8ClampedArray store: NaN -> 0, clamp, round half to even
    if not (v > 0):
        return 0
    if v >= 255:
        return 255
    return int(round(v))  # Python's round() is half-to-even


# ---- samplers ------------------------------------------------------------------------------------------------
def sampleTexture(texture, u, v):  # main.js:343-351
    x = max(0, math.ceil(u * texture.width) - 1)
    y = max(0, math.ceil(v * texture.height) - 1)
    i = (y * texture.width + x) * 4
    t = texture.texels
    return [int(t[i + 0]) / 255, int(t[i + 1]) / 255, int(t[i + 2]) / 255]


def _sampler(obj, hit):
    s = obj.mtl.sampler
    kind = type(s).__name__
    if kind == "textureSampler":                       # main.js:143-157
        return sampleTexture(s.texture, hit["u"], hit["v"])
    if kind == "checkerSampler":                       # main.js:126-133
        u = math.atan2(-hit["n"][1], -hit["n"][0]) / math.pi / 2 + 0.5
        v = math.asin(-hit["n"][2]) / (math.pi / 2) / 2 + 0.5
        a = _toint32(u * s.freq[0]) & 1
        b = _toint32(v * s.freq[1]) & 1
        return list(s.colors[a ^ b])
    if kind == "starSampler":                          # main.js:135-139 with Math.random = () => 0.5
        c = 0.5
        if c >= s.threshold:
            return [0, 0, 0]
        c *= s.gain
        return [c, c, c]
    return list(obj.mtl.color)                         # main.js:404


# ---- intersectSphere main.js:420-451 -----------------------------------------------------------------------
def intersectSphere(obj, org, dir_, ext, eps=0.001):
    t = Infinity
    L = between3D(org, obj.origin)
    tca = dot3D(dir_, L)
    d2 = mag3D(L) - tca * tca
    if d2 > obj.r2:
        return t
    thc = math.sqrt(obj.r2 - d2)
    t0 = tca - thc
    t1 = tca + thc
    if t0 < t1:
        if t0 < eps:
            if t1 < eps:
                return t
            else:
                t = t1
        else:
            t = t0
    else:
        if t1 < eps:
            if t0 < eps:
                return t
            else:
                t = t0
        else:
            t = t1
    if ext is not None:
        ext["t"] = t
        ext["p"] = project3D(org, dir_, ext["t"])
        ext["n"] = between3D(obj.origin, ext["p"])
        ext["n"] = normal3D(ext["n"])
        ext["l"] = oppose3D(ext["n"]) if ((t0 < eps) or (t1 < eps)) else ext["n"]
        ext["u"] = math.atan2(-ext["n"][2], -ext["n"][0]) / math.pi / 2 + 0.5
        ext["v"] = math.asin(-ext["n"][1]) / (math.pi / 2) / 2 + 0.5
        ext["m"] = obj
    return t


def createIntersect():  # main.js:216-218
    return {"t": Infinity, "p": [], "n": [], "l": [], "u": 0, "v": 0, "m": None}


LIGHTS = [[5.0, 10.0, 5.0], [5.0, 10.0, 0.0]]          # main.js:283


# ---- intersectWorld main.js:220-337 ------------------------------------------------------------------------
def intersectWorld(segs, objs, org, dir_):
    if segs == 0:
        return [0, 0, 0]
    hit_i = -1
    hit = createIntersect()
    for i in range(len(objs)):
        o = objs[i]
        check = createIntersect()
        check["t"] = intersectSphere(o, org, dir_, check)
        if check["t"] < hit["t"]:
            hit = check
            hit_i = i
    if hit["t"] == Infinity:
        return [1, 0, 0]
    m = hit["m"].mtl

    reflect_dir = [0, 0, 0]
    reflect_len = 0
    if m.albedo[3] > 0:
        reflect_dir = reflect3D(dir_, hit["n"])
        reflect_len = length3D(reflect_dir)
        if reflect_len != 0:
            reflect_dir = scale3D(reflect_dir, 1 / reflect_len)

    refract_dir = [0, 0, 0]
    refract_len = 0
    if m.albedo[4] > 0:
        dot = dot3D(dir_, hit["n"])
        cosi = -_jsmax(-1, _jsmin(1, dot))
        if cosi < 0:
            cosi = -cosi
            norm = oppose3D(hit["n"])
            eta = m.refract_index
        else:
            norm = list(hit["n"])
            eta = 1 / m.refract_index
        k = 1 - eta * eta * (1 - cosi * cosi)
        if k > 0:
            q = eta * cosi - math.sqrt(k)
            refract_dir[0] = dir_[0] * eta + norm[0] * q
            refract_dir[1] = dir_[1] * eta + norm[1] * q
            refract_dir[2] = dir_[2] * eta + norm[2] * q
        else:
            refract_dir = reflect3D(dir_, norm)
        refract_len = length3D(refract_dir)
        if refract_len != 0:
            refract_dir = scale3D(refract_dir, 1 / refract_len)

    reflect_color = [0, 0, 0]
    if reflect_len != 0:
        reflect_color = intersectWorld(segs - 1, objs, hit["p"], reflect_dir)
        reflect_color = scale3D(reflect_color, m.albedo[3])

    refract_color = [0, 0, 0]
    if refract_len != 0:
        refract_color = intersectWorld(segs - 1, objs, hit["p"], refract_dir)
        refract_color = scale3D(refract_color, m.albedo[4])

    diffuse_intensity = 0
    specular_intensity = 0
    if (m.albedo[1] > 0) or (m.albedo[2] > 0):
        light_intensity = 50                               # outside the loop on purpose (main.js:284)
        for k in range(len(LIGHTS)):
            light = LIGHTS[k]
            shadow_vec = between3D(hit["p"], light)
            light_mag = mag3D(shadow_vec)
            light_len = math.sqrt(light_mag)
            if light_len != 0:
                shadow_vec = scale3D(shadow_vec, 1 / light_len)
            shadow_dot = dot3D(shadow_vec, hit["l"])
            if shadow_dot <= 0:
                continue
            for j in range(len(objs)):
                if j == hit_i:
                    continue
                o = objs[j]
                t = intersectSphere(o, hit["p"], shadow_vec, None)
                if t < light_len:
                    if o.mtl.albedo[4] != 0:
                        light_intensity /= o.mtl.albedo[4]
                    else:
                        light_intensity = 0
                        break
            if light_intensity == 0:
                continue
            diffuse_intensity += light_intensity * shadow_dot / light_mag
            if m.albedo[2] > 0:
                light_dir = oppose3D(shadow_vec)
                light_ref = reflect3D(light_dir, hit["l"])
                light_ref = normal3D(light_ref)
                spec_dir = oppose3D(light_ref)
                spec_dot = dot3D(dir_, spec_dir)
                if spec_dot > 0:
                    specular_intensity += math.pow(spec_dot, m.specular_exponent)
        diffuse_intensity = _jsmin(1, diffuse_intensity) * m.albedo[1]
        specular_intensity = _jsmin(1, specular_intensity) * m.albedo[2]

    rgb = _sampler(hit["m"], hit)

    ambient_color = scale3D(rgb, m.albedo[0])
    diffuse_color = scale3D(rgb, diffuse_intensity)
    specular_color = scale3D(rgb, specular_intensity)
    shader_color = [
        diffuse_color[0] + specular_color[0] + reflect_color[0] + refract_color[0],
        diffuse_color[1] + specular_color[1] + reflect_color[1] + refract_color[1],
        diffuse_color[2] + specular_color[2] + reflect_color[2] + refract_color[2],
    ]
    return [
        _jsmax(ambient_color[0], _jsmin(1, shader_color[0])),
        _jsmax(ambient_color[1], _jsmin(1, shader_color[1])),
        _jsmax(ambient_color[2], _jsmin(1, shader_color[2])),
    ]


# ---- spanish(y) main.js:183-199 --------------------------------------------------------------------------------
def render(objects, cam, segs=8):
    W, H = cam.width, cam.height
    data = np.zeros((H, W, 4), np.uint8)
    origin, axisX, axisY, axisZ = cam.origin, cam.axisX, cam.axisY, cam.axisZ
    projW, projH, projD = cam.projW, cam.projH, cam.projD
    for y in range(H):
        for x in range(W):
            dist = [x - projW + 0.5, projH - y - 0.5, projD]
            target = [
                origin[0] + axisX[0] * dist[0] + axisY[0] * dist[0] + axisZ[0] * dist[0],
                origin[1] + axisX[1] * dist[1] + axisY[1] * dist[1] + axisZ[1] * dist[1],
                origin[2] + axisX[2] * dist[2] + axisY[2] * dist[2] + axisZ[2] * dist[2],
            ]
            ray = between3D(origin, target)
            ray = normal3D(ray)
            rgb = intersectWorld(segs, objects, origin, ray)
            data[y, x, 0] = _clamped_u8(255 * rgb[0])
            data[y, x, 1] = _clamped_u8(255 * rgb[1])
            data[y, x, 2] = _clamped_u8(255 * rgb[2])
            data[y, x, 3] = 255
    return data
