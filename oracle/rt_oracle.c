/*
 * rt_oracle.c — CPU float64 RESTATEMENT of the per-pixel hot path of
 * termuxinator/html5-canvas-raytracer main.js (build '741').
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path (librtb200.so) never links
 * or calls anything in oracle/ and has no CPU fallback.
 *
 * Parity status: the reference ships NO tests, golden vectors or fixtures and
 * this container has no JavaScript engine, so the restatement is pinned by executing the
 * UNMODIFIED /root/reference/main.js under oracle/minijs (a small ECMAScript-subset interpreter
 * written for this purpose; IEEE doubles via Python floats) and committing its framebuffers under
 * tests/golden/ (see oracle/gen_golden.py).  That is the reference's own source run here, but not
 * under V8: Math.atan2/asin/pow/tan come from the host libm instead of V8's fdlibm port, which can
 * differ in the last ulp (the committed frames do not move under +-2^20 ulps of such differences: see
 * oracle_set_libm_jitter below).  A run under real Node (oracle/run_main_js.js) is the remaining step.
 *
 * Rules followed: IEEE binary64, one rounding per JS operator in source association order, no FMA
 * contraction (build with -ffp-contract=off, no -ffast-math), Math.max/min NaN-propagating,
 * `&` = ToInt32, Uint8ClampedArray store = NaN->0, clamp, round-half-to-even.
 *
 * Each function cites the main.js lines it restates.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/rtb200.h"

/* Sensitivity probe for the one open caveat above (host libm vs V8's fdlibm port, last ulp): with a non-zero
 * seed every Math.atan2 / Math.asin / Math.pow result is moved by -max_ulps .. +max_ulps ulps, chosen by a hash of the
 * result's bits and the seed (deterministic, thread-safe).  tests/test_oracle.py renders the golden frames
 * with it and counts the bytes that move: a bound on what a different-but-faithful libm can change.
 * 0 (the default) = exact libm results; nothing but that test ever sets it. */
static uint64_t g_libm_jitter = 0;
static int64_t g_libm_jitter_ulps = 1;
void oracle_set_libm_jitter(uint64_t seed, int64_t max_ulps) { g_libm_jitter = seed; g_libm_jitter_ulps = max_ulps > 0 ? max_ulps : 1; }
static inline double jitter_ulp(double r) {
  if (g_libm_jitter == 0 || !isfinite(r)) return r;
  uint64_t b;
  memcpy(&b, &r, 8);
  uint64_t h = (b ^ g_libm_jitter) * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  const int64_t k = (int64_t)(h % (uint64_t)(2 * g_libm_jitter_ulps + 1)) - g_libm_jitter_ulps;   /* -max .. +max ulps */
  if ((b << 1) == 0) return r;                       /* +-0 stays */
  b += (uint64_t)k;                                  /* neighbouring doubles have neighbouring bit patterns */
  double q;
  memcpy(&q, &b, 8);
  return isfinite(q) && (q < 0) == (r < 0) ? q : r;
}
static inline double m_atan2(double y, double x) { return jitter_ulp(atan2(y, x)); }
static inline double m_asin(double x) { return jitter_ulp(asin(x)); }
static inline double m_pow(double x, double y) { return jitter_ulp(pow(x, y)); }

typedef struct OracleCounters {
  uint64_t nodes;         /* intersectWorld calls with segs > 0                 main.js:220-221 */
  uint64_t shadow_rays;   /* light iterations entering the occluder loop        main.js:293     */
  uint64_t sphere_tests;  /* intersectSphere evaluations (closest-hit + occluder) main.js:228,296 */
  uint64_t sampler_calls; /* main.js:320                                                        */
  uint64_t tir;           /* total internal reflections                         main.js:261-263 */
  uint64_t depth_cap_px;  /* pixels (sub-samples) whose tree reached segs == 0  main.js:221     */
  uint64_t max_nodes_px;  /* max tree nodes in one pixel sub-sample                             */
  uint64_t star_draws;    /* Math.random() draws                                main.js:136     */
} OracleCounters;

typedef struct Hit {      /* createIntersect() main.js:216-218 */
  double t;
  double p[3], n[3], l[3];
  double u, v;
  const RtbSphere* m;     /* ext.m = obj.mtl (main.js:448); we keep the sphere */
} Hit;

typedef struct Ctx {
  const RtbSceneDesc* s;
  int star_mode;
  uint64_t star_seed;
  const double* rand_stream;   /* optional supplied Math.random stream (raster + post-order)     */
  uint64_t rand_len;
  uint64_t* rand_pos;
  /* per-pixel */
  uint32_t px, py, sub;
  uint32_t node_in_px;
  int hit_cap;
  OracleCounters c;
} Ctx;

/* ---- JS Math helpers ------------------------------------------------------------------------ */
static double js_max(double a, double b) { /* Math.max: NaN if either NaN; +0 > -0 */
  if (isnan(a) || isnan(b)) return NAN;
  if (a == 0 && b == 0) return signbit(a) ? b : a;
  return a > b ? a : b;
}
static double js_min(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  if (a == 0 && b == 0) return signbit(a) ? a : b;
  return a < b ? a : b;
}
static int32_t js_toint32(double d) { /* ECMA-262 ToInt32 */
  if (!isfinite(d)) return 0;
  double t = trunc(d);
  double m = fmod(t, 4294967296.0);
  if (m < 0) m += 4294967296.0;
  uint32_t u = (uint32_t)m;
  return (int32_t)u;
}
static uint8_t js_clamped_u8(double v) { /* Uint8ClampedArray store, main.js:195-198 */
  if (!(v > 0)) return 0;      /* NaN, -x, 0 */
  if (v >= 255) return 255;
  return (uint8_t)nearbyint(v); /* default rounding mode = round-half-to-even */
}

/* ---- vec3 helpers main.js:20-75 --------------------------------------------------------------- */
static double dot3D(const double a[3], const double b[3]) { /* main.js:49-51 */
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static void scale3D(double o[3], const double v[3], double t) { /* main.js:32-34 */
  o[0] = v[0] * t; o[1] = v[1] * t; o[2] = v[2] * t;
}
static void project3D(double r[3], const double o[3], const double v[3], double t) { /* 36-38 */
  r[0] = o[0] + v[0] * t; r[1] = o[1] + v[1] * t; r[2] = o[2] + v[2] * t;
}
static void reflect3D(double r[3], const double v[3], const double n[3]) { /* main.js:40-43 */
  const double t = -(2 * dot3D(v, n));
  project3D(r, v, n, t);
}
static void between3D(double r[3], const double a[3], const double b[3]) { /* main.js:68-70 */
  r[0] = b[0] - a[0]; r[1] = b[1] - a[1]; r[2] = b[2] - a[2];
}
static void normal3D(double r[3], const double v[3]) { /* main.js:62-66: scale by 1/l */
  const double l = sqrt(dot3D(v, v));
  if (l == 0) { r[0] = v[0]; r[1] = v[1]; r[2] = v[2]; return; }
  scale3D(r, v, 1 / l);
}
static void oppose3D(double r[3], const double v[3]) { r[0] = -v[0]; r[1] = -v[1]; r[2] = -v[2]; }

/* ---- intersectSphere main.js:420-451 ---------------------------------------------------------- */
static double intersectSphere(Ctx* cx, const RtbSphere* obj, const double org[3],
                              const double dir[3], Hit* ext) {
  const double eps = cx->s->eps;
  double t = INFINITY;
  double L[3];
  cx->c.sphere_tests++;
  between3D(L, org, obj->origin);
  const double tca = dot3D(dir, L);
  const double d2 = dot3D(L, L) - tca * tca;
  if (d2 > obj->r2) return t;
  const double thc = sqrt(obj->r2 - d2);
  const double t0 = tca - thc;
  const double t1 = tca + thc;
  if (t0 < t1) {
    if (t0 < eps) {
      if (t1 < eps) return t;
      else t = t1;
    } else t = t0;
  } else {
    if (t1 < eps) {
      if (t0 < eps) return t;
      else t = t0;
    } else t = t1;
  }
  if (ext != NULL) {
    double d[3];
    ext->t = t;
    project3D(ext->p, org, dir, ext->t);
    between3D(d, obj->origin, ext->p);
    normal3D(ext->n, d);
    if ((t0 < eps) || (t1 < eps)) oppose3D(ext->l, ext->n);
    else { ext->l[0] = ext->n[0]; ext->l[1] = ext->n[1]; ext->l[2] = ext->n[2]; }
    ext->u = m_atan2(-ext->n[2], -ext->n[0]) / M_PI / 2 + 0.5;
    ext->v = m_asin(-ext->n[1]) / (M_PI / 2) / 2 + 0.5;
    ext->m = obj;
  }
  return t;
}

/* ---- samplers --------------------------------------------------------------------------------- */
/* sampleTexture main.js:343-351 */
static void sampleTexture(const RtbTexture* tex, double u, double v, double rgb[3]) {
  const double x = js_max(0, ceil(u * tex->width) - 1);
  const double y = js_max(0, ceil(v * tex->height) - 1);
  /* JS indexes with a double; NaN/out-of-range would read `undefined` -> NaN/255 = NaN -> byte 0.
   * In-range for every finite (u,v) in [0,1]. */
  if (!(x >= 0 && y >= 0 && x < tex->width && y < tex->height)) {
    rgb[0] = rgb[1] = rgb[2] = NAN;
    return;
  }
  const size_t i = ((size_t)y * (size_t)tex->width + (size_t)x) * 4;
  rgb[0] = tex->texels[i + 0] / 255.0;
  rgb[1] = tex->texels[i + 1] / 255.0;
  rgb[2] = tex->texels[i + 2] / 255.0;
}

/* splitmix64 finaliser; used only by RTB_STARS_HASH (a harness-defined stand-in for Math.random). */
static uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
double oracle_star_uniform(uint64_t seed, uint32_t px, uint32_t py, uint32_t sub, uint32_t node) {
  uint64_t k = mix64(seed ^ mix64(((uint64_t)py << 32) | px) ^ mix64(((uint64_t)sub << 32) | node) * 3);
  return (double)(k >> 11) * (1.0 / 9007199254740992.0); /* [0,1) with 53 bits */
}

static void sampler(Ctx* cx, const Hit* hit, uint32_t node_id, double rgb[3]) {
  const RtbMaterial* m = &hit->m->mtl;
  cx->c.sampler_calls++;
  switch (m->sampler_kind) {
    default:
    case RTB_SAMPLER_COLOR: /* main.js:404 */
      rgb[0] = m->color[0]; rgb[1] = m->color[1]; rgb[2] = m->color[2];
      return;
    case RTB_SAMPLER_TEXTURE: /* main.js:143-145, 149-151, 155-157 */
      sampleTexture(&cx->s->textures[m->texture_unit], hit->u, hit->v, rgb);
      return;
    case RTB_SAMPLER_CHECKER: { /* main.js:126-133 — note the axis convention differs from 446-447 */
      const double u = m_atan2(-hit->n[1], -hit->n[0]) / M_PI / 2 + 0.5;
      const double v = m_asin(-hit->n[2]) / (M_PI / 2) / 2 + 0.5;
      const int32_t s = js_toint32(u * m->checker_freq[0]) & 1;
      const int32_t t = js_toint32(v * m->checker_freq[1]) & 1;
      const double* c = m->checker_color[s ^ t];
      rgb[0] = c[0]; rgb[1] = c[1]; rgb[2] = c[2];
      return;
    }
    case RTB_SAMPLER_STARS: { /* main.js:135-139 */
      double c;
      if (cx->rand_stream) {       /* supplied stream, consumed in call order (single thread only) */
        c = cx->rand_stream[(*cx->rand_pos)++ % cx->rand_len];
        cx->c.star_draws++;
      } else if (cx->star_mode == RTB_STARS_HASH) {
        c = oracle_star_uniform(cx->star_seed, cx->px, cx->py, cx->sub, node_id);
        cx->c.star_draws++;
      } else {
        c = 0.5;                   /* P0 policy: Math.random = () => 0.5 */
        cx->c.star_draws++;
      }
      if (c >= m->star_threshold) { rgb[0] = rgb[1] = rgb[2] = 0; return; }
      c *= m->star_gain;
      rgb[0] = rgb[1] = rgb[2] = c;
      return;
    }
  }
}

/* ---- intersectWorld main.js:220-337 ------------------------------------------------------------ */
static void intersectWorld(Ctx* cx, int segs, const double org[3], const double dir[3],
                           uint32_t path, double out[3]) {
  const RtbSceneDesc* s = cx->s;
  const int n = s->n_spheres;
  if (segs == 0) { out[0] = out[1] = out[2] = 0; cx->hit_cap = 1; return; } /* main.js:221 */
  cx->c.nodes++;
  /* path code of this node inside the pixel: root 1, reflect child 2k, refract child 2k+1 — the
   * traversal-order-independent key of RTB_STARS_HASH (not part of the reference) */
  const uint32_t node_id = path;
  cx->node_in_px++;

  int hit_i = -1;                              /* main.js:223-230 */
  Hit hit; hit.t = INFINITY;
  for (int i = 0; i < n; i++) {
    Hit check; check.t = INFINITY;
    check.t = intersectSphere(cx, &s->spheres[i], org, dir, &check);
    if (check.t < hit.t) { hit = check; hit_i = i; }
  }
  if (hit.t == INFINITY) { out[0] = 1; out[1] = 0; out[2] = 0; return; } /* main.js:231 */
  const RtbMaterial* m = &hit.m->mtl;

  double reflect_dir[3] = {0, 0, 0};           /* main.js:233-239 */
  double reflect_len = 0;
  if (m->albedo[3] > 0) {
    reflect3D(reflect_dir, dir, hit.n);
    reflect_len = sqrt(dot3D(reflect_dir, reflect_dir));
    if (reflect_len != 0) scale3D(reflect_dir, reflect_dir, 1 / reflect_len);
  }

  double refract_dir[3] = {0, 0, 0};           /* main.js:241-266 */
  double refract_len = 0;
  if (m->albedo[4] > 0) {
    double norm[3], eta;
    const double dot = dot3D(dir, hit.n);
    double cosi = -js_max(-1, js_min(1, dot));
    if (cosi < 0) {
      cosi = -cosi;
      oppose3D(norm, hit.n);
      eta = m->refract_index;
    } else {
      norm[0] = hit.n[0]; norm[1] = hit.n[1]; norm[2] = hit.n[2];
      eta = 1 / m->refract_index;
    }
    const double k = 1 - eta * eta * (1 - cosi * cosi);
    if (k > 0) {
      const double q = eta * cosi - sqrt(k);
      refract_dir[0] = dir[0] * eta + norm[0] * q;
      refract_dir[1] = dir[1] * eta + norm[1] * q;
      refract_dir[2] = dir[2] * eta + norm[2] * q;
    } else {
      reflect3D(refract_dir, dir, norm);
      cx->c.tir++;
    }
    refract_len = sqrt(dot3D(refract_dir, refract_dir));
    if (refract_len != 0) scale3D(refract_dir, refract_dir, 1 / refract_len);
  }

  double reflect_color[3] = {0, 0, 0};         /* main.js:268-272 */
  if (reflect_len != 0) {
    intersectWorld(cx, segs - 1, hit.p, reflect_dir, 2u * path, reflect_color);
    scale3D(reflect_color, reflect_color, m->albedo[3]);
  }
  double refract_color[3] = {0, 0, 0};         /* main.js:274-278 */
  if (refract_len != 0) {
    intersectWorld(cx, segs - 1, hit.p, refract_dir, 2u * path + 1u, refract_color);
    scale3D(refract_color, refract_color, m->albedo[4]);
  }

  double diffuse_intensity = 0;                /* main.js:280-318 */
  double specular_intensity = 0;
  if ((m->albedo[1] > 0) || (m->albedo[2] > 0)) {
    double light_intensity = s->light_intensity; /* ONCE, outside the loop (main.js:284) */
    for (int k = 0; k < s->n_lights; k++) {
      const double* light = s->lights[k];
      double shadow_vec[3];
      between3D(shadow_vec, hit.p, light);
      const double light_mag = dot3D(shadow_vec, shadow_vec);
      const double light_len = sqrt(light_mag);
      if (light_len != 0) scale3D(shadow_vec, shadow_vec, 1 / light_len);
      const double shadow_dot = dot3D(shadow_vec, hit.l);
      if (shadow_dot <= 0) continue;
      cx->c.shadow_rays++;
      for (int j = 0; j < n; j++) {
        if (j == hit_i) continue;
        const RtbSphere* o = &s->spheres[j];
        const double t = intersectSphere(cx, o, hit.p, shadow_vec, NULL);
        if (t < light_len) {
          if (o->mtl.albedo[4] != 0) {
            light_intensity /= o->mtl.albedo[4];
          } else {
            light_intensity = 0; break;
          }
        }
      }
      if (light_intensity == 0) continue;
      diffuse_intensity += light_intensity * shadow_dot / light_mag;
      if (m->albedo[2] > 0) {
        double light_dir[3], light_ref[3], spec_dir[3];
        oppose3D(light_dir, shadow_vec);
        reflect3D(light_ref, light_dir, hit.l);
        normal3D(light_ref, light_ref);
        oppose3D(spec_dir, light_ref);
        const double spec_dot = dot3D(dir, spec_dir);
        if (spec_dot > 0) specular_intensity += m_pow(spec_dot, m->specular_exponent);
      }
    }
    diffuse_intensity = js_min(1, diffuse_intensity) * m->albedo[1];
    specular_intensity = js_min(1, specular_intensity) * m->albedo[2];
  }

  double rgb[3];                               /* main.js:320 */
  sampler(cx, &hit, node_id, rgb);

  double ambient_color[3], diffuse_color[3], specular_color[3];  /* main.js:322-336 */
  scale3D(ambient_color, rgb, m->albedo[0]);
  scale3D(diffuse_color, rgb, diffuse_intensity);
  scale3D(specular_color, rgb, specular_intensity);
  for (int c = 0; c < 3; c++) {
    const double shader = diffuse_color[c] + specular_color[c] + reflect_color[c] + refract_color[c];
    out[c] = js_max(ambient_color[c], js_min(1, shader));
  }
}

/* ---- pixel loop main.js:183-199 ---------------------------------------------------------------- */
static void render_pixel(Ctx* cx, const RtbCamera* cam, const RtbRenderOpts* o, int x, int y,
                         uint8_t* px_out, double* rgb_out_or_null) {
  const int g = o->spp_grid > 0 ? o->spp_grid : 1;
  double acc[3] = {0, 0, 0};
  double rgb[3];
  for (int j = 0; j < g; j++) {
    for (int i = 0; i < g; i++) {
      /* main.js:186 with +0.5 / -0.5 generalised to the (i,j) stratum centre;
       * g == 1 gives (0+0.5)/1 == 0.5 exactly. */
      const double ox = (i + 0.5) / g, oy = (j + 0.5) / g;
      const double dist[3] = {x - cam->projW + ox, cam->projH - y - oy, cam->projD};
      double target[3], ray[3];
      if (cam->camera_mode == RTB_CAMERA_REFERENCE) { /* main.js:187-191, per-component dist */
        for (int k = 0; k < 3; k++)
          target[k] = cam->origin[k] + cam->axisX[k] * dist[k] + cam->axisY[k] * dist[k] +
                      cam->axisZ[k] * dist[k];
      } else {
        for (int k = 0; k < 3; k++)
          target[k] = cam->origin[k] + cam->axisX[k] * dist[0] + cam->axisY[k] * dist[1] +
                      cam->axisZ[k] * dist[2];
      }
      between3D(ray, cam->origin, target);       /* main.js:192 — round trip kept on purpose */
      normal3D(ray, ray);                        /* main.js:193 */
      cx->px = (uint32_t)x; cx->py = (uint32_t)y; cx->sub = (uint32_t)(j * g + i);
      cx->node_in_px = 0; cx->hit_cap = 0;
      intersectWorld(cx, o->segs, cam->origin, ray, 1u, rgb);  /* main.js:194 */
      if (cx->node_in_px > cx->c.max_nodes_px) cx->c.max_nodes_px = cx->node_in_px;
      if (cx->hit_cap) cx->c.depth_cap_px++;
      if (g == 1) { acc[0] = rgb[0]; acc[1] = rgb[1]; acc[2] = rgb[2]; }
      else { acc[0] += rgb[0]; acc[1] += rgb[1]; acc[2] += rgb[2]; }
    }
  }
  if (g > 1) { const double q = (double)(g * g); acc[0] /= q; acc[1] /= q; acc[2] /= q; }
  px_out[0] = js_clamped_u8(255 * acc[0]);       /* main.js:195-198 */
  px_out[1] = js_clamped_u8(255 * acc[1]);
  px_out[2] = js_clamped_u8(255 * acc[2]);
  px_out[3] = 255;
  if (rgb_out_or_null) { rgb_out_or_null[0] = acc[0]; rgb_out_or_null[1] = acc[1]; rgb_out_or_null[2] = acc[2]; }
}

static void counters_add(OracleCounters* a, const OracleCounters* b) {
  a->nodes += b->nodes; a->shadow_rays += b->shadow_rays; a->sphere_tests += b->sphere_tests;
  a->sampler_calls += b->sampler_calls; a->tir += b->tir; a->depth_cap_px += b->depth_cap_px;
  a->star_draws += b->star_draws;
  if (b->max_nodes_px > a->max_nodes_px) a->max_nodes_px = b->max_nodes_px;
}

static int row_selected(const RtbRenderOpts* o, int y) {
  if (o->stripe_world > 1 && o->stripe_rows > 0)
    return ((y / o->stripe_rows) % o->stripe_world) == o->stripe_rank;
  return 1;
}

/*
 * Render rows [y0,y1) (and, when striping is on, only this rank's stripes) into a full-frame
 * row-major RGBA8 buffer.  `row_step` > 1 renders only every row_step-th row (bounded CPU samples
 * for the bench).  `rgb_f64_or_null` (W*H*3 doubles) receives the pre-quantisation
 * colour.  `threads` <= 0 = all cores; threads != 1 requires rand_stream == NULL.
 * Returns 0, or -1 on bad arguments.
 */
int oracle_render(const RtbSceneDesc* s, const RtbCamera* cam, const RtbRenderOpts* o,
                  uint8_t* out, double* rgb_f64_or_null, OracleCounters* counters_or_null,
                  int row_step, int threads, const double* rand_stream, uint64_t rand_len) {
  if (!s || !cam || !o || !out || s->n_spheres < 0 || o->segs < 0) return -1;
  const int W = cam->width, H = cam->height;
  int y0 = o->y0 < 0 ? 0 : o->y0;
  int y1 = (o->y1 <= 0 || o->y1 > H) ? H : o->y1;
  if (row_step < 1) row_step = 1;
  OracleCounters total; memset(&total, 0, sizeof total);
  uint64_t rand_pos = 0;
  if (rand_stream) threads = 1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel num_threads(threads)
  {
    Ctx cx; memset(&cx, 0, sizeof cx);
    cx.s = s; cx.star_mode = o->star_mode; cx.star_seed = o->star_seed;
    cx.rand_stream = rand_stream; cx.rand_len = rand_len; cx.rand_pos = &rand_pos;
#pragma omp for schedule(dynamic, 1)
    for (int y = y0; y < y1; y++) {
      if (((y - y0) % row_step) != 0 || !row_selected(o, y)) continue;
      for (int x = 0; x < W; x++) {
        const size_t p = (size_t)y * W + x;
        render_pixel(&cx, cam, o, x, y, out + p * 4, rgb_f64_or_null ? rgb_f64_or_null + p * 3 : NULL);
      }
    }
#pragma omp critical
    counters_add(&total, &cx.c);
  }
  if (counters_or_null) *counters_or_null = total;
  return 0;
}

/* Single ray, for unit tests of intersectWorld: returns the colour of intersectWorld(segs,org,dir). */
int oracle_trace_ray(const RtbSceneDesc* s, int segs, const double org[3], const double dir[3],
                     double out_rgb[3], OracleCounters* counters_or_null) {
  Ctx cx; memset(&cx, 0, sizeof cx);
  cx.s = s;
  intersectWorld(&cx, segs, org, dir, 1u, out_rgb);
  if (counters_or_null) *counters_or_null = cx.c;
  return 0;
}

/* intersectSphere alone (main.js:420-451): returns t; fills p,n,l,u,v when ext != 0. */
double oracle_intersect_sphere(const RtbSceneDesc* s, int sphere, const double org[3],
                               const double dir[3], double* ext13_or_null) {
  Ctx cx; memset(&cx, 0, sizeof cx);
  cx.s = s;
  Hit h; memset(&h, 0, sizeof h); h.t = INFINITY;
  double t = intersectSphere(&cx, &s->spheres[sphere], org, dir, ext13_or_null ? &h : NULL);
  if (ext13_or_null && t != INFINITY) {
    double* e = ext13_or_null;
    e[0] = h.t; memcpy(e + 1, h.p, 24); memcpy(e + 4, h.n, 24); memcpy(e + 7, h.l, 24);
    e[10] = h.u; e[11] = h.v; e[12] = 0;
  }
  return t;
}

void oracle_sample_texture(const RtbTexture* tex, double u, double v, double rgb[3]) {
  sampleTexture(tex, u, v, rgb);
}
uint8_t oracle_clamped_u8(double v) { return js_clamped_u8(v); }
int32_t oracle_toint32(double v) { return js_toint32(v); }
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
