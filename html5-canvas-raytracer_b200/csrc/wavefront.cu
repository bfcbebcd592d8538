// wavefront.cu — fp32 THROUGHPUT path: the Whitted tree of main.js:220-337 evaluated
// breadth-first over a whole row band instead of depth-first per pixel.
//
//   TRACE  level L : one ray per lane; closest-hit sweep over ALL spheres in array order from shared
//                    memory (main.js:223-230), then the hit record, reflect/refract directions
//                    (main.js:233-266), the sampler (main.js:320) and the occlusion-independent
//                    part of the light loop (main.js:286-292, 306-313).  Live child rays and
//                    shadow rays are compacted into queues with warp ballots (one atomic per warp).
//   SHADOW light k : occluder sweep main.js:293-304 for every queued node, traced FROM the light so
//                    that all rays of the launch share an origin: per-sphere constants
//                    (C - light, |C - light|^2 - r2) are precomputed in fp64 and the per-test work
//                    drops to 4 FP32 instructions.  The light_intensity carry-over
//                    between lights (main.js:284, H1-b) is the queue hand-off k -> next facing light.
//   COMPOSE level L: post-order clamp-compose main.js:322-336, deepest level first; each child
//                    writes its colour (scaled by the parent's albedo[3|4]) into its parent; level 1
//                    averages sub-samples and stores RGBA8 as uint4 (4 pixels per thread).
//
// Every level is a grid-stride kernel whose bounds are read from device counters, so a band is a
// fixed launch sequence without host round trips.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "rtb_internal.h"
#include "pixel_kernel.cuh"   // primary_ray<>

namespace {

struct WfSticky {             // survives across bands of one call
  uint32_t overflow;
  uint32_t _pad;
  unsigned long long nodes, shadow_rays, sphere_tests;
};

struct WfParams {
  DevSceneT<float> sc;
  DevFrameT<double> fr64;     // primary rays are generated in fp64 and rounded once
  RowMap rm;
  WfBuffers b;
  WfSticky* sticky;
  const float4* geom;         // Cx, Cy, Cz, r2
  const float4* primary_tab;  // C - origin, |C - origin|^2 - r2
  const float4* light_tab[RTB_MAX_LIGHTS];
  float origin[3];
  int32_t band_l0;            // first local row of the band
  int32_t band_rows;
  uint32_t n_sub;             // band_rows * W * spp  == number of level-1 nodes
  int32_t spp;                // g*g
  int32_t aligned4;           // W % 4 == 0 and output 16-byte aligned: uint4 stores
};

constexpr int WF_THREADS = 256;
constexpr float WF_INF = __builtin_huge_valf();

RTB_DEVINL uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// Level bounds: level 1 = [0, n_sub); level L >= 2 = n_sub + [snap[L-2], snap[L-1]).
RTB_DEVINL void level_range(const WfParams& P, int level, uint32_t& begin, uint32_t& end) {
  if (level == 1) { begin = 0; end = P.n_sub; return; }
  const volatile uint32_t* snap = P.b.ctr->level_end;
  begin = P.n_sub + snap[level - 2];
  end = P.n_sub + snap[level - 1];
}

// Last CTA out records how many children exist now: they are the next level.
RTB_DEVINL void publish_level(const WfParams& P, int level) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(&P.b.ctr->blocks_done, 1u);
    if (t == gridDim.x - 1) {
      uint32_t nc = atomicAdd(&P.b.ctr->node_count, 0u);
      const uint32_t room = P.b.cap_nodes - P.n_sub;
      if (nc > room) nc = room;
      P.b.ctr->level_end[level] = nc;
      P.b.ctr->blocks_done = 0;
      __threadfence();
    }
  }
}

// ---- TRACE ----------------------------------------------------------------------------------------------
template <bool PRIMARY>
__global__ void __launch_bounds__(WF_THREADS)
wf_trace_kernel(const __grid_constant__ WfParams P, int level) {
  extern __shared__ float4 tab[];
  const int n = P.sc.n_spheres;
  {
    const float4* src = PRIMARY ? P.primary_tab : P.geom;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tab[i] = src[i];
  }
  __syncthreads();
  uint32_t begin, end;
  level_range(P, level, begin, end);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const float eps = P.sc.eps;
  const int W = P.fr64.width;
  unsigned long long my_nodes = 0;

  for (uint64_t base = (uint64_t)begin + (uint64_t)gwarp * 32u; base < end; base += (uint64_t)nwarps * 32u) {
    const uint32_t node = (uint32_t)base + lane;
    const bool active = node < end;
    float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f};
    int segs_left = 0;
    uint32_t path = 1, px = 0, py = 0, sub = 0;
    if (active) {
      if (PRIMARY) {
        const uint32_t pix = node / (uint32_t)P.spp;
        sub = node - pix * (uint32_t)P.spp;
        const int lrow = P.band_l0 + (int)(pix / (uint32_t)W);
        px = pix % (uint32_t)W;
        py = (uint32_t)P.rm.local_to_y(lrow);
        const int g = P.fr64.spp_grid;
        double ray[3];
        primary_ray<double>(P.fr64, (int)px, (int)py, (int)(sub % g), (int)(sub / g), ray);
        d[0] = (float)ray[0]; d[1] = (float)ray[1]; d[2] = (float)ray[2];
        o[0] = P.origin[0]; o[1] = P.origin[1]; o[2] = P.origin[2];
        segs_left = P.fr64.segs;
      } else {
        const float4 ro = P.b.ray_o[node], rd = P.b.ray_d[node];
        const int2 lk = P.b.link[node];
        o[0] = ro.x; o[1] = ro.y; o[2] = ro.z;
        d[0] = ro.w; d[1] = rd.x; d[2] = rd.y;
        path = (uint32_t)lk.y & 0xffffffu;
        segs_left = (int)((uint32_t)lk.y >> 24);
        // px/py/sub (star hash key) ride in the pixel id stored by the level-1 ancestor
        const uint32_t pixsub = __float_as_uint(rd.w);
        const uint32_t pix = pixsub / (uint32_t)P.spp;
        sub = pixsub - pix * (uint32_t)P.spp;
        px = pix % (uint32_t)W;
        py = (uint32_t)P.rm.local_to_y(P.band_l0 + (int)(pix / (uint32_t)W));
      }
    }

    // closest-hit sweep, array order, strict '<' keeps the lowest index on ties (main.js:229)
    float best = WF_INF;
    int hit_i = -1;
    bool flip = false;
    if (active) {
      my_nodes++;
#pragma unroll 4
      for (int j = 0; j < n; j++) {
        const float4 s = tab[j];
        float tca, disc;
        if (PRIMARY) {
          tca = d[0] * s.x + d[1] * s.y + d[2] * s.z;
          disc = tca * tca - s.w;
        } else {
          const float L0 = s.x - o[0], L1 = s.y - o[1], L2 = s.z - o[2];
          tca = d[0] * L0 + d[1] * L1 + d[2] * L2;
          const float d2 = (L0 * L0 + L1 * L1 + L2 * L2) - tca * tca;
          disc = s.w - d2;
          if (disc >= 0.f) {   // candidate: redo d2 as |L - tca d|^2, which does not cancel for r = 500
            const float v0 = L0 - tca * d[0], v1 = L1 - tca * d[1], v2 = L2 - tca * d[2];
            disc = s.w - (v0 * v0 + v1 * v1 + v2 * v2);
          }
        }
        if (disc >= 0.f) {
          const float thc = sqrtf(disc);
          const float t0 = tca - thc, t1 = tca + thc;
          const float t = (t0 >= eps) ? t0 : ((t1 >= eps) ? t1 : WF_INF);
          if (t < best) { best = t; hit_i = j; flip = (t0 < eps); }
        }
      }
    }
    __syncwarp();

    // ---- hit record, directions, sampler, light terms --------------------------------------------
    bool want_refl = false, want_refr = false;
    float p[3] = {0.f, 0.f, 0.f}, rdir[3] = {0.f, 0.f, 0.f}, fdir[3] = {0.f, 0.f, 0.f};
    float w_refl = 0.f, w_refr = 0.f;
    int first_light = -1;
    if (active) {
      if (hit_i < 0) {                                    // main.js:231
        P.b.pos[node] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        P.b.base[node] = make_float4(1.f, 0.f, 0.f, __int_as_float(0));
      } else {
        const float4 g = __ldg(P.geom + hit_i);
        const DevMaterial<float>& m = P.sc.mtl[hit_i];
        p[0] = o[0] + d[0] * best; p[1] = o[1] + d[1] * best; p[2] = o[2] + d[2] * best;
        float nn[3] = {p[0] - g.x, p[1] - g.y, p[2] - g.z};
        const float nl = sqrtf(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
        if (nl != 0.f) { const float r = 1.f / nl; nn[0] *= r; nn[1] *= r; nn[2] *= r; }
        const float l[3] = {flip ? -nn[0] : nn[0], flip ? -nn[1] : nn[1], flip ? -nn[2] : nn[2]};
        const float a1 = m.albedo[1], a2 = m.albedo[2], a3 = m.albedo[3], a4 = m.albedo[4];
        const float dn = d[0] * nn[0] + d[1] * nn[1] + d[2] * nn[2];
        if (a3 > 0.f) {                                   // main.js:233-239
          const float tt = -(2.f * dn);
          rdir[0] = d[0] + nn[0] * tt; rdir[1] = d[1] + nn[1] * tt; rdir[2] = d[2] + nn[2] * tt;
          const float len = sqrtf(rdir[0] * rdir[0] + rdir[1] * rdir[1] + rdir[2] * rdir[2]);
          if (len != 0.f) {
            const float r = 1.f / len; rdir[0] *= r; rdir[1] *= r; rdir[2] *= r;
            want_refl = segs_left > 1;                    // a child with segs == 0 returns [0,0,0]
            w_refl = a3;
          }
        }
        if (a4 > 0.f) {                                   // main.js:241-266
          float norm[3], eta;
          float cosi = -js_max<float>(-1.f, js_min<float>(1.f, dn));
          if (cosi < 0.f) { cosi = -cosi; norm[0] = -nn[0]; norm[1] = -nn[1]; norm[2] = -nn[2]; eta = m.refract_index; }
          else { norm[0] = nn[0]; norm[1] = nn[1]; norm[2] = nn[2]; eta = 1.f / m.refract_index; }
          const float k = 1.f - eta * eta * (1.f - cosi * cosi);
          if (k > 0.f) {
            const float q = eta * cosi - sqrtf(k);
            fdir[0] = d[0] * eta + norm[0] * q; fdir[1] = d[1] * eta + norm[1] * q; fdir[2] = d[2] * eta + norm[2] * q;
          } else {
            const float tt = -(2.f * (d[0] * norm[0] + d[1] * norm[1] + d[2] * norm[2]));
            fdir[0] = d[0] + norm[0] * tt; fdir[1] = d[1] + norm[1] * tt; fdir[2] = d[2] + norm[2] * tt;
          }
          const float len = sqrtf(fdir[0] * fdir[0] + fdir[1] * fdir[1] + fdir[2] * fdir[2]);
          if (len != 0.f) {
            const float r = 1.f / len; fdir[0] *= r; fdir[1] *= r; fdir[2] *= r;
            want_refr = segs_left > 1;
            w_refr = a4;
          }
        }
        // sampler (main.js:320); u,v only matter to the texture sampler
        float u = 0.f, v = 0.f, rgb[3];
        if (m.sampler_kind == RTB_SAMPLER_TEXTURE) {
          u = atan2f(-nn[2], -nn[0]) / 3.14159265358979323846f / 2.f + 0.5f;
          v = asinf(-nn[1]) / (3.14159265358979323846f / 2.f) / 2.f + 0.5f;
        }
        run_sampler<float>(P.sc, m, nn, u, v, P.fr64.star_mode, P.fr64.star_seed, px, py, sub, path, rgb);
        // occlusion-independent light terms (main.js:286-292, 306-313)
        uint32_t mask = 0;
        float q[RTB_MAX_LIGHTS] = {0.f, 0.f, 0.f, 0.f}, sp[RTB_MAX_LIGHTS] = {0.f, 0.f, 0.f, 0.f};
        if (a1 > 0.f || a2 > 0.f) {
          for (int k = 0; k < P.sc.n_lights; k++) {
            float sv[3] = {P.sc.lights[k][0] - p[0], P.sc.lights[k][1] - p[1], P.sc.lights[k][2] - p[2]};
            const float mag = sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2];
            const float len = sqrtf(mag);
            if (len != 0.f) { const float r = 1.f / len; sv[0] *= r; sv[1] *= r; sv[2] *= r; }
            const float sdot = sv[0] * l[0] + sv[1] * l[1] + sv[2] * l[2];
            if (!(sdot > 0.f)) continue;
            mask |= 1u << k;
            q[k] = sdot / mag;
            if (a2 > 0.f) {
              // light_ref = normal3D(reflect3D(-sv, l)); spec_dot = dir . (-light_ref)
              const float tt = 2.f * sdot;                 // -(2 * dot(-sv, l))
              float lr[3] = {-sv[0] + l[0] * tt, -sv[1] + l[1] * tt, -sv[2] + l[2] * tt};
              const float ll = sqrtf(lr[0] * lr[0] + lr[1] * lr[1] + lr[2] * lr[2]);
              if (ll != 0.f) { const float r = 1.f / ll; lr[0] *= r; lr[1] *= r; lr[2] *= r; }
              const float sd = -(d[0] * lr[0] + d[1] * lr[1] + d[2] * lr[2]);
              if (sd > 0.f) sp[k] = powf(sd, m.specular_exponent);
            }
          }
        }
        if (mask) first_light = __ffs((int)mask) - 1;
        P.b.pos[node] = make_float4(p[0], p[1], p[2], __int_as_float(hit_i));
        P.b.base[node] = make_float4(rgb[0], rgb[1], rgb[2], __int_as_float((int)mask));
        P.b.lq[node] = make_float4(q[0], sp[0], q[1], sp[1]);
        if (P.b.lq2) P.b.lq2[node] = make_float4(q[2], sp[2], q[3], sp[3]);
        P.b.acc[node] = make_float4(P.sc.light_intensity, 0.f, 0.f, 0.f);
      }
      P.b.child_a[node] = make_float4(0.f, 0.f, 0.f, 0.f);
      P.b.child_b[node] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // ---- warp-compacted child allocation (north star: ballot/shuffle compaction of live rays) -------
    const uint32_t b1 = __ballot_sync(0xffffffffu, want_refl);
    const uint32_t b2 = __ballot_sync(0xffffffffu, want_refr);
    const uint32_t total = __popc(b1) + __popc(b2);
    if (total) {
      uint32_t cbase = 0;
      if (lane == 0) cbase = atomicAdd(&P.b.ctr->node_count, total);
      cbase = __shfl_sync(0xffffffffu, cbase, 0) + P.n_sub;
      const uint32_t pixsub = PRIMARY ? node : __float_as_uint(P.b.ray_d[active ? node : begin].w);
      if (want_refl) {
        const uint32_t c = cbase + __popc(b1 & lanemask_lt());
        if (c < P.b.cap_nodes) {
          P.b.ray_o[c] = make_float4(p[0], p[1], p[2], rdir[0]);
          P.b.ray_d[c] = make_float4(rdir[1], rdir[2], w_refl, __uint_as_float(pixsub));
          P.b.link[c] = make_int2((int)node, (int)((2u * path) | ((uint32_t)(segs_left - 1) << 24)));
        } else { P.b.ctr->overflow = 1u; P.sticky->overflow = 1u; }
      }
      if (want_refr) {
        const uint32_t c = cbase + __popc(b1) + __popc(b2 & lanemask_lt());
        if (c < P.b.cap_nodes) {
          P.b.ray_o[c] = make_float4(p[0], p[1], p[2], fdir[0]);
          P.b.ray_d[c] = make_float4(fdir[1], fdir[2], w_refr, __uint_as_float(pixsub));
          P.b.link[c] = make_int2((int)node, (int)((2u * path + 1u) | ((uint32_t)(segs_left - 1) << 24)));
        } else { P.b.ctr->overflow = 1u; P.sticky->overflow = 1u; }
      }
    }
    // ---- warp-compacted shadow queue push: first light that faces the surface -----------------------
    for (int k = 0; k < P.sc.n_lights; k++) {
      const uint32_t bk = __ballot_sync(0xffffffffu, first_light == k);
      if (!bk) continue;
      uint32_t qb = 0;
      if (lane == (uint32_t)(__ffs((int)bk) - 1)) qb = atomicAdd(&P.b.ctr->sq_count[k], (uint32_t)__popc(bk));
      qb = __shfl_sync(0xffffffffu, qb, __ffs((int)bk) - 1);
      if (first_light == k) P.b.sq[k][qb + __popc(bk & lanemask_lt())] = node;
    }
  }
  // stats
  for (int o2 = 16; o2 > 0; o2 >>= 1) my_nodes += __shfl_down_sync(0xffffffffu, my_nodes, o2);
  if (lane == 0 && my_nodes) {
    atomicAdd(&P.sticky->nodes, my_nodes);
    atomicAdd(&P.sticky->sphere_tests, my_nodes * (unsigned long long)n);
  }
  publish_level(P, level);
}

// ---- SHADOW ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WF_THREADS)
wf_shadow_kernel(const __grid_constant__ WfParams P, int k) {
  extern __shared__ float4 tab[];
  const int n = P.sc.n_spheres;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tab[i] = P.light_tab[k][i];
  __syncthreads();
  const uint32_t count = *((volatile uint32_t*)&P.b.ctr->sq_count[k]);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const float eps = P.sc.eps;
  const float lx = P.sc.lights[k][0], ly = P.sc.lights[k][1], lz = P.sc.lights[k][2];
  unsigned long long my_rays = 0, my_tests = 0;

  for (uint64_t base = (uint64_t)gwarp * 32u; base < count; base += (uint64_t)nwarps * 32u) {
    const uint32_t e = (uint32_t)base + lane;
    const bool active = e < count;
    uint32_t node = 0;
    int hit_i = -1;
    float w[3] = {0.f, 0.f, 1.f}, len = 0.f, I = 0.f;
    if (active) {
      node = P.b.sq[k][e];
      const float4 ps = P.b.pos[node];
      hit_i = __float_as_int(ps.w);
      // direction light -> p (= -shadow_vec) and light_len (main.js:287-290)
      w[0] = ps.x - lx; w[1] = ps.y - ly; w[2] = ps.z - lz;
      len = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
      if (len != 0.f) { const float r = 1.f / len; w[0] *= r; w[1] *= r; w[2] *= r; }
      I = P.b.acc[node].x;
      my_rays++;
    }
    bool live = active && I != 0.f;
    for (int j0 = 0; j0 < n; j0 += 64) {                  // main.js:293-304, all lanes share the table
      if (!__any_sync(0xffffffffu, live)) break;          // every lane already fully occluded
      const int j1 = min(n, j0 + 64);
      if (live) {
        my_tests += (unsigned long long)(j1 - j0);
#pragma unroll 4
        for (int j = j0; j < j1; j++) {
          const float4 s = tab[j];
          const float tca = w[0] * s.x + w[1] * s.y + w[2] * s.z;
          const float disc = tca * tca - s.w;
          if (disc >= 0.f && j != hit_i) {
            const float thc = sqrtf(disc);
            // parameters along the reference's ray (from hit.p toward the light): t = len - s
            const float t0 = len - (tca + thc), t1 = len - (tca - thc);
            const float t = (t0 >= eps) ? t0 : ((t1 >= eps) ? t1 : WF_INF);
            if (t < len) {
              const float a4 = P.sc.mtl[j].albedo[4];
              if (a4 != 0.f) I /= a4; else I = 0.f;      // main.js:298-302
            }
          }
        }
        if (I == 0.f) live = false;
      }
    }
    int next = -1;
    if (active) {
      float4 a = P.b.acc[node];
      a.x = I;
      const uint32_t mask = (uint32_t)__float_as_int(P.b.base[node].w);
      if (I != 0.f) {                                     // main.js:305-313
        const float4 lq = (k < 2) ? P.b.lq[node] : P.b.lq2[node];
        const float qk = (k & 1) ? lq.z : lq.x, sk = (k & 1) ? lq.w : lq.y;
        a.y += I * qk;
        a.z += sk;
      }
      P.b.acc[node] = a;
      const uint32_t rest = mask & ~((2u << k) - 1u);     // facing lights after k
      if (I != 0.f && rest) next = __ffs((int)rest) - 1;  // carry light_intensity to that light
    }
    __syncwarp();
    // warp-compacted hand-off; lanes may target different queues, so one ballot per later light
    for (int k2 = k + 1; k2 < P.sc.n_lights; k2++) {
      const uint32_t bk = __ballot_sync(0xffffffffu, next == k2);
      if (!bk) continue;
      const int leader = __ffs((int)bk) - 1;
      uint32_t qb = 0;
      if ((int)lane == leader) qb = atomicAdd(&P.b.ctr->sq_count[k2], (uint32_t)__popc(bk));
      qb = __shfl_sync(0xffffffffu, qb, leader);
      if (next == k2) P.b.sq[k2][qb + __popc(bk & lanemask_lt())] = node;
    }
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    my_rays += __shfl_down_sync(0xffffffffu, my_rays, o2);
    my_tests += __shfl_down_sync(0xffffffffu, my_tests, o2);
  }
  if (lane == 0 && my_rays) {
    atomicAdd(&P.sticky->shadow_rays, my_rays);
    atomicAdd(&P.sticky->sphere_tests, my_tests);
  }
}

// ---- COMPOSE --------------------------------------------------------------------------------------------
// main.js:316-336 for one node whose children have already deposited their colours.
RTB_DEVINL void node_colour(const WfParams& P, uint32_t node, float col[3]) {
  const float4 ps = P.b.pos[node];
  const float4 bs = P.b.base[node];
  const int hit_i = __float_as_int(ps.w);
  if (hit_i < 0) { col[0] = bs.x; col[1] = bs.y; col[2] = bs.z; return; }
  const DevMaterial<float>& m = P.sc.mtl[hit_i];
  const float4 a = P.b.acc[node];
  const float4 ca = P.b.child_a[node], cb = P.b.child_b[node];
  const float di = js_min<float>(1.f, a.y) * m.albedo[1];
  const float si = js_min<float>(1.f, a.z) * m.albedo[2];
  const float rgb[3] = {bs.x, bs.y, bs.z};
  const float ra[3] = {ca.x, ca.y, ca.z}, rb[3] = {cb.x, cb.y, cb.z};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float shader = rgb[c] * di + rgb[c] * si + ra[c] + rb[c];
    col[c] = js_max<float>(rgb[c] * m.albedo[0], js_min<float>(1.f, shader));
  }
}

__global__ void __launch_bounds__(WF_THREADS)
wf_compose_kernel(const __grid_constant__ WfParams P, int level) {
  uint32_t begin, end;
  level_range(P, level, begin, end);
  for (uint64_t i = (uint64_t)begin + blockIdx.x * blockDim.x + threadIdx.x; i < end;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t node = (uint32_t)i;
    float col[3];
    node_colour(P, node, col);
    const int2 lk = P.b.link[node];
    const float wgt = P.b.ray_d[node].z;                  // parent's albedo[3] or albedo[4]
    const float4 v = make_float4(col[0] * wgt, col[1] * wgt, col[2] * wgt, 0.f);   // main.js:271, 277
    if (lk.y & 1) P.b.child_b[lk.x] = v; else P.b.child_a[lk.x] = v;
  }
}

// Level 1: average the sub-samples of each pixel, quantise as a Uint8ClampedArray store would
// (main.js:195-198) and write 4 pixels per thread as one uint4.
__global__ void __launch_bounds__(WF_THREADS)
wf_resolve_kernel(const __grid_constant__ WfParams P, uint8_t* __restrict__ out) {
  const int W = P.fr64.width;
  const int groups_per_row = (W + 3) / 4;
  const long long total = (long long)P.band_rows * groups_per_row;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total;
       gi += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(gi / groups_per_row), x0 = (int)(gi % groups_per_row) * 4;
    const int lrow = P.band_l0 + r;
    uint32_t packed[4] = {0u, 0u, 0u, 0u};
    const int nx = min(4, W - x0);
    for (int i = 0; i < nx; i++) {
      const uint32_t pix = (uint32_t)r * (uint32_t)W + (uint32_t)(x0 + i);
      float acc[3] = {0.f, 0.f, 0.f};
      for (int s = 0; s < P.spp; s++) {
        float col[3];
        node_colour(P, pix * (uint32_t)P.spp + (uint32_t)s, col);
        if (P.spp == 1) { acc[0] = col[0]; acc[1] = col[1]; acc[2] = col[2]; }
        else { acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2]; }
      }
      if (P.spp > 1) { const float q = (float)P.spp; acc[0] /= q; acc[1] /= q; acc[2] /= q; }
      packed[i] = clamped_u8(255.f * acc[0]) | (clamped_u8(255.f * acc[1]) << 8) |
                  (clamped_u8(255.f * acc[2]) << 16) | (255u << 24);
    }
    uint32_t* row = reinterpret_cast<uint32_t*>(out) + P.rm.out_row(lrow) * (size_t)W;
    if (P.aligned4 && nx == 4) {
      *reinterpret_cast<uint4*>(row + x0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    } else {
      for (int i = 0; i < nx; i++) row[x0 + i] = packed[i];
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------
size_t per_node_bytes(int n_lights) {
  return 16 * 7 + 8 + (n_lights > 2 ? 16 : 0) + 4 * (size_t)n_lights;
}

int ensure_workspace(int n_lights, uint32_t cap_nodes) {
  WfBuffers& b = g_rtb.wf;
  const bool need_lq2 = n_lights > 2;
  if (b.cap_nodes >= cap_nodes && b.ctr && g_rtb.wf_lights >= n_lights && (!need_lq2 || b.lq2)) return RTB_OK;
  const uint32_t cap = std::max(cap_nodes, b.cap_nodes);
  const int lights = std::max(n_lights, g_rtb.wf_lights);
  wf_free();
  const size_t n = cap;
  RTB_CUDA(cudaMalloc(&b.ray_o, n * 16));
  RTB_CUDA(cudaMalloc(&b.ray_d, n * 16));
  RTB_CUDA(cudaMalloc(&b.link, n * 8));
  RTB_CUDA(cudaMalloc(&b.pos, n * 16));
  RTB_CUDA(cudaMalloc(&b.base, n * 16));
  RTB_CUDA(cudaMalloc(&b.lq, n * 16));
  if (lights > 2) RTB_CUDA(cudaMalloc(&b.lq2, n * 16));
  RTB_CUDA(cudaMalloc(&b.acc, n * 16));
  RTB_CUDA(cudaMalloc(&b.child_a, n * 16));
  RTB_CUDA(cudaMalloc(&b.child_b, n * 16));
  for (int k = 0; k < lights; k++) RTB_CUDA(cudaMalloc(&b.sq[k], n * 4));
  RTB_CUDA(cudaMalloc(&b.ctr, sizeof(WfCounters) + sizeof(WfSticky)));
  b.cap_nodes = cap;
  g_rtb.wf_lights = lights;
  g_rtb.wf_bytes = n * per_node_bytes(lights);
  return RTB_OK;
}

}  // namespace

void wf_free() {
  WfBuffers& b = g_rtb.wf;
  cudaFree(b.ray_o); cudaFree(b.ray_d); cudaFree(b.link); cudaFree(b.pos); cudaFree(b.base);
  cudaFree(b.lq); cudaFree(b.lq2); cudaFree(b.acc); cudaFree(b.child_a); cudaFree(b.child_b);
  for (int k = 0; k < RTB_MAX_LIGHTS; k++) cudaFree(b.sq[k]);
  cudaFree(b.ctr);
  b = WfBuffers{};
  g_rtb.wf_lights = 0;
  g_rtb.wf_bytes = 0;
}

int wf_render(RtbScene_* s, const RtbCamera* cam, const DevFrameT<float>& fr32, const RowMap& rm,
              uint8_t* d_out, cudaStream_t st, RtbStats* stats, bool want_counts) {
  (void)fr32; (void)want_counts;
  const int W = cam->width;
  const int n_rows = rm.n_local();
  if (n_rows <= 0) return RTB_OK;
  const int n = s->n_spheres;
  const size_t smem = (size_t)n * sizeof(float4);
  if (smem > g_rtb.smem_optin) {
    rtb_set_error("wavefront kernels stage all %d spheres in shared memory (%zu B > %zu B available)",
                  n, smem, g_rtb.smem_optin);
    return RTB_ERR_UNSUPPORTED;
  }
  if (n == 0) {
    rtb_set_error("wavefront path needs at least one sphere (use RTB_KERNEL_PIXEL)");
    return RTB_ERR_UNSUPPORTED;
  }

  // primary table: per-sphere constants of rays leaving the camera origin, in fp64, rounded once
  if (!s->primary_valid || s->primary_origin[0] != cam->origin[0] || s->primary_origin[1] != cam->origin[1] ||
      s->primary_origin[2] != cam->origin[2]) {
    for (int i = 0; i < n; i++) {
      const RtbSphere& sp = s->spheres[i];
      const double L0 = sp.origin[0] - cam->origin[0], L1 = sp.origin[1] - cam->origin[1],
                   L2 = sp.origin[2] - cam->origin[2];
      s->h_primary_tab[i] = make_float4((float)L0, (float)L1, (float)L2,
                                        (float)((L0 * L0 + L1 * L1 + L2 * L2) - sp.r2));
    }
    RTB_CUDA(cudaMemcpyAsync(s->d_primary_tab, s->h_primary_tab, smem, cudaMemcpyHostToDevice, st));
    // the pinned staging buffer is rewritten by the next camera: wait for the copy
    RTB_CUDA(cudaStreamSynchronize(st));
    for (int c = 0; c < 3; c++) s->primary_origin[c] = cam->origin[c];
    s->primary_valid = true;
  }

  const int g = fr32.spp_grid, spp = g * g;
  const uint64_t sub_per_row = (uint64_t)W * spp;
  const uint64_t n_sub_frame = sub_per_row * (uint64_t)n_rows;
  // queue capacity: room for ~2 extra nodes per primary ray, bounded; RTB_WF_CAP_NODES overrides (tests)
  uint64_t cap = std::min<uint64_t>(std::max<uint64_t>(3 * n_sub_frame, 1u << 20), 1u << 25);
  if (const char* e = getenv("RTB_WF_CAP_NODES")) { const long long v = atoll(e); if (v > 0) cap = (uint64_t)v; }
  if (cap < 4 * sub_per_row) cap = 4 * sub_per_row;      // at least one row per band with headroom
  if (cap > 0xfffffff0ull) { rtb_set_error("frame too wide for the wavefront queues"); return RTB_ERR_UNSUPPORTED; }
  int rc = ensure_workspace(s->n_lights, (uint32_t)cap);
  if (rc != RTB_OK) return rc;
  // the shared workspace may be larger than this call asked for; band sizes and the overflow
  // check follow the REQUESTED capacity so that results do not depend on earlier calls

  WfParams P;
  memset(&P, 0, sizeof P);
  P.sc = s->sc32;
  // fp64 camera for primary rays
  DevFrameT<double>& f = P.fr64;
  f.width = cam->width; f.height = cam->height; f.segs = fr32.segs; f.spp_grid = g;
  f.star_mode = fr32.star_mode; f.camera_mode = cam->camera_mode; f.star_seed = fr32.star_seed;
  for (int c = 0; c < 3; c++) {
    f.origin[c] = cam->origin[c]; f.axisX[c] = cam->axisX[c]; f.axisY[c] = cam->axisY[c]; f.axisZ[c] = cam->axisZ[c];
    P.origin[c] = (float)cam->origin[c];
  }
  f.projW = cam->projW; f.projH = cam->projH; f.projD = cam->projD;
  P.rm = rm;
  P.b = g_rtb.wf;
  P.b.cap_nodes = (uint32_t)cap;
  if (s->n_lights <= 2) P.b.lq2 = nullptr;                // kernels key the lights-2/3 store on this
  P.sticky = reinterpret_cast<WfSticky*>(reinterpret_cast<char*>(g_rtb.wf.ctr) + sizeof(WfCounters));
  P.geom = reinterpret_cast<const float4*>(s->d_geom32);
  P.primary_tab = s->d_primary_tab;
  for (int k = 0; k < s->n_lights; k++) P.light_tab[k] = s->d_light_tab[k];
  P.spp = spp;
  P.aligned4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15u) == 0);

  // persistent grids: as many CTAs as fit, per kernel
  auto grid_for = [&](const void* fn, size_t sm) -> int {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, WF_THREADS, sm);
    if (per_sm < 1) per_sm = 1;
    return per_sm * g_rtb.sm_count;
  };
  if (smem > 48 * 1024) {
    RTB_CUDA(cudaFuncSetAttribute(wf_trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RTB_CUDA(cudaFuncSetAttribute(wf_trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RTB_CUDA(cudaFuncSetAttribute(wf_shadow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int grid_p = grid_for((const void*)wf_trace_kernel<true>, smem);
  const int grid_s = grid_for((const void*)wf_trace_kernel<false>, smem);
  const int grid_l = grid_for((const void*)wf_shadow_kernel, smem);
  const int grid_c = g_rtb.sm_count * 8;

  int rows_per_band = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)n_rows, (cap / 3) / sub_per_row));
  WfSticky h_sticky;
  uint32_t launches = 0, bands = 0;
  // per-class timing (stats only): an event after every class of every band, classes tagged
  enum { T_START = 0, T_PRIMARY, T_SECONDARY, T_SHADOW, T_COMPOSE };
  std::vector<int> tags;
  auto mark = [&](int tag) -> int {
    if (!stats) return RTB_OK;
    if ((int)tags.size() >= RtbGlobal::kMaxTimers) return RTB_OK;   // long band lists: later bands untimed
    if ((int)tags.size() >= g_rtb.n_timers) {
      RTB_CUDA(cudaEventCreate(&g_rtb.timers[g_rtb.n_timers]));
      g_rtb.n_timers++;
    }
    RTB_CUDA(cudaEventRecord(g_rtb.timers[tags.size()], st));
    tags.push_back(tag);
    return RTB_OK;
  };
  for (int attempt = 0;; attempt++) {
    launches = 0; bands = 0;
    tags.clear();
    RTB_CUDA(cudaMemsetAsync(P.sticky, 0, sizeof(WfSticky), st));
    for (int l0 = 0; l0 < n_rows; l0 += rows_per_band) {
      P.band_l0 = l0;
      P.band_rows = std::min(rows_per_band, n_rows - l0);
      P.n_sub = (uint32_t)((uint64_t)P.band_rows * sub_per_row);
      bands++;
      RTB_CUDA(cudaMemsetAsync(g_rtb.wf.ctr, 0, sizeof(WfCounters), st));
      if ((rc = mark(T_START)) != RTB_OK) return rc;
      wf_trace_kernel<true><<<grid_p, WF_THREADS, smem, st>>>(P, 1);
      launches++;
      if ((rc = mark(T_PRIMARY)) != RTB_OK) return rc;
      for (int L = 2; L <= fr32.segs; L++) {
        wf_trace_kernel<false><<<grid_s, WF_THREADS, smem, st>>>(P, L);
        launches++;
      }
      if ((rc = mark(T_SECONDARY)) != RTB_OK) return rc;
      for (int k = 0; k < s->n_lights; k++) {
        wf_shadow_kernel<<<grid_l, WF_THREADS, smem, st>>>(P, k);
        launches++;
      }
      if ((rc = mark(T_SHADOW)) != RTB_OK) return rc;
      for (int L = fr32.segs; L >= 2; L--) {
        wf_compose_kernel<<<grid_c, WF_THREADS, 0, st>>>(P, L);
        launches++;
      }
      wf_resolve_kernel<<<grid_c, WF_THREADS, 0, st>>>(P, d_out);
      launches++;
      if ((rc = mark(T_COMPOSE)) != RTB_OK) return rc;
      RTB_CUDA(cudaGetLastError());
    }
    RTB_CUDA(cudaMemcpyAsync(&h_sticky, P.sticky, sizeof h_sticky, cudaMemcpyDeviceToHost, st));
    RTB_CUDA(cudaStreamSynchronize(st));
    if (!h_sticky.overflow) break;
    if (rows_per_band == 1) {
      rtb_set_error("ray queues overflow even with one-row bands (cap_nodes=%llu, row=%llu sub-samples)",
                    (unsigned long long)cap, (unsigned long long)sub_per_row);
      return RTB_ERR_OVERFLOW;
    }
    rows_per_band = std::max(1, rows_per_band / 2);       // deterministic: smaller bands need fewer nodes
  }
  if (stats) {
    stats->nodes += h_sticky.nodes;
    stats->shadow_rays += h_sticky.shadow_rays;
    stats->sphere_tests += h_sticky.sphere_tests;
    stats->launches += launches;
    stats->bands += bands;
    for (size_t i = 1; i < tags.size(); i++) {
      if (tags[i] == T_START) continue;
      float ms = 0.f;
      RTB_CUDA(cudaEventElapsedTime(&ms, g_rtb.timers[i - 1], g_rtb.timers[i]));
      if (tags[i] == T_PRIMARY) stats->ms_trace_primary += ms;
      else if (tags[i] == T_SECONDARY) stats->ms_trace_secondary += ms;
      else if (tags[i] == T_SHADOW) stats->ms_shadow += ms;
      else stats->ms_compose += ms;
    }
  }
  return RTB_OK;
}
