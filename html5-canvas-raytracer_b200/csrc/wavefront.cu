// wavefront.cu — fp32 THROUGHPUT path: the Whitted tree of main.js:220-337 evaluated
// breadth-first over a whole row band — of one frame, or of a GROUP of up to 8 frames of a sequence at once
// (rtb_render_device_group: level 1 holds the primary rays of every frame, all later work is shared) —
// instead of depth-first per pixel.  One recursion level = a SWEEP kernel + a HIT kernel:
//
//   RAYGEN         : primary rays (main.js:186-193) in fp64 with each frame's own camera, rounded once.
//   SWEEP  level L : several rays per lane; closest-hit sweep over ALL spheres in array order from shared
//                    memory (main.js:223-230); writes key[node] = (t, sphere, flip).  Nothing else: every
//                    resident warp is in the FMA-bound loop.
//   HIT    level L : one thread per ray at high occupancy: the hit record, reflect/refract directions
//                    (main.js:233-266), the sampler (main.js:320) and the occlusion-independent part of the
//                    light loop (main.js:286-292, 306-313).  Live child rays and shadow-ray RECORDS are
//                    compacted into queues with warp ballots, queue space reserved once per 256-node tile.
//   SSWEEP light k : occluder sweep main.js:293-304 for every queued record, traced FROM the light so that
//                    all rays of the launch share an origin.
//   SPOST  light k : diffuse / specular sums (main.js:305-313) and the light_intensity carry-over between
//                    lights (main.js:284, H1-b) = a new record in the queue of the next facing light.
//   COMPOSE level L: post-order clamp-compose main.js:322-336, deepest level first; each child
//                    writes its colour (scaled by the parent's albedo[3|4]) into its parent; level 1
//                    averages sub-samples and stores RGBA8 as uint4 (4 pixels per thread).
//
// The round-1 FUSED forms (wf_trace_kernel = sweep + hit processing in one persistent kernel, wf_shadow_kernel =
// sweep + post) are kept behind RTB_WF_SPLIT as the cross-check: both pipelines give byte-identical frames.
//
// The sweeps are written for the B200 SM as measured (tools/fp32_peak.cu, profiles/):
//   * a warp-wide LDS.128 costs 4 LSU cycles even when all lanes read the same 16 bytes, so every
//     sphere fetched from shared memory is tested against R rays held in registers (R = 8 for the
//     shared-origin sweeps, 4 for general rays) — otherwise the kernel is LSU-bound, not FP32-bound;
//   * tests run two spheres at a time on the packed FP32 pipe (FFMA2 / FADD2 / FMUL2, sm_100 f32x2):
//     half the issue slots of scalar FFMA and no 3-source register-bank conflicts
//     (measured: scalar dot-shaped FFMA 76 % of peak, FFMA2 89 %);
//   * the fast path never branches per test: the sign bit of each test value is funnel-shifted into
//     a per-ray 32-bit mask (one ALU-pipe instruction, off the FMA pipe) and only after 32 spheres, when the
//     masks are full, does the warp look at them (a single branch); the slow path visits exactly the set bits
//     (exact discriminant from shared memory, sqrt, the 0.001 window of main.js:429-439) and keeps the closest
//     hit as a (t, index) key.
//
// Every sweep is a persistent kernel that pulls fixed-size chunks from a device counter; bounds are
// read from device counters too, so a band is a fixed launch sequence without host round trips.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <map>
#include <utility>
#include "rtb_internal.h"
#include "pixel_kernel.cuh"   // primary_ray<>

namespace {

struct WfSticky {             // survives across bands of one call
  uint32_t overflow;
  uint32_t _pad;
  unsigned long long nodes, shadow_rays, sphere_tests;
};

// One camera of a frame group (the closure variables spanish() reads, main.js:85-105), fp64: primary rays
// are generated in fp64 and rounded once.  Field names are those primary_ray<>() reads.
struct WfCam {
  double origin[3], axisX[3], axisY[3], axisZ[3];
  double projW, projH, projD;
  int32_t spp_grid, camera_mode;
};

// A wavefront pass renders the same row band of n_frames frames of a sequence at once (a FRAME GROUP):
// level 1 holds the primary rays of all of them, frame by frame, and every later level, shadow queue and
// compose pass is shared - so that the kernels of a striped multi-GPU frame (1/world of the rows) are as
// large as those of a whole frame.  Frames of a group share scene, size and options; cameras may differ.
struct WfParams {
  DevSceneT<float> sc;
  DevFrameT<double> fr64;     // size / options shared by the group (camera fields unused: see cam[])
  WfCam cam[RTB_MAX_GROUP];
  float origin32[RTB_MAX_GROUP][3];   // camera origins rounded once (ray origins of level 1)
  uint8_t* out[RTB_MAX_GROUP];   // one framebuffer per frame
  RowMap rm;
  WfBuffers b;
  WfSticky* sticky;
  const float4* geom;         // exact Cx, Cy, Cz, r2                                  (n_spheres)
  // packed sweep tables (see wf_make_table): per sphere PAIR two float4, padded with never-hit dummies
  const float4* geom_tab;     // (x0,x1,y0,y1) (z0,z1,-r2_0',-r2_1')
  // primary tables, one per run of consecutive frames that share a camera origin:
  const float4* primary_tab[RTB_MAX_GROUP];  // (Lx0,Lx1,Ly0,Ly1) (Lz0,Lz1,sc0,sc1), L = C - origin
  const float4* primary_x[RTB_MAX_GROUP];    // exact (L, |L|^2 - r2) for the slow path  (n_spheres)
  const float4* light_tab[RTB_MAX_LIGHTS];
  const float4* light_x[RTB_MAX_LIGHTS];
  int32_t n_pad;
  int32_t n_frames;           // frames in the group (1 .. RTB_MAX_GROUP)
  int32_t n_runs;             // origin runs; run g = frames [run_begin[g], run_begin[g + 1])
  int32_t run_begin[RTB_MAX_GROUP + 1];
  int32_t band_l0;            // first local row of the band
  int32_t band_rows;
  uint32_t n_sub_frame;       // band_rows * W * spp == level-1 nodes of ONE frame
  uint32_t n_sub;             // n_frames * n_sub_frame == number of level-1 nodes
  int32_t spp;                // g*g
  int32_t aligned4;           // W % 4 == 0 and outputs 16-byte aligned: uint4 stores
  int32_t fixed_r;            // 0: rays per lane from the cost model; 1/2/3: always R, R/2, R/4 and no guided tail
};

constexpr int WF_THREADS = 256;
#ifndef WF_MIN_CTAS
#define WF_MIN_CTAS 2
#endif
#ifndef WF_MIN_CTAS_SHADOW
#define WF_MIN_CTAS_SHADOW 2
#endif
#ifndef RTB_WF_SPLIT_DEFAULT
#define RTB_WF_SPLIT_DEFAULT 7
#endif
#ifndef WF_MIN_CTAS_SWEEP_PRIMARY
#define WF_MIN_CTAS_SWEEP_PRIMARY 2
#endif
#ifndef WF_MIN_CTAS_HIT
#define WF_MIN_CTAS_HIT 3
#endif
#ifndef WF_MIN_CTAS_GENERAL
#define WF_MIN_CTAS_GENERAL 3            // general rays: 4 per lane fit 80 registers (measured 3.93 -> 3.66 ms per 4K frame)
#endif
constexpr float WF_INF = __builtin_huge_valf();
constexpr int WF_CHECK = 32;             // spheres per CANDIDATE CHECK: the sign bits of 32 tests fill a ray's mask, and only
                                         // then does the warp look at them (one OR + branch, and the per-ray visit of set
                                         // bits if any) - with a check after every 8 the bookkeeping around the branch was
                                         // ~13 % of the shadow sweep's and ~6 % of the trace sweeps' warp time
constexpr int WF_GROUP = 8;              // spheres per unrolled body; the body (4 sphere pairs x R
                                         // rays) must stay well inside the ~6 KB L0 instruction cache
static_assert(RTB_WF_GROUP % WF_CHECK == 0 && WF_CHECK % WF_GROUP == 0, "table padding must be a multiple of the check interval");
#ifndef WF_R_SHARED
#define WF_R_SHARED 8
#endif
#ifndef WF_R_GENERAL
#define WF_R_GENERAL 4
#endif
constexpr int R_SHARED = WF_R_SHARED;    // rays per lane, shared-origin sweeps (primary, shadow)
constexpr int R_GENERAL = WF_R_GENERAL;  // rays per lane, general sweep (reflect / refract rays)
constexpr unsigned long long KEY_NONE = 0x7f800000ffffffffull;   // t = +inf, no sphere
constexpr int WF_WARPS = WF_THREADS / 32;

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
      if (level == 1) P.b.ctr->sq_snap0 = atomicAdd(&P.b.ctr->sq_count[0], 0u);   // early shadow pass range
      P.b.ctr->blocks_done = 0;
      __threadfence();
    }
  }
}

// Persistent-kernel work distribution: one atomic per warp per chunk.
RTB_DEVINL uint32_t next_chunk(uint32_t* counter, uint32_t lane) {
  uint32_t c = 0;
  if (lane == 0) c = atomicAdd(counter, 1u);
  return __shfl_sync(0xffffffffu, c, 0);
}

// First claim of a persistent CTA: eight consecutive chunks at once, one per warp.  A level that does
// not fill the grid is then worked on by ceil(chunks / 8) CTAs with all their warps busy, and the other
// CTAs leave immediately, instead of every CTA staying resident for the sake of one or two busy warps —
// which matters when the kernels of several frames share the GPU.  (Contains a CTA barrier.)
RTB_DEVINL uint32_t cta_claim(uint32_t* counter) {
  __shared__ uint32_t s_first;
  if (threadIdx.x == 0) s_first = atomicAdd(counter, (uint32_t)WF_WARPS);
  __syncthreads();
  return s_first + (threadIdx.x >> 5);
}

RTB_DEVINL float2 f2(float a, float b) { return make_float2(a, b); }
// (mask << 1) | sign(v): one SHF on the ALU pipe
RTB_DEVINL uint32_t push_sign(uint32_t mask, float v) { return __funnelshift_l(__float_as_uint(v), mask, 1); }
// closest-hit key: (t bits, sphere index, flip) — min over keys == smallest t, then lowest index
RTB_DEVINL unsigned long long hit_key(float t, int idx, bool flip) {
  return ((unsigned long long)__float_as_uint(t) << 32) | ((uint32_t)idx << 1) | (flip ? 1u : 0u);
}

// Exact part of intersectSphere (main.js:426-439) from tca and the discriminant r2 - d2.
RTB_DEVINL void exact_update(float tca, float disc, float eps, int jj, unsigned long long& best) {
  if (disc >= 0.f) {
    const float thc = sqrtf(disc);
    const float t0 = tca - thc, t1 = tca + thc;
    const float t = (t0 >= eps) ? t0 : ((t1 >= eps) ? t1 : WF_INF);
    if (t < WF_INF) {
      const unsigned long long k = hit_key(t, jj, t0 < eps);   // main.js:445: l flips when t0 < eps
      if (k < best) best = k;
    }
  }
}

// ---- slow paths (out of line: the hot loops must stay small for the instruction cache) ----------------------
// Visit exactly the candidate spheres of one ray in group j (mask bit b <-> sphere j + GROUP-1-b).
__device__ __noinline__ unsigned long long slow_primary(const float4* __restrict__ exact, int j, uint32_t m,
                                                        float ndx, float ndy, float ndz, float eps,
                                                        unsigned long long best) {
  while (m) {
    const int b = 31 - __clz((int)m);
    m ^= 1u << b;
    const int jj = j + (WF_CHECK - 1) - b;
    const float4 x = exact[jj];
    const float tca = -(ndx * x.x + ndy * x.y + ndz * x.z);
    exact_update(tca, tca * tca - x.w, eps, jj, best);
  }
  return best;
}

// The ray itself (origin, direction) is re-read from the queue: the fast path only keeps its (u, v, ku, kv).
__device__ __noinline__ unsigned long long slow_general(const float4* __restrict__ geom,
                                                        const float4* __restrict__ ray_o,
                                                        const float4* __restrict__ ray_d, int j, uint32_t m,
                                                        uint32_t node, int self, float eps,
                                                        unsigned long long best) {
  const float4 ro = ray_o[node], rd = ray_d[node];
  const float nox = -ro.x, noy = -ro.y, noz = -ro.z, dx = ro.w, dy = rd.x, dz = rd.y;
  while (m) {
    const int b = 31 - __clz((int)m);
    m ^= 1u << b;
    const int jj = j + (WF_CHECK - 1) - b;
    const float4 s = geom[jj];
    const float L0 = s.x + nox, L1 = s.y + noy, L2 = s.z + noz;
    const float tca = dx * L0 + dy * L1 + dz * L2;
    if (jj == self) {
      // The ray starts ON this sphere (hit.p of its parent).  With p on the surface the roots of
      // main.js:427-428 are exactly {0, 2 tca}: t0 < 0.001 always, so the hit is t1 = 2 tca when that
      // clears the 0.001 window (the ray crosses the inside) and l is flipped.  The float64 reference
      // computes precisely this (p is on the surface to 1e-15); in fp32 p sits up to ~5e-5 off the
      // surface, and the generic formula then reports false grazing self-hits beyond the window.
      const float t1 = 2.f * tca;
      if (t1 >= eps) { const unsigned long long k = hit_key(t1, jj, true); if (k < best) best = k; }
      continue;
    }
    // d2 as |L - tca d|^2: no cancellation between |L|^2 and tca^2 for the r = 500 / 5000 spheres
    const float v0 = L0 - tca * dx, v1 = L1 - tca * dy, v2 = L2 - tca * dz;
    exact_update(tca, s.w - (v0 * v0 + v1 * v1 + v2 * v2), eps, jj, best);
  }
  return best;
}

// Occluders of one shadow ray among the candidates of group j; returns the updated light_intensity
// (main.js:297-302).  t runs along the reference's ray (hit.p -> light): t = len - s.
__device__ __noinline__ float slow_shadow(const float4* __restrict__ exact, const DevMaterial<float>* __restrict__ mtl,
                                          int j, uint32_t m, float nwx, float nwy, float nwz, float len, int hit_i,
                                          float eps, float I) {
  while (m) {
    const int b = 31 - __clz((int)m);
    m ^= 1u << b;
    const int jj = j + (WF_CHECK - 1) - b;
    if (jj == hit_i) continue;                              // main.js:294 — no shadow on self
    const float4 x = exact[jj];
    const float tca = -(nwx * x.x + nwy * x.y + nwz * x.z);
    const float disc = tca * tca - x.w;
    if (disc >= 0.f) {
      const float thc = sqrtf(disc);
      const float t0 = len - (tca + thc), t1 = len - (tca - thc);
      const float t = (t0 >= eps) ? t0 : ((t1 >= eps) ? t1 : WF_INF);
      if (t < len) {
        const float a4 = mtl[jj].albedo[4];
        if (a4 != 0.f) I /= a4; else return 0.f;
      }
    }
  }
  return I;
}

// ---- sweeps ------------------------------------------------------------------------------------------------
// Shared-origin closest hit (primary rays).  nd = -dir.  Fast test per sphere: sc - dir.L <= 0, i.e.
// "the line hits the sphere and the sphere is ahead" (sc = sqrt(|L|^2 - r2), fp64 on the host, -inf
// when the origin is inside).  3 FFMA2 per sphere pair per ray.
template <int R>
RTB_DEVINL void sweep_primary(const float4* __restrict__ tab, const float4* __restrict__ exact, int j_begin,
                              int j_end, const float (&nd)[R][3], const bool (&active)[R],
                              unsigned long long (&best)[R], float eps) {
  uint32_t mask[R];
#pragma unroll
  for (int q = 0; q < R; q++) { best[q] = KEY_NONE; mask[q] = 0; }
  for (int j = j_begin; j < j_end; j += WF_CHECK) {
#pragma unroll 1
    for (int jg = j; jg < j + WF_CHECK; jg += WF_GROUP) {
#pragma unroll
      for (int u = 0; u < WF_GROUP / 2; u++) {
        const float4 A = tab[jg + 2 * u], B = tab[jg + 2 * u + 1];
#pragma unroll
        for (int q = 0; q < R; q++) {
          float2 t = __ffma2_rn(f2(nd[q][2], nd[q][2]), f2(B.x, B.y), f2(B.z, B.w));
          t = __ffma2_rn(f2(nd[q][1], nd[q][1]), f2(A.z, A.w), t);
          t = __ffma2_rn(f2(nd[q][0], nd[q][0]), f2(A.x, A.y), t);
          mask[q] = push_sign(push_sign(mask[q], t.x), t.y);
        }
      }
    }
    uint32_t any = 0;                                      // the 32 pushes have replaced every bit of every mask
#pragma unroll
    for (int q = 0; q < R; q++) any |= mask[q];
    if (any) {
#pragma unroll
      for (int q = 0; q < R; q++) {
        const uint32_t m = mask[q];
        if (m && active[q]) best[q] = slow_primary(exact, j, m, nd[q][0], nd[q][1], nd[q][2], eps, best[q]);
      }
    }
  }
}

// General closest hit (reflect / refract rays).  The reference tests d2 = |L|^2 - tca^2 <= r2
// (main.js:422-425); with two unit vectors u, v that complete dir to an orthonormal basis,
// d2 = (u.L)^2 + (v.L)^2 EXACTLY (L = tca dir + (u.L) u + (v.L) v), which needs neither L nor tca and has
// no cancellation between |L|^2 and tca^2 (r = 500 / 5000 spheres).  With ku = u.org, kv = v.org per
// ray:  a = u.C - ku, b = v.C - kv, e = a^2 + b^2 - r2'  ->  8 FFMA2 per sphere pair per ray.
template <int R>
RTB_DEVINL void sweep_general(const float4* __restrict__ tab, const float4* __restrict__ geom,
                              const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, int j_begin,
                              int j_end, const float (&u)[R][3], const float (&v)[R][3], const float (&nku)[R],
                              const float (&nkv)[R], const uint32_t (&node)[R], const int (&self)[R],
                              const bool (&active)[R], unsigned long long (&best)[R], float eps) {
  uint32_t mask[R];
#pragma unroll
  for (int q = 0; q < R; q++) { best[q] = KEY_NONE; mask[q] = 0; }
  for (int j = j_begin; j < j_end; j += WF_CHECK) {
#pragma unroll 1
    for (int jg = j; jg < j + WF_CHECK; jg += WF_GROUP) {
#pragma unroll
    for (int p = 0; p < WF_GROUP / 2; p++) {
      const float4 A = tab[jg + 2 * p], B = tab[jg + 2 * p + 1];
#pragma unroll
      for (int q = 0; q < R; q++) {
        float2 a = __ffma2_rn(f2(u[q][2], u[q][2]), f2(B.x, B.y), f2(nku[q], nku[q]));
        a = __ffma2_rn(f2(u[q][1], u[q][1]), f2(A.z, A.w), a);
        a = __ffma2_rn(f2(u[q][0], u[q][0]), f2(A.x, A.y), a);
        float2 b = __ffma2_rn(f2(v[q][2], v[q][2]), f2(B.x, B.y), f2(nkv[q], nkv[q]));
        b = __ffma2_rn(f2(v[q][1], v[q][1]), f2(A.z, A.w), b);
        b = __ffma2_rn(f2(v[q][0], v[q][0]), f2(A.x, A.y), b);
        float2 e = __ffma2_rn(a, a, f2(B.z, B.w));                             // a^2 - r2'
        e = __ffma2_rn(b, b, e);                                               // d2 - r2' <= 0: candidate
        mask[q] = push_sign(push_sign(mask[q], e.x), e.y);
      }
    }
    }
    uint32_t any = 0;
#pragma unroll
    for (int q = 0; q < R; q++) any |= mask[q];
    if (any) {
#pragma unroll
      for (int q = 0; q < R; q++) {
        const uint32_t m = mask[q];
        if (m && active[q]) best[q] = slow_general(geom, ray_o, ray_d, j, m, node[q], self[q], eps, best[q]);
      }
    }
  }
}

// ---- per-node work after the sweep ---------------------------------------------------------------------------
struct WfRay {
  float o[3], d[3];
  uint32_t node, path, pixsub;
  int segs_left;
  bool active;
};

// Everything intersectWorld does with the closest hit except the occluder sweeps and the final
// compose: hit record (main.js:441-447), reflect / refract directions (233-266), sampler (320), the
// occlusion-independent light terms (286-292, 306-313); then warp-compacted queue pushes.
// Must be called by all 32 lanes together (ballots inside).
// Shadow ray of light k toward p as the occluder sweep wants it: a = (nw, light_len), nw = -(direction
// light -> p) (main.js:287-290; the sweep runs from the light).
RTB_DEVINL float4 shadow_record(const WfParams& P, int k, float px, float py, float pz) {
  float wx = px - P.sc.lights[k][0], wy = py - P.sc.lights[k][1], wz = pz - P.sc.lights[k][2];
  const float len = sqrtf(wx * wx + wy * wy + wz * wz);
  if (len != 0.f) { const float rr = 1.f / len; wx *= rr; wy *= rr; wz *= rr; }
  return make_float4(-wx, -wy, -wz, len);
}

// Queue space for a warp's pushes: counts[0] = child nodes, counts[1 + k] = entries of shadow queue k; on
// return counts[] holds the warp's first index in each queue.
// AGG = false (fused persistent kernels: warps run independently): one atomic per warp and counter.
// AGG = true (hit / post kernels: the warps of a CTA walk the same 256-node tile together): the counts of the
// CTA's warps are summed in shared memory and ONE atomic per counter reserves space for the whole tile - the
// level counters are single addresses, and with one atomic per warp a kernel that does nothing but hit
// processing runs at the L2's same-address atomic rate (measured: 8.3 M nodes in 532 us, 252 us per tile).
// Contains CTA barriers when AGG.
template <bool AGG>
RTB_DEVINL void reserve(const WfParams& P, uint32_t (&counts)[1 + RTB_MAX_LIGHTS], int n_lights, int first_light,
                        uint32_t lane) {
  if (!AGG) {
#pragma unroll
    for (int j = 0; j < 1 + RTB_MAX_LIGHTS; j++) {
      if (j > n_lights || j < first_light) continue;
      uint32_t* counter = j == 0 ? &P.b.ctr->node_count : &P.b.ctr->sq_count[j - 1];
      uint32_t base = 0;
      if (lane == 0 && counts[j]) base = atomicAdd(counter, counts[j]);
      counts[j] = __shfl_sync(0xffffffffu, base, 0);
    }
    return;
  }
  __shared__ uint32_t s_cnt[1 + RTB_MAX_LIGHTS][WF_WARPS];
  const uint32_t w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 1 + RTB_MAX_LIGHTS; j++) s_cnt[j][w] = counts[j];
  }
  __syncthreads();
  if ((int)threadIdx.x <= n_lights && (int)threadIdx.x >= first_light) {   // thread j owns counter j
    const int j = threadIdx.x;
    uint32_t sum = 0;
    for (int i = 0; i < WF_WARPS; i++) { const uint32_t c = s_cnt[j][i]; s_cnt[j][i] = sum; sum += c; }
    uint32_t* counter = j == 0 ? &P.b.ctr->node_count : &P.b.ctr->sq_count[j - 1];
    const uint32_t base = sum ? atomicAdd(counter, sum) : 0u;
    for (int i = 0; i < WF_WARPS; i++) s_cnt[j][i] += base;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 1 + RTB_MAX_LIGHTS; j++) counts[j] = s_cnt[j][w];
  // no third barrier: a warp only ever rewrites its OWN column before the next tile's first barrier, and the
  // scanning threads touch the other columns only after that barrier, which every reader has then passed
}

// RECORDS = false: a queued shadow ray is the node index (the fused shadow kernel follows queue -> pos -> acc).
// RECORDS = true : it is a complete record (direction, length, node, sphere it leaves, light_intensity carried
//                  so far) that the occluder sweep kernel reads with two coalesced loads.
// AGG: see reserve().  Must be called by all 32 lanes together (ballots inside), with AGG by the whole CTA.
template <bool RECORDS, bool AGG>
RTB_DEVINL void process_hit_impl(const WfParams& P, const WfRay& r, unsigned long long key, uint32_t lane) {
  const bool active = r.active;
  const uint32_t node = r.node;
  const float* o = r.o;
  const float* d = r.d;
  const float best = __uint_as_float((uint32_t)(key >> 32));
  const int hit_i = (best < WF_INF) ? (int)(((uint32_t)key) >> 1) : -1;
  const bool flip = (key & 1ull) != 0;
  const int W = P.fr64.width;
  bool want_refl = false, want_refr = false;
  float p[3] = {0.f, 0.f, 0.f}, rdir[3] = {0.f, 0.f, 0.f}, fdir[3] = {0.f, 0.f, 0.f};
  float w_refl = 0.f, w_refr = 0.f;
  int first_light = -1;
  if (active) {
    if (hit_i < 0) {                                      // main.js:231
      P.b.pos[node] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      P.b.base[node] = make_float4(1.f, 0.f, 0.f, __int_as_float(0));
    } else {
      const float4 g = __ldg(P.geom + hit_i);
      const DevMaterial<float>& m = P.sc.mtl[hit_i];
      p[0] = o[0] + d[0] * best; p[1] = o[1] + d[1] * best; p[2] = o[2] + d[2] * best;
      float nn[3] = {p[0] - g.x, p[1] - g.y, p[2] - g.z};
      const float nl = nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2];
      if (nl != 0.f) { const float q = rsqrtf(nl); nn[0] *= q; nn[1] *= q; nn[2] *= q; }
      const float l[3] = {flip ? -nn[0] : nn[0], flip ? -nn[1] : nn[1], flip ? -nn[2] : nn[2]};
      const float a1 = m.albedo[1], a2 = m.albedo[2], a3 = m.albedo[3], a4 = m.albedo[4];
      const float dn = d[0] * nn[0] + d[1] * nn[1] + d[2] * nn[2];
      if (a3 > 0.f) {                                     // main.js:233-239
        const float tt = -(2.f * dn);
        rdir[0] = d[0] + nn[0] * tt; rdir[1] = d[1] + nn[1] * tt; rdir[2] = d[2] + nn[2] * tt;
        const float len = rdir[0] * rdir[0] + rdir[1] * rdir[1] + rdir[2] * rdir[2];
        if (len != 0.f) {
          const float q = rsqrtf(len); rdir[0] *= q; rdir[1] *= q; rdir[2] *= q;
          want_refl = r.segs_left > 1;                    // a child with segs == 0 returns [0,0,0]
          w_refl = a3;
        }
      }
      if (a4 > 0.f) {                                     // main.js:241-266
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
        const float len = fdir[0] * fdir[0] + fdir[1] * fdir[1] + fdir[2] * fdir[2];
        if (len != 0.f) {
          const float q = rsqrtf(len); fdir[0] *= q; fdir[1] *= q; fdir[2] *= q;
          want_refr = r.segs_left > 1;
          w_refr = a4;
        }
      }
      // sampler (main.js:320); u,v only matter to the texture sampler
      float u = 0.f, v = 0.f, rgb[3];
      if (m.sampler_kind == RTB_SAMPLER_TEXTURE) {
        u = atan2f(-nn[2], -nn[0]) * 0.15915494309189535f + 0.5f;     // / pi / 2
        v = asinf(-nn[1]) * 0.31830988618379069f + 0.5f;              // / (pi / 2) / 2
      }
      uint32_t px = 0, py = 0, sub = 0;
      if (m.sampler_kind == RTB_SAMPLER_STARS && P.fr64.star_mode == RTB_STARS_HASH) {
        const uint32_t in_frame = r.pixsub % P.n_sub_frame;   // star hash key: pixel, sub-sample, path
        const uint32_t pix = in_frame / (uint32_t)P.spp;
        sub = in_frame - pix * (uint32_t)P.spp;
        px = pix % (uint32_t)W;
        py = (uint32_t)P.rm.local_to_y(P.band_l0 + (int)(pix / (uint32_t)W));
      }
      run_sampler<float>(P.sc, m, nn, u, v, P.fr64.star_mode, P.fr64.star_seed, px, py, sub, r.path, rgb);
      // occlusion-independent light terms (main.js:286-292, 306-313)
      uint32_t mask = 0;
      float q[RTB_MAX_LIGHTS] = {0.f, 0.f, 0.f, 0.f}, sp[RTB_MAX_LIGHTS] = {0.f, 0.f, 0.f, 0.f};
      if (a1 > 0.f || a2 > 0.f) {
        for (int k = 0; k < P.sc.n_lights; k++) {
          float sv[3] = {P.sc.lights[k][0] - p[0], P.sc.lights[k][1] - p[1], P.sc.lights[k][2] - p[2]};
          const float mag = sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2];
          float inv_mag = 0.f;
          if (mag != 0.f) { const float qq = rsqrtf(mag); sv[0] *= qq; sv[1] *= qq; sv[2] *= qq; inv_mag = qq * qq; }
          const float sdot = sv[0] * l[0] + sv[1] * l[1] + sv[2] * l[2];
          if (!(sdot > 0.f)) continue;
          mask |= 1u << k;
          q[k] = (mag != 0.f) ? sdot * inv_mag : sdot / mag;
          if (a2 > 0.f) {
            // light_ref = normal3D(reflect3D(-sv, l)); spec_dot = dir . (-light_ref)
            const float tt = 2.f * sdot;                   // -(2 * dot(-sv, l))
            float lr[3] = {-sv[0] + l[0] * tt, -sv[1] + l[1] * tt, -sv[2] + l[2] * tt};
            const float ll = lr[0] * lr[0] + lr[1] * lr[1] + lr[2] * lr[2];
            if (ll != 0.f) { const float qq = rsqrtf(ll); lr[0] *= qq; lr[1] *= qq; lr[2] *= qq; }
            const float sd = -(d[0] * lr[0] + d[1] * lr[1] + d[2] * lr[2]);
            if (sd > 0.f) sp[k] = __powf(sd, m.specular_exponent);
          }
        }
      }
      if (mask) first_light = __ffs((int)mask) - 1;
      P.b.pos[node] = make_float4(p[0], p[1], p[2], __int_as_float(hit_i));
      // base.w: facing-light mask (bits 0-7) + which child slots will be filled (bits 8, 9) — the compose
      // pass reads a child colour only when its bit is set, so the slots need no zero-initialisation
      P.b.base[node] = make_float4(rgb[0], rgb[1], rgb[2],
                                   __int_as_float((int)(mask | (want_refl ? 0x100u : 0u) | (want_refr ? 0x200u : 0u))));
      if (mask) {                                         // no facing light: never queued, never accumulated
        P.b.lq[node] = make_float4(q[0], sp[0], q[1], sp[1]);
        if (P.b.lq2) P.b.lq2[node] = make_float4(q[2], sp[2], q[3], sp[3]);
        P.b.acc[node] = make_float4(P.sc.light_intensity, 0.f, 0.f, 0.f);
      }
    }
  }

  // ---- warp-compacted allocation (north star: ballot/shuffle compaction of live rays): children, and the
  //      shadow queue of the first light that faces the surface ----------------------------------------------
  const uint32_t b1 = __ballot_sync(0xffffffffu, want_refl);
  const uint32_t b2 = __ballot_sync(0xffffffffu, want_refr);
  uint32_t bk[RTB_MAX_LIGHTS];
  uint32_t counts[1 + RTB_MAX_LIGHTS];
  counts[0] = __popc(b1) + __popc(b2);
#pragma unroll
  for (int k = 0; k < RTB_MAX_LIGHTS; k++) {
    bk[k] = k < P.sc.n_lights ? __ballot_sync(0xffffffffu, first_light == k) : 0u;
    counts[1 + k] = __popc(bk[k]);
  }
  reserve<AGG>(P, counts, P.sc.n_lights, 0, lane);
  const uint32_t cbase = counts[0] + P.n_sub;
  if (want_refl) {
    const uint32_t c = cbase + __popc(b1 & lanemask_lt());
    if (c < P.b.cap_nodes) {
      P.b.ray_o[c] = make_float4(p[0], p[1], p[2], rdir[0]);
      P.b.ray_d[c] = make_float4(rdir[1], rdir[2], w_refl, __int_as_float(hit_i));
      P.b.link[c] = make_int4((int)node, (int)((2u * r.path) | ((uint32_t)(r.segs_left - 1) << 24)), (int)r.pixsub, 0);
    } else { P.b.ctr->overflow = 1u; P.sticky->overflow = 1u; }
  }
  if (want_refr) {
    const uint32_t c = cbase + __popc(b1) + __popc(b2 & lanemask_lt());
    if (c < P.b.cap_nodes) {
      P.b.ray_o[c] = make_float4(p[0], p[1], p[2], fdir[0]);
      P.b.ray_d[c] = make_float4(fdir[1], fdir[2], w_refr, __int_as_float(hit_i));
      P.b.link[c] = make_int4((int)node, (int)((2u * r.path + 1u) | ((uint32_t)(r.segs_left - 1) << 24)), (int)r.pixsub, 0);
    } else { P.b.ctr->overflow = 1u; P.sticky->overflow = 1u; }
  }
#pragma unroll
  for (int k = 0; k < RTB_MAX_LIGHTS; k++) {
    if (first_light == k) {
      const uint32_t e = counts[1 + k] + __popc(bk[k] & lanemask_lt());
      if (RECORDS) {
        P.b.srec_a[k][e] = shadow_record(P, k, p[0], p[1], p[2]);
        P.b.srec_b[k][e] = make_int4((int)node, hit_i, __float_as_int(P.sc.light_intensity), 0);
      } else {
        P.b.sq[k][e] = node;
      }
    }
  }
}

// The fused trace kernels call hit processing out of line (their hot loops must stay small).
template <bool RECORDS>
__device__ __noinline__ void process_hit(const WfParams& P, const WfRay& r, unsigned long long key, uint32_t lane) {
  process_hit_impl<RECORDS, false>(P, r, key, lane);
}

// ---- RAYGEN ---------------------------------------------------------------------------------------------
// Primary rays of the band (main.js:186-193) for every frame of the group, generated in fp64 with the
// frame's own camera and rounded ONCE; direction stored in ray_o[node] (level-1 entries of the ray arrays
// are otherwise unused).  A kernel of its own: one thread per ray at full occupancy, which keeps the fp64
// division / square-root code and the per-frame camera loads out of the register-bound sweep kernel.
__global__ void __launch_bounds__(WF_THREADS)
wf_raygen_kernel(const __grid_constant__ WfParams P) {
  const int W = P.fr64.width;
  const int g = P.fr64.spp_grid;
  for (uint32_t node = blockIdx.x * blockDim.x + threadIdx.x; node < P.n_sub; node += gridDim.x * blockDim.x) {
    const uint32_t f = node / P.n_sub_frame;               // frame of the group
    const uint32_t in_frame = node - f * P.n_sub_frame;
    const uint32_t pix = in_frame / (uint32_t)P.spp;
    const uint32_t sub = in_frame - pix * (uint32_t)P.spp;
    const int lrow = P.band_l0 + (int)(pix / (uint32_t)W);
    double ray[3];
    primary_ray<double>(P.cam[f], (int)(pix % (uint32_t)W), P.rm.local_to_y(lrow), (int)(sub % g), (int)(sub / g), ray);
    P.b.ray_o[node] = make_float4((float)ray[0], (float)ray[1], (float)ray[2], 0.f);
  }
}

// ---- TRACE ----------------------------------------------------------------------------------------------
// Rays per lane are a launch-time decision made on the device from the size of the level: R_SHARED /
// R_GENERAL when the level fills the machine several times over, half or a quarter of that for small
// levels (deep recursion levels, multi-GPU stripes), so that chunks stay plentiful and the tail short.
// Cost model (SMSP cycles per sphere pair, measured with tools/sweep_micro*.cu): a pair costs
// c(R) = max(R * f, 32) + 4 per warp — f = FMA+ALU cycles per ray (9 shared-origin, 20 general), 32 = the
// two broadcast LDS.128 when 16 warps share the SM's LSU — and with chunks pulled dynamically a launch
// lasts about c(R) * (chunks / warps + 1): the work, plus one chunk of tail.
RTB_DEVINL int pick_rays_per_lane(uint32_t count, int r_big, uint32_t f) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  int best_r = r_big;
  uint64_t best_cost = ~0ull;
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const uint32_t r = (uint32_t)r_big >> s;
    const uint32_t chunks = (count + 32u * r - 1u) / (32u * r);
    const uint64_t cost = (uint64_t)(max(r * f, 32u) + 4u) * ((uint64_t)chunks + warps);
    if (cost < best_cost) { best_cost = cost; best_r = (int)r; }   // ties keep the larger R
  }
  return best_r;
}

// Guided scheduling: the last part of a level is cut into SMALL chunks (r_small rays per lane) so that
// the tail of the launch is one small chunk, not one big one.  Returns how many rays to leave for that
// tail phase (a multiple of the small chunk; 0 when the big chunks are already the smallest).
RTB_DEVINL uint32_t tail_rays(uint32_t count, int r, int r_small) {
  if (r <= r_small) return 0;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  uint32_t t = min(count / 4u, warps * 16u * (uint32_t)r);        // about half a big chunk per warp
  const uint32_t q = 32u * (uint32_t)r_small;
  return (t / q) * q;
}

// Per-ray sweep operands.  PRIMARY: a = -dir (ray generated in fp64, rounded once).  General: a = u, b = v
// (unit vectors completing dir to an orthonormal basis), nku = -(u.org), nkv = -(v.org), self = the sphere
// the ray starts on (hit_i of its parent node).
template <bool PRIMARY>
RTB_DEVINL void ray_setup(const WfParams& P, uint32_t node, float (&a)[3], float (&b)[3], float& nku, float& nkv,
                          int& self) {
  if (PRIMARY) {
    const float4 d = P.b.ray_o[node];                      // written by wf_raygen_kernel
    a[0] = -d.x; a[1] = -d.y; a[2] = -d.z;
  } else {
    const float4 ro = P.b.ray_o[node], rd = P.b.ray_d[node];
    self = __float_as_int(rd.w);                           // hit_i of the parent node, stored with the ray
    const float dx = ro.w, dy = rd.x, dz = rd.y;
    // branchless orthonormal basis (Duff et al. 2017): u, v unit, u.v = u.d = v.d = 0 to ~1 ulp
    const float sg = copysignf(1.f, dz);
    const float ka = -1.f / (sg + dz);
    const float kb = dx * dy * ka;
    a[0] = 1.f + sg * dx * dx * ka; a[1] = sg * kb; a[2] = -sg * dx;
    b[0] = kb; b[1] = sg + dy * dy * ka; b[2] = -dy;
    nku = -(a[0] * ro.x + a[1] * ro.y + a[2] * ro.z);
    nkv = -(b[0] * ro.x + b[1] * ro.y + b[2] * ro.z);
  }
}

// The ray as process_hit wants it (origin, direction, tree position).  nd = -dir of a primary ray.
template <bool PRIMARY>
RTB_DEVINL WfRay ray_record(const WfParams& P, uint32_t node, bool active, float ndx, float ndy, float ndz) {
  WfRay r;
  r.node = node;
  r.active = active;
  r.path = 1; r.pixsub = node; r.segs_left = P.fr64.segs;
  r.o[0] = r.o[1] = r.o[2] = 0.f;
  r.d[0] = -ndx; r.d[1] = -ndy; r.d[2] = -ndz;
  if (PRIMARY) {
    const float* o = P.origin32[min(node / P.n_sub_frame, (uint32_t)(RTB_MAX_GROUP - 1))];
    r.o[0] = o[0]; r.o[1] = o[1]; r.o[2] = o[2];
  } else {
    r.d[0] = r.d[1] = r.d[2] = 0.f;
    if (active) {
      const float4 ro = P.b.ray_o[node], rd = P.b.ray_d[node];
      const int4 lk = P.b.link[node];
      r.o[0] = ro.x; r.o[1] = ro.y; r.o[2] = ro.z;
      r.d[0] = ro.w; r.d[1] = rd.x; r.d[2] = rd.y;
      r.path = (uint32_t)lk.y & 0xffffffu;
      r.segs_left = (int)((uint32_t)lk.y >> 24);
      r.pixsub = (uint32_t)lk.z;                         // level-1 node (frame, pixel, sub-sample) of the ancestor
    }
  }
  return r;
}

// `first` is the warp's first chunk, claimed for the whole CTA at once (cta_claim); the following ones
// are pulled per warp.
template <bool PRIMARY, int R, bool SPLIT, bool RECORDS>
RTB_DEVINL void trace_chunks(const WfParams& P, uint32_t* work, uint32_t first, const float4* __restrict__ tab,
                             const float4* __restrict__ exact, uint32_t begin, uint32_t end, uint32_t lane,
                             unsigned long long& my_nodes) {
  const int n_pad = P.n_pad;
  const float eps = P.sc.eps;
  constexpr uint32_t CHUNK = 32u * R;

  for (uint32_t c = first;; c = next_chunk(work, lane)) {
    const uint64_t base = (uint64_t)begin + (uint64_t)c * CHUNK;
    if (base >= end) break;
    float a[R][3], b[R][3];          // PRIMARY: a = -dir;  general: a = u, b = v (basis completing dir)
    float nku[R], nkv[R];            // general: -(u.org), -(v.org)
    uint32_t nodes[R];
    int self[R];                     // general: sphere the ray starts on (hit_i of its parent node)
    bool active[R];
    unsigned long long best[R];
#pragma unroll
    for (int q = 0; q < R; q++) {
      const uint32_t node = (uint32_t)base + (uint32_t)q * 32u + lane;
      nodes[q] = node;
      active[q] = node < end;
      // idle slots must never look like candidates: zero vectors (shared origin: test value = sc,
      // positive unless the origin is inside the sphere; general: a = -ku huge, d2 huge)
      a[q][0] = a[q][1] = a[q][2] = 0.f;
      b[q][0] = b[q][1] = b[q][2] = 0.f;
      nku[q] = 3.0e18f; nkv[q] = 0.f;
      self[q] = -1;
      if (active[q]) {
        my_nodes++;
        ray_setup<PRIMARY>(P, node, a[q], b[q], nku[q], nkv[q], self[q]);
      }
    }
    if (PRIMARY) sweep_primary<R>(tab, exact, 0, n_pad, a, active, best, eps);
    else sweep_general<R>(tab, exact, P.b.ray_o, P.b.ray_d, 0, n_pad, a, b, nku, nkv, nodes, self, active, best, eps);
    if (SPLIT) {                                          // the hit kernel takes it from here
#pragma unroll
      for (int q = 0; q < R; q++)
        if (active[q]) P.b.key[nodes[q]] = best[q];
    } else {
      __syncwarp();
#pragma unroll
      for (int q = 0; q < R; q++)
        process_hit<RECORDS>(P, ray_record<PRIMARY>(P, nodes[q], active[q], a[q][0], a[q][1], a[q][2]), best[q], lane);
    }
  }
}

// One contiguous range of rays that share a sweep table: a whole level of general rays, or the primary rays
// of one run of frames with a common camera origin.  Contains CTA barriers: called by all threads.
template <bool PRIMARY, bool SPLIT, bool RECORDS>
RTB_DEVINL void trace_range(const WfParams& P, float4* tab, const float4* __restrict__ src,
                            const float4* __restrict__ exact, uint32_t begin, uint32_t end, uint32_t* wa,
                            uint32_t* wt, uint32_t lane, unsigned long long& my_nodes) {
  constexpr int RB = PRIMARY ? R_SHARED : R_GENERAL;
  const int n_pad = P.n_pad;
  const int r = P.fixed_r ? (RB >> (P.fixed_r - 1)) : pick_rays_per_lane(end - begin, RB, PRIMARY ? 9u : 20u);
  const uint32_t split = P.fixed_r ? end : end - tail_rays(end - begin, r, RB / 4);
  const uint32_t first = cta_claim(wa);
  // a CTA whose claim lies beyond the range has nothing to do: leave at once (the slot goes to the
  // kernels of the other frames in flight) without even loading the table
  if ((uint64_t)begin + (uint64_t)(first - (threadIdx.x >> 5)) * (32u * (uint32_t)r) < split || split < end) {
    // both tables in shared memory: the packed sweep table and the exact one the slow path reads
    // (general rays: the spheres themselves, n entries; the padding is never a candidate)
    float4* xt = tab + n_pad;
    const int n_x = PRIMARY ? n_pad : P.sc.n_spheres;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
      tab[i] = src[i];
      xt[i] = i < n_x ? exact[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (r == RB) trace_chunks<PRIMARY, RB, SPLIT, RECORDS>(P, wa, first, tab, xt, begin, split, lane, my_nodes);
    else if (r == RB / 2) trace_chunks<PRIMARY, RB / 2, SPLIT, RECORDS>(P, wa, first, tab, xt, begin, split, lane, my_nodes);
    else trace_chunks<PRIMARY, RB / 4, SPLIT, RECORDS>(P, wa, first, tab, xt, begin, split, lane, my_nodes);
    if (split < end) trace_chunks<PRIMARY, RB / 4, SPLIT, RECORDS>(P, wt, next_chunk(wt, lane), tab, xt, split, end, lane, my_nodes);
  }
}

template <bool PRIMARY, bool SPLIT, bool RECORDS>
RTB_DEVINL void trace_body(const WfParams& P, int level) {
  extern __shared__ float4 tab[];
  const int n = P.sc.n_spheres;
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long my_nodes = 0;
  if (PRIMARY) {
    // level 1: the frames of the group, run by run (frames of a run share the camera origin, hence the
    // table); a CTA moves on to the next run while others still finish the previous one
    for (int g = 0; g < P.n_runs; g++) {
      if (g > 0) __syncthreads();                          // table and claim word are reused
      trace_range<true, SPLIT, RECORDS>(P, tab, P.primary_tab[g], P.primary_x[g], (uint32_t)P.run_begin[g] * P.n_sub_frame,
                               (uint32_t)P.run_begin[g + 1] * P.n_sub_frame, &P.b.ctr->work_primary[g],
                               &P.b.ctr->work_primary_tail[g], lane, my_nodes);
    }
  } else {
    uint32_t begin, end;
    level_range(P, level, begin, end);
    trace_range<false, SPLIT, RECORDS>(P, tab, P.geom_tab, P.geom, begin, end, &P.b.ctr->work_trace[level],
                              &P.b.ctr->work_trace_tail[level], lane, my_nodes);
  }
  // stats
  for (int o2 = 16; o2 > 0; o2 >>= 1) my_nodes += __shfl_down_sync(0xffffffffu, my_nodes, o2);
  if (lane == 0 && my_nodes) {
    atomicAdd(&P.sticky->nodes, my_nodes);
    atomicAdd(&P.sticky->sphere_tests, my_nodes * (unsigned long long)n);
  }
  if (!SPLIT) publish_level(P, level);                    // split pipeline: the hit kernel creates the children
}

// Fused pipeline: sweep + hit processing of a level in one persistent kernel.
template <bool PRIMARY, bool RECORDS>
__global__ void __launch_bounds__(WF_THREADS, PRIMARY ? WF_MIN_CTAS : WF_MIN_CTAS_GENERAL)
wf_trace_kernel(const __grid_constant__ WfParams P, int level) {
  trace_body<PRIMARY, false, RECORDS>(P, level);
}

// Split pipeline, step 1 of a level: the closest-hit SWEEP alone.  key[node] = (t bits, sphere, flip) or
// KEY_NONE.  Every resident warp is in the FMA-bound sweep nearly all the time; nothing here waits on the
// scattered loads and stores of hit processing.
template <bool PRIMARY>
__global__ void __launch_bounds__(WF_THREADS, PRIMARY ? WF_MIN_CTAS_SWEEP_PRIMARY : WF_MIN_CTAS_GENERAL)
wf_sweep_kernel(const __grid_constant__ WfParams P, int level) {
  trace_body<PRIMARY, true, false>(P, level);
}

// Split pipeline, step 2 of a level: HIT PROCESSING, one thread per ray at high occupancy - the work is
// scattered loads (material, texels) and stores (node record, child rays, shadow records), i.e. latency
// bound, and wants many warps, not many registers.  Warps take 32 consecutive nodes at a time, so the
// ballot-compacted queue pushes see the same neighbours as in the fused kernels.
template <bool PRIMARY, bool RECORDS>
__global__ void __launch_bounds__(WF_THREADS, WF_MIN_CTAS_HIT)
wf_hit_kernel(const __grid_constant__ WfParams P, int level) {
  uint32_t begin, end;
  level_range(P, level, begin, end);
  const uint32_t lane = threadIdx.x & 31u;
  // the CTA walks tiles of 256 consecutive nodes: queue space is reserved once per tile (reserve<true>)
  for (uint64_t base = (uint64_t)begin + (uint64_t)blockIdx.x * WF_THREADS; base < end;
       base += (uint64_t)gridDim.x * WF_THREADS) {
    const uint32_t node = (uint32_t)base + threadIdx.x;
    const bool active = node < end;
    float ndx = 0.f, ndy = 0.f, ndz = 0.f;
    unsigned long long key = KEY_NONE;
    if (active) {
      key = P.b.key[node];
      if (PRIMARY) { const float4 d = P.b.ray_o[node]; ndx = -d.x; ndy = -d.y; ndz = -d.z; }
    }
    process_hit_impl<RECORDS, true>(P, ray_record<PRIMARY>(P, node, active, ndx, ndy, ndz), key, lane);
  }
  publish_level(P, level);
}

// ---- SHADOW ---------------------------------------------------------------------------------------------
// Occluder sweep of main.js:293-304 for light k, traced from the light toward hit.p.  Same fast test as
// sweep_primary (table = C - light, sc); the slow path applies the reference's [0.001, light_len)
// window along the reference's own ray (from hit.p toward the light): t = light_len - s.
template <int R, bool SPLIT>
RTB_DEVINL void shadow_chunks(const WfParams& P, int k, const float4* __restrict__ tab, uint32_t first,
                              uint32_t count, uint32_t* work, uint32_t first_chunk, uint32_t lane,
                              unsigned long long& my_rays, unsigned long long& my_tests) {
  const int n = P.sc.n_spheres, n_pad = P.n_pad;
  const float4* __restrict__ exact = tab + n_pad;          // light_x[k], staged behind the packed table
  const float eps = P.sc.eps;
  const float lx = P.sc.lights[k][0], ly = P.sc.lights[k][1], lz = P.sc.lights[k][2];
  constexpr uint32_t CHUNK = 32u * R;

  for (uint32_t c = first_chunk;; c = next_chunk(work, lane)) {
    const uint64_t base = (uint64_t)first + (uint64_t)c * CHUNK;
    if (base >= count) break;
    uint32_t node[R];
    int hit_i[R];
    float nw[R][3], len[R], I[R];      // nw = -(direction light -> p)
    uint32_t mask[R];
    bool active[R];
#pragma unroll
    for (int q = 0; q < R; q++) {
      const uint32_t e = (uint32_t)base + (uint32_t)q * 32u + lane;
      active[q] = e < count;
      node[q] = 0; hit_i[q] = -1; len[q] = 0.f; I[q] = 0.f; mask[q] = 0;
      nw[q][0] = nw[q][1] = nw[q][2] = 0.f;                // idle / dead slot: never a candidate
      if (active[q] && SPLIT) {                             // complete ray record, two coalesced loads
        const float4 ra = P.b.srec_a[k][e];
        const int4 rb = P.b.srec_b[k][e];
        node[q] = (uint32_t)rb.x;
        hit_i[q] = rb.y;
        len[q] = ra.w;
        I[q] = __int_as_float(rb.z);
        my_rays++;
        if (I[q] != 0.f) { nw[q][0] = ra.x; nw[q][1] = ra.y; nw[q][2] = ra.z; }
      } else if (active[q]) {
        node[q] = P.b.sq[k][e];
        const float4 ps = P.b.pos[node[q]];
        hit_i[q] = __float_as_int(ps.w);
        // direction light -> p (= -shadow_vec) and light_len (main.js:287-290)
        float wx = ps.x - lx, wy = ps.y - ly, wz = ps.z - lz;
        len[q] = sqrtf(wx * wx + wy * wy + wz * wz);
        if (len[q] != 0.f) { const float rr = 1.f / len[q]; wx *= rr; wy *= rr; wz *= rr; }
        I[q] = P.b.acc[node[q]].x;
        my_rays++;
        if (I[q] != 0.f) { nw[q][0] = -wx; nw[q][1] = -wy; nw[q][2] = -wz; }
      }
    }
    for (int j = 0; j < n_pad; j += WF_CHECK) {
      if ((j & 127) == 0) {                                 // all rays of the warp fully occluded?
        bool l = false;
#pragma unroll
        for (int q = 0; q < R; q++) l |= (I[q] != 0.f);
        if (!__any_sync(0xffffffffu, l)) break;
      }
#pragma unroll 1
      for (int jg = j; jg < j + WF_CHECK; jg += WF_GROUP) {
#pragma unroll
        for (int u = 0; u < WF_GROUP / 2; u++) {
          const float4 A = tab[jg + 2 * u], B = tab[jg + 2 * u + 1];
#pragma unroll
          for (int q = 0; q < R; q++) {
            float2 t = __ffma2_rn(f2(nw[q][2], nw[q][2]), f2(B.x, B.y), f2(B.z, B.w));
            t = __ffma2_rn(f2(nw[q][1], nw[q][1]), f2(A.z, A.w), t);
            t = __ffma2_rn(f2(nw[q][0], nw[q][0]), f2(A.x, A.y), t);
            mask[q] = push_sign(push_sign(mask[q], t.x), t.y);
          }
        }
      }
      uint32_t any = 0;
#pragma unroll
      for (int q = 0; q < R; q++) any |= mask[q];
      if (any) {
#pragma unroll
        for (int q = 0; q < R; q++) {
          uint32_t m = mask[q];
          // the ray ends ON sphere hit_i, so that sphere is always a candidate; drop its bit here
          // (main.js:294 skips it) instead of paying a divergent slow-path call to find that out
          if (j == (hit_i[q] & ~(WF_CHECK - 1))) m &= ~(0x80000000u >> (hit_i[q] & (WF_CHECK - 1)));
          if (m && I[q] != 0.f) {
            I[q] = slow_shadow(exact, P.sc.mtl, j, m, nw[q][0], nw[q][1], nw[q][2], len[q], hit_i[q], eps, I[q]);
            if (I[q] == 0.f) {                               // fully occluded: stop testing
              nw[q][0] = nw[q][1] = nw[q][2] = 0.f;
              my_tests += (unsigned long long)min(j + WF_CHECK, n);   // device statistic: spheres swept while lit
            }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < R; q++) if (active[q] && I[q] != 0.f) my_tests += (unsigned long long)n;
    if (SPLIT) {                                            // wf_spost_kernel accumulates and hands on
#pragma unroll
      for (int q = 0; q < R; q++)
        if (active[q]) P.b.sq[k][(uint32_t)base + (uint32_t)q * 32u + lane] = __float_as_uint(I[q]);
      continue;
    }
#pragma unroll
    for (int q = 0; q < R; q++) {
      int next = -1;
      if (active[q]) {
        float4 a = P.b.acc[node[q]];
        a.x = I[q];
        const uint32_t lmask = (uint32_t)__float_as_int(P.b.base[node[q]].w) & 0xffu;
        if (I[q] != 0.f) {                                 // main.js:305-313
          const float4 lq = (k < 2) ? P.b.lq[node[q]] : P.b.lq2[node[q]];
          const float qk = (k & 1) ? lq.z : lq.x, sk = (k & 1) ? lq.w : lq.y;
          a.y += I[q] * qk;
          a.z += sk;
        }
        P.b.acc[node[q]] = a;
        const uint32_t rest = lmask & ~((2u << k) - 1u);   // facing lights after k
        if (I[q] != 0.f && rest) next = __ffs((int)rest) - 1;   // carry light_intensity to that light
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
        if (next == k2) P.b.sq[k2][qb + __popc(bk & lanemask_lt())] = node[q];
      }
    }
  }
}

// phase 0: everything queued for light k.  phase 1 (k = 0 only): the EARLY pass over the entries the
// primary trace queued ([0, sq_snap0)), launched on a second stream so that it overlaps the secondary
// trace levels and fills their tails.  phase 2 (k = 0): the rest, [sq_snap0, end).
RTB_DEVINL void shadow_range(const WfParams& P, int k, int phase, uint32_t& first, uint32_t& count) {
  first = 0;
  count = *((volatile uint32_t*)&P.b.ctr->sq_count[k]);
  if (phase == 1) count = *((volatile uint32_t*)&P.b.ctr->sq_snap0);
  else if (phase == 2) first = *((volatile uint32_t*)&P.b.ctr->sq_snap0);
}

template <bool SPLIT>
RTB_DEVINL void shadow_body(const WfParams& P, int k, int phase) {
  extern __shared__ float4 tab[];
  const int n_pad = P.n_pad;
  uint32_t first, count;
  shadow_range(P, k, phase, first, count);
  uint32_t* work = (phase == 1) ? &P.b.ctr->work_shadow_early : &P.b.ctr->work_shadow[k];
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long my_rays = 0, my_tests = 0;
  uint32_t* work_tail = (phase == 1) ? &P.b.ctr->work_shadow_early_tail : &P.b.ctr->work_shadow_tail[k];
  const int r = P.fixed_r ? (R_SHARED >> (P.fixed_r - 1)) : pick_rays_per_lane(count - first, R_SHARED, 9u);
  const uint32_t split = P.fixed_r ? count : count - tail_rays(count - first, r, R_SHARED / 4);
  const uint32_t fc = cta_claim(work);
  if ((uint64_t)first + (uint64_t)(fc - (threadIdx.x >> 5)) * (32u * (uint32_t)r) < split || split < count) {
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) { tab[i] = P.light_tab[k][i]; tab[n_pad + i] = P.light_x[k][i]; }
    __syncthreads();
    if (r == R_SHARED) shadow_chunks<R_SHARED, SPLIT>(P, k, tab, first, split, work, fc, lane, my_rays, my_tests);
    else if (r == R_SHARED / 2) shadow_chunks<R_SHARED / 2, SPLIT>(P, k, tab, first, split, work, fc, lane, my_rays, my_tests);
    else shadow_chunks<R_SHARED / 4, SPLIT>(P, k, tab, first, split, work, fc, lane, my_rays, my_tests);
    if (split < count)
      shadow_chunks<R_SHARED / 4, SPLIT>(P, k, tab, split, count, work_tail, next_chunk(work_tail, lane), lane, my_rays, my_tests);
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

// Fused pipeline: occluder sweep + accumulation + hand-off to the next facing light.
__global__ void __launch_bounds__(WF_THREADS, WF_MIN_CTAS_SHADOW)
wf_shadow_kernel(const __grid_constant__ WfParams P, int k, int phase) {
  shadow_body<false>(P, k, phase);
}

// Split pipeline: the occluder SWEEP alone, from complete ray records; the surviving light_intensity of
// entry e goes to sq[k][e] (float bits).
__global__ void __launch_bounds__(WF_THREADS, WF_MIN_CTAS_SHADOW)
wf_ssweep_kernel(const __grid_constant__ WfParams P, int k, int phase) {
  shadow_body<true>(P, k, phase);
}

// Split pipeline: what follows the sweep of light k for every entry - main.js:305-313 (diffuse / specular
// sums) and the hand-off of the carried light_intensity to the next light that faces the surface
// (main.js:284: declared once outside the loop), as a new ray record.  One thread per entry.
__global__ void __launch_bounds__(WF_THREADS, WF_MIN_CTAS_HIT)
wf_spost_kernel(const __grid_constant__ WfParams P, int k, int phase) {
  uint32_t first, count;
  shadow_range(P, k, phase, first, count);
  const uint32_t lane = threadIdx.x & 31u;
  for (uint64_t base = (uint64_t)first + (uint64_t)blockIdx.x * WF_THREADS; base < count;
       base += (uint64_t)gridDim.x * WF_THREADS) {
    const uint32_t e = (uint32_t)base + threadIdx.x;
    int next = -1;
    uint32_t node = 0;
    int hit_i = -1;
    float I = 0.f;
    if (e < count) {
      const int4 rb = P.b.srec_b[k][e];
      node = (uint32_t)rb.x;
      hit_i = rb.y;
      I = __uint_as_float(P.b.sq[k][e]);
      float4 a = P.b.acc[node];
      a.x = I;
      const uint32_t lmask = (uint32_t)__float_as_int(P.b.base[node].w) & 0xffu;
      if (I != 0.f) {                                      // main.js:305-313
        const float4 lq = (k < 2) ? P.b.lq[node] : P.b.lq2[node];
        const float qk = (k & 1) ? lq.z : lq.x, sk = (k & 1) ? lq.w : lq.y;
        a.y += I * qk;
        a.z += sk;
      }
      P.b.acc[node] = a;
      const uint32_t rest = lmask & ~((2u << k) - 1u);     // facing lights after k
      if (I != 0.f && rest) next = __ffs((int)rest) - 1;   // carry light_intensity to that light
    }
    if (k + 1 >= P.sc.n_lights) continue;                   // last light: nobody to hand on to
    uint32_t bk[RTB_MAX_LIGHTS];
    uint32_t counts[1 + RTB_MAX_LIGHTS];
    counts[0] = 0;
#pragma unroll
    for (int k2 = 0; k2 < RTB_MAX_LIGHTS; k2++) {
      bk[k2] = (k2 > k && k2 < P.sc.n_lights) ? __ballot_sync(0xffffffffu, next == k2) : 0u;
      counts[1 + k2] = __popc(bk[k2]);
    }
    reserve<true>(P, counts, P.sc.n_lights, k + 2, lane);
#pragma unroll
    for (int k2 = 0; k2 < RTB_MAX_LIGHTS; k2++) {
      if (next == k2) {
        const uint32_t e2 = counts[1 + k2] + __popc(bk[k2] & lanemask_lt());
        const float4 ps = P.b.pos[node];
        P.b.srec_a[k2][e2] = shadow_record(P, k2, ps.x, ps.y, ps.z);
        P.b.srec_b[k2][e2] = make_int4((int)node, hit_i, __float_as_int(I), 0);
      }
    }
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
  const uint32_t bits = (uint32_t)__float_as_int(bs.w);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 a = (bits & 0xffu) ? P.b.acc[node] : zero;   // diffuse / specular sums exist only for lit nodes
  const float4 ca = (bits & 0x100u) ? P.b.child_a[node] : zero, cb = (bits & 0x200u) ? P.b.child_b[node] : zero;
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
    const int4 lk = P.b.link[node];
    const float wgt = P.b.ray_d[node].z;                  // parent's albedo[3] or albedo[4]
    const float4 v = make_float4(col[0] * wgt, col[1] * wgt, col[2] * wgt, 0.f);   // main.js:271, 277
    if (lk.y & 1) P.b.child_b[lk.x] = v; else P.b.child_a[lk.x] = v;
  }
}

// Level 1: average the sub-samples of each pixel, quantise as a Uint8ClampedArray store would
// (main.js:195-198) and write 4 pixels per thread as one uint4 - for every frame of the group.
__global__ void __launch_bounds__(WF_THREADS)
wf_resolve_kernel(const __grid_constant__ WfParams P) {
  const int W = P.fr64.width;
  const int groups_per_row = (W + 3) / 4;
  const long long per_frame = (long long)P.band_rows * groups_per_row;
  const long long total = per_frame * P.n_frames;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total;
       gi += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(gi / per_frame);
    const long long gf = gi - (long long)f * per_frame;
    const int r = (int)(gf / groups_per_row), x0 = (int)(gf % groups_per_row) * 4;
    const int lrow = P.band_l0 + r;
    uint32_t packed[4] = {0u, 0u, 0u, 0u};
    const int nx = min(4, W - x0);
    for (int i = 0; i < nx; i++) {
      const uint32_t pix = (uint32_t)r * (uint32_t)W + (uint32_t)(x0 + i);
      float acc[3] = {0.f, 0.f, 0.f};
      for (int s = 0; s < P.spp; s++) {
        float col[3];
        node_colour(P, (uint32_t)f * P.n_sub_frame + pix * (uint32_t)P.spp + (uint32_t)s, col);
        if (P.spp == 1) { acc[0] = col[0]; acc[1] = col[1]; acc[2] = col[2]; }
        else { acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2]; }
      }
      if (P.spp > 1) { const float q = (float)P.spp; acc[0] /= q; acc[1] /= q; acc[2] /= q; }
      packed[i] = clamped_u8(255.f * acc[0]) | (clamped_u8(255.f * acc[1]) << 8) |
                  (clamped_u8(255.f * acc[2]) << 16) | (255u << 24);
    }
    uint32_t* row = reinterpret_cast<uint32_t*>(P.out[f]) + P.rm.out_row(lrow) * (size_t)W;
    if (P.aligned4 && nx == 4) {
      *reinterpret_cast<uint4*>(row + x0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    } else {
      for (int i = 0; i < nx; i++) row[x0 + i] = packed[i];
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------
size_t per_node_bytes(int n_lights) {
  return 16 * 8 + (n_lights > 2 ? 16 : 0) + 4 * (size_t)n_lights;
}

void slot_free(WfSlot& sl) {
  WfBuffers& b = sl.wf;
  cudaFree(b.ray_o); cudaFree(b.ray_d); cudaFree(b.link); cudaFree(b.pos); cudaFree(b.base);
  cudaFree(b.lq); cudaFree(b.lq2); cudaFree(b.acc); cudaFree(b.child_a); cudaFree(b.child_b);
  for (int k = 0; k < RTB_MAX_LIGHTS; k++) { cudaFree(b.sq[k]); cudaFree(b.srec_a[k]); cudaFree(b.srec_b[k]); }
  cudaFree(b.key);
  cudaFree(b.ctr);
  b = WfBuffers{};
  sl.wf_lights = 0;
  sl.wf_bytes = 0;
}

int ensure_workspace(WfSlot& sl, int n_lights, uint32_t cap_nodes, bool split) {
  WfBuffers& b = sl.wf;
  if (!sl.stream2) {
    RTB_CUDA(cudaStreamCreateWithFlags(&sl.stream2, cudaStreamNonBlocking));
    RTB_CUDA(cudaEventCreateWithFlags(&sl.ev_primary, cudaEventDisableTiming));
    RTB_CUDA(cudaEventCreateWithFlags(&sl.ev_early, cudaEventDisableTiming));
    RTB_CUDA(cudaEventCreateWithFlags(&sl.ev_async, cudaEventDisableTiming));
    RTB_CUDA(cudaEventCreateWithFlags(&sl.ev_half, cudaEventDisableTiming));
    RTB_CUDA(cudaMallocHost(&sl.h_sticky, sizeof(WfSticky)));
  }
  const bool need_lq2 = n_lights > 2;
  if (b.cap_nodes >= cap_nodes && b.ctr && sl.wf_lights >= n_lights && (!need_lq2 || b.lq2) && (!split || b.key))
    return RTB_OK;
  split = split || b.key != nullptr;
  const uint32_t cap = std::max(cap_nodes, b.cap_nodes);
  const int lights = std::max(n_lights, sl.wf_lights);
  slot_free(sl);
  const size_t n = cap;
  RTB_CUDA(cudaMalloc(&b.ray_o, n * 16));
  RTB_CUDA(cudaMalloc(&b.ray_d, n * 16));
  RTB_CUDA(cudaMalloc(&b.link, n * 16));
  RTB_CUDA(cudaMalloc(&b.pos, n * 16));
  RTB_CUDA(cudaMalloc(&b.base, n * 16));
  RTB_CUDA(cudaMalloc(&b.lq, n * 16));
  if (lights > 2) RTB_CUDA(cudaMalloc(&b.lq2, n * 16));
  RTB_CUDA(cudaMalloc(&b.acc, n * 16));
  RTB_CUDA(cudaMalloc(&b.child_a, n * 16));
  RTB_CUDA(cudaMalloc(&b.child_b, n * 16));
  for (int k = 0; k < lights; k++) RTB_CUDA(cudaMalloc(&b.sq[k], n * 4));
  if (split) {
    RTB_CUDA(cudaMalloc(&b.key, n * 8));
    for (int k = 0; k < lights; k++) {
      RTB_CUDA(cudaMalloc(&b.srec_a[k], n * 16));
      RTB_CUDA(cudaMalloc(&b.srec_b[k], n * 16));
    }
  }
  RTB_CUDA(cudaMalloc(&b.ctr, sizeof(WfCounters) + sizeof(WfSticky)));
  b.cap_nodes = cap;
  sl.wf_lights = lights;
  sl.wf_bytes = n * (per_node_bytes(lights) + (split ? 8 + 32 * (size_t)lights : 0));
  return RTB_OK;
}

}  // namespace

void wf_free() {
  for (WfSlot& sl : g_rtb.slot) {
    slot_free(sl);
    if (sl.stream2) cudaStreamDestroy(sl.stream2);
    if (sl.ev_primary) cudaEventDestroy(sl.ev_primary);
    if (sl.ev_early) cudaEventDestroy(sl.ev_early);
    if (sl.ev_async) cudaEventDestroy(sl.ev_async);
    if (sl.ev_half) cudaEventDestroy(sl.ev_half);
    if (sl.h_sticky) cudaFreeHost(sl.h_sticky);
    if (sl.h_tab) cudaFreeHost(sl.h_tab);
    if (sl.d_ptab) cudaFree(sl.d_ptab);
    sl = WfSlot{};
  }
}

static int finish_slot(WfSlot& sl) {
  RtbScene_* s = sl.pending_scene;
  if (!s) return RTB_OK;
  sl.pending_scene = nullptr;
  RTB_CUDA(cudaEventSynchronize(sl.ev_async));
  const WfSticky& h = *static_cast<const WfSticky*>(sl.h_sticky);
  if (h.overflow) {
    s->deferred_status = RTB_ERR_OVERFLOW;
    if (sl.pending_rows <= 1) {
      rtb_set_error("ray queues overflow even with one-row bands");
      return RTB_OK;
    }
    // the same call will now use smaller bands (two frames in flight may both report: keep the smaller)
    const int hint = std::max(1, sl.pending_rows / 2);
    s->wf_rows_hint = s->wf_rows_hint > 0 ? std::min(s->wf_rows_hint, hint) : hint;
  }
  return RTB_OK;
}

int rtb_finish_pending(int slot) {
  if (slot >= 0) return finish_slot(g_rtb.slot[slot % RTB_MAX_IN_FLIGHT]);
  // oldest first: next_slot is the one that would be reused next
  for (int i = 0; i < RTB_MAX_IN_FLIGHT; i++) {
    const int rc = finish_slot(g_rtb.slot[(g_rtb.next_slot + i) % RTB_MAX_IN_FLIGHT]);
    if (rc != RTB_OK) return rc;
  }
  return RTB_OK;
}

// Kernel launch with a per-launch priority (cudaLaunchAttributePriority): with several passes in flight the short
// HBM-bound kernels (raygen, hit, post, compose, resolve) sit on their pass's critical path between two long
// sweeps; at high priority their blocks are dispatched before the pending sweep blocks of the other passes as
// soon as an SM has room.  prio 0 = plain launch.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), int grid, size_t smem, cudaStream_t st, int prio, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(WF_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (prio != 0) {
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = prio;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int wf_render(RtbScene_* s, const RtbCamera* cams, int n_frames, const DevFrameT<float>& fr32, const RowMap& rm,
              uint8_t* const* d_outs, cudaStream_t st, RtbStats* stats, bool async, int slot) {
  if (stats) async = false;
  const int ps = slot % RTB_MAX_IN_FLIGHT;
  WfSlot& sl = g_rtb.slot[ps];
  const RtbCamera* cam = &cams[0];                         // size and options are common to the group
  const int W = cam->width;
  const int n_rows = rm.n_local();
  if (n_rows <= 0 || n_frames <= 0) return RTB_OK;
  if (n_frames > RTB_MAX_GROUP) { rtb_set_error("a frame group holds at most %d frames", RTB_MAX_GROUP); return RTB_ERR_INVALID; }
  const int n = s->n_spheres;
  const int n_pad = s->n_pad;
  const size_t tab_bytes = (size_t)n_pad * sizeof(float4);
  const size_t smem = 2 * tab_bytes;                       // packed sweep table + exact table of the slow path
  if (smem > g_rtb.smem_optin) {
    rtb_set_error("wavefront kernels stage all %d spheres in shared memory (%zu B > %zu B available)",
                  n, smem, g_rtb.smem_optin);
    return RTB_ERR_UNSUPPORTED;
  }
  if (n == 0) {
    rtb_set_error("wavefront path needs at least one sphere (use RTB_KERNEL_PIXEL)");
    return RTB_ERR_UNSUPPORTED;
  }

  WfParams P;
  memset(&P, 0, sizeof P);

  // primary tables: per-sphere constants of rays leaving a camera origin, in fp64, rounded once; one table
  // per run of consecutive frames with the same origin (a fixed camera: one run).  Tables and their pinned
  // staging belong to the workspace slot: the slot's previous pass has been collected before it is reused,
  // so neither is still being read and no host wait is needed here.  Cached per (scene, origin).
  {
    const size_t need = (size_t)RTB_MAX_GROUP * 2 * tab_bytes;
    if (sl.ptab_bytes < need) {
      if (sl.d_ptab) { RTB_CUDA(cudaDeviceSynchronize()); cudaFree(sl.d_ptab); cudaFreeHost(sl.h_tab); }
      sl.d_ptab = nullptr; sl.h_tab = nullptr; sl.ptab_bytes = 0;
      RTB_CUDA(cudaMalloc(&sl.d_ptab, need));
      RTB_CUDA(cudaMallocHost(&sl.h_tab, need));
      sl.ptab_bytes = need;
      for (int g = 0; g < RTB_MAX_GROUP; g++) sl.tab_scene[g] = 0;
    }
    int n_runs = 0;
    bool uploaded = false;
    for (int f = 0; f < n_frames; f++) {
      const double* o = cams[f].origin;
      if (f > 0 && o[0] == cams[f - 1].origin[0] && o[1] == cams[f - 1].origin[1] && o[2] == cams[f - 1].origin[2]) continue;
      const int g = n_runs++;
      P.run_begin[g] = f;
      float4* d_tab = reinterpret_cast<float4*>(sl.d_ptab + (size_t)g * 2 * tab_bytes);
      float4* d_x = d_tab + n_pad;
      if (sl.tab_scene[g] != s->uid || sl.tab_n_pad[g] != n_pad || sl.tab_origin[g][0] != o[0] ||
          sl.tab_origin[g][1] != o[1] || sl.tab_origin[g][2] != o[2]) {
        float4* h_tab = reinterpret_cast<float4*>(sl.h_tab + (size_t)g * 2 * tab_bytes);
        wf_make_table(s->spheres.data(), n, n_pad, o, h_tab, h_tab + n_pad);
        RTB_CUDA(cudaMemcpyAsync(d_tab, h_tab, 2 * tab_bytes, cudaMemcpyHostToDevice, st));
        uploaded = true;
        sl.tab_scene[g] = s->uid; sl.tab_n_pad[g] = n_pad;
        for (int c = 0; c < 3; c++) sl.tab_origin[g][c] = o[c];
      }
      P.primary_tab[g] = d_tab;
      P.primary_x[g] = d_x;
    }
    P.run_begin[n_runs] = n_frames;
    P.n_runs = n_runs;
    if (uploaded && !async) RTB_CUDA(cudaStreamSynchronize(st));   // synchronous calls may come back on another stream
  }

  const int g = fr32.spp_grid, spp = g * g;
  const uint64_t sub_per_row = (uint64_t)W * spp * (uint64_t)n_frames;   // level-1 nodes per band row, all frames
  const uint64_t n_sub_total = sub_per_row * (uint64_t)n_rows;
  // queue capacity: room for ~2 extra nodes per primary ray, bounded; RTB_WF_CAP_NODES overrides (tests)
  uint64_t cap = std::min<uint64_t>(std::max<uint64_t>(3 * n_sub_total, 1u << 20), 1u << 25);
  if (const char* e = getenv("RTB_WF_CAP_NODES")) { const long long v = atoll(e); if (v > 0) cap = (uint64_t)v; }
  if (cap < 4 * sub_per_row) cap = 4 * sub_per_row;      // at least one row per band with headroom
  if (cap > 0xfffffff0ull) { rtb_set_error("frame too wide for the wavefront queues"); return RTB_ERR_UNSUPPORTED; }
  // Pipeline choice: RTB_WF_SPLIT=1 separates every level into a sweep kernel and a hit-processing kernel
  // (and every light into an occluder sweep and a post kernel); 0 = the fused persistent kernels.
  // bit 0: primary level, bit 1: levels >= 2 (sweep kernel + hit kernel instead of the fused trace kernel),
  // bit 2: lights (occluder sweep from ray records + post kernel instead of the fused shadow kernel)
  const char* split_env = getenv("RTB_WF_SPLIT");            // read per call: tests compare the pipelines
  const int split_bits = (split_env ? atoi(split_env) : RTB_WF_SPLIT_DEFAULT) & 7;
  const bool split_p = (split_bits & 1) != 0, split_s = (split_bits & 2) != 0, records = (split_bits & 4) != 0;
  const bool split = split_bits != 0;
  int rc = ensure_workspace(sl, s->n_lights, (uint32_t)cap, split);
  if (rc != RTB_OK) return rc;
  // the shared workspace may be larger than this call asked for; band sizes and the overflow
  // check follow the REQUESTED capacity so that results do not depend on earlier calls

  P.sc = s->sc32;
  // fp64 cameras for the primary rays
  DevFrameT<double>& f = P.fr64;
  f.width = cam->width; f.height = cam->height; f.segs = fr32.segs; f.spp_grid = g;
  f.star_mode = fr32.star_mode; f.camera_mode = cam->camera_mode; f.star_seed = fr32.star_seed;
  for (int c = 0; c < 3; c++) {
    f.origin[c] = cam->origin[c]; f.axisX[c] = cam->axisX[c]; f.axisY[c] = cam->axisY[c]; f.axisZ[c] = cam->axisZ[c];
  }
  f.projW = cam->projW; f.projH = cam->projH; f.projD = cam->projD;
  bool aligned = (W % 4 == 0);
  for (int i = 0; i < n_frames; i++) {
    WfCam& c = P.cam[i];
    for (int k = 0; k < 3; k++) {
      c.origin[k] = cams[i].origin[k]; c.axisX[k] = cams[i].axisX[k];
      c.axisY[k] = cams[i].axisY[k]; c.axisZ[k] = cams[i].axisZ[k];
    }
    c.projW = cams[i].projW; c.projH = cams[i].projH; c.projD = cams[i].projD;
    c.spp_grid = g; c.camera_mode = cams[i].camera_mode;
    for (int k = 0; k < 3; k++) P.origin32[i][k] = (float)cams[i].origin[k];
    P.out[i] = d_outs[i];
    aligned = aligned && ((reinterpret_cast<uintptr_t>(d_outs[i]) & 15u) == 0);
  }
  P.n_frames = n_frames;
  P.rm = rm;
  P.b = sl.wf;
  P.b.cap_nodes = (uint32_t)cap;
  if (s->n_lights <= 2) P.b.lq2 = nullptr;                // kernels key the lights-2/3 store on this
  P.sticky = reinterpret_cast<WfSticky*>(reinterpret_cast<char*>(sl.wf.ctr) + sizeof(WfCounters));
  P.geom = reinterpret_cast<const float4*>(s->d_geom32);
  P.geom_tab = s->d_geom_tab;
  for (int k = 0; k < s->n_lights; k++) { P.light_tab[k] = s->d_light_tab[k]; P.light_x[k] = s->d_light_x[k]; }
  P.n_pad = n_pad;
  P.spp = spp;
  P.aligned4 = aligned ? 1 : 0;
  {
    static int fixed_r = -1;
    if (fixed_r < 0) { const char* e = getenv("RTB_WF_FIXED_R"); fixed_r = e ? std::max(0, std::min(3, atoi(e))) : 0; }
    // Three or more passes in flight: other passes' kernels fill this one's tails, so the chunk size
    // that is cheapest per ray wins over the one that ends a lone launch soonest (measured, 1/8 of a 4K
    // frame per GPU: 1.37 -> 1.28 ms per frame).
    P.fixed_r = fixed_r ? fixed_r : (async && g_rtb.frames_in_flight >= 3 ? 1 : 0);
  }

  // persistent grids: as many CTAs as fit, per kernel
  // (the occupancy query costs ~10 us of host time per call: remembered per kernel and table size)
  static std::map<std::pair<const void*, size_t>, int> occupancy;
  auto grid_for = [&](const void* fn, size_t sm) -> int {
    const auto key = std::make_pair(fn, sm);
    auto it = occupancy.find(key);
    if (it == occupancy.end()) {
      int per_sm = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, WF_THREADS, sm);
      if (per_sm < 1) per_sm = 1;
      it = occupancy.emplace(key, per_sm).first;
    }
    return it->second * g_rtb.sm_count;
  };
  // launch priorities (see launch_k): only with several passes in flight
  static int prio_hi = 1, prio_lo = 1;
  if (prio_hi == 1) { cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi); if (prio_hi == 1) prio_hi = 0; }
  int use_prio = (async && g_rtb.frames_in_flight >= 3) ? 1 : 0;
  if (const char* e = getenv("RTB_WF_PRIO")) use_prio = atoi(e);
  const int prio_light = use_prio ? prio_hi : 0, prio_sweep = use_prio ? prio_lo : 0;
  typedef void (*LevelKernel)(WfParams, int);
  const LevelKernel k_trace_p = records ? wf_trace_kernel<true, true> : wf_trace_kernel<true, false>;
  const LevelKernel k_trace_s = records ? wf_trace_kernel<false, true> : wf_trace_kernel<false, false>;
  const LevelKernel k_sweep_p = wf_sweep_kernel<true>, k_sweep_s = wf_sweep_kernel<false>;
  const LevelKernel k_hit_p = records ? wf_hit_kernel<true, true> : wf_hit_kernel<true, false>;
  const LevelKernel k_hit_s = records ? wf_hit_kernel<false, true> : wf_hit_kernel<false, false>;
  if (smem > 48 * 1024) {
    for (const void* fn : {(const void*)k_trace_p, (const void*)k_trace_s, (const void*)k_sweep_p, (const void*)k_sweep_s,
                           (const void*)wf_shadow_kernel, (const void*)wf_ssweep_kernel})
      RTB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int grid_p = grid_for((const void*)(split_p ? k_sweep_p : k_trace_p), smem);
  int grid_s = grid_for((const void*)(split_s ? k_sweep_s : k_trace_s), smem);
  int grid_l = grid_for(records ? (const void*)wf_ssweep_kernel : (const void*)wf_shadow_kernel, smem);
  // Three or more passes in flight: ONE persistent sweep CTA per SM instead of as many as fit.  A lone sweep
  // loses ~3 % (the FMA pipe is almost saturated by 8 warps), but the registers it leaves free let the kernels
  // of the OTHER passes in flight run beside it on the same SM - a second pass's sweep, or its HBM-bound hit /
  // post / compose kernels, whose latency then hides under FMA work instead of owning the SMs between two
  // sweeps (measured, 4K / 1024 spheres, 4 passes in flight: 8.06 -> 7.75 ms per frame; lone frame 8.5 -> 9.7 ms,
  // which is why single frames keep the full grids).  RTB_WF_SWEEP_CTAS="p,s,l" (or one number) overrides.
  {
    int c[3] = {0, 0, 0};
    if (async && g_rtb.frames_in_flight >= 3) c[0] = c[1] = c[2] = 1;
    if (const char* e = getenv("RTB_WF_SWEEP_CTAS")) {
      const int n = sscanf(e, "%d,%d,%d", &c[0], &c[1], &c[2]);
      if (n == 1) c[1] = c[2] = c[0];
    }
    if (c[0] > 0) grid_p = std::min(grid_p, c[0] * g_rtb.sm_count);
    if (c[1] > 0) grid_s = std::min(grid_s, c[1] * g_rtb.sm_count);
    if (c[2] > 0) grid_l = std::min(grid_l, c[2] * g_rtb.sm_count);
  }
  const int grid_h = split ? grid_for((const void*)k_hit_s, 0) : 0;
  const int grid_c = g_rtb.sm_count * 8;

  // bands are sized for 3 nodes per primary ray; RTB_WF_NODES_PER_RAY overrides (tests force a first overflow)
  double per_ray = 3.0;
  if (const char* e = getenv("RTB_WF_NODES_PER_RAY")) { const double v = atof(e); if (v >= 1.0) per_ray = v; }
  int rows_per_band = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)n_rows, (uint64_t)((double)cap / per_ray) / sub_per_row));
  if (s->wf_rows_hint > 0 && s->wf_rows_hint < rows_per_band) rows_per_band = s->wf_rows_hint;   // learnt earlier
  WfSticky& h_sticky = *static_cast<WfSticky*>(sl.h_sticky);
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
  // Passes in flight share the SMs equally (their kernels come from different streams), so the first
  // `frames_in_flight` passes of a sequence all complete together, at the end.  RTB_WF_STAGGER=1 makes a
  // pass wait until the previous one has traced its last recursion level: passes then complete in order, one
  // pass time apart (results leave for the host while the next pass renders), at the price of less mutual
  // tail filling - measured 4 % slower per frame on one B200 (8.77 -> 9.16 ms at 4K), so it is off by default.
  static int stagger = -1;
  if (stagger < 0) { const char* e = getenv("RTB_WF_STAGGER"); stagger = e ? atoi(e) : 0; }
  const bool pipelined = async && g_rtb.frames_in_flight > 1;
  if (pipelined && stagger && g_rtb.half_slot >= 0 && g_rtb.half_slot != ps)
    RTB_CUDA(cudaStreamWaitEvent(st, g_rtb.slot[g_rtb.half_slot].ev_half, 0));
  for (int attempt = 0;; attempt++) {
    launches = 0; bands = 0;
    tags.clear();
    RTB_CUDA(cudaMemsetAsync(P.sticky, 0, sizeof(WfSticky), st));
    for (int l0 = 0; l0 < n_rows; l0 += rows_per_band) {
      P.band_l0 = l0;
      P.band_rows = std::min(rows_per_band, n_rows - l0);
      P.n_sub = (uint32_t)((uint64_t)P.band_rows * sub_per_row);
      P.n_sub_frame = P.n_sub / (uint32_t)n_frames;
      bands++;
      RTB_CUDA(cudaMemsetAsync(sl.wf.ctr, 0, sizeof(WfCounters), st));
      if ((rc = mark(T_START)) != RTB_OK) return rc;
      RTB_CUDA(launch_k(wf_raygen_kernel, grid_c, 0, st, prio_light, P));
      if (split_p) {
        RTB_CUDA(launch_k(k_sweep_p, grid_p, smem, st, prio_sweep, P, 1));
        RTB_CUDA(launch_k(k_hit_p, grid_h, 0, st, prio_light, P, 1));
        launches += 3;
      } else {
        RTB_CUDA(launch_k(k_trace_p, grid_p, smem, st, prio_sweep, P, 1));
        launches += 2;
      }
      if ((rc = mark(T_PRIMARY)) != RTB_OK) return rc;
      // Without instrumentation the light-0 sweep of the level-1 nodes runs on a second stream,
      // concurrently with the secondary trace levels (independent work; fills their tails).
      const bool overlap = !stats && s->n_lights > 0 && fr32.segs > 1;
      if (overlap) {
        RTB_CUDA(cudaEventRecord(sl.ev_primary, st));
        RTB_CUDA(cudaStreamWaitEvent(sl.stream2, sl.ev_primary, 0));
        if (records) {
          RTB_CUDA(launch_k(wf_ssweep_kernel, grid_l, smem, sl.stream2, prio_sweep, P, 0, 1));
          RTB_CUDA(launch_k(wf_spost_kernel, grid_h, 0, sl.stream2, prio_light, P, 0, 1));
          launches += 2;
        } else {
          RTB_CUDA(launch_k(wf_shadow_kernel, grid_l, smem, sl.stream2, prio_sweep, P, 0, 1));
          launches++;
        }
        RTB_CUDA(cudaEventRecord(sl.ev_early, sl.stream2));
      }
      for (int L = 2; L <= fr32.segs; L++) {
        if (split_s) {
          RTB_CUDA(launch_k(k_sweep_s, grid_s, smem, st, prio_sweep, P, L));
          RTB_CUDA(launch_k(k_hit_s, grid_h, 0, st, prio_light, P, L));
          launches += 2;
        } else {
          RTB_CUDA(launch_k(k_trace_s, grid_s, smem, st, prio_sweep, P, L));
          launches++;
        }
      }
      if ((rc = mark(T_SECONDARY)) != RTB_OK) return rc;
      if (pipelined && l0 + rows_per_band >= n_rows) {       // last band: the next pass may start now
        RTB_CUDA(cudaEventRecord(sl.ev_half, st));
        g_rtb.half_slot = ps;
      }
      if (overlap) RTB_CUDA(cudaStreamWaitEvent(st, sl.ev_early, 0));
      for (int k = 0; k < s->n_lights; k++) {
        const int phase = (k == 0 && overlap) ? 2 : 0;
        if (records) {
          RTB_CUDA(launch_k(wf_ssweep_kernel, grid_l, smem, st, prio_sweep, P, k, phase));
          RTB_CUDA(launch_k(wf_spost_kernel, grid_h, 0, st, prio_light, P, k, phase));
          launches += 2;
        } else {
          RTB_CUDA(launch_k(wf_shadow_kernel, grid_l, smem, st, prio_sweep, P, k, phase));
          launches++;
        }
      }
      if ((rc = mark(T_SHADOW)) != RTB_OK) return rc;
      for (int L = fr32.segs; L >= 2; L--) {
        RTB_CUDA(launch_k(wf_compose_kernel, grid_c, 0, st, prio_light, P, L));
        launches++;
      }
      RTB_CUDA(launch_k(wf_resolve_kernel, grid_c, 0, st, prio_light, P));
      launches++;
      if ((rc = mark(T_COMPOSE)) != RTB_OK) return rc;
      RTB_CUDA(cudaGetLastError());
    }
    RTB_CUDA(cudaMemcpyAsync(&h_sticky, P.sticky, sizeof h_sticky, cudaMemcpyDeviceToHost, st));
    if (async) {                                          // verdict collected by rtb_finish_pending()
      RTB_CUDA(cudaEventRecord(sl.ev_async, st));
      sl.pending_scene = s;
      sl.pending_rows = rows_per_band;
      g_rtb.last_slot = ps;
      return RTB_OK;
    }
    RTB_CUDA(cudaStreamSynchronize(st));
    if (!h_sticky.overflow) { if (attempt > 0) s->wf_rows_hint = rows_per_band; break; }
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
