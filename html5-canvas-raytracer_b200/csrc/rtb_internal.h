// rtb_internal.h — library-private state shared by abi.cu, pixel_f32/f64.cu and wavefront.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>
#include "rtb_device.cuh"

// ---- wavefront (fp32 throughput path) device buffers --------------------------------------------------
// Nodes = intersectWorld invocations (main.js:220) of the current row band, numbered level by level:
// level 1 (primary rays) occupies [0, n_sub); the children created while level L is traced are
// appended contiguously and form level L+1.  All arrays are SoA over the node index.
struct WfCounters {          // one small device struct, zeroed per band
  uint32_t node_count;       // next free node
  uint32_t blocks_done;      // last-block ticket for level bookkeeping
  uint32_t overflow;         // set when a child did not fit in cap_nodes
  uint32_t _pad;
  uint32_t level_end[RTB_MAX_SEGS + 1];   // level L = [level_end[L-1], level_end[L]); level_end[0] = 0
  uint32_t sq_count[RTB_MAX_LIGHTS];      // shadow queue fill per light
  uint32_t work_trace[RTB_MAX_SEGS + 2];   // next chunk of level L (persistent-kernel work counters)
  uint32_t work_primary[RTB_MAX_GROUP];    // level 1: one counter pair per run of frames sharing a camera origin
  uint32_t work_primary_tail[RTB_MAX_GROUP];
  uint32_t work_shadow[RTB_MAX_LIGHTS];
  uint32_t work_trace_tail[RTB_MAX_SEGS + 2];   // small-chunk tail phase of the same launches
  uint32_t work_shadow_tail[RTB_MAX_LIGHTS];
  uint32_t work_shadow_early_tail;
  uint32_t work_shadow_early;              // chunks of the early (level-1) pass of light 0
  uint32_t sq_snap0;                       // fill of shadow queue 0 when the primary trace finished
  unsigned long long stat_nodes, stat_shadow, stat_tests;
};

struct WfBuffers {
  uint32_t cap_nodes;
  float4* ray_o;   // ox, oy, oz, dx                        (levels >= 2)
  float4* ray_d;   // dy, dz, weight (albedo[3|4] of parent), sphere the ray starts on (hit_i of the parent, int bits)
  int4* link;      // parent node, meta = path code | segs_left << 24, level-1 ancestor (frame, pixel, sub-sample), unused
  float4* pos;     // hit.p.xyz, hit_i (int bits); hit_i < 0: no hit (main.js:231)
  float4* base;    // sampler rgb, facing-light mask (int bits)
  float4* lq;      // q0, s0, q1, s1   (q = shadow_dot / light_mag, s = pow(spec_dot, se))
  float4* lq2;     // same for lights 2, 3 (only allocated when n_lights > 2)
  float4* acc;     // light_intensity carry, diffuse sum, specular sum, unused
  float4* child_a; // reflect_color (already scaled by albedo[3]) written by the child's compose
  float4* child_b; // refract_color (scaled by albedo[4])
  uint32_t* sq[RTB_MAX_LIGHTS];   // shadow queues: node indices waiting for light k's occluder sweep (fused pipeline);
                                  // split pipeline: surviving light_intensity of queue entry e (float bits)
  // split pipeline only (sweep kernels + hit / post kernels):
  unsigned long long* key;        // closest hit of the node's ray: (t bits, sphere index, flip) or KEY_NONE
  float4* srec_a[RTB_MAX_LIGHTS]; // shadow ray records per light: nw.xyz (= -(direction light -> p)), light_len
  int4* srec_b[RTB_MAX_LIGHTS];   //   node, sphere the ray leaves (hit_i), light_intensity carried so far (float bits), unused
  WfCounters* ctr;
};

constexpr int RTB_MAX_IN_FLIGHT = 8;   // workspace slots: asynchronous wavefront frames that may overlap

struct RtbScene_ {
  int n_spheres = 0, n_textures = 0, n_lights = 0;
  std::vector<RtbSphere> spheres;         // host copy (primary table is camera dependent)
  double lights[RTB_MAX_LIGHTS][3];
  // device, both precisions
  double* d_geom64 = nullptr;
  float* d_geom32 = nullptr;
  DevMaterial<double>* d_mtl64 = nullptr;
  DevMaterial<float>* d_mtl32 = nullptr;
  DevTexture* d_tex = nullptr;
  uchar4* d_texels = nullptr;
  DevSceneT<double> sc64;
  DevSceneT<float> sc32;
  // wavefront tables: per light (C - light, |C - light|^2 - r2), computed in fp64 once per scene;
  // the primary table (C - origin, |C - origin|^2 - r2) is rebuilt per camera.
  int n_pad = 0;                          // n_spheres rounded up to RTB_WF_GROUP
  float4* d_geom_tab = nullptr;           // packed general-ray table (wf_make_geom_table)
  float4* d_light_tab[RTB_MAX_LIGHTS] = {nullptr, nullptr, nullptr, nullptr};   // packed, per light
  float4* d_light_x[RTB_MAX_LIGHTS] = {nullptr, nullptr, nullptr, nullptr};     // exact (L, c), per light
  uint64_t uid = 0;                       // never reused: keys the per-slot primary-table cache (WfSlot)
  char* arena = nullptr;                  // single device allocation all the pointers above point into
  size_t arena_bytes = 0;
  size_t h2d_bytes = 0;                   // bytes rtb_scene_create uploaded
  // asynchronous device renders (rtb_render_device without stats): the queue-overflow verdict of the
  // last enqueued frame is collected by rtb_finish() / the next call on the library
  int wf_rows_hint = 0;                   // band size known to fit (0 = none yet)
  int deferred_status = 0;                // RTB_OK or RTB_ERR_OVERFLOW, reported by rtb_finish()
};

// Everything one wavefront frame in flight owns.  With rtb_set_frames_in_flight(n) consecutive
// asynchronous device renders rotate through n slots, so that (issued on different streams) the kernels
// of frame i+1 fill the tails of frame i's short per-level launches.
struct WfSlot {
  WfBuffers wf = {};
  int wf_lights = 0;                          // shadow queues allocated
  size_t wf_bytes = 0;
  cudaStream_t stream2 = nullptr;             // overlapped early shadow pass
  cudaEvent_t ev_primary = nullptr, ev_early = nullptr, ev_async = nullptr;
  cudaEvent_t ev_half = nullptr;              // recorded after the last trace level: the next pass in flight may start
  void* h_sticky = nullptr;                   // pinned: overflow flag + device counters of the last call
  // per-camera-origin primary tables of the slot's frame group (packed + exact, RTB_MAX_GROUP sets), their
  // pinned staging, and what each set currently holds (scene uid, padded size, origin)
  char* d_ptab = nullptr;
  char* h_tab = nullptr;
  size_t ptab_bytes = 0;
  uint64_t tab_scene[RTB_MAX_GROUP] = {};
  int tab_n_pad[RTB_MAX_GROUP] = {};
  double tab_origin[RTB_MAX_GROUP][3] = {};
  RtbScene_* pending_scene = nullptr;         // scene with an asynchronous frame in flight in this slot
  int pending_rows = 0;                       // band size of that frame
};

struct RtbGlobal {
  bool inited = false;
  int device = -1;
  int sm_count = 0;
  int clock_khz = 0;
  size_t smem_optin = 0;
  char name[128] = {0};
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  cudaEvent_t frame_done[RTB_MAX_IN_FLIGHT] = {}, copy_done[RTB_MAX_IN_FLIGHT] = {};
  cudaStream_t batch_stream[RTB_MAX_IN_FLIGHT] = {};   // rtb_render_batch: one per frame in flight (created on first use)
  // library-owned output staging + wavefront workspaces (one per frame in flight), shared by all scenes
  // (the entry points serialise on a mutex), so creating a scene stays cheap
  uint8_t* d_frame[RTB_MAX_IN_FLIGHT] = {};
  size_t frame_cap[RTB_MAX_IN_FLIGHT] = {};
  WfSlot slot[RTB_MAX_IN_FLIGHT];
  int frames_in_flight = 1;                   // 1 .. RTB_MAX_IN_FLIGHT (rtb_set_frames_in_flight)
  int next_slot = 0, last_slot = 0;
  int half_slot = -1;                         // slot of the asynchronous pass enqueued last (its ev_half), -1: none
  unsigned long long* d_counters = nullptr;   // 3 x u64 for the pixel kernels
  uint32_t* d_peer_counter = nullptr;         // last-block election words of rtb_peer_scatter
  char* cached_arena = nullptr;               // last destroyed scene's device arena, reused by the next create
  size_t cached_arena_bytes = 0;
  char* h_stage = nullptr;                    // pinned staging for scene uploads and per-camera tables
  size_t stage_bytes = 0;
  char* d_star = nullptr;                     // RTB_STARS_STREAM scratch: counts, bases, stream, scan temp
  size_t star_bytes = 0;
  // per-kernel-class event pool for RtbStats.ms_* (only used when stats are requested)
  static const int kMaxTimers = 256;
  cudaEvent_t timers[kMaxTimers] = {};
  int n_timers = 0;
};
extern RtbGlobal g_rtb;

void rtb_set_error(const char* fmt, ...);
// One render / scene / peer operation at a time per process: workspaces, staging buffers, streams and the
// lazily created peer counters are shared library state.
std::recursive_mutex& rtb_api_mutex();
#define RTB_API_LOCK() std::lock_guard<std::recursive_mutex> api_lock_(rtb_api_mutex())
#define RTB_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      rtb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return e_ == cudaErrorMemoryAllocation ? RTB_ERR_OOM : RTB_ERR_CUDA;                 \
    }                                                                                      \
  } while (0)

// pixel kernels (one TU per precision so that only the fp64 one is built with -fmad=false)
int launch_pixel_f64(const DevSceneT<double>& sc, const DevFrameT<double>& fr, const RowMap& rm,
                     uint8_t* d_out, unsigned long long* d_counters, cudaStream_t st, uint32_t* launches);
int launch_pixel_f32(const DevSceneT<float>& sc, const DevFrameT<float>& fr, const RowMap& rm,
                     uint8_t* d_out, unsigned long long* d_counters, cudaStream_t st, uint32_t* launches);

// Packed sweep tables of the wavefront kernels.  Spheres are stored in PAIRS so that one packed FP32
// instruction (FFMA2, sm_100 f32x2) tests two spheres: pair p = spheres (2p, 2p+1) occupies
//   tab[2p]   = (x0, x1, y0, y1)        tab[2p+1] = (z0, z1, w0, w1)
// and the table is padded to a multiple of RTB_WF_GROUP spheres with dummies that never test positive.
// A test value whose SIGN BIT IS SET marks a candidate; all table arithmetic is fp64, one rounding.
//
// wf_make_table — rays that all start at `org` (camera origin, or a light for the occluder sweeps):
//   (x,y,z) = L = C - org,  w = sc = sqrt(|L|^2 - r2);  test value  sc - dir.L  <= 0  <=>  the line hits
//   the sphere and the sphere is ahead of org.  sc is shrunk slightly so that fp32 rounding of the dot
//   product can only ADD candidates (the slow path re-decides exactly).  org inside the sphere: -inf
//   (always a candidate); dummy: +inf.   exact[i] = (L, |L|^2 - r2) for the slow path.
// wf_make_geom_table — general rays: (x,y,z) = C, w = -r2 (1 + 2e-6); test value d2 - r2' (main.js:424-425).
constexpr int RTB_WF_GROUP = 32;
inline int wf_pad(int n) { return (n + RTB_WF_GROUP - 1) / RTB_WF_GROUP * RTB_WF_GROUP; }
inline void wf_pack(float4* tab, int i, float x, float y, float z, float w) {
  float* a = reinterpret_cast<float*>(&tab[2 * (i / 2)]);
  float* b = reinterpret_cast<float*>(&tab[2 * (i / 2) + 1]);
  const int h = i & 1;
  a[h] = x; a[2 + h] = y; b[h] = z; b[2 + h] = w;
}
inline void wf_make_table(const RtbSphere* sp, int n, int n_pad, const double org[3], float4* tab, float4* exact) {
  const float inf = __builtin_huge_valf();
  for (int i = 0; i < n_pad; i++) {
    if (i >= n) { wf_pack(tab, i, 0.f, 0.f, 0.f, inf); exact[i] = make_float4(0.f, 0.f, 0.f, inf); continue; }
    const double L0 = sp[i].origin[0] - org[0], L1 = sp[i].origin[1] - org[1], L2 = sp[i].origin[2] - org[2];
    const double cc = (L0 * L0 + L1 * L1 + L2 * L2) - sp[i].r2;
    float sc = -inf;
    if (cc > 0) {
      const double len = __builtin_sqrt(L0 * L0 + L1 * L1 + L2 * L2);
      sc = (float)(__builtin_sqrt(cc) * (1.0 - 1e-6) - 4e-7 * len);
    }
    wf_pack(tab, i, (float)L0, (float)L1, (float)L2, sc);
    exact[i] = make_float4((float)L0, (float)L1, (float)L2, (float)cc);
  }
}
// The fast test evaluates e = a^2 + b^2 - r2' with a = u.C - u.org, b = v.C - v.org in fp32.  Near the
// decision boundary |a|,|b| <= r and the absolute error of a, b is a few ulp of (|C| + |org|), so
// err(e) <~ 1e-6 r (|C| + R), R = bound on |org| (ray origins lie on sphere surfaces).  r2' carries that
// margin so that rounding can only ADD candidates.
inline void wf_make_geom_table(const RtbSphere* sp, int n, int n_pad, float4* tab) {
  const float inf = __builtin_huge_valf();
  double R = 0;
  for (int i = 0; i < n; i++) {
    const double c = __builtin_sqrt(sp[i].origin[0] * sp[i].origin[0] + sp[i].origin[1] * sp[i].origin[1] +
                                    sp[i].origin[2] * sp[i].origin[2]);
    const double e = c + __builtin_sqrt(sp[i].r2 > 0 ? sp[i].r2 : 0);
    if (e > R) R = e;
  }
  for (int i = 0; i < n_pad; i++) {
    if (i >= n) { wf_pack(tab, i, 0.f, 0.f, 0.f, inf); continue; }
    const double c = __builtin_sqrt(sp[i].origin[0] * sp[i].origin[0] + sp[i].origin[1] * sp[i].origin[1] +
                                    sp[i].origin[2] * sp[i].origin[2]);
    const double r = __builtin_sqrt(sp[i].r2 > 0 ? sp[i].r2 : 0);
    wf_pack(tab, i, (float)sp[i].origin[0], (float)sp[i].origin[1], (float)sp[i].origin[2],
            -(float)(sp[i].r2 * (1.0 + 2e-6) + 2e-6 * r * (c + R)));
  }
}

// wavefront pipeline: renders the rows of `rm` into d_out, splitting into bands as needed.
// async = true: enqueue only; the overflow check is deferred to rtb_finish_pending().
// `cams` / `d_outs`: the n_frames (<= RTB_MAX_GROUP) frames of one group - same scene, size and options.
int wf_render(RtbScene_* s, const RtbCamera* cams, int n_frames, const DevFrameT<float>& fr, const RowMap& rm,
              uint8_t* const* d_outs, cudaStream_t st, RtbStats* stats, bool async, int slot);
// Wait for the asynchronous frame in flight in `slot` (-1: in every slot, oldest first) and file its
// verdict in its scene.
int rtb_finish_pending(int slot);
void wf_free();
