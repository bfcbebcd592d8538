// abi.cu — implementation of include/rtb200.h: device setup, scene upload, the render entry points
// (host buffer, device buffer, frame batch), stripe bookkeeping for the multi-GPU layer, errors.
// No CPU fallback anywhere: without a CUDA device every computing entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <cub/device/device_scan.cuh>
#include "rtb_internal.h"

RtbGlobal g_rtb;
static const int RTB_AUTO_PIXEL_MAX_SPHERES = 32;
static const int RTB_BATCH_IN_FLIGHT = 4;     // frames rtb_render_batch keeps in flight
static thread_local char t_err[512] = "";
// One render / scene operation at a time per process: the workspace, staging buffers and streams are
// shared library state (the reference's caller is the single JS thread; this keeps other hosts safe).
static std::recursive_mutex g_api_mutex;
std::recursive_mutex& rtb_api_mutex() { return g_api_mutex; }

void rtb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof t_err, fmt, ap);
  va_end(ap);
}

namespace {

template <typename T>
void fill_material(DevMaterial<T>& d, const RtbMaterial& m) {
  for (int i = 0; i < 3; i++) d.color[i] = (T)m.color[i];
  for (int i = 0; i < 5; i++) d.albedo[i] = (T)m.albedo[i];
  d.specular_exponent = (T)m.specular_exponent;
  d.refract_index = (T)m.refract_index;
  d.checker_freq[0] = (T)m.checker_freq[0];
  d.checker_freq[1] = (T)m.checker_freq[1];
  for (int c = 0; c < 2; c++)
    for (int i = 0; i < 3; i++) d.checker_color[c][i] = (T)m.checker_color[c][i];
  d.star_threshold = (T)m.star_threshold;
  d.star_gain = (T)m.star_gain;
  d.sampler_kind = m.sampler_kind;
  d.texture_unit = m.texture_unit;
}

template <typename T>
void fill_scene_view(DevSceneT<T>& v, const RtbScene_* s, const RtbSceneDesc* d, const T* geom,
                     const DevMaterial<T>* mtl) {
  v.n_spheres = s->n_spheres;
  v.n_textures = s->n_textures;
  v.n_lights = s->n_lights;
  v._pad = 0;
  v.geom = geom;
  v.mtl = mtl;
  v.tex = s->d_tex;
  v.texels = s->d_texels;
  for (int k = 0; k < RTB_MAX_LIGHTS; k++)
    for (int c = 0; c < 3; c++) v.lights[k][c] = k < s->n_lights ? (T)d->lights[k][c] : (T)0;
  v.light_intensity = (T)d->light_intensity;
  v.eps = (T)d->eps;
}

int validate_scene(const RtbSceneDesc* d) {
  if (!d) { rtb_set_error("scene description is NULL"); return RTB_ERR_INVALID; }
  if (d->abi_version != RTB_ABI_VERSION) {
    rtb_set_error("abi_version %u, library speaks %u", d->abi_version, RTB_ABI_VERSION);
    return RTB_ERR_INVALID;
  }
  if (d->n_spheres < 0 || d->n_spheres > (1 << 20)) { rtb_set_error("n_spheres out of range"); return RTB_ERR_INVALID; }
  if (d->n_spheres > 0 && !d->spheres) { rtb_set_error("spheres is NULL"); return RTB_ERR_INVALID; }
  if (d->n_textures < 0 || d->n_textures > 4096) { rtb_set_error("n_textures out of range"); return RTB_ERR_INVALID; }
  if (d->n_textures > 0 && !d->textures) { rtb_set_error("textures is NULL"); return RTB_ERR_INVALID; }
  if (d->n_lights < 0 || d->n_lights > RTB_MAX_LIGHTS) {
    rtb_set_error("n_lights must be 0..%d", RTB_MAX_LIGHTS);
    return RTB_ERR_INVALID;
  }
  if (!(d->eps == d->eps) || !(d->light_intensity == d->light_intensity)) {
    rtb_set_error("eps / light_intensity is NaN");
    return RTB_ERR_INVALID;
  }
  for (int t = 0; t < d->n_textures; t++) {
    const RtbTexture& tx = d->textures[t];
    if (tx.width <= 0 || tx.height <= 0 || tx.width > 32768 || tx.height > 32768 || !tx.texels) {
      rtb_set_error("texture %d: bad size %dx%d or NULL texels", t, tx.width, tx.height);
      return RTB_ERR_INVALID;
    }
  }
  for (int i = 0; i < d->n_spheres; i++) {
    const RtbMaterial& m = d->spheres[i].mtl;
    if (m.sampler_kind < RTB_SAMPLER_COLOR || m.sampler_kind > RTB_SAMPLER_STARS) {
      rtb_set_error("sphere %d: unknown sampler_kind %d", i, m.sampler_kind);
      return RTB_ERR_INVALID;
    }
    if (m.sampler_kind == RTB_SAMPLER_TEXTURE && (m.texture_unit < 0 || m.texture_unit >= d->n_textures)) {
      rtb_set_error("sphere %d: texture_unit %d outside 0..%d", i, m.texture_unit, d->n_textures - 1);
      return RTB_ERR_INVALID;
    }
  }
  return RTB_OK;
}

int validate_render(const RtbScene_* s, const RtbCamera* cam, const RtbRenderOpts* o, const void* out) {
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!s || !cam || !o || !out) { rtb_set_error("NULL argument"); return RTB_ERR_INVALID; }
  if (cam->width <= 0 || cam->height <= 0 || (long long)cam->width * cam->height > (1ll << 31)) {
    rtb_set_error("bad canvas size %dx%d", cam->width, cam->height);
    return RTB_ERR_INVALID;
  }
  if (o->segs < 0 || o->segs > RTB_MAX_SEGS) { rtb_set_error("segs must be 0..%d", RTB_MAX_SEGS); return RTB_ERR_INVALID; }
  if (o->spp_grid < 1 || o->spp_grid > 16) { rtb_set_error("spp_grid must be 1..16"); return RTB_ERR_INVALID; }
  if (o->precision != RTB_PRECISION_FP32 && o->precision != RTB_PRECISION_FP64) {
    rtb_set_error("unknown precision %d", o->precision);
    return RTB_ERR_INVALID;
  }
  if (o->kernel < RTB_KERNEL_AUTO || o->kernel > RTB_KERNEL_WAVEFRONT) {
    rtb_set_error("unknown kernel %d", o->kernel);
    return RTB_ERR_INVALID;
  }
  if (o->kernel == RTB_KERNEL_WAVEFRONT && o->precision == RTB_PRECISION_FP64) {
    rtb_set_error("the wavefront pipeline is fp32 only");
    return RTB_ERR_UNSUPPORTED;
  }
  if (o->star_mode < RTB_STARS_OFF || o->star_mode > RTB_STARS_STREAM) {
    rtb_set_error("unknown star_mode %d", o->star_mode);
    return RTB_ERR_INVALID;
  }
  if (o->star_mode == RTB_STARS_STREAM) {
    if (!o->star_stream || o->star_stream_len == 0) { rtb_set_error("RTB_STARS_STREAM needs star_stream"); return RTB_ERR_INVALID; }
    const bool whole = o->y0 == 0 && (o->y1 <= 0 || o->y1 == cam->height) && o->stripe_world <= 1 && !o->out_packed;
    if ((long long)cam->width * cam->height > 0x7fffffffll) {    // the raster-order scan counts pixels in an int
      rtb_set_error("RTB_STARS_STREAM: canvas of %dx%d pixels exceeds 2^31 - 1", cam->width, cam->height);
      return RTB_ERR_UNSUPPORTED;
    }
    if (!whole || o->kernel == RTB_KERNEL_WAVEFRONT || o->precision != RTB_PRECISION_FP64) {
      rtb_set_error("RTB_STARS_STREAM replays the reference's call order: whole frames on the fp64 validation "
                    "kernel only (one differing hit/miss in fp32 would shift every later stream index)");
      return RTB_ERR_UNSUPPORTED;
    }
  }
  if (cam->camera_mode != RTB_CAMERA_REFERENCE && cam->camera_mode != RTB_CAMERA_BASIS) {
    rtb_set_error("unknown camera_mode %d", cam->camera_mode);
    return RTB_ERR_INVALID;
  }
  if (o->y0 < 0 || o->y0 > cam->height || o->y1 > cam->height || (o->y1 > 0 && o->y1 < o->y0)) {
    rtb_set_error("bad row range [%d,%d) for height %d", o->y0, o->y1, cam->height);
    return RTB_ERR_INVALID;
  }
  if (o->stripe_world > 1 && (o->stripe_rows <= 0 || o->stripe_rank < 0 || o->stripe_rank >= o->stripe_world)) {
    rtb_set_error("bad stripe spec rows=%d rank=%d world=%d", o->stripe_rows, o->stripe_rank, o->stripe_world);
    return RTB_ERR_INVALID;
  }
  return RTB_OK;
}

RowMap make_rowmap(const RtbCamera* cam, const RtbRenderOpts* o) {
  RowMap rm;
  rm.y0 = o->y0;
  rm.y1 = o->y1 <= 0 ? cam->height : o->y1;
  rm.S = o->stripe_rows;
  rm.rank = o->stripe_rank;
  rm.world = o->stripe_world;
  rm.packed = o->out_packed;
  rm.width = cam->width;
  return rm;
}

template <typename T>
DevFrameT<T> make_frame(const RtbCamera* cam, const RtbRenderOpts* o) {
  DevFrameT<T> f;
  memset(&f, 0, sizeof f);
  f.width = cam->width; f.height = cam->height;
  f.segs = o->segs; f.spp_grid = o->spp_grid;
  f.star_mode = o->star_mode; f.camera_mode = cam->camera_mode;
  f.y0 = o->y0; f.y1 = o->y1 <= 0 ? cam->height : o->y1;
  f.stripe_rows = o->stripe_rows; f.stripe_rank = o->stripe_rank; f.stripe_world = o->stripe_world;
  f.out_packed = o->out_packed;
  f.star_seed = o->star_seed;
  for (int c = 0; c < 3; c++) {
    f.origin[c] = (T)cam->origin[c]; f.axisX[c] = (T)cam->axisX[c];
    f.axisY[c] = (T)cam->axisY[c]; f.axisZ[c] = (T)cam->axisZ[c];
  }
  f.projW = (T)cam->projW; f.projH = (T)cam->projH; f.projD = (T)cam->projD;
  return f;
}

// Launch the kernels of one frame into d_out on stream st.  Synchronises only on the wavefront path
// (queue-overflow check) or when counters are wanted.
int render_group_to_device(RtbScene_* s, const RtbCamera* cams, int n_frames, const RtbRenderOpts* o,
                           uint8_t* const* d_outs, cudaStream_t st, RtbStats* stats, bool async);

int render_to_device(RtbScene_* s, const RtbCamera* cam, const RtbRenderOpts* o, uint8_t* d_out,
                     cudaStream_t st, RtbStats* stats, bool async = false) {
  return render_group_to_device(s, cam, 1, o, &d_out, st, stats, async);
}

// Frames of a group share the scene, the canvas size and the options; the wavefront path renders them in
// one pass (wf_render), the per-pixel kernels one launch per frame.
int render_group_to_device(RtbScene_* s, const RtbCamera* cams, int n_frames, const RtbRenderOpts* o,
                           uint8_t* const* d_outs, cudaStream_t st, RtbStats* stats, bool async) {
  const RtbCamera* cam = &cams[0];
  // Asynchronous frames may still be in flight.  A synchronous call collects them all; an asynchronous
  // one takes the next workspace slot and collects only the frame that used it last (with two frames in
  // flight that is the frame before the previous one, which therefore keeps running).  If a collected
  // frame was THIS scene's and overflowed, the caller has not been told yet that it is invalid.
  int slot = 0;
  if (async && !stats && g_rtb.frames_in_flight > 1) {
    slot = g_rtb.next_slot % g_rtb.frames_in_flight;
    g_rtb.next_slot = (slot + 1) % g_rtb.frames_in_flight;
  }
  int prc = rtb_finish_pending((async && !stats && g_rtb.frames_in_flight > 1) ? slot : -1);
  if (prc != RTB_OK) return prc;
  if (s->deferred_status != RTB_OK) {
    const int d = s->deferred_status;
    s->deferred_status = RTB_OK;
    rtb_set_error("the previous asynchronous frame overflowed its ray queues and is invalid; the band size "
                  "has been reduced - render it again (rtb_finish reports this per frame)");
    return d;
  }
  const RowMap rm = make_rowmap(cam, o);
  int kernel = o->kernel;
  // AUTO: the fused per-pixel kernel wins while a sweep is a handful of spheres (demo scene, N = 14:
  // 0.37 ms vs 0.63 ms at 1080p, 19.7 vs 67.6 ms at 4K/16 spp — no node arrays to write and re-read);
  // the wavefront wins as soon as the sweeps dominate (N = 1024 at 4K: 9.9 ms vs 200 ms).  Measured
  // crossover ~ 24-32 spheres.
  if (kernel == RTB_KERNEL_AUTO)
    kernel = (o->precision == RTB_PRECISION_FP64 || s->n_spheres <= RTB_AUTO_PIXEL_MAX_SPHERES ||
              o->star_mode == RTB_STARS_STREAM ||
              2 * (size_t)s->n_pad * sizeof(float4) > g_rtb.smem_optin)   // tables must fit shared memory
                 ? RTB_KERNEL_PIXEL : RTB_KERNEL_WAVEFRONT;
  uint32_t launches = 0;
  if (kernel == RTB_KERNEL_WAVEFRONT) {
    const DevFrameT<float> fr = make_frame<float>(cam, o);
    return wf_render(s, cams, n_frames, fr, rm, d_outs, st, stats, async, slot);
  }
  if (n_frames > 1) {                                      // per-pixel kernels: frame after frame
    for (int f = 0; f < n_frames; f++) {
      RtbStats one;
      memset(&one, 0, sizeof one);
      const int rc1 = render_group_to_device(s, &cams[f], 1, o, &d_outs[f], st, stats ? &one : nullptr, async);
      if (rc1 != RTB_OK) return rc1;
      if (stats) {
        stats->nodes += one.nodes; stats->shadow_rays += one.shadow_rays; stats->sphere_tests += one.sphere_tests;
        stats->launches += one.launches; stats->bands += one.bands;
      }
    }
    return RTB_OK;
  }
  uint8_t* d_out = d_outs[0];
  unsigned long long* ctr = stats ? g_rtb.d_counters : nullptr;
  if (ctr) RTB_CUDA(cudaMemsetAsync(ctr, 0, 3 * sizeof(unsigned long long), st));
  int rc;
  // RTB_STARS_STREAM: counting pass -> exclusive scan over pixels in raster order -> render pass
  const double* d_stream = nullptr;
  const unsigned long long* d_base = nullptr;
  if (o->star_mode == RTB_STARS_STREAM) {
    const size_t npx = (size_t)cam->width * cam->height;
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)npx, st);
    const size_t need = o->star_stream_len * sizeof(double) + 2 * npx * sizeof(unsigned long long) + temp_bytes + 1024;
    if (g_rtb.star_bytes < need) {
      cudaFree(g_rtb.d_star);
      g_rtb.d_star = nullptr; g_rtb.star_bytes = 0;
      RTB_CUDA(cudaMalloc(&g_rtb.d_star, need));
      g_rtb.star_bytes = need;
    }
    char* base = g_rtb.d_star;
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(base);
    unsigned long long* d_b = d_counts + npx;
    double* d_s = reinterpret_cast<double*>(d_b + npx);
    void* d_temp = reinterpret_cast<char*>(d_s + o->star_stream_len) + ((256 - ((uintptr_t)(d_s + o->star_stream_len) & 255)) & 255);
    RTB_CUDA(cudaMemcpyAsync(d_s, o->star_stream, o->star_stream_len * sizeof(double), cudaMemcpyHostToDevice, st));
    if (o->precision == RTB_PRECISION_FP64) {
      DevFrameT<double> fr = make_frame<double>(cam, o);
      fr.star_counts = d_counts;
      rc = launch_pixel_f64(s->sc64, fr, rm, d_out, nullptr, st, &launches);
    } else {
      DevFrameT<float> fr = make_frame<float>(cam, o);
      fr.star_counts = d_counts;
      rc = launch_pixel_f32(s->sc32, fr, rm, d_out, nullptr, st, &launches);
    }
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_counts, d_b, (int)npx, st));
    unsigned long long last[2];
    RTB_CUDA(cudaMemcpyAsync(&last[0], d_counts + npx - 1, sizeof last[0], cudaMemcpyDeviceToHost, st));
    RTB_CUDA(cudaMemcpyAsync(&last[1], d_b + npx - 1, sizeof last[1], cudaMemcpyDeviceToHost, st));
    RTB_CUDA(cudaStreamSynchronize(st));
    if (last[0] + last[1] > o->star_stream_len) {
      rtb_set_error("star_stream holds %llu values but the frame draws %llu", (unsigned long long)o->star_stream_len,
                    last[0] + last[1]);
      return RTB_ERR_INVALID;
    }
    d_stream = d_s;
    d_base = d_b;
  }
  if (o->precision == RTB_PRECISION_FP64) {
    DevFrameT<double> fr = make_frame<double>(cam, o);
    fr.star_stream = d_stream; fr.star_stream_len = o->star_stream_len; fr.star_base = d_base;
    rc = launch_pixel_f64(s->sc64, fr, rm, d_out, ctr, st, &launches);
  } else {
    DevFrameT<float> fr = make_frame<float>(cam, o);
    fr.star_stream = d_stream; fr.star_stream_len = o->star_stream_len; fr.star_base = d_base;
    rc = launch_pixel_f32(s->sc32, fr, rm, d_out, ctr, st, &launches);
  }
  if (rc != RTB_OK) return rc;
  if (stats) {
    unsigned long long h[3];
    RTB_CUDA(cudaMemcpyAsync(h, ctr, sizeof h, cudaMemcpyDeviceToHost, st));
    RTB_CUDA(cudaStreamSynchronize(st));
    stats->nodes += h[0]; stats->shadow_rays += h[1]; stats->sphere_tests += h[2];
    stats->launches += launches;
    stats->bands += 1;
  }
  return RTB_OK;
}

int ensure_frame(int slot, size_t bytes) {
  if (g_rtb.frame_cap[slot] >= bytes) return RTB_OK;
  if (g_rtb.d_frame[slot]) cudaFree(g_rtb.d_frame[slot]);
  g_rtb.d_frame[slot] = nullptr; g_rtb.frame_cap[slot] = 0;
  RTB_CUDA(cudaMalloc(&g_rtb.d_frame[slot], bytes));
  g_rtb.frame_cap[slot] = bytes;
  return RTB_OK;
}

// D2H of the rows this call rendered.  Packed: one contiguous copy.  Natural layout: the rows sit at
// their frame position; stripes become one strided 2-D copy (+ a ragged tail).
int copy_rows_to_host(const RowMap& rm, const uint8_t* d_frame, uint8_t* h_out, int height, cudaStream_t st) {
  const size_t row_bytes = (size_t)rm.width * 4;
  const int n_local = rm.n_local();
  if (n_local <= 0) return RTB_OK;
  if (rm.packed) {
    RTB_CUDA(cudaMemcpyAsync(h_out, d_frame, row_bytes * n_local, cudaMemcpyDeviceToHost, st));
    return RTB_OK;
  }
  if (rm.world <= 1 || rm.S <= 0) {
    RTB_CUDA(cudaMemcpyAsync(h_out + row_bytes * rm.y0, d_frame + row_bytes * rm.y0,
                             row_bytes * (size_t)(rm.y1 - rm.y0), cudaMemcpyDeviceToHost, st));
    return RTB_OK;
  }
  (void)height;
  // runs of consecutive owned rows
  int l = 0;
  while (l < n_local) {
    const int y = rm.local_to_y(l);
    int run = 1;
    while (l + run < n_local && rm.local_to_y(l + run) == y + run) run++;
    // try to batch equally long, equally spaced runs into one 2-D copy
    int reps = 1;
    const int pitch_rows = rm.S * rm.world;
    while (l + (reps + 1) * run <= n_local && rm.local_to_y(l + reps * run) == y + reps * pitch_rows) {
      bool same = true;
      for (int i = 1; i < run && same; i++) same = rm.local_to_y(l + reps * run + i) == y + reps * pitch_rows + i;
      if (!same) break;
      reps++;
    }
    if (reps > 1) {
      RTB_CUDA(cudaMemcpy2DAsync(h_out + row_bytes * y, row_bytes * pitch_rows, d_frame + row_bytes * y,
                                 row_bytes * pitch_rows, row_bytes * run, reps, cudaMemcpyDeviceToHost, st));
    } else {
      RTB_CUDA(cudaMemcpyAsync(h_out + row_bytes * y, d_frame + row_bytes * y, row_bytes * run,
                               cudaMemcpyDeviceToHost, st));
    }
    l += run * reps;
  }
  return RTB_OK;
}

// One (clock64, globaltimer) pair per SM: two probes bracket a region, cycles / ns = the SM clock the region
// actually ran at - measured on the device, without asking the driver (NVML polling during a multi-GPU
// sequence stalls the exchange for milliseconds).
__global__ void clock_probe_kernel(unsigned long long* out, int max_sms) {
  if (threadIdx.x != 0) return;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if ((int)smid >= max_sms) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  out[2 * smid] = (unsigned long long)clock64();
  out[2 * smid + 1] = t;
}

__global__ void unstripe_kernel(const uint32_t* __restrict__ gathered, uint32_t* __restrict__ frame, int W,
                                int H, int S, int world, int padded_rows) {
  // gathered: [rank][padded_rows][W]; local row l of rank r is frame row (l / S) * S * world + r * S + l % S
  const long long total = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i % W);
    const int stripe = y / S, r = stripe % world, l = (stripe / world) * S + y % S;
    frame[i] = gathered[((size_t)r * padded_rows + l) * W + x];
  }
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------------------
// The library is built with -fvisibility=hidden; only the entry points of include/rtb200.h are exported.
#pragma GCC visibility push(default)
extern "C" {

int rtb_init(int device) {
  RTB_API_LOCK();
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    rtb_set_error("no CUDA device visible (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    (void)cudaGetLastError();
    return RTB_ERR_NO_DEVICE;
  }
  if (device < 0) {
    if (g_rtb.inited) return RTB_OK;
    RTB_CUDA(cudaGetDevice(&device));
  }
  if (device >= count) { rtb_set_error("device %d of %d", device, count); return RTB_ERR_INVALID; }
  if (g_rtb.inited && g_rtb.device == device) { RTB_CUDA(cudaSetDevice(device)); return RTB_OK; }
  if (g_rtb.inited) rtb_shutdown();
  RTB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RTB_CUDA(cudaGetDeviceProperties(&prop, device));
  g_rtb.device = device;
  g_rtb.sm_count = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  g_rtb.clock_khz = khz;
  g_rtb.smem_optin = prop.sharedMemPerBlockOptin;
  snprintf(g_rtb.name, sizeof g_rtb.name, "%s", prop.name);
  RTB_CUDA(cudaStreamCreateWithFlags(&g_rtb.stream, cudaStreamNonBlocking));
  RTB_CUDA(cudaStreamCreateWithFlags(&g_rtb.copy_stream, cudaStreamNonBlocking));
  RTB_CUDA(cudaEventCreate(&g_rtb.ev0));
  RTB_CUDA(cudaEventCreate(&g_rtb.ev1));
  RTB_CUDA(cudaEventCreate(&g_rtb.ev2));
  RTB_CUDA(cudaEventCreate(&g_rtb.ev3));
  for (int i = 0; i < RTB_MAX_IN_FLIGHT; i++) {
    RTB_CUDA(cudaEventCreateWithFlags(&g_rtb.frame_done[i], cudaEventDisableTiming));
    RTB_CUDA(cudaEventCreateWithFlags(&g_rtb.copy_done[i], cudaEventDisableTiming));
  }
  RTB_CUDA(cudaMalloc(&g_rtb.d_counters, 3 * sizeof(unsigned long long)));
  g_rtb.inited = true;
  return RTB_OK;
}

void rtb_shutdown(void) {
  RTB_API_LOCK();
  if (!g_rtb.inited) return;
  cudaDeviceSynchronize();
  cudaStreamDestroy(g_rtb.stream);
  cudaStreamDestroy(g_rtb.copy_stream);
  cudaEventDestroy(g_rtb.ev0); cudaEventDestroy(g_rtb.ev1);
  cudaEventDestroy(g_rtb.ev2); cudaEventDestroy(g_rtb.ev3);
  for (int i = 0; i < RTB_MAX_IN_FLIGHT; i++) {
    cudaEventDestroy(g_rtb.frame_done[i]); cudaEventDestroy(g_rtb.copy_done[i]);
    if (g_rtb.batch_stream[i]) cudaStreamDestroy(g_rtb.batch_stream[i]);
    cudaFree(g_rtb.d_frame[i]);
  }
  for (int i = 0; i < g_rtb.n_timers; i++) cudaEventDestroy(g_rtb.timers[i]);
  cudaFree(g_rtb.d_counters);
  cudaFree(g_rtb.cached_arena);
  cudaFree(g_rtb.d_star);
  cudaFree(g_rtb.d_peer_counter);
  if (g_rtb.h_stage) cudaFreeHost(g_rtb.h_stage);
  wf_free();
  (void)cudaGetLastError();
  g_rtb = RtbGlobal{};
}

int rtb_scene_create(const RtbSceneDesc* d, rtb_scene_t* out) {
  RTB_API_LOCK();
  if (!out) { rtb_set_error("out is NULL"); return RTB_ERR_INVALID; }
  *out = nullptr;
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  int rc = validate_scene(d);
  if (rc != RTB_OK) return rc;
  RtbScene_* s = new (std::nothrow) RtbScene_();
  if (!s) { rtb_set_error("host allocation failed"); return RTB_ERR_OOM; }
  struct Guard { RtbScene_* s; ~Guard() { if (s) rtb_scene_destroy(s); } } guard{s};
  const int n = d->n_spheres;
  s->n_spheres = n; s->n_textures = d->n_textures; s->n_lights = d->n_lights;
  s->spheres.assign(d->spheres, d->spheres + n);
  memcpy(s->lights, d->lights, sizeof s->lights);
  const int na = n > 0 ? n : 1;
  const int np = wf_pad(na);
  const int nt = d->n_textures > 0 ? d->n_textures : 1;
  s->n_pad = np;

  // One device arena for everything the scene owns, filled from one pinned staging buffer with one
  // H2D copy.  Both are cached in the library between scenes, so handing a scene over per frame costs
  // a memcpy + a DMA, not a dozen cudaMalloc/cudaFree calls.
  std::vector<DevTexture> tex(nt);
  size_t pool = 0;
  for (int t = 0; t < d->n_textures; t++) {
    tex[t].width = d->textures[t].width; tex[t].height = d->textures[t].height;
    tex[t].offset = (uint32_t)pool; tex[t]._pad = 0;
    pool += (size_t)tex[t].width * tex[t].height;
  }
  if (pool > 0xffffffffull) { rtb_set_error("texture pool too large"); return RTB_ERR_INVALID; }
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_g64 = take((size_t)na * 4 * sizeof(double));
  const size_t o_g32 = take((size_t)na * 4 * sizeof(float));
  const size_t o_m64 = take((size_t)na * sizeof(DevMaterial<double>));
  const size_t o_m32 = take((size_t)na * sizeof(DevMaterial<float>));
  const size_t o_tex = take((size_t)nt * sizeof(DevTexture));
  const size_t o_txl = take((pool > 0 ? pool : 1) * sizeof(uchar4));
  const size_t tb = (size_t)np * sizeof(float4);
  size_t o_ltab[RTB_MAX_LIGHTS], o_lx[RTB_MAX_LIGHTS];
  for (int k = 0; k < d->n_lights; k++) { o_ltab[k] = take(tb); o_lx[k] = take(tb); }
  const size_t o_gtab = take(tb);
  const size_t upload_bytes = off;                       // everything is written by this call (the per-camera
  const size_t total = off;                              // primary tables live in the workspace slots)

  if (g_rtb.cached_arena && g_rtb.cached_arena_bytes >= total) {
    s->arena = g_rtb.cached_arena; s->arena_bytes = g_rtb.cached_arena_bytes;
    g_rtb.cached_arena = nullptr; g_rtb.cached_arena_bytes = 0;
  } else {
    RTB_CUDA(cudaMalloc(&s->arena, total));
    s->arena_bytes = total;
  }
  const size_t stage_need = upload_bytes > 2 * tb ? upload_bytes : 2 * tb;
  if (g_rtb.stage_bytes < stage_need) {
    if (g_rtb.h_stage) { RTB_CUDA(cudaStreamSynchronize(g_rtb.stream)); cudaFreeHost(g_rtb.h_stage); }
    g_rtb.h_stage = nullptr; g_rtb.stage_bytes = 0;
    RTB_CUDA(cudaMallocHost(&g_rtb.h_stage, stage_need));
    g_rtb.stage_bytes = stage_need;
  }
  char* hs = g_rtb.h_stage;
  char* ds = s->arena;
  double* g64 = reinterpret_cast<double*>(hs + o_g64);
  float* g32 = reinterpret_cast<float*>(hs + o_g32);
  DevMaterial<double>* m64 = reinterpret_cast<DevMaterial<double>*>(hs + o_m64);
  DevMaterial<float>* m32 = reinterpret_cast<DevMaterial<float>*>(hs + o_m32);
  for (int i = 0; i < n; i++) {
    const RtbSphere& sp = d->spheres[i];
    for (int c = 0; c < 3; c++) { g64[4 * i + c] = sp.origin[c]; g32[4 * i + c] = (float)sp.origin[c]; }
    g64[4 * i + 3] = sp.r2; g32[4 * i + 3] = (float)sp.r2;
    fill_material(m64[i], sp.mtl);
    fill_material(m32[i], sp.mtl);
  }
  memcpy(hs + o_tex, tex.data(), (size_t)nt * sizeof(DevTexture));
  for (int t = 0; t < d->n_textures; t++)                // byte-exact copies of the caller's RGBA8 texels
    memcpy(hs + o_txl + (size_t)tex[t].offset * 4, d->textures[t].texels, (size_t)tex[t].width * tex[t].height * 4);
  // wavefront per-light tables (fp64 arithmetic, one rounding to fp32) and the general-ray table
  for (int k = 0; k < d->n_lights; k++)
    wf_make_table(d->spheres, n, np, d->lights[k], reinterpret_cast<float4*>(hs + o_ltab[k]),
                  reinterpret_cast<float4*>(hs + o_lx[k]));
  wf_make_geom_table(d->spheres, n, np, reinterpret_cast<float4*>(hs + o_gtab));
  RTB_CUDA(cudaMemcpyAsync(ds, hs, upload_bytes, cudaMemcpyHostToDevice, g_rtb.stream));
  RTB_CUDA(cudaStreamSynchronize(g_rtb.stream));         // staging is reused; the scene is usable on any stream
  s->h2d_bytes = upload_bytes;

  s->d_geom64 = reinterpret_cast<double*>(ds + o_g64);
  s->d_geom32 = reinterpret_cast<float*>(ds + o_g32);
  s->d_mtl64 = reinterpret_cast<DevMaterial<double>*>(ds + o_m64);
  s->d_mtl32 = reinterpret_cast<DevMaterial<float>*>(ds + o_m32);
  s->d_tex = reinterpret_cast<DevTexture*>(ds + o_tex);
  s->d_texels = reinterpret_cast<uchar4*>(ds + o_txl);
  for (int k = 0; k < d->n_lights; k++) {
    s->d_light_tab[k] = reinterpret_cast<float4*>(ds + o_ltab[k]);
    s->d_light_x[k] = reinterpret_cast<float4*>(ds + o_lx[k]);
  }
  s->d_geom_tab = reinterpret_cast<float4*>(ds + o_gtab);
  static uint64_t next_uid = 0;
  s->uid = ++next_uid;
  fill_scene_view<double>(s->sc64, s, d, s->d_geom64, s->d_mtl64);
  fill_scene_view<float>(s->sc32, s, d, s->d_geom32, s->d_mtl32);
  guard.s = nullptr;
  *out = s;
  return RTB_OK;
}

void rtb_scene_destroy(rtb_scene_t s) {
  RTB_API_LOCK();
  if (!s) return;
  for (int i = 0; i < RTB_MAX_IN_FLIGHT; i++)
    if (g_rtb.slot[i].pending_scene == s) (void)rtb_finish_pending(i);
  if (s->arena) {
    if (g_rtb.inited) {
      // renders are synchronous w.r.t. the host except rtb_render_device on a caller stream: drain
      cudaDeviceSynchronize();
      if (!g_rtb.cached_arena) {                          // keep one arena for the next scene
        g_rtb.cached_arena = s->arena; g_rtb.cached_arena_bytes = s->arena_bytes;
      } else if (g_rtb.cached_arena_bytes < s->arena_bytes) {
        cudaFree(g_rtb.cached_arena);
        g_rtb.cached_arena = s->arena; g_rtb.cached_arena_bytes = s->arena_bytes;
      } else {
        cudaFree(s->arena);
      }
    } else {
      cudaFree(s->arena);
    }
  }
  (void)cudaGetLastError();
  delete s;
}


uint64_t rtb_scene_h2d_bytes(rtb_scene_t s) { return s ? (uint64_t)s->h2d_bytes : 0; }

int rtb_copy_status_device(rtb_scene_t s, void* d_dst, void* cuda_stream) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!s || !d_dst) { rtb_set_error("NULL argument"); return RTB_ERR_INVALID; }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  const WfSlot& sl = g_rtb.slot[g_rtb.last_slot % RTB_MAX_IN_FLIGHT];   // the frame enqueued last
  if (sl.pending_scene == s && sl.wf.ctr) {
    // WfSticky.overflow is the first word after the per-band counter block
    const char* flag = reinterpret_cast<const char*>(sl.wf.ctr) + sizeof(WfCounters);
    RTB_CUDA(cudaMemcpyAsync(d_dst, flag, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  } else {
    RTB_CUDA(cudaMemsetAsync(d_dst, 0, sizeof(uint32_t), st));     // nothing in flight that could overflow
  }
  return RTB_OK;
}

int rtb_set_frames_in_flight(int n) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (n < 1 || n > RTB_MAX_IN_FLIGHT) {
    rtb_set_error("frames in flight must be 1..%d (got %d)", RTB_MAX_IN_FLIGHT, n);
    return RTB_ERR_INVALID;
  }
  int rc = rtb_finish_pending(-1);                         // statuses stay filed in their scenes
  if (rc != RTB_OK) return rc;
  g_rtb.frames_in_flight = n;
  g_rtb.next_slot = 0;
  g_rtb.half_slot = -1;
  return RTB_OK;
}

int rtb_finish(rtb_scene_t s) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!s) { rtb_set_error("NULL argument"); return RTB_ERR_INVALID; }
  int rc = rtb_finish_pending(-1);
  if (rc != RTB_OK) return rc;
  const int d = s->deferred_status;
  s->deferred_status = RTB_OK;
  if (d != RTB_OK)
    rtb_set_error("the asynchronous frame overflowed its ray queues and is invalid; the band size has been "
                  "reduced - issue the same render again");
  return d;
}

int rtb_render_device(rtb_scene_t s, const RtbCamera* cam, const RtbRenderOpts* o, void* d_out,
                      void* cuda_stream, RtbStats* stats) {
  RTB_API_LOCK();
  int rc = validate_render(s, cam, o, d_out);
  if (rc != RTB_OK) return rc;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  if (stats) { memset(stats, 0, sizeof *stats); RTB_CUDA(cudaEventRecord(g_rtb.ev0, st)); }
  rc = render_to_device(s, cam, o, (uint8_t*)d_out, st, stats, /*async=*/stats == nullptr);
  if (rc != RTB_OK) return rc;
  if (stats) {
    RTB_CUDA(cudaEventRecord(g_rtb.ev1, st));
    RTB_CUDA(cudaEventSynchronize(g_rtb.ev1));
    float ms = 0;
    RTB_CUDA(cudaEventElapsedTime(&ms, g_rtb.ev0, g_rtb.ev1));
    stats->kernel_ms = ms;
  }
  return RTB_OK;
}

int rtb_render_device_group(rtb_scene_t s, const RtbCamera* cams, int32_t nframes, const RtbRenderOpts* o,
                            void* const* d_outs, void* cuda_stream, RtbStats* stats) {
  RTB_API_LOCK();
  if (nframes < 1 || nframes > RTB_MAX_GROUP || !cams || !d_outs) {
    rtb_set_error("a frame group holds 1..%d frames (got %d)", RTB_MAX_GROUP, nframes);
    return RTB_ERR_INVALID;
  }
  int rc;
  uint8_t* outs[RTB_MAX_GROUP];
  for (int f = 0; f < nframes; f++) {
    if ((rc = validate_render(s, &cams[f], o, d_outs[f])) != RTB_OK) return rc;
    if (cams[f].width != cams[0].width || cams[f].height != cams[0].height) {
      rtb_set_error("all frames of a group have the same size");
      return RTB_ERR_UNSUPPORTED;
    }
    outs[f] = static_cast<uint8_t*>(d_outs[f]);
  }
  if (o->star_mode == RTB_STARS_STREAM && nframes > 1) {
    rtb_set_error("RTB_STARS_STREAM renders one frame per call");
    return RTB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  if (stats) { memset(stats, 0, sizeof *stats); RTB_CUDA(cudaEventRecord(g_rtb.ev0, st)); }
  rc = render_group_to_device(s, cams, nframes, o, outs, st, stats, /*async=*/stats == nullptr);
  if (rc != RTB_OK) return rc;
  if (stats) {
    RTB_CUDA(cudaEventRecord(g_rtb.ev1, st));
    RTB_CUDA(cudaEventSynchronize(g_rtb.ev1));
    float ms = 0;
    RTB_CUDA(cudaEventElapsedTime(&ms, g_rtb.ev0, g_rtb.ev1));
    stats->kernel_ms = ms;
  }
  return RTB_OK;
}

int rtb_render(rtb_scene_t s, const RtbCamera* cam, const RtbRenderOpts* o, uint8_t* out, RtbStats* stats) {
  RTB_API_LOCK();
  int rc = validate_render(s, cam, o, out);
  if (rc != RTB_OK) return rc;
  const size_t bytes = (size_t)cam->width * cam->height * 4;
  rc = ensure_frame(0, bytes);
  if (rc != RTB_OK) return rc;
  cudaStream_t st = g_rtb.stream;
  if (stats) memset(stats, 0, sizeof *stats);
  RTB_CUDA(cudaEventRecord(g_rtb.ev0, st));
  rc = render_to_device(s, cam, o, g_rtb.d_frame[0], st, stats);
  if (rc != RTB_OK) return rc;
  RTB_CUDA(cudaEventRecord(g_rtb.ev1, st));
  const RowMap rm = make_rowmap(cam, o);
  rc = copy_rows_to_host(rm, g_rtb.d_frame[0], out, cam->height, st);
  if (rc != RTB_OK) return rc;
  RTB_CUDA(cudaEventRecord(g_rtb.ev2, st));
  RTB_CUDA(cudaEventSynchronize(g_rtb.ev2));
  if (stats) {
    float ms = 0;
    RTB_CUDA(cudaEventElapsedTime(&ms, g_rtb.ev0, g_rtb.ev1));
    stats->kernel_ms = ms;
    RTB_CUDA(cudaEventElapsedTime(&ms, g_rtb.ev1, g_rtb.ev2));
    stats->d2h_ms = ms;
  }
  return RTB_OK;
}

int rtb_render_batch(rtb_scene_t s, const RtbCamera* cams, int32_t nframes, const RtbRenderOpts* o,
                     uint8_t* out, RtbStats* stats) {
  RTB_API_LOCK();
  if (nframes < 0 || (nframes > 0 && !cams)) { rtb_set_error("bad frame batch"); return RTB_ERR_INVALID; }
  if (stats) memset(stats, 0, sizeof *stats);
  if (nframes == 0) return RTB_OK;
  if (!o) { rtb_set_error("NULL argument"); return RTB_ERR_INVALID; }
  if (o->out_packed || o->stripe_world > 1 || o->y0 != 0 || (o->y1 > 0 && o->y1 != cams[0].height)) {
    rtb_set_error("rtb_render_batch renders whole frames (no stripes / row ranges)");
    return RTB_ERR_UNSUPPORTED;
  }
  int rc = validate_render(s, &cams[0], o, out);
  if (rc != RTB_OK) return rc;
  const size_t bytes = (size_t)cams[0].width * cams[0].height * 4;
  for (int f = 1; f < nframes; f++) {
    if ((rc = validate_render(s, &cams[f], o, out)) != RTB_OK) return rc;
    if (cams[f].width != cams[0].width || cams[f].height != cams[0].height) {
      rtb_set_error("rtb_render_batch: all frames of a batch have the same size");
      return RTB_ERR_UNSUPPORTED;
    }
  }
  if (stats) {
    // instrumented: one frame at a time on the compute stream (per-class event times), the D2H of frame
    // f-1 on the copy stream under the render of frame f
    RTB_CUDA(cudaEventRecord(g_rtb.ev0, g_rtb.stream));
    for (int f = 0; f < nframes; f++) {
      const int slot = f & 1;
      if ((rc = ensure_frame(slot, bytes)) != RTB_OK) return rc;
      if (f >= 2) RTB_CUDA(cudaStreamWaitEvent(g_rtb.stream, g_rtb.copy_done[slot], 0));
      RtbStats one;
      memset(&one, 0, sizeof one);
      if ((rc = render_to_device(s, &cams[f], o, g_rtb.d_frame[slot], g_rtb.stream, &one)) != RTB_OK) return rc;
      stats->nodes += one.nodes; stats->shadow_rays += one.shadow_rays; stats->sphere_tests += one.sphere_tests;
      stats->launches += one.launches; stats->bands += one.bands;
      stats->ms_trace_primary += one.ms_trace_primary; stats->ms_trace_secondary += one.ms_trace_secondary;
      stats->ms_shadow += one.ms_shadow; stats->ms_compose += one.ms_compose;
      RTB_CUDA(cudaEventRecord(g_rtb.frame_done[slot], g_rtb.stream));
      RTB_CUDA(cudaStreamWaitEvent(g_rtb.copy_stream, g_rtb.frame_done[slot], 0));
      RTB_CUDA(cudaMemcpyAsync(out + (size_t)f * bytes, g_rtb.d_frame[slot], bytes, cudaMemcpyDeviceToHost, g_rtb.copy_stream));
      RTB_CUDA(cudaEventRecord(g_rtb.copy_done[slot], g_rtb.copy_stream));
    }
    RTB_CUDA(cudaEventRecord(g_rtb.ev1, g_rtb.stream));
    RTB_CUDA(cudaStreamSynchronize(g_rtb.copy_stream));
    RTB_CUDA(cudaStreamSynchronize(g_rtb.stream));
    float ms = 0;
    RTB_CUDA(cudaEventElapsedTime(&ms, g_rtb.ev0, g_rtb.ev1));
    stats->kernel_ms = ms;
    return RTB_OK;
  }

  // Throughput path: D frames in flight, each on its own stream with its own workspace and staging
  // framebuffer; the D2H of a finished frame runs on the copy stream under the kernels of the next ones.
  // The queue-overflow verdict of frame f is collected just before its workspace is reused (frame f + D);
  // an overflow (the library has shrunk that frame's bands) restarts the batch at that frame.
  const int D = std::min(nframes, RTB_BATCH_IN_FLIGHT);
  const int saved_in_flight = g_rtb.frames_in_flight;
  if ((rc = rtb_finish_pending(-1)) != RTB_OK) return rc;
  g_rtb.frames_in_flight = D;
  g_rtb.next_slot = 0;
  g_rtb.half_slot = -1;
  struct Restore {
    int v;
    ~Restore() { (void)rtb_finish_pending(-1); g_rtb.frames_in_flight = v; g_rtb.next_slot = 0; g_rtb.half_slot = -1; }
  } restore{saved_in_flight};
  for (int i = 0; i < D; i++) {
    if (!g_rtb.batch_stream[i]) RTB_CUDA(cudaStreamCreateWithFlags(&g_rtb.batch_stream[i], cudaStreamNonBlocking));
    if ((rc = ensure_frame(i, bytes)) != RTB_OK) return rc;
  }
  struct InFlight { int frame, lib_slot; };
  std::vector<InFlight> q;                                // oldest first
  std::vector<char> buffer_used(D, 0);
  int f = 0, restarts = 0;
  while (f < nframes || !q.empty()) {
    if (f < nframes && (int)q.size() < D) {
      const int b = f % D;                                // staging framebuffer + stream of this frame
      cudaStream_t st = g_rtb.batch_stream[b];
      if (buffer_used[b]) RTB_CUDA(cudaStreamWaitEvent(st, g_rtb.copy_done[b], 0));   // its last D2H has drained
      const int lib_slot = g_rtb.next_slot % g_rtb.frames_in_flight;
      if ((rc = render_to_device(s, &cams[f], o, g_rtb.d_frame[b], st, nullptr, /*async=*/true)) != RTB_OK) return rc;
      RTB_CUDA(cudaEventRecord(g_rtb.frame_done[b], st));
      RTB_CUDA(cudaStreamWaitEvent(g_rtb.copy_stream, g_rtb.frame_done[b], 0));
      RTB_CUDA(cudaMemcpyAsync(out + (size_t)f * bytes, g_rtb.d_frame[b], bytes, cudaMemcpyDeviceToHost, g_rtb.copy_stream));
      RTB_CUDA(cudaEventRecord(g_rtb.copy_done[b], g_rtb.copy_stream));
      buffer_used[b] = 1;
      q.push_back({f, lib_slot});
      f++;
      continue;
    }
    const InFlight h = q.front();
    q.erase(q.begin());
    if ((rc = rtb_finish_pending(h.lib_slot)) != RTB_OK) return rc;
    if (s->deferred_status != RTB_OK) {                   // frame h.frame is invalid: redo it and what followed
      if ((rc = rtb_finish_pending(-1)) != RTB_OK) return rc;
      s->deferred_status = RTB_OK;
      if (++restarts > 64) {
        rtb_set_error("frame %d of the batch keeps overflowing its ray queues", h.frame);
        return RTB_ERR_OVERFLOW;
      }
      q.clear();
      f = h.frame;
    }
  }
  RTB_CUDA(cudaStreamSynchronize(g_rtb.copy_stream));
  return RTB_OK;
}

int32_t rtb_stripe_local_rows(int32_t height, int32_t stripe_rows, int32_t rank, int32_t world) {
  RowMap rm{0, height, stripe_rows, rank, world, 0, 0};
  return rm.n_local();
}

int32_t rtb_stripe_padded_rows(int32_t height, int32_t stripe_rows, int32_t world) {
  int32_t m = 0;
  for (int r = 0; r < (world > 1 ? world : 1); r++) {
    const int32_t v = rtb_stripe_local_rows(height, stripe_rows, r, world);
    if (v > m) m = v;
  }
  return m;
}

int rtb_unstripe_device(const void* d_gathered, void* d_frame, int32_t width, int32_t height,
                        int32_t stripe_rows, int32_t stripe_world, int32_t rank_stride_rows, void* cuda_stream) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!d_gathered || !d_frame || width <= 0 || height <= 0 || stripe_rows <= 0 || stripe_world < 1) {
    rtb_set_error("bad unstripe arguments");
    return RTB_ERR_INVALID;
  }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  int padded = rtb_stripe_padded_rows(height, stripe_rows, stripe_world);
  if (rank_stride_rows != 0) {
    if (rank_stride_rows < padded) { rtb_set_error("rank_stride_rows smaller than the padded stripe rows"); return RTB_ERR_INVALID; }
    padded = rank_stride_rows;
  }
  unstripe_kernel<<<g_rtb.sm_count * 8, 256, 0, st>>>((const uint32_t*)d_gathered, (uint32_t*)d_frame, width,
                                                     height, stripe_rows, stripe_world, padded);
  RTB_CUDA(cudaGetLastError());
  return RTB_OK;
}

int rtb_stripe_rows_to_host(const void* d_packed, void* h_frame, int32_t width, int32_t height, int32_t stripe_rows,
                            int32_t stripe_rank, int32_t stripe_world, void* cuda_stream) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!d_packed || !h_frame || width <= 0 || height <= 0 || stripe_world < 1 ||
      (stripe_world > 1 && (stripe_rows <= 0 || stripe_rank < 0 || stripe_rank >= stripe_world))) {
    rtb_set_error("bad stripe copy arguments");
    return RTB_ERR_INVALID;
  }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  const size_t row_bytes = (size_t)width * 4;
  const uint8_t* src = static_cast<const uint8_t*>(d_packed);
  uint8_t* dst = static_cast<uint8_t*>(h_frame);
  if (stripe_world <= 1) {
    RTB_CUDA(cudaMemcpyAsync(dst, src, row_bytes * (size_t)height, cudaMemcpyDeviceToHost, st));
    return RTB_OK;
  }
  // stripe i of this rank = rows [(i * world + rank) * S, +S): full stripes as ONE strided 2-D copy (packed
  // source pitch S rows, destination pitch S * world rows), a ragged last stripe on its own
  const int S = stripe_rows, period = S * stripe_world;
  const int first = stripe_rank * S;
  if (first >= height) return RTB_OK;
  const int full = (height - first) / period + (((height - first) % period) >= S ? 1 : 0);   // complete stripes
  if (full > 0)
    RTB_CUDA(cudaMemcpy2DAsync(dst + row_bytes * first, row_bytes * period, src, row_bytes * S, row_bytes * S,
                               (size_t)full, cudaMemcpyDeviceToHost, st));
  const int y_tail = first + full * period;                  // a stripe cut short by the frame's last row
  if (y_tail < height) {
    const int rows = std::min(S, height - y_tail);
    RTB_CUDA(cudaMemcpyAsync(dst + row_bytes * y_tail, src + row_bytes * (size_t)full * S, row_bytes * rows,
                             cudaMemcpyDeviceToHost, st));
  }
  return RTB_OK;
}

int rtb_host_register(void* p, uint64_t bytes) {
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!p || bytes == 0) { rtb_set_error("NULL argument"); return RTB_ERR_INVALID; }
  RTB_CUDA(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable));
  return RTB_OK;
}

int rtb_host_unregister(void* p) {
  if (!p) return RTB_OK;
  RTB_CUDA(cudaHostUnregister(p));
  return RTB_OK;
}

void rtb_default_opts(RtbRenderOpts* o) {
  if (!o) return;
  memset(o, 0, sizeof *o);
  o->segs = 8;                 // main.js:194
  o->spp_grid = 1;
  o->precision = RTB_PRECISION_FP32;
  o->kernel = RTB_KERNEL_AUTO;
  o->star_mode = RTB_STARS_OFF;
}

const char* rtb_last_error(void) { return t_err; }
const char* rtb_version(void) { return "rtb200 abi 1 sm_100a"; }
uint32_t rtb_abi_version(void) { return RTB_ABI_VERSION; }

int rtb_clock_probe(uint64_t* d_out, int32_t max_sms, void* cuda_stream) {
  RTB_API_LOCK();
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (!d_out || max_sms < 1) { rtb_set_error("bad clock probe arguments"); return RTB_ERR_INVALID; }
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g_rtb.stream;
  clock_probe_kernel<<<g_rtb.sm_count * 4, 32, 0, st>>>(reinterpret_cast<unsigned long long*>(d_out), max_sms);
  RTB_CUDA(cudaGetLastError());
  return RTB_OK;
}

int rtb_device_info(int32_t* sm_count, int32_t* clock_khz, char* name_buf, int32_t name_len) {
  if (!g_rtb.inited) { rtb_set_error("rtb_init() has not been called"); return RTB_ERR_NOT_INIT; }
  if (sm_count) *sm_count = g_rtb.sm_count;
  if (clock_khz) *clock_khz = g_rtb.clock_khz;
  if (name_buf && name_len > 0) snprintf(name_buf, (size_t)name_len, "%s", g_rtb.name);
  return RTB_OK;
}

}  // extern "C"
#pragma GCC visibility pop
