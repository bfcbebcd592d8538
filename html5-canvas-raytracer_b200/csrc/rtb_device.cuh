// rtb_device.cuh — device-side scene views and scalar helpers shared by the per-pixel kernels
// (pixel_f64.cu / pixel_f32.cu) and the fp32 wavefront pipeline (wavefront.cu).
//
// Everything is templated on the scalar type T (double = validation kernel following main.js
// operation for operation; float = throughput kernels).  The double instantiation is compiled with
// -fmad=false so that no multiply-add is contracted (JS rounds after every operator).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rtb200.h"

#define RTB_DEVINL __device__ __forceinline__

// Material as the kernels read it (createMaterial main.js:397-406 + sampler kind).
template <typename T>
struct DevMaterial {
  T color[3];
  T albedo[5];          // ambient, diffuse, specular, reflect, refract
  T specular_exponent;
  T refract_index;
  T checker_freq[2];
  T checker_color[2][3];
  T star_threshold, star_gain;
  int32_t sampler_kind;
  int32_t texture_unit;
};

struct DevTexture {     // createTexture main.js:339-341; texels live in one shared uchar4 pool
  int32_t width, height;
  uint32_t offset;      // first texel in the pool
  uint32_t _pad;
};

// Scene as uploaded by rtb_scene_create.  One copy per precision (geometry + materials).
template <typename T>
struct DevSceneT {
  int32_t n_spheres;
  int32_t n_textures;
  int32_t n_lights;
  int32_t _pad;
  const T* geom;                  // n_spheres x 4: Cx, Cy, Cz, r2   (createSphere main.js:408-418)
  const DevMaterial<T>* mtl;      // n_spheres
  const DevTexture* tex;          // n_textures
  const uchar4* texels;           // RGBA8 pool
  T lights[RTB_MAX_LIGHTS][3];    // main.js:283
  T light_intensity;              // main.js:284
  T eps;                          // main.js:430-445
};

// Per-call camera/projection (closure variables of main(), main.js:85-105) + option scalars.
template <typename T>
struct DevFrameT {
  int32_t width, height;
  int32_t segs, spp_grid;
  int32_t star_mode, camera_mode;
  int32_t y0, y1;                 // rows [y0,y1)
  int32_t stripe_rows, stripe_rank, stripe_world, out_packed;
  uint64_t star_seed;
  // RTB_STARS_STREAM (pixel kernels only, whole frames): Math.random() values in the reference's call
  // order.  Pass 1 counts the star-sampler calls of every pixel, an exclusive scan turns the counts into
  // each pixel's first stream index, pass 2 renders.
  const double* star_stream;              // device copy of the supplied stream
  unsigned long long star_stream_len;
  const unsigned long long* star_base;    // per pixel (y*W + x): index of its first draw      (pass 2)
  unsigned long long* star_counts;        // per pixel: number of draws                        (pass 1)
  T origin[3], axisX[3], axisY[3], axisZ[3];
  T projW, projH, projD;
};

// ---- row bookkeeping shared by all kernels (striping / packing) -----------------------
// Rows this call renders are the rows y in [y0,y1) with (y / stripe_rows) % world == rank (all rows
// when world <= 1), enumerated in ascending y as "local rows" 0..n-1.
struct RowMap {
  int32_t y0, y1, S, rank, world, packed, width;
  __host__ __device__ int32_t owned_before(int32_t y) const {
    // number of owned rows with index < y (y >= 0), ignoring y0
    if (world <= 1 || S <= 0) return y;
    const int32_t period = S * world;
    const int32_t full = y / period, rem = y % period;
    int32_t r = rem - rank * S;
    r = r < 0 ? 0 : (r > S ? S : r);
    return full * S + r;
  }
  __host__ __device__ int32_t n_local() const { return owned_before(y1) - owned_before(y0); }
  __host__ __device__ int32_t local_to_y(int32_t l) const {
    if (world <= 1 || S <= 0) return y0 + l;
    const int32_t g = l + owned_before(y0);       // index among all owned rows from row 0
    return (g / S) * (S * world) + rank * S + (g % S);
  }
  __host__ __device__ size_t out_row(int32_t l) const {   // row index inside the output buffer
    return packed ? (size_t)l : (size_t)local_to_y(l);
  }
};

// ---- JS number semantics -------------------------------------------------------------------------------
template <typename T> RTB_DEVINL T js_max(T a, T b) {   // Math.max: NaN-propagating
  if (a != a || b != b) return a != a ? a : b;
  return a > b ? a : b;
}
template <typename T> RTB_DEVINL T js_min(T a, T b) {
  if (a != a || b != b) return a != a ? a : b;
  return a < b ? a : b;
}
RTB_DEVINL int32_t js_toint32(double d) {                // `&` operand conversion
  if (!isfinite(d)) return 0;
  double t = trunc(d);
  double m = fmod(t, 4294967296.0);
  if (m < 0) m += 4294967296.0;
  return (int32_t)(uint32_t)m;
}
RTB_DEVINL int32_t js_toint32(float f) {
  if (!isfinite(f)) return 0;
  if (fabsf(f) < 2147483520.f) return (int32_t)f;      // truncation toward zero
  return js_toint32((double)f);
}
// Uint8ClampedArray store (main.js:195-198): NaN -> 0, clamp, round half to even.
RTB_DEVINL uint32_t clamped_u8(double v) {
  if (!(v > 0)) return 0u;
  if (v >= 255) return 255u;
  return (uint32_t)__double2int_rn(v);
}
RTB_DEVINL uint32_t clamped_u8(float v) {
  if (!(v > 0)) return 0u;
  if (v >= 255.f) return 255u;
  return (uint32_t)__float2int_rn(v);
}

template <typename T> struct Num;
template <> struct Num<double> {
  static RTB_DEVINL double sqrt(double x) { return ::sqrt(x); }
  static RTB_DEVINL double atan2(double y, double x) { return ::atan2(y, x); }
  static RTB_DEVINL double asin(double x) { return ::asin(x); }
  static RTB_DEVINL double pow(double x, double y) { return ::pow(x, y); }
  static RTB_DEVINL double ceil(double x) { return ::ceil(x); }
  static RTB_DEVINL double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
  static RTB_DEVINL double pi() { return 3.141592653589793; }
};
template <> struct Num<float> {
  static RTB_DEVINL float sqrt(float x) { return ::sqrtf(x); }
  static RTB_DEVINL float atan2(float y, float x) { return ::atan2f(y, x); }
  static RTB_DEVINL float asin(float x) { return ::asinf(x); }
  static RTB_DEVINL float pow(float x, float y) { return ::powf(x, y); }
  static RTB_DEVINL float ceil(float x) { return ::ceilf(x); }
  static RTB_DEVINL float inf() { return __int_as_float(0x7f800000); }
  static RTB_DEVINL float pi() { return 3.14159265358979323846f; }
};

// splitmix64 finaliser: RTB_STARS_HASH stand-in for Math.random (same function as the oracle's).
RTB_DEVINL uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
RTB_DEVINL double star_uniform(uint64_t seed, uint32_t px, uint32_t py, uint32_t sub, uint32_t node) {
  uint64_t k = mix64(seed ^ mix64(((uint64_t)py << 32) | px) ^ mix64(((uint64_t)sub << 32) | node) * 3);
  return (double)(k >> 11) * (1.0 / 9007199254740992.0);
}

// sampleTexture main.js:343-351 — nearest texel with ceil(u*w)-1 addressing (NOT floor), /255.
template <typename T>
RTB_DEVINL void sample_texture(const DevTexture& tx, const uchar4* __restrict__ pool, T u, T v, T rgb[3]) {
  const T x = js_max<T>(T(0), Num<T>::ceil(u * T(tx.width)) - T(1));
  const T y = js_max<T>(T(0), Num<T>::ceil(v * T(tx.height)) - T(1));
  if (!(x >= T(0) && y >= T(0) && x < T(tx.width) && y < T(tx.height))) {   // undefined/255 = NaN
    const T nan = Num<T>::inf() - Num<T>::inf();
    rgb[0] = rgb[1] = rgb[2] = nan;
    return;
  }
  const uchar4 c = __ldg(pool + tx.offset + (uint32_t)y * (uint32_t)tx.width + (uint32_t)x);
  rgb[0] = T(c.x) / T(255);
  rgb[1] = T(c.y) / T(255);
  rgb[2] = T(c.z) / T(255);
}

// mtl.sampler(hit) main.js:320 — dispatch on the enumerated kind.  n = hit.n, (u,v) = hit.u/v.
template <typename T>
RTB_DEVINL void run_sampler(const DevSceneT<T>& sc, const DevMaterial<T>& m, const T n[3], T u, T v,
                            int star_mode, uint64_t star_seed, uint32_t px, uint32_t py, uint32_t sub,
                            uint32_t node_id, T rgb[3], const double* star_stream = nullptr,
                            unsigned long long star_idx = 0, unsigned long long star_len = 0) {
  switch (m.sampler_kind) {
    default:
    case RTB_SAMPLER_COLOR:                                   // main.js:404
      rgb[0] = m.color[0]; rgb[1] = m.color[1]; rgb[2] = m.color[2];
      return;
    case RTB_SAMPLER_TEXTURE:                                 // main.js:143-157
      sample_texture<T>(sc.tex[m.texture_unit], sc.texels, u, v, rgb);
      return;
    case RTB_SAMPLER_CHECKER: {                               // main.js:126-133 (axes differ from 446-447)
      T cu, cv;
      if (sizeof(T) == sizeof(float)) {                       // throughput kernels: reciprocal constants
        cu = Num<T>::atan2(-n[1], -n[0]) * T(0.15915494309189535) + T(0.5);
        cv = Num<T>::asin(-n[2]) * T(0.31830988618379069) + T(0.5);
      } else {                                                // validation kernel: the reference's divisions
        cu = Num<T>::atan2(-n[1], -n[0]) / Num<T>::pi() / T(2) + T(0.5);
        cv = Num<T>::asin(-n[2]) / (Num<T>::pi() / T(2)) / T(2) + T(0.5);
      }
      const int32_t s = js_toint32(cu * m.checker_freq[0]) & 1;
      const int32_t t = js_toint32(cv * m.checker_freq[1]) & 1;
      const T* c = m.checker_color[s ^ t];
      rgb[0] = c[0]; rgb[1] = c[1]; rgb[2] = c[2];
      return;
    }
    case RTB_SAMPLER_STARS: {                                 // main.js:135-139
      T c = T(0.5);                                           // RTB_STARS_OFF: Math.random = () => 0.5
      if (star_mode == RTB_STARS_HASH) c = T(star_uniform(star_seed, px, py, sub, node_id));
      else if (star_mode == RTB_STARS_STREAM && star_stream && star_idx < star_len) c = T(star_stream[star_idx]);
      if (c >= m.star_threshold) { rgb[0] = rgb[1] = rgb[2] = T(0); return; }
      c *= m.star_gain;
      rgb[0] = rgb[1] = rgb[2] = c;
      return;
    }
  }
}
