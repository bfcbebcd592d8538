// host_prep.cu — the host-side inputs that fix the result, for hosts that bind the C ABI
// directly and would otherwise restate them: main.js's closure-private lookAt + projection constants
// (main.js:92-105) and the object sort (main.js:159-163).  Plain host code in float64, one rounding per JS
// operator in source order; no device work, no rtb_init needed.
#include <algorithm>
#include <cmath>
#include <vector>
#include "rtb_internal.h"

namespace {

inline void between3(const double a[3], const double b[3], double o[3]) {   // main.js:68-70: b - a
  o[0] = b[0] - a[0]; o[1] = b[1] - a[1]; o[2] = b[2] - a[2];
}
inline void cross3(const double a[3], const double b[3], double o[3]) {     // main.js:53-55
  o[0] = a[1] * b[2] - b[1] * a[2];
  o[1] = a[2] * b[0] - b[2] * a[0];
  o[2] = a[0] * b[1] - b[0] * a[1];
}
inline double length3(const double v[3]) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }   // 45-47, 57-60
inline void normal3(double v[3]) {                                           // main.js:62-66: v * (1 / l), v itself if l == 0
  const double l = length3(v);
  if (l == 0) return;
  const double q = 1 / l;
  v[0] *= q; v[1] *= q; v[2] *= q;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int rtb_look_at(const double org[3], const double tgt[3], const double up[3], int32_t width, int32_t height,
                int32_t camera_mode, RtbCamera* out) {
  if (!org || !tgt || !up || !out || width <= 0 || height <= 0) { rtb_set_error("bad look_at arguments"); return RTB_ERR_INVALID; }
  if (camera_mode != RTB_CAMERA_REFERENCE && camera_mode != RTB_CAMERA_BASIS) {
    rtb_set_error("unknown camera_mode %d", camera_mode);
    return RTB_ERR_INVALID;
  }
  RtbCamera c;
  c.width = width; c.height = height; c.camera_mode = camera_mode; c._pad = 0;
  for (int k = 0; k < 3; k++) c.origin[k] = org[k];                          // main.js:93
  between3(org, tgt, c.axisZ);                                               // 94
  cross3(up, c.axisZ, c.axisX);                                              // 95
  cross3(c.axisZ, c.axisX, c.axisY);                                         // 96
  normal3(c.axisX); normal3(c.axisY); normal3(c.axisZ);                      // 97-99
  const double projA = 60 * 3.141592653589793 / 180;                         // 102
  c.projW = (double)width / 2;                                               // 103
  c.projH = (double)height / 2;                                              // 104
  c.projD = c.projW / std::tan(projA / 2);                                   // 105
  *out = c;
  return RTB_OK;
}

int rtb_sort_spheres(RtbSphere* spheres, int32_t n, const double camera_origin[3]) {
  if (n < 0 || (n > 0 && !spheres) || !camera_origin) { rtb_set_error("bad sort arguments"); return RTB_ERR_INVALID; }
  // key = surface_area / distance3D(origin, centre), surface_area = 4 * Math.PI * r2 (main.js:160-161, 416);
  // ascending, ties keep their order (Array.prototype.sort is stable since V8 7.0)
  std::vector<std::pair<double, int>> key((size_t)n);
  for (int i = 0; i < n; i++) {
    double v[3];
    between3(camera_origin, spheres[i].origin, v);
    key[i] = {4 * 3.141592653589793 * spheres[i].r2 / length3(v), i};
  }
  std::stable_sort(key.begin(), key.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) {
    return a.first - b.first < 0;                                            // comparator returns aa - bb (162)
  });
  std::vector<RtbSphere> sorted((size_t)n);
  for (int i = 0; i < n; i++) sorted[i] = spheres[key[i].second];
  std::copy(sorted.begin(), sorted.end(), spheres);
  return RTB_OK;
}

// Diagnostics (host only): the packed sweep tables the wavefront kernels stage in shared memory, as built by
// wf_make_table / wf_make_geom_table for `origin` - so that their safety margins ("fp32 rounding can only ADD
// candidates, never lose a true hit") can be checked on the CPU against float64 for arbitrary scenes
// (tests/test_sweep_tables.py).  Layout per sphere PAIR p: tab[8p..8p+7] = x0 x1 y0 y1 z0 z1 w0 w1.
// Buffers hold rtb_debug_table_pad(n_spheres) * 4 floats each; exact = (Lx, Ly, Lz, |L|^2 - r2) per sphere.
int32_t rtb_debug_table_pad(int32_t n_spheres) { return wf_pad(n_spheres > 0 ? n_spheres : 1); }

int rtb_debug_sweep_tables(const RtbSceneDesc* d, const double origin[3], float* origin_tab, float* origin_exact,
                           float* general_tab) {
  if (!d || !origin || !origin_tab || !origin_exact || !general_tab || d->n_spheres < 0 ||
      (d->n_spheres > 0 && !d->spheres)) {
    rtb_set_error("bad table arguments");
    return RTB_ERR_INVALID;
  }
  const int n = d->n_spheres, np = wf_pad(n > 0 ? n : 1);
  wf_make_table(d->spheres, n, np, origin, reinterpret_cast<float4*>(origin_tab), reinterpret_cast<float4*>(origin_exact));
  wf_make_geom_table(d->spheres, n, np, reinterpret_cast<float4*>(general_tab));
  return RTB_OK;
}

}  // extern "C"
#pragma GCC visibility pop
