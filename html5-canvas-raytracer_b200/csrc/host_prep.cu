// host_prep.cu — the host-side inputs that fix the result, for hosts that bind the C ABI
// directly and would otherwise restate them: main.js's closure-private lookAt + projection constants
// (main.js:92-105) and the object sort (main.js:159-163).  Plain host code in float64, one rounding per JS
// operator in source order; no device work, no rtb_init needed.
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
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

// zlib's one-shot inflate, bound at run time (dlopen) so that the library keeps no link-time dependency on it
typedef int (*UncompressFn)(unsigned char* dest, unsigned long* dest_len, const unsigned char* src, unsigned long src_len);
UncompressFn zlib_uncompress() {
  static bool looked = false;
  static UncompressFn fn = nullptr;
  if (!looked) {
    looked = true;
    void* h = dlopen("libz.so.1", RTLD_NOW);
    if (!h) h = dlopen("libz.so", RTLD_NOW);
    if (h) fn = (UncompressFn)dlsym(h, "uncompress");
  }
  return fn;
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

// loadTexture without the browser (main.js:375-395: Image -> offscreen canvas -> getImageData): PNG bytes ->
// the RGBA8 texels `getImageData` returns.  Non-interlaced 8-bit greyscale / RGB / palette (+ tRNS) /
// grey+alpha / RGBA - the reference's earth.png and mars.png are 8-bit palette images with PLTE only, so the
// palette colours come back unchanged with alpha 255 (no gamma / ICC chunk, no colour management).
int rtb_decode_png(const uint8_t* png, uint64_t len, int32_t* width, int32_t* height, uint8_t* out_rgba8,
                   uint64_t out_capacity) try {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  if (!png || !width || !height || len < 8 || memcmp(png, sig, 8) != 0) { rtb_set_error("not a PNG"); return RTB_ERR_INVALID; }
  const uint8_t *plte = nullptr, *trns = nullptr;
  uint32_t n_plte = 0, n_trns = 0, w = 0, h = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat;
  bool have_ihdr = false;
  for (uint64_t pos = 8; pos + 12 <= len;) {
    const uint32_t n = be32(png + pos);
    const uint8_t* typ = png + pos + 4;
    const uint8_t* body = png + pos + 8;
    if ((uint64_t)n + 12 > len - pos) { rtb_set_error("truncated PNG chunk"); return RTB_ERR_INVALID; }
    pos += 12 + (uint64_t)n;
    if (!memcmp(typ, "IHDR", 4)) {
      if (n < 13) { rtb_set_error("bad IHDR"); return RTB_ERR_INVALID; }
      w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
      have_ihdr = true;
    } else if (!memcmp(typ, "PLTE", 4)) { plte = body; n_plte = n / 3; }
    else if (!memcmp(typ, "tRNS", 4)) { trns = body; n_trns = n; }
    else if (!memcmp(typ, "IDAT", 4)) idat.insert(idat.end(), body, body + n);
    else if (!memcmp(typ, "IEND", 4)) break;
  }
  if (!have_ihdr || w == 0 || h == 0 || w > 32768 || h > 32768) { rtb_set_error("PNG without a usable IHDR"); return RTB_ERR_INVALID; }
  if (depth != 8 || interlace != 0) { rtb_set_error("only 8-bit non-interlaced PNGs are supported"); return RTB_ERR_UNSUPPORTED; }
  int ch;
  switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 3: ch = 1; break; case 4: ch = 2; break; case 6: ch = 4; break;
    default: rtb_set_error("bad PNG colour type %d", ctype); return RTB_ERR_INVALID; }
  if (ctype == 3 && !plte) { rtb_set_error("palette PNG without PLTE"); return RTB_ERR_INVALID; }
  *width = (int32_t)w; *height = (int32_t)h;
  if (!out_rgba8) return RTB_OK;                           // size query
  if (out_capacity < (uint64_t)w * h * 4) { rtb_set_error("output buffer smaller than width*height*4"); return RTB_ERR_INVALID; }
  UncompressFn inflate = zlib_uncompress();
  if (!inflate) { rtb_set_error("libz.so.1 not found: PNG textures must be decoded by the host"); return RTB_ERR_UNSUPPORTED; }
  const size_t stride = (size_t)w * ch;
  std::vector<uint8_t> raw((stride + 1) * h);
  unsigned long raw_len = (unsigned long)raw.size();
  if (inflate(raw.data(), &raw_len, idat.data(), (unsigned long)idat.size()) != 0 || raw_len != raw.size()) {
    rtb_set_error("PNG image data does not inflate to height * (1 + width * channels) bytes");
    return RTB_ERR_INVALID;
  }
  // undo the scanline filters in place (PNG spec 9.2): None, Sub, Up, Average, Paeth
  std::vector<uint8_t> zero(stride, 0);
  for (uint32_t y = 0; y < h; y++) {
    uint8_t* cur = raw.data() + (size_t)y * (stride + 1) + 1;
    const uint8_t* prev = y ? cur - (stride + 1) : zero.data();
    const int ft = cur[-1];
    if (ft < 0 || ft > 4) { rtb_set_error("bad PNG filter type %d", ft); return RTB_ERR_INVALID; }
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0;
      int p = 0;
      if (ft == 1) p = a;
      else if (ft == 2) p = b;
      else if (ft == 3) p = (a + b) >> 1;
      else if (ft == 4) {
        const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c);
        p = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
      }
      cur[i] = (uint8_t)(cur[i] + p);
    }
  }
  for (uint32_t y = 0; y < h; y++) {
    const uint8_t* px = raw.data() + (size_t)y * (stride + 1) + 1;
    uint8_t* o = out_rgba8 + (size_t)y * w * 4;
    for (uint32_t x = 0; x < w; x++, o += 4) {
      switch (ctype) {
        case 0: o[0] = o[1] = o[2] = px[x]; o[3] = 255; break;
        case 2: o[0] = px[3 * x]; o[1] = px[3 * x + 1]; o[2] = px[3 * x + 2]; o[3] = 255; break;
        case 3: {
          const uint32_t i = px[x];
          if (i >= n_plte) { rtb_set_error("PNG palette index %u outside PLTE", i); return RTB_ERR_INVALID; }
          o[0] = plte[3 * i]; o[1] = plte[3 * i + 1]; o[2] = plte[3 * i + 2]; o[3] = i < n_trns ? trns[i] : 255;
          break;
        }
        case 4: o[0] = o[1] = o[2] = px[2 * x]; o[3] = px[2 * x + 1]; break;
        default: memcpy(o, px + 4 * x, 4);
      }
    }
  }
  return RTB_OK;
} catch (const std::bad_alloc&) {
  rtb_set_error("PNG decoder: host allocation failed");
  return RTB_ERR_OOM;
} catch (...) {
  rtb_set_error("PNG decoder: unexpected C++ exception");
  return RTB_ERR_INVALID;
}

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

int rtb_sort_spheres(RtbSphere* spheres, int32_t n, const double camera_origin[3]) try {
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
} catch (const std::bad_alloc&) {
  rtb_set_error("sort: host allocation failed");
  return RTB_ERR_OOM;
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
