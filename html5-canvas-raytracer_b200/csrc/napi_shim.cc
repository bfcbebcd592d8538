// napi_shim.cc — thin Node-API addon over the C ABI of include/rtb200.h.
//
//   const rtb = require('./rtb200.node');
//   rtb.render(scene, camera, opts, colorbuf.data);      // fills the Uint8ClampedArray in place
//   await rtb.renderAsync(scene, camera, opts, colorbuf.data);   // same on a worker thread, returns a Promise
//   const tex = rtb.decodeTexture(fs.readFileSync('earth.png'));  // loadTexture without a canvas (main.js:375-395)
//
// It replaces the pixel loops of main.js's redraw()/spanish() (main.js:180-199): JavaScript hands over
// the same scene description main() builds (main.js:85-163) and gets back the bytes it would have
// passed to putImageData (main.js:200).  See INTEGRATION.md for the patch to main.js.
//
// HEADER-FREE: this image has neither Node nor node_api.h, so the handful of Node-API prototypes used
// here are declared below from the ABI-stable Node-API (C linkage, napi_status return, napi_ok == 0).
// The napi_* symbols stay undefined in the .node file and are bound by the `node` executable when it
// dlopen()s the addon.  NOT EXECUTED in this container (no Node); verify against a real node_api.h on
// first contact with a Node install.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>
#include "rtb200.h"

extern "C" {
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0 } napi_status;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
  napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array,
  napi_biguint64_array
} napi_typedarray_type;
typedef enum {
  napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object,
  napi_function, napi_external, napi_bigint
} napi_valuetype;

napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_named_property(napi_env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_has_named_property(napi_env, napi_value object, const char* utf8name, bool* result);
napi_status napi_get_array_length(napi_env, napi_value value, uint32_t* result);
napi_status napi_get_element(napi_env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_get_value_double(napi_env, napi_value value, double* result);
napi_status napi_get_value_int32(napi_env, napi_value value, int32_t* result);
napi_status napi_typeof(napi_env, napi_value value, napi_valuetype* result);
napi_status napi_is_typedarray(napi_env, napi_value value, bool* result);
napi_status napi_is_array(napi_env, napi_value value, bool* result);
napi_status napi_get_typedarray_info(napi_env, napi_value typedarray, napi_typedarray_type* type, size_t* length,
                                     void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_function(napi_env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_set_named_property(napi_env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_create_object(napi_env, napi_value* result);
napi_status napi_create_double(napi_env, double value, napi_value* result);
napi_status napi_create_string_utf8(napi_env, const char* str, size_t length, napi_value* result);
napi_status napi_get_undefined(napi_env, napi_value* result);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
// promise-returning render (renderAsync): libuv worker + deferred
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_async_work__* napi_async_work;
typedef struct napi_ref__* napi_ref;
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);
napi_status napi_create_promise(napi_env, napi_deferred* deferred, napi_value* promise);
napi_status napi_resolve_deferred(napi_env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_error(napi_env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_create_async_work(napi_env, napi_value async_resource, napi_value async_resource_name,
                                   napi_async_execute_callback execute, napi_async_complete_callback complete,
                                   void* data, napi_async_work* result);
napi_status napi_queue_async_work(napi_env, napi_async_work work);
napi_status napi_delete_async_work(napi_env, napi_async_work work);
napi_status napi_create_reference(napi_env, napi_value value, uint32_t initial_refcount, napi_ref* result);
// decodeTexture: a fresh Uint8ClampedArray for the texels
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type type, size_t length, napi_value arraybuffer,
                                   size_t byte_offset, napi_value* result);
napi_status napi_delete_reference(napi_env, napi_ref ref);
}
#define NAPI_AUTO_LENGTH SIZE_MAX

namespace {

struct Fail { std::string msg; };
#define NAPI_OK(call) do { if ((call) != napi_ok) throw Fail{std::string("Node-API call failed: ") + #call}; } while (0)

bool has(napi_env env, napi_value o, const char* k) {
  bool r = false;
  if (napi_has_named_property(env, o, k, &r) != napi_ok) return false;
  if (!r) return false;
  napi_value v; napi_valuetype t;
  NAPI_OK(napi_get_named_property(env, o, k, &v));
  NAPI_OK(napi_typeof(env, v, &t));
  return t != napi_undefined && t != napi_null;
}
napi_value prop(napi_env env, napi_value o, const char* k) {
  napi_value v;
  NAPI_OK(napi_get_named_property(env, o, k, &v));
  return v;
}
double num(napi_env env, napi_value v) {
  double d = 0;
  if (napi_get_value_double(env, v, &d) != napi_ok) throw Fail{"expected a number"};
  return d;
}
double num_prop(napi_env env, napi_value o, const char* k, double dflt, bool required = false) {
  if (!has(env, o, k)) {
    if (required) throw Fail{std::string("missing property '") + k + "'"};
    return dflt;
  }
  return num(env, prop(env, o, k));
}
void vec(napi_env env, napi_value o, const char* k, double* out, uint32_t n) {
  if (!has(env, o, k)) throw Fail{std::string("missing property '") + k + "'"};
  napi_value a = prop(env, o, k);
  uint32_t len = 0;
  NAPI_OK(napi_get_array_length(env, a, &len));
  if (len < n) throw Fail{std::string("property '") + k + "' is too short"};
  for (uint32_t i = 0; i < n; i++) {
    napi_value e;
    NAPI_OK(napi_get_element(env, a, i, &e));
    out[i] = num(env, e);
  }
}

// texels: Uint8ClampedArray / Uint8Array (loadTexture, main.js:391) or a plain Array of numbers
// (checkerTexture writes 255*colour into `texels: []`, main.js:340, 358-366).
void read_texels(napi_env env, napi_value t, size_t want, std::vector<uint8_t>& out) {
  bool typed = false;
  NAPI_OK(napi_is_typedarray(env, t, &typed));
  out.resize(want);
  if (typed) {
    napi_typedarray_type ty; size_t len = 0; void* data = nullptr; napi_value ab; size_t off = 0;
    NAPI_OK(napi_get_typedarray_info(env, t, &ty, &len, &data, &ab, &off));
    if ((ty != napi_uint8_clamped_array && ty != napi_uint8_array) || len < want) throw Fail{"texels: need >= width*height*4 bytes"};
    memcpy(out.data(), data, want);
    return;
  }
  uint32_t len = 0;
  NAPI_OK(napi_get_array_length(env, t, &len));
  if (len < want) throw Fail{"texels: need >= width*height*4 entries"};
  for (size_t i = 0; i < want; i++) {
    napi_value e;
    NAPI_OK(napi_get_element(env, t, (uint32_t)i, &e));
    const double v = num(env, e);
    out[i] = (uint8_t)(!(v > 0) ? 0 : (v >= 255 ? 255 : std::nearbyint(v)));
  }
}

// render(scene, camera, opts, out): scene = { objects: [{origin, r2, mtl:{color, albedo, specular_exponent,
// refract_index, sampler_kind?, texture_unit?, checker_freq?, checker_color?, star_threshold?, star_gain?}}],
// textures: [{width, height, texels}], lights?, light_intensity?, eps? }
//
// A call is three steps: marshal() walks the JS objects (main thread only), run() talks to the C ABI and
// touches no JS value (any thread), result() builds the stats object (main thread).  render() does them in a
// row; renderAsync() runs the middle one on a libuv worker and returns a Promise, so that the page keeps its
// event loop - the reason the reference renders one scanline per setTimeout tick (main.js:182, 201).
struct Job {
  std::vector<RtbSphere> spheres;
  std::vector<RtbTexture> textures;
  std::vector<std::vector<uint8_t>> texel_store;
  RtbSceneDesc desc;
  RtbCamera cam;
  RtbRenderOpts opts;
  uint8_t* out = nullptr;
  RtbStats st;
  std::string err;              // empty = success
  // renderAsync only
  napi_deferred deferred = nullptr;
  napi_async_work work = nullptr;
  napi_ref out_ref = nullptr;   // keeps the Uint8ClampedArray (and its ArrayBuffer) alive until completion
};

void marshal(napi_env env, napi_value* argv, Job& J) {
  {
    // ---- scene ------------------------------------------------------------------------------------
    napi_value objects = prop(env, argv[0], "objects");
    uint32_t n = 0;
    NAPI_OK(napi_get_array_length(env, objects, &n));
    if (n > (1u << 20)) throw Fail{"more than 2^20 spheres"};
    J.spheres.assign(n, RtbSphere{});
    std::vector<RtbSphere>& spheres = J.spheres;
    for (uint32_t i = 0; i < n; i++) {
      napi_value o, m;
      NAPI_OK(napi_get_element(env, objects, i, &o));
      RtbSphere& s = spheres[i];
      memset(&s, 0, sizeof s);
      vec(env, o, "origin", s.origin, 3);
      s.r2 = num_prop(env, o, "r2", 0, true);                       // createSphere stores r*r (main.js:413)
      m = prop(env, o, "mtl");
      vec(env, m, "color", s.mtl.color, 3);
      vec(env, m, "albedo", s.mtl.albedo, 5);
      s.mtl.specular_exponent = num_prop(env, m, "specular_exponent", 0, true);
      s.mtl.refract_index = num_prop(env, m, "refract_index", 1, true);
      s.mtl.sampler_kind = (int32_t)num_prop(env, m, "sampler_kind", RTB_SAMPLER_COLOR);
      s.mtl.texture_unit = (int32_t)num_prop(env, m, "texture_unit", 0);
      s.mtl.checker_freq[0] = 5000; s.mtl.checker_freq[1] = 2500;    // main.js:129-131 defaults
      const double c0[3] = {1, 1, 0}, c1[3] = {1, 0, 1};
      memcpy(s.mtl.checker_color[0], c0, sizeof c0);
      memcpy(s.mtl.checker_color[1], c1, sizeof c1);
      s.mtl.star_threshold = num_prop(env, m, "star_threshold", 0.001);
      s.mtl.star_gain = num_prop(env, m, "star_gain", 1000);
      if (has(env, m, "checker_freq")) vec(env, m, "checker_freq", s.mtl.checker_freq, 2);
    }
    std::vector<RtbTexture>& textures = J.textures;
    std::vector<std::vector<uint8_t>>& texel_store = J.texel_store;
    if (has(env, argv[0], "textures")) {
      napi_value ts = prop(env, argv[0], "textures");
      uint32_t nt = 0;
      NAPI_OK(napi_get_array_length(env, ts, &nt));
      if (nt > 4096) throw Fail{"more than 4096 textures"};       // the library's own limits, checked before
      textures.resize(nt);                                          // anything is sized from JS-supplied numbers
      texel_store.resize(nt);
      for (uint32_t i = 0; i < nt; i++) {
        napi_value t;
        NAPI_OK(napi_get_element(env, ts, i, &t));
        textures[i].width = (int32_t)num_prop(env, t, "width", 0, true);
        textures[i].height = (int32_t)num_prop(env, t, "height", 0, true);
        if (textures[i].width <= 0 || textures[i].height <= 0) throw Fail{"texture not loaded (width/height 0)"};
        if (textures[i].width > 32768 || textures[i].height > 32768) throw Fail{"texture larger than 32768 x 32768"};
        read_texels(env, prop(env, t, "texels"), (size_t)textures[i].width * textures[i].height * 4, texel_store[i]);
        textures[i].texels = texel_store[i].data();
      }
    }
    RtbSceneDesc& desc = J.desc;
    memset(&desc, 0, sizeof desc);
    desc.abi_version = RTB_ABI_VERSION;
    desc.n_spheres = (int32_t)n;
    desc.spheres = spheres.data();
    desc.n_textures = (int32_t)textures.size();
    desc.textures = textures.data();
    const double dl[2][3] = {{5, 10, 5}, {5, 10, 0}};                 // main.js:283
    desc.n_lights = 2;
    memcpy(desc.lights, dl, sizeof dl);
    if (has(env, argv[0], "lights")) {
      napi_value ls = prop(env, argv[0], "lights");
      uint32_t nl = 0;
      NAPI_OK(napi_get_array_length(env, ls, &nl));
      if (nl > RTB_MAX_LIGHTS) throw Fail{"too many lights"};
      desc.n_lights = (int32_t)nl;
      for (uint32_t i = 0; i < nl; i++) {
        napi_value l;
        NAPI_OK(napi_get_element(env, ls, i, &l));
        for (uint32_t c = 0; c < 3; c++) {
          napi_value e;
          NAPI_OK(napi_get_element(env, l, c, &e));
          desc.lights[i][c] = num(env, e);
        }
      }
    }
    desc.light_intensity = num_prop(env, argv[0], "light_intensity", 50);   // main.js:284
    desc.eps = num_prop(env, argv[0], "eps", 0.001);                        // main.js:430-445

    // ---- camera -----------------------------------------------------------------------------------
    RtbCamera& cam = J.cam;
    memset(&cam, 0, sizeof cam);
    cam.width = (int32_t)num_prop(env, argv[1], "width", 0, true);
    cam.height = (int32_t)num_prop(env, argv[1], "height", 0, true);
    vec(env, argv[1], "origin", cam.origin, 3);
    vec(env, argv[1], "axisX", cam.axisX, 3);
    vec(env, argv[1], "axisY", cam.axisY, 3);
    vec(env, argv[1], "axisZ", cam.axisZ, 3);
    cam.projW = num_prop(env, argv[1], "projW", 0, true);
    cam.projH = num_prop(env, argv[1], "projH", 0, true);
    cam.projD = num_prop(env, argv[1], "projD", 0, true);
    cam.camera_mode = (int32_t)num_prop(env, argv[1], "camera_mode", RTB_CAMERA_REFERENCE);

    // ---- options ----------------------------------------------------------------------------------
    RtbRenderOpts& o = J.opts;
    rtb_default_opts(&o);
    o.segs = (int32_t)num_prop(env, argv[2], "segs", 8);                    // main.js:194
    o.spp_grid = (int32_t)num_prop(env, argv[2], "spp_grid", 1);
    o.precision = (int32_t)num_prop(env, argv[2], "precision", RTB_PRECISION_FP32);
    o.kernel = (int32_t)num_prop(env, argv[2], "kernel", RTB_KERNEL_AUTO);
    o.star_mode = (int32_t)num_prop(env, argv[2], "star_mode", RTB_STARS_OFF);
    o.y0 = (int32_t)num_prop(env, argv[2], "y0", 0);                        // progressive rows (main.js:200-201)
    o.y1 = (int32_t)num_prop(env, argv[2], "y1", 0);

    // ---- output: colorbuf.data, a Uint8ClampedArray(W*H*4) (main.js:83) ----------------------------
    bool typed = false;
    NAPI_OK(napi_is_typedarray(env, argv[3], &typed));
    if (!typed) throw Fail{"out must be a Uint8ClampedArray"};
    napi_typedarray_type ty; size_t len = 0; void* data = nullptr; napi_value ab; size_t off = 0;
    NAPI_OK(napi_get_typedarray_info(env, argv[3], &ty, &len, &data, &ab, &off));
    if (ty != napi_uint8_clamped_array && ty != napi_uint8_array) throw Fail{"out must be a Uint8ClampedArray"};
    if (len < (size_t)cam.width * cam.height * 4) throw Fail{"out is smaller than width*height*4"};
    J.out = static_cast<uint8_t*>(data);

  }
}

// No Node-API call in here: this is what runs on the worker thread.
void run(Job& J) {
  rtb_scene_t scene = nullptr;
  memset(&J.st, 0, sizeof J.st);
  try {
    if (rtb_init(-1) != RTB_OK) { J.err = rtb_last_error(); return; }
    if (rtb_scene_create(&J.desc, &scene) != RTB_OK) { J.err = rtb_last_error(); return; }
    if (rtb_render(scene, &J.cam, &J.opts, J.out, &J.st) != RTB_OK) J.err = rtb_last_error();
  } catch (const std::exception& e) {
    J.err = e.what();
  } catch (...) {
    J.err = "rtb200: unknown C++ exception";
  }
  if (scene) rtb_scene_destroy(scene);
}

napi_value result(napi_env env, const Job& J) {
  napi_value res, v;
  NAPI_OK(napi_create_object(env, &res));
  NAPI_OK(napi_create_double(env, J.st.kernel_ms, &v)); NAPI_OK(napi_set_named_property(env, res, "kernel_ms", v));
  NAPI_OK(napi_create_double(env, J.st.d2h_ms, &v)); NAPI_OK(napi_set_named_property(env, res, "d2h_ms", v));
  NAPI_OK(napi_create_double(env, (double)J.st.nodes, &v)); NAPI_OK(napi_set_named_property(env, res, "nodes", v));
  NAPI_OK(napi_create_double(env, (double)J.st.shadow_rays, &v)); NAPI_OK(napi_set_named_property(env, res, "shadow_rays", v));
  NAPI_OK(napi_create_double(env, (double)J.st.launches, &v)); NAPI_OK(napi_set_named_property(env, res, "launches", v));
  return res;
}

napi_value Render(napi_env env, napi_callback_info info) {
  napi_value undef;
  napi_get_undefined(env, &undef);
  try {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    if (argc < 4) throw Fail{"render(scene, camera, opts, out) takes 4 arguments"};
    Job J;
    marshal(env, argv, J);
    run(J);
    if (!J.err.empty()) throw Fail{J.err};
    return result(env, J);
  } catch (const Fail& f) {
    napi_throw_error(env, nullptr, f.msg.c_str());          // JS `Error`; nothing crosses as a C++ exception
    return undef;
  } catch (const std::exception& e) {                       // std::bad_alloc / length_error of the staging vectors
    napi_throw_error(env, nullptr, e.what());
    return undef;
  } catch (...) {
    napi_throw_error(env, nullptr, "rtb200: unknown C++ exception");
    return undef;
  }
}

// renderAsync(scene, camera, opts, out) -> Promise<stats>.  `out` must not be read or
// resized by JavaScript until the promise settles; rows outside opts.y0..y1 are left untouched, so a page can
// keep the reference's row-by-row cadence with one promise per band.
void AsyncExecute(napi_env, void* data) { run(*static_cast<Job*>(data)); }

void AsyncComplete(napi_env env, napi_status status, void* data) {
  Job* J = static_cast<Job*>(data);
  napi_value v = nullptr;
  bool ok = status == napi_ok && J->err.empty();
  if (ok) {
    try { v = result(env, *J); } catch (...) { ok = false; J->err = "rtb200: could not build the result object"; }
  }
  if (ok) {
    napi_resolve_deferred(env, J->deferred, v);
  } else {
    napi_value msg, err;
    const std::string m = J->err.empty() ? std::string("rtb200: the async work was cancelled") : J->err;
    napi_create_string_utf8(env, m.c_str(), NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &err);
    napi_reject_deferred(env, J->deferred, err);
  }
  if (J->out_ref) napi_delete_reference(env, J->out_ref);
  if (J->work) napi_delete_async_work(env, J->work);
  delete J;
}

napi_value RenderAsync(napi_env env, napi_callback_info info) {
  napi_value undef;
  napi_get_undefined(env, &undef);
  Job* J = nullptr;
  try {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    if (argc < 4) throw Fail{"renderAsync(scene, camera, opts, out) takes 4 arguments"};
    J = new Job();
    marshal(env, argv, *J);                                 // copies everything but the framebuffer
    napi_value promise, name;
    NAPI_OK(napi_create_reference(env, argv[3], 1, &J->out_ref));
    NAPI_OK(napi_create_promise(env, &J->deferred, &promise));
    NAPI_OK(napi_create_string_utf8(env, "rtb200.renderAsync", NAPI_AUTO_LENGTH, &name));
    NAPI_OK(napi_create_async_work(env, nullptr, name, AsyncExecute, AsyncComplete, J, &J->work));
    NAPI_OK(napi_queue_async_work(env, J->work));
    return promise;                                         // J now belongs to AsyncComplete
  } catch (const Fail& f) {
    if (J) { if (J->out_ref) napi_delete_reference(env, J->out_ref); if (J->work) napi_delete_async_work(env, J->work); delete J; }
    napi_throw_error(env, nullptr, f.msg.c_str());
    return undef;
  } catch (const std::exception& e) {
    if (J) { if (J->out_ref) napi_delete_reference(env, J->out_ref); if (J->work) napi_delete_async_work(env, J->work); delete J; }
    napi_throw_error(env, nullptr, e.what());
    return undef;
  } catch (...) {
    if (J) { if (J->out_ref) napi_delete_reference(env, J->out_ref); if (J->work) napi_delete_async_work(env, J->work); delete J; }
    napi_throw_error(env, nullptr, "rtb200: unknown C++ exception");
    return undef;
  }
}

// decodeTexture(pngBytes: Uint8Array | Buffer) -> {width, height, texels: Uint8ClampedArray, loaded: 1}: the object
// loadTexture leaves behind once img.onload has run (main.js:385-393), for hosts without Image / canvas.
napi_value DecodeTexture(napi_env env, napi_callback_info info) {
  napi_value undef;
  napi_get_undefined(env, &undef);
  try {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    if (argc < 1) throw Fail{"decodeTexture(pngBytes) takes 1 argument"};
    bool typed = false;
    NAPI_OK(napi_is_typedarray(env, argv[0], &typed));
    if (!typed) throw Fail{"decodeTexture: expected a Uint8Array / Buffer of PNG bytes"};
    napi_typedarray_type ty; size_t len = 0; void* data = nullptr; napi_value ab; size_t off = 0;
    NAPI_OK(napi_get_typedarray_info(env, argv[0], &ty, &len, &data, &ab, &off));
    if (ty != napi_uint8_array && ty != napi_uint8_clamped_array) throw Fail{"decodeTexture: expected a Uint8Array / Buffer"};
    int32_t w = 0, h = 0;
    if (rtb_decode_png((const uint8_t*)data, len, &w, &h, nullptr, 0) != RTB_OK) throw Fail{rtb_last_error()};
    const size_t bytes = (size_t)w * (size_t)h * 4;
    void* texels = nullptr;
    napi_value buf, arr, res, v;
    NAPI_OK(napi_create_arraybuffer(env, bytes, &texels, &buf));
    if (rtb_decode_png((const uint8_t*)data, len, &w, &h, (uint8_t*)texels, bytes) != RTB_OK) throw Fail{rtb_last_error()};
    NAPI_OK(napi_create_typedarray(env, napi_uint8_clamped_array, bytes, buf, 0, &arr));
    NAPI_OK(napi_create_object(env, &res));
    NAPI_OK(napi_create_double(env, w, &v)); NAPI_OK(napi_set_named_property(env, res, "width", v));
    NAPI_OK(napi_create_double(env, h, &v)); NAPI_OK(napi_set_named_property(env, res, "height", v));
    NAPI_OK(napi_create_double(env, 1, &v)); NAPI_OK(napi_set_named_property(env, res, "loaded", v));
    NAPI_OK(napi_set_named_property(env, res, "texels", arr));
    return res;
  } catch (const Fail& f) {
    napi_throw_error(env, nullptr, f.msg.c_str());
    return undef;
  } catch (const std::exception& e) {
    napi_throw_error(env, nullptr, e.what());
    return undef;
  } catch (...) {
    napi_throw_error(env, nullptr, "rtb200: unknown C++ exception");
    return undef;
  }
}

napi_value Version(napi_env env, napi_callback_info) {
  napi_value v;
  napi_create_string_utf8(env, rtb_version(), NAPI_AUTO_LENGTH, &v);
  return v;
}

}  // namespace

// The well-known entry symbol Node resolves with dlsym() after dlopen()ing the .node file.
extern "C" __attribute__((visibility("default"))) napi_value napi_register_module_v1(napi_env env, napi_value exports) {
  napi_value fn;
  if (napi_create_function(env, "render", NAPI_AUTO_LENGTH, Render, nullptr, &fn) == napi_ok)
    napi_set_named_property(env, exports, "render", fn);
  if (napi_create_function(env, "renderAsync", NAPI_AUTO_LENGTH, RenderAsync, nullptr, &fn) == napi_ok)
    napi_set_named_property(env, exports, "renderAsync", fn);
  if (napi_create_function(env, "decodeTexture", NAPI_AUTO_LENGTH, DecodeTexture, nullptr, &fn) == napi_ok)
    napi_set_named_property(env, exports, "decodeTexture", fn);
  if (napi_create_function(env, "version", NAPI_AUTO_LENGTH, Version, nullptr, &fn) == napi_ok)
    napi_set_named_property(env, exports, "version", fn);
  return exports;
}
