// mock_node.cc — TEST INFRASTRUCTURE: a minimal stand-in for the Node-API surface csrc/napi_shim.cc uses, so that
// the addon's marshalling, error and promise logic can be EXECUTED where no Node exists (this image has none).
// It is not Node: values are a small tagged struct, there is no GC, the "event loop" is mock_run_loop().
// Loaded with RTLD_GLOBAL before rtb200.node, so that the addon's undefined napi_* symbols bind here — the way
// they bind to the `node` executable in production.  Python drives it through the mock_* helpers (ctypes).
#include <dlfcn.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

extern "C" {
typedef enum { napi_ok = 0, napi_invalid_arg = 1, napi_generic_failure = 9 } napi_status;
typedef enum {
  napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object,
  napi_function, napi_external, napi_bigint
} napi_valuetype;
struct Val;
struct Env;
typedef Env* napi_env;
typedef Val* napi_value;
struct CbInfo { std::vector<Val*> args; };
typedef CbInfo* napi_callback_info;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef void (*napi_async_execute_callback)(napi_env, void*);
typedef void (*napi_async_complete_callback)(napi_env, napi_status, void*);
}

enum Kind { K_UNDEF, K_NULL, K_NUM, K_STR, K_OBJ, K_ARR, K_TYPED, K_FN, K_PROMISE, K_ERROR };
struct Val {
  Kind kind = K_UNDEF;
  double num = 0;
  std::string str;
  std::map<std::string, Val*> props;
  std::vector<Val*> elems;
  void* data = nullptr; size_t len = 0; int ta_type = 0;
  napi_callback cb = nullptr;
  int promise_state = 0;          // 0 pending, 1 resolved, 2 rejected
  Val* promise_value = nullptr;
};
struct Work {
  napi_async_execute_callback exec; napi_async_complete_callback done; void* data;
  std::thread th; bool queued = false, deleted = false;
};
struct Env {
  Val undef;
  std::string exception; bool has_exception = false;
  std::vector<Work*> works;
  int live_refs = 0;
};
struct napi_deferred__ { Val* promise; };
struct napi_ref__ { Val* v; };

static Val* mk(Kind k) { Val* v = new Val(); v->kind = k; return v; }

extern "C" {
typedef napi_deferred__* napi_deferred;
typedef Work* napi_async_work;
typedef napi_ref__* napi_ref;

// ---- the Node-API subset -----------------------------------------------------------------------------------
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value*, void**) {
  const size_t want = *argc;
  for (size_t i = 0; i < want && i < info->args.size(); i++) argv[i] = info->args[i];
  *argc = info->args.size();
  return napi_ok;
}
napi_status napi_get_named_property(napi_env env, napi_value o, const char* k, napi_value* r) {
  if (!o || (o->kind != K_OBJ && o->kind != K_ARR && o->kind != K_TYPED && o->kind != K_FN)) return napi_invalid_arg;
  auto it = o->props.find(k);
  *r = it == o->props.end() ? &env->undef : it->second;
  return napi_ok;
}
napi_status napi_has_named_property(napi_env, napi_value o, const char* k, bool* r) {
  if (!o || (o->kind != K_OBJ && o->kind != K_ARR && o->kind != K_TYPED && o->kind != K_FN)) return napi_invalid_arg;
  *r = o->props.count(k) != 0;
  return napi_ok;
}
napi_status napi_get_array_length(napi_env, napi_value v, uint32_t* r) {
  if (!v || v->kind != K_ARR) return napi_invalid_arg;
  *r = (uint32_t)v->elems.size();
  return napi_ok;
}
napi_status napi_get_element(napi_env env, napi_value o, uint32_t i, napi_value* r) {
  if (!o || o->kind != K_ARR) return napi_invalid_arg;
  *r = i < o->elems.size() ? o->elems[i] : &env->undef;
  return napi_ok;
}
napi_status napi_get_value_double(napi_env, napi_value v, double* r) {
  if (!v || v->kind != K_NUM) return napi_invalid_arg;      // napi_number_expected
  *r = v->num;
  return napi_ok;
}
napi_status napi_get_value_int32(napi_env, napi_value v, int32_t* r) {
  if (!v || v->kind != K_NUM) return napi_invalid_arg;
  *r = (int32_t)v->num;
  return napi_ok;
}
napi_status napi_typeof(napi_env, napi_value v, napi_valuetype* r) {
  switch (v->kind) {
    case K_UNDEF: *r = napi_undefined; break;
    case K_NULL: *r = napi_null; break;
    case K_NUM: *r = napi_number; break;
    case K_STR: *r = napi_string; break;
    case K_FN: *r = napi_function; break;
    default: *r = napi_object;
  }
  return napi_ok;
}
napi_status napi_is_typedarray(napi_env, napi_value v, bool* r) { *r = v && v->kind == K_TYPED; return napi_ok; }
napi_status napi_is_array(napi_env, napi_value v, bool* r) { *r = v && v->kind == K_ARR; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env, napi_value v, int* type, size_t* length, void** data, napi_value* ab, size_t* off) {
  if (!v || v->kind != K_TYPED) return napi_invalid_arg;
  if (type) *type = v->ta_type;
  if (length) *length = v->len;
  if (data) *data = v->data;
  if (ab) *ab = v;
  if (off) *off = 0;
  return napi_ok;
}
napi_status napi_create_function(napi_env, const char*, size_t, napi_callback cb, void*, napi_value* r) {
  *r = mk(K_FN); (*r)->cb = cb; return napi_ok;
}
napi_status napi_set_named_property(napi_env, napi_value o, const char* k, napi_value v) {
  if (!o) return napi_invalid_arg;
  o->props[k] = v;
  return napi_ok;
}
napi_status napi_create_object(napi_env, napi_value* r) { *r = mk(K_OBJ); return napi_ok; }
napi_status napi_create_double(napi_env, double d, napi_value* r) { *r = mk(K_NUM); (*r)->num = d; return napi_ok; }
napi_status napi_create_string_utf8(napi_env, const char* s, size_t n, napi_value* r) {
  *r = mk(K_STR); (*r)->str = n == SIZE_MAX ? std::string(s) : std::string(s, n); return napi_ok;
}
napi_status napi_get_undefined(napi_env env, napi_value* r) { *r = &env->undef; return napi_ok; }
napi_status napi_throw_error(napi_env env, const char*, const char* msg) {
  env->has_exception = true; env->exception = msg ? msg : ""; return napi_ok;
}
napi_status napi_create_error(napi_env, napi_value, napi_value msg, napi_value* r) {
  *r = mk(K_ERROR); (*r)->str = msg ? msg->str : ""; return napi_ok;
}
napi_status napi_create_promise(napi_env, napi_deferred* d, napi_value* p) {
  *p = mk(K_PROMISE); *d = new napi_deferred__{*p}; return napi_ok;
}
napi_status napi_resolve_deferred(napi_env, napi_deferred d, napi_value v) {
  d->promise->promise_state = 1; d->promise->promise_value = v; delete d; return napi_ok;
}
napi_status napi_reject_deferred(napi_env, napi_deferred d, napi_value v) {
  d->promise->promise_state = 2; d->promise->promise_value = v; delete d; return napi_ok;
}
napi_status napi_create_async_work(napi_env env, napi_value, napi_value, napi_async_execute_callback e,
                                   napi_async_complete_callback c, void* data, napi_async_work* r) {
  Work* w = new Work(); w->exec = e; w->done = c; w->data = data; env->works.push_back(w); *r = w; return napi_ok;
}
napi_status napi_queue_async_work(napi_env env, napi_async_work w) {
  w->queued = true;
  w->th = std::thread([env, w] { w->exec(env, w->data); });   // the libuv worker
  return napi_ok;
}
napi_status napi_delete_async_work(napi_env, napi_async_work w) { w->deleted = true; return napi_ok; }
napi_status napi_create_reference(napi_env env, napi_value v, uint32_t, napi_ref* r) { *r = new napi_ref__{v}; env->live_refs++; return napi_ok; }
napi_status napi_create_arraybuffer(napi_env, size_t n, void** data, napi_value* r) {
  *r = mk(K_TYPED); (*r)->ta_type = -1; (*r)->len = n; (*r)->data = n ? calloc(n, 1) : nullptr;   // never freed: no GC here
  if (data) *data = (*r)->data;
  return napi_ok;
}
napi_status napi_create_typedarray(napi_env, int type, size_t length, napi_value ab, size_t off, napi_value* r) {
  if (!ab || ab->kind != K_TYPED || ab->ta_type != -1 || off + length > ab->len) return napi_invalid_arg;
  *r = mk(K_TYPED); (*r)->ta_type = type; (*r)->len = length; (*r)->data = (char*)ab->data + off;
  return napi_ok;
}
napi_status napi_delete_reference(napi_env env, napi_ref r) { delete r; env->live_refs--; return napi_ok; }

// ---- helpers the Python test drives (ctypes) ---------------------------------------------------------------
Env* mock_env_new() { return new Env(); }
Val* mock_undefined(Env* e) { return &e->undef; }
Val* mock_object() { return mk(K_OBJ); }
Val* mock_array() { return mk(K_ARR); }
Val* mock_number(double d) { Val* v = mk(K_NUM); v->num = d; return v; }
Val* mock_string(const char* s) { Val* v = mk(K_STR); v->str = s; return v; }
Val* mock_typedarray(int type, void* data, size_t len) { Val* v = mk(K_TYPED); v->ta_type = type; v->data = data; v->len = len; return v; }
void mock_set(Val* o, const char* k, Val* v) { o->props[k] = v; }
void mock_push(Val* a, Val* v) { a->elems.push_back(v); }
Val* mock_get(Env* e, Val* o, const char* k) { auto it = o->props.find(k); return it == o->props.end() ? &e->undef : it->second; }
int mock_kind(Val* v) { return v->kind; }
void* mock_typed_data(Val* v) { return v->kind == K_TYPED ? v->data : nullptr; }
size_t mock_typed_len(Val* v) { return v->kind == K_TYPED ? v->len : 0; }
int mock_typed_type(Val* v) { return v->kind == K_TYPED ? v->ta_type : -2; }
double mock_number_value(Val* v) { return v->num; }
const char* mock_string_value(Val* v) { return v->str.c_str(); }
int mock_promise_state(Val* v) { return v->promise_state; }
Val* mock_promise_value(Val* v) { return v->promise_value; }
int mock_live_refs(Env* e) { return e->live_refs; }
const char* mock_exception(Env* e) { return e->has_exception ? e->exception.c_str() : nullptr; }
void mock_clear_exception(Env* e) { e->has_exception = false; e->exception.clear(); }

// dlopen the addon and run its napi_register_module_v1: returns `exports`
Val* mock_load_addon(Env* e, const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { e->has_exception = true; e->exception = dlerror(); return nullptr; }
  typedef Val* (*Reg)(Env*, Val*);
  Reg reg = (Reg)dlsym(h, "napi_register_module_v1");
  if (!reg) { e->has_exception = true; e->exception = "napi_register_module_v1 not found"; return nullptr; }
  return reg(e, mk(K_OBJ));
}
// call a JS-visible function of the addon; returns NULL when it threw (see mock_exception)
Val* mock_call(Env* e, Val* fn, int argc, Val** argv) {
  if (!fn || fn->kind != K_FN) { e->has_exception = true; e->exception = "not a function"; return nullptr; }
  CbInfo info;
  for (int i = 0; i < argc; i++) info.args.push_back(argv[i]);
  Val* r = fn->cb(e, &info);
  return e->has_exception ? nullptr : r;
}
// the "event loop": wait for the workers, run their completion callbacks on THIS thread; returns how many ran
int mock_run_loop(Env* e) {
  int n = 0;
  std::vector<Work*> works;
  works.swap(e->works);
  for (Work* w : works) {
    if (!w->queued) continue;
    w->th.join();
    w->done(e, napi_ok, w->data);
    n++;
  }
  return n;
}
}  // extern "C"
