// sweep_micro4.cu — where should the sphere table of the shared-origin sweep live?  Per sphere pair the sweep
// needs 32 bytes that are UNIFORM across the warp (every lane tests the same spheres) and, per ray, 3 FFMA2 +
// 2 funnel shifts.  From shared memory that is 2 LDS.128 per pair: 2 x 512 B of register-file writes for 32 B of
// information.  Alternatives: the constant bank (a __constant__ array, or the table passed BY VALUE as a
// __grid_constant__ kernel parameter), read through the uniform datapath.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sweep_micro4 tools/sweep_micro4.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int NS = 1024, REPS = 16, GROUP = 8, CHECKN = 32;

struct Table { float4 t[NS]; };
__constant__ float4 ctab[NS];

template <int R, int SRC>
__global__ void __launch_bounds__(256, 2) k(float* out, const float4* gtab, const float* dirs, const __grid_constant__ Table ptab) {
  __shared__ float4 stab[SRC == 0 ? NS : 1];
  if (SRC == 0) {
    for (int i = threadIdx.x; i < NS; i += blockDim.x) stab[i] = gtab[i];
    __syncthreads();
  }
  float nd[R][3];
  unsigned mask[R];
#pragma unroll
  for (int q = 0; q < R; q++) {
    nd[q][0] = dirs[q * 3] + threadIdx.x * 1e-6f; nd[q][1] = dirs[q * 3 + 1]; nd[q][2] = dirs[q * 3 + 2];
    mask[q] = 0;
  }
  unsigned acc = 0;
  for (int rep = 0; rep < REPS; rep++) {
    for (int j = 0; j < NS; j += CHECKN) {
#pragma unroll 1
      for (int jg = j; jg < j + CHECKN; jg += GROUP) {
#pragma unroll
        for (int u = 0; u < GROUP / 2; u++) {
          float4 A, B;
          if (SRC == 0) { A = stab[jg + 2 * u]; B = stab[jg + 2 * u + 1]; }
          else if (SRC == 1) { A = ctab[jg + 2 * u]; B = ctab[jg + 2 * u + 1]; }
          else { A = ptab.t[jg + 2 * u]; B = ptab.t[jg + 2 * u + 1]; }
#pragma unroll
          for (int q = 0; q < R; q++) {
            float2 t = __ffma2_rn(make_float2(nd[q][2], nd[q][2]), make_float2(B.x, B.y), make_float2(B.z, B.w));
            t = __ffma2_rn(make_float2(nd[q][1], nd[q][1]), make_float2(A.z, A.w), t);
            t = __ffma2_rn(make_float2(nd[q][0], nd[q][0]), make_float2(A.x, A.y), t);
            mask[q] = __funnelshift_l(__float_as_uint(t.x), mask[q], 1);
            mask[q] = __funnelshift_l(__float_as_uint(t.y), mask[q], 1);
          }
        }
      }
      unsigned any = 0;
#pragma unroll
      for (int q = 0; q < R; q++) any |= mask[q];
      if (any == 0x5a5a5a5au && rep == 1000) acc += any;
    }
  }
#pragma unroll
  for (int q = 0; q < R; q++) acc += mask[q] + __float_as_uint(nd[q][0]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

// scalar form: one sphere per float4 (Lx, Ly, Lz, sc) in the constant bank, read through the UNIFORM datapath
// (LDCU into uniform registers; FFMA takes one uniform-register source): per test 3 FFMA + 1 SHF, no register
// bandwidth for the table at all.  ORDER 0: t = fma(ndz, Lz, sc) first (two uniform sources in one FFMA: the
// compiler moves one through a vector register); ORDER 1: t = ndz * Lz; ...; t += sc (4 ops, all uniform).
__constant__ float4 ctab1[NS];
template <int R, int ORDER>
__global__ void __launch_bounds__(256, 2) ks(float* out, const float* dirs) {
  float nd[R][3];
  unsigned mask[R];
#pragma unroll
  for (int q = 0; q < R; q++) {
    nd[q][0] = dirs[q * 3] + threadIdx.x * 1e-6f; nd[q][1] = dirs[q * 3 + 1]; nd[q][2] = dirs[q * 3 + 2];
    mask[q] = 0;
  }
  unsigned acc = 0;
  for (int rep = 0; rep < REPS; rep++) {
    for (int j = 0; j < NS; j += CHECKN) {
#pragma unroll 1
      for (int jg = j; jg < j + CHECKN; jg += GROUP) {
#pragma unroll
        for (int u = 0; u < GROUP; u++) {
          const float4 S = ctab1[jg + u];
#pragma unroll
          for (int q = 0; q < R; q++) {
            float t;
            if (ORDER == 0) {
              t = fmaf(nd[q][2], S.z, S.w);
              t = fmaf(nd[q][1], S.y, t);
              t = fmaf(nd[q][0], S.x, t);
            } else {
              t = nd[q][2] * S.z;
              t = fmaf(nd[q][1], S.y, t);
              t = fmaf(nd[q][0], S.x, t);
              t += S.w;
            }
            mask[q] = __funnelshift_l(__float_as_uint(t), mask[q], 1);
          }
        }
      }
      unsigned any = 0;
#pragma unroll
      for (int q = 0; q < R; q++) any |= mask[q];
      if (any == 0x5a5a5a5au && rep == 1000) acc += any;
    }
  }
#pragma unroll
  for (int q = 0; q < R; q++) acc += mask[q] + __float_as_uint(nd[q][0]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

template <typename F> float time_ms(F f) {
  cudaEvent_t a, b; CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
  f(); f(); CHECK(cudaDeviceSynchronize());
  CHECK(cudaEventRecord(a)); for (int i = 0; i < 5; i++) f(); CHECK(cudaEventRecord(b)); CHECK(cudaEventSynchronize(b));
  float ms; CHECK(cudaEventElapsedTime(&ms, a, b)); return ms / 5;
}
static Table h;
template <int R, int SRC>
void run(const char* name, float* out, const float4* tab, const float* dirs, int sms, double hz, int ctas) {
  float ms = time_ms([&] { k<R, SRC><<<sms * ctas, 256>>>(out, tab, dirs, h); });
  CHECK(cudaGetLastError());
  const double tests = (ctas * 8 / 4.0) * REPS * (double)NS * R;        // test-warps per SMSP
  printf("{\"table\": \"%s\", \"R\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"cycles_per_pair_per_ray\": %.2f, \"cycles_per_ffma2\": %.2f}\n",
         name, R, ctas, ms, 2 * ms * 1e-3 * hz / tests, 2 * ms * 1e-3 * hz / tests / 3);
}
template <int R, int ORDER>
void run_s(const char* name, float* out, const float* dirs, int sms, double hz, int ctas) {
  float ms = time_ms([&] { ks<R, ORDER><<<sms * ctas, 256>>>(out, dirs); });
  CHECK(cudaGetLastError());
  const double tests = (ctas * 8 / 4.0) * REPS * (double)NS * R;        // test-warps per SMSP
  printf("{\"table\": \"%s\", \"R\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"cycles_per_pair_per_ray\": %.2f, \"cycles_per_test\": %.2f}\n",
         name, R, ctas, ms, 2 * ms * 1e-3 * hz / tests, ms * 1e-3 * hz / tests);
}
int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double hz = khz * 1e3;
  float4* tab; float* dirs; float* out;
  CHECK(cudaMalloc(&tab, NS * sizeof(float4))); CHECK(cudaMalloc(&dirs, 64 * 4));
  CHECK(cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4));
  for (int i = 0; i < NS; i++) h.t[i] = make_float4(1.f + i, 2.f, -3.f + i, 1e30f);
  float hd[64]; for (int i = 0; i < 64; i++) hd[i] = 0.01f * i;
  CHECK(cudaMemcpy(tab, h.t, sizeof h.t, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dirs, hd, sizeof hd, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpyToSymbol(ctab, h.t, sizeof h.t));
  const int sms = p.multiProcessorCount;
  for (int ctas = 1; ctas <= 2; ctas++) {
    run<8, 0>("shared (LDS.128)", out, tab, dirs, sms, hz, ctas);
    run<8, 1>("__constant__", out, tab, dirs, sms, hz, ctas);
    run<8, 2>("kernel parameter", out, tab, dirs, sms, hz, ctas);
  }
  CHECK(cudaMemcpyToSymbol(ctab1, h.t, sizeof h.t));
  for (int ctas = 1; ctas <= 2; ctas++) {
    run_s<8, 0>("scalar FFMA, uniform-register table, 3 FFMA + SHF", out, dirs, sms, hz, ctas);
    run_s<8, 1>("scalar FFMA, uniform-register table, FMUL + 2 FFMA + FADD + SHF", out, dirs, sms, hz, ctas);
  }
  run_s<12, 0>("scalar FFMA, uniform-register table, 3 FFMA + SHF", out, dirs, sms, hz, 2);
  run_s<16, 0>("scalar FFMA, uniform-register table, 3 FFMA + SHF", out, dirs, sms, hz, 2);
  run_s<4, 0>("scalar FFMA, uniform-register table, 3 FFMA + SHF", out, dirs, sms, hz, 2);
  run<4, 0>("shared (LDS.128)", out, tab, dirs, sms, hz, 2);
  run<4, 1>("__constant__", out, tab, dirs, sms, hz, 2);
  run<12, 1>("__constant__", out, tab, dirs, sms, hz, 2);
  run<16, 1>("__constant__", out, tab, dirs, sms, hz, 2);
  return 0;
}
