// sweep_micro3.cu — what does collecting the candidate information cost?  Packed shared-origin sweep
// (per sphere pair 2 LDS.128, per ray 3 FFMA2) with three ways of remembering "some test went negative":
//   shf2 : 2 funnel shifts per pair per ray  (sign bits -> per-sphere mask; what wavefront.cu does)
//   lop3 : 1 LOP3 per pair per ray           (m |= bits(t.x) | bits(t.y): sign bit = "any of the group")
//   min3 : 1 FMNMX3 per pair per ray         (m = min(m, t.x, t.y))
//   none : nothing (t values xor-folded every 4th pair only; lower bound)
// lop3 / min3 only say THAT a group of 8 has a candidate; the slow path would have to find out which.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int NS = 1024, REPS = 16, GROUP = 8;

template <int R, int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, const float4* gtab, const float* dirs) {
  __shared__ float4 tab[NS];
  for (int i = threadIdx.x; i < NS; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  float nd[R][3];
  unsigned mask[R];
  float mn[R];
#pragma unroll
  for (int q = 0; q < R; q++) {
    nd[q][0] = dirs[q * 3] + threadIdx.x * 1e-6f; nd[q][1] = dirs[q * 3 + 1]; nd[q][2] = dirs[q * 3 + 2];
    mask[q] = 0; mn[q] = 1e30f;
  }
  unsigned acc = 0;
  for (int rep = 0; rep < REPS; rep++) {
    for (int j = 0; j < NS; j += GROUP) {
#pragma unroll
      for (int u = 0; u < GROUP / 2; u++) {
        const float4 A = tab[j + 2 * u], B = tab[j + 2 * u + 1];
#pragma unroll
        for (int q = 0; q < R; q++) {
          float2 t = __ffma2_rn(make_float2(nd[q][2], nd[q][2]), make_float2(B.x, B.y), make_float2(B.z, B.w));
          t = __ffma2_rn(make_float2(nd[q][1], nd[q][1]), make_float2(A.z, A.w), t);
          t = __ffma2_rn(make_float2(nd[q][0], nd[q][0]), make_float2(A.x, A.y), t);
          if (MODE == 0) {
            mask[q] = __funnelshift_l(__float_as_uint(t.x), mask[q], 1);
            mask[q] = __funnelshift_l(__float_as_uint(t.y), mask[q], 1);
          } else if (MODE == 1) {
            mask[q] |= __float_as_uint(t.x) | __float_as_uint(t.y);
          } else if (MODE == 2) {
            mn[q] = fminf(fminf(mn[q], t.x), t.y);
          } else {
            if (u == 3) mask[q] ^= __float_as_uint(t.x) ^ __float_as_uint(t.y);
            else { nd[q][0] += t.x * 1e-30f; }   // keep the chain alive with an FMA instead of an ALU op
          }
        }
      }
      unsigned any = 0;
      if (MODE == 2) {
        float m = mn[0];
#pragma unroll
        for (int q = 1; q < R; q++) m = fminf(m, mn[q]);
        any = __float_as_uint(m) >> 31;
        if (any && rep == 1000) acc += 1;
      } else {
#pragma unroll
        for (int q = 0; q < R; q++) any |= mask[q];
        if (MODE == 0) { if ((any & 0xffu) == 0x5au && rep == 1000) acc += any; }
        else if ((any >> 31) && rep == 1000) acc += any;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < R; q++) acc += mask[q] + __float_as_uint(mn[q]) + __float_as_uint(nd[q][0]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

template <typename F> float time_ms(F f) {
  cudaEvent_t a, b; CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
  f(); f(); CHECK(cudaDeviceSynchronize());
  CHECK(cudaEventRecord(a)); for (int i = 0; i < 5; i++) f(); CHECK(cudaEventRecord(b)); CHECK(cudaEventSynchronize(b));
  float ms; CHECK(cudaEventElapsedTime(&ms, a, b)); return ms / 5;
}
template <int R, int MODE>
void run(const char* name, float* out, const float4* tab, const float* dirs, int sms, double hz) {
  const int ctas = 2;
  float ms = time_ms([&] { k<R, MODE><<<sms * ctas, 256>>>(out, tab, dirs); });
  const double tests = (ctas * 8 / 4.0) * REPS * (double)NS * R;        // test-warps per SMSP
  printf("{\"collect\": \"%s\", \"R\": %d, \"ms\": %.3f, \"smsp_cycles_per_test_warp\": %.3f, \"cycles_per_pair_per_ray\": %.2f}\n",
         name, R, ms, ms * 1e-3 * hz / tests, 2 * ms * 1e-3 * hz / tests * 4 / 4);
}
int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double hz = khz * 1e3;
  float4* tab; float* dirs; float* out;
  CHECK(cudaMalloc(&tab, NS * sizeof(float4))); CHECK(cudaMalloc(&dirs, 64 * 4));
  CHECK(cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4));
  static float4 h[NS]; for (int i = 0; i < NS; i++) h[i] = make_float4(1.f + i, 2.f, -3.f + i, 1e30f);
  float hd[64]; for (int i = 0; i < 64; i++) hd[i] = 0.01f * i;
  CHECK(cudaMemcpy(tab, h, sizeof h, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dirs, hd, sizeof hd, cudaMemcpyHostToDevice));
  const int sms = p.multiProcessorCount;
  run<8, 0>("shf2", out, tab, dirs, sms, hz);
  run<8, 1>("lop3", out, tab, dirs, sms, hz);
  run<8, 2>("min3", out, tab, dirs, sms, hz);
  run<8, 3>("none", out, tab, dirs, sms, hz);
  run<4, 0>("shf2", out, tab, dirs, sms, hz);
  run<4, 1>("lop3", out, tab, dirs, sms, hz);
  run<4, 2>("min3", out, tab, dirs, sms, hz);
  return 0;
}
