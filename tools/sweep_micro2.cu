// sweep_micro2.cu — scalar-FFMA vs packed-FFMA2 form of the shared-origin sweep, same work per test:
//   packed : per sphere PAIR  2 LDS.128, per ray 3 FFMA2 + 2 SHF     (what wavefront.cu does)
//   scalar : per SPHERE       1 LDS.128, per ray 3 FFMA  + 1 SHF
// Output: SMSP cycles per (ray x sphere) test-warp.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int NS = 1024, REPS = 16;

template <int R, bool PACKED, int UNROLL>
__global__ void __launch_bounds__(256) k(float* out, const float4* gtab, const float* dirs) {
  __shared__ float4 tab[NS];
  for (int i = threadIdx.x; i < NS; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  float nd[R][3];
  unsigned mask[R];
#pragma unroll
  for (int q = 0; q < R; q++) {
    nd[q][0] = dirs[q * 3] + threadIdx.x * 1e-6f; nd[q][1] = dirs[q * 3 + 1]; nd[q][2] = dirs[q * 3 + 2];
    mask[q] = 0;
  }
  unsigned acc = 0;
  for (int rep = 0; rep < REPS; rep++) {
    for (int j = 0; j < NS; j += UNROLL) {
      if (PACKED) {
#pragma unroll
        for (int u = 0; u < UNROLL / 2; u++) {
          const float4 A = tab[j + 2 * u], B = tab[j + 2 * u + 1];
#pragma unroll
          for (int q = 0; q < R; q++) {
            float2 t = __ffma2_rn(make_float2(nd[q][2], nd[q][2]), make_float2(B.x, B.y), make_float2(B.z, B.w));
            t = __ffma2_rn(make_float2(nd[q][1], nd[q][1]), make_float2(A.z, A.w), t);
            t = __ffma2_rn(make_float2(nd[q][0], nd[q][0]), make_float2(A.x, A.y), t);
            mask[q] = __funnelshift_l(__float_as_uint(t.x), mask[q], 1);
            mask[q] = __funnelshift_l(__float_as_uint(t.y), mask[q], 1);
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
          const float4 s = tab[j + u];
#pragma unroll
          for (int q = 0; q < R; q++) {
            const float t = fmaf(nd[q][0], s.x, fmaf(nd[q][1], s.y, fmaf(nd[q][2], s.z, s.w)));
            mask[q] = __funnelshift_l(__float_as_uint(t), mask[q], 1);
          }
        }
      }
      unsigned any = 0;
#pragma unroll
      for (int q = 0; q < R; q++) any |= mask[q];
      if ((any & 0xffu) == 0x5au && rep == 1000) acc += any;
    }
  }
#pragma unroll
  for (int q = 0; q < R; q++) acc += mask[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

template <typename F> float time_ms(F f) {
  cudaEvent_t a, b; CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
  f(); f(); CHECK(cudaDeviceSynchronize());
  CHECK(cudaEventRecord(a)); for (int i = 0; i < 5; i++) f(); CHECK(cudaEventRecord(b)); CHECK(cudaEventSynchronize(b));
  float ms; CHECK(cudaEventElapsedTime(&ms, a, b)); return ms / 5;
}
template <int R, bool P, int U>
void run(const char* name, int ctas, float* out, const float4* tab, const float* dirs, int sms, double hz) {
  float ms = time_ms([&] { k<R, P, U><<<sms * ctas, 256>>>(out, tab, dirs); });
  const double warps_per_smsp = ctas * 8 / 4.0;
  const double tests = warps_per_smsp * REPS * (double)NS * R;        // test-warps per SMSP
  printf("{\"variant\": \"%s\", \"R\": %d, \"unroll_spheres\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"smsp_cycles_per_test_warp\": %.2f}\n",
         name, R, U, ctas, ms, ms * 1e-3 * hz / tests);
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
  run<8, true, 8>("packed FFMA2", 2, out, tab, dirs, sms, hz);
  run<8, false, 8>("scalar FFMA", 2, out, tab, dirs, sms, hz);
  run<8, false, 4>("scalar FFMA", 2, out, tab, dirs, sms, hz);
  run<8, false, 8>("scalar FFMA", 3, out, tab, dirs, sms, hz);
  run<12, false, 4>("scalar FFMA", 2, out, tab, dirs, sms, hz);
  run<6, false, 8>("scalar FFMA", 3, out, tab, dirs, sms, hz);
  run<4, false, 8>("scalar FFMA", 4, out, tab, dirs, sms, hz);
  run<16, false, 4>("scalar FFMA", 1, out, tab, dirs, sms, hz);
  return 0;
}
