// sweep_micro.cu — what does the wavefront sweep's instruction mix cost on this SM?  One "pair body" =
// 2 LDS.128 (sphere pair) + R x 3 FFMA2 (scalar ray operand x packed pair + packed acc) + R x 2 SHF.
// Variants switch the ingredients on one at a time; output = SMSP cycles per pair body per warp-slot.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sweep_micro tools/sweep_micro.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NPAIR = 512;   // 1024 spheres
constexpr int REPS = 16;

// MODE bit0: duplicate ray directions into float2 pairs (no .F32 broadcast operand)
//      bit1: 2 SHF per pair per ray (else 1 LOP3)      bit2: only 2 FFMA2 per pair per ray (marginal cost)
//      bit3: no mask op at all (results folded by one LOP3 per body)
template <int R, int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float4* gtab, const float* dirs) {
  __shared__ float4 tab[2 * NPAIR];
  for (int i = threadIdx.x; i < 2 * NPAIR; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  float2 nd[R][3];
  unsigned mask[R];
#pragma unroll
  for (int q = 0; q < R; q++) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float v = dirs[q * 3 + c] + threadIdx.x * 1e-6f;
      nd[q][c] = (MODE & 1) ? make_float2(v, v * 1.0000001f) : make_float2(v, v);
    }
    mask[q] = 0;
  }
  unsigned acc = 0;
  for (int rep = 0; rep < REPS; rep++) {
    for (int j = 0; j < NPAIR; j += 4) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float4 A = tab[2 * (j + u)], B = tab[2 * (j + u) + 1];
#pragma unroll
        for (int q = 0; q < R; q++) {
          float2 t = __ffma2_rn(nd[q][2], make_float2(B.x, B.y), make_float2(B.z, B.w));
          t = __ffma2_rn(nd[q][1], make_float2(A.z, A.w), t);
          if (!(MODE & 4)) t = __ffma2_rn(nd[q][0], make_float2(A.x, A.y), t);
          if (MODE & 8) { if (q == R - 1) mask[0] |= __float_as_uint(t.x); else mask[q + 1] = __float_as_uint(t.y) ^ mask[q + 1]; }
          else if (MODE & 2) {
            mask[q] = __funnelshift_l(__float_as_uint(t.x), mask[q], 1);
            mask[q] = __funnelshift_l(__float_as_uint(t.y), mask[q], 1);
          } else {
            mask[q] |= __float_as_uint(t.x) | __float_as_uint(t.y);
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

template <int R, int MODE>
void run(const char* name, int ctas_per_sm, float* out, const float4* tab, const float* dirs, int sms, double hz) {
  const int blocks = sms * ctas_per_sm;
  float ms = time_ms([&] { k<R, MODE><<<blocks, 256>>>(out, tab, dirs); });
  const double warps_per_smsp = ctas_per_sm * 8 / 4.0;
  const double bodies = warps_per_smsp * REPS * NPAIR;
  const double cyc = ms * 1e-3 * hz / bodies;
  const int nf = (MODE & 4) ? 2 : 3;
  printf("{\"variant\": \"%s\", \"R\": %d, \"ms\": %.3f, \"smsp_cycles_per_pair_body\": %.1f, \"cycles_per_ffma2\": %.2f}\n",
         name, R, ms, cyc, cyc / (R * nf));
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double hz = khz * 1e3;
  float4* tab; float* dirs; float* out;
  CHECK(cudaMalloc(&tab, 2 * NPAIR * sizeof(float4)));
  CHECK(cudaMalloc(&dirs, 64 * 4));
  CHECK(cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4));
  static float4 h[2 * NPAIR]; for (int i = 0; i < 2 * NPAIR; i++) h[i] = make_float4(1.f + i, 2.f, -3.f + i, 1e30f);
  float hd[64]; for (int i = 0; i < 64; i++) hd[i] = 0.01f * i;
  CHECK(cudaMemcpy(tab, h, sizeof h, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dirs, hd, sizeof hd, cudaMemcpyHostToDevice));
  const int sms = p.multiProcessorCount;
  const int c = 2;
  run<8, 8>("scalar dirs, 3 ffma2, no mask ops", c, out, tab, dirs, sms, hz);
  run<8, 8 | 4>("scalar dirs, 2 ffma2, no mask ops", c, out, tab, dirs, sms, hz);
  run<8, 8 | 1>("pair dirs, 3 ffma2, no mask ops", c, out, tab, dirs, sms, hz);
  run<8, 0>("scalar dirs, 3 ffma2, LOP3", c, out, tab, dirs, sms, hz);
  run<8, 1>("pair dirs, 3 ffma2, LOP3", c, out, tab, dirs, sms, hz);
  run<8, 2>("scalar dirs, 3 ffma2, 2 SHF (kernel mix)", c, out, tab, dirs, sms, hz);
  run<8, 3>("pair dirs, 3 ffma2, 2 SHF", c, out, tab, dirs, sms, hz);
  run<4, 2>("scalar dirs, R=4, 2 SHF", c, out, tab, dirs, sms, hz);
  run<4, 3>("pair dirs, R=4, 2 SHF", c, out, tab, dirs, sms, hz);
  run<4, 1>("pair dirs, R=4, LOP3", 4, out, tab, dirs, sms, hz);
  return 0;
}
