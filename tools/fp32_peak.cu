// fp32_peak.cu — measured FP32 FMA throughput of the device, scalar FFMA vs packed FFMA2 (sm_100
// f32x2), with operands laid out the way the sweep kernels use them.  The roofline denominator in
// bench.py is the nominal SMs x 128 x 2 x f; this tool tells how close real instruction streams get.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp32_peak tools/fp32_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int ACC = 16;

// independent accumulator chains: acc[i] = acc[i] * b + c   (2 distinct extra registers)
__global__ void __launch_bounds__(256) k_ffma(float* out, float b, float c) {
  float acc[ACC];
#pragma unroll
  for (int i = 0; i < ACC; i++) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = fmaf(acc[i], b, c);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dot-product shape of the sweep: acc[i] = d * s[i] + acc[i], s[i] varying registers
__global__ void __launch_bounds__(256) k_ffma_dot(float* out, const float* in) {
  float acc[ACC], s[ACC];
  const float d0 = in[0], d1 = in[1], d2 = in[2];
#pragma unroll
  for (int i = 0; i < ACC; i++) { acc[i] = 0.f; s[i] = in[3 + i] + threadIdx.x; }
  for (int it = 0; it < ITERS / 3; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = fmaf(d0, s[i], acc[i]);
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = fmaf(d1, s[(i + 1) % ACC], acc[i]);
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = fmaf(d2, s[(i + 2) % ACC], acc[i]);
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < ACC; i++) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float b, float c) {
  float2 acc[ACC];
  const float2 bb = make_float2(b, b * 1.0001f), cc = make_float2(c, c * 0.999f);
#pragma unroll
  for (int i = 0; i < ACC; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = __ffma2_rn(acc[i], bb, cc);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ACC; i++) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma2_dot(float* out, const float* in) {
  float2 acc[ACC], s[ACC];
  const float2 d0 = make_float2(in[0], in[0]), d1 = make_float2(in[1], in[1]), d2 = make_float2(in[2], in[2]);
#pragma unroll
  for (int i = 0; i < ACC; i++) { acc[i] = make_float2(0.f, 0.f); s[i] = make_float2(in[3 + i] + threadIdx.x, in[4 + i]); }
  for (int it = 0; it < ITERS / 3; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = __ffma2_rn(d0, s[i], acc[i]);
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = __ffma2_rn(d1, s[(i + 1) % ACC], acc[i]);
#pragma unroll
    for (int i = 0; i < ACC; i++) acc[i] = __ffma2_rn(d2, s[(i + 2) % ACC], acc[i]);
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < ACC; i++) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
float time_ms(F launch) {
  cudaEvent_t a, b;
  CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
  for (int i = 0; i < 3; i++) launch();
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaEventRecord(a));
  for (int i = 0; i < 10; i++) launch();
  CHECK(cudaEventRecord(b));
  CHECK(cudaEventSynchronize(b));
  float ms; CHECK(cudaEventElapsedTime(&ms, a, b));
  return ms / 10;
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  float *out, *in;
  CHECK(cudaMalloc(&out, (size_t)blocks * threads * 4));
  CHECK(cudaMalloc(&in, 64 * 4));
  float h[64]; for (int i = 0; i < 64; i++) h[i] = 0.5f + i * 0.01f;
  CHECK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
  const double lanes = (double)blocks * threads;
  const double nominal = p.multiProcessorCount * 128.0 * 2 * khz * 1e3 / 1e12;
  double f;
  float ms;
  ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  f = lanes * ITERS * ACC * 2 / (ms * 1e-3) / 1e12;
  printf("{\"kernel\": \"ffma scalar (acc*b+c)\", \"ms\": %.4f, \"tflops\": %.2f, \"frac_nominal\": %.3f}\n", ms, f, f / nominal);
  ms = time_ms([&] { k_ffma_dot<<<blocks, threads>>>(out, in); });
  f = lanes * (ITERS / 3 * 3) * ACC * 2 / (ms * 1e-3) / 1e12;
  printf("{\"kernel\": \"ffma scalar dot-shape (d*s+acc)\", \"ms\": %.4f, \"tflops\": %.2f, \"frac_nominal\": %.3f}\n", ms, f, f / nominal);
  ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  f = lanes * ITERS * ACC * 4 / (ms * 1e-3) / 1e12;
  printf("{\"kernel\": \"ffma2 packed (acc*b+c)\", \"ms\": %.4f, \"tflops\": %.2f, \"frac_nominal\": %.3f}\n", ms, f, f / nominal);
  ms = time_ms([&] { k_ffma2_dot<<<blocks, threads>>>(out, in); });
  f = lanes * (ITERS / 3 * 3) * ACC * 4 / (ms * 1e-3) / 1e12;
  printf("{\"kernel\": \"ffma2 packed dot-shape (d*s+acc)\", \"ms\": %.4f, \"tflops\": %.2f, \"frac_nominal\": %.3f}\n", ms, f, f / nominal);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"nominal_fp32_tflops\": %.2f}\n", p.name, p.multiProcessorCount, khz, nominal);
  return 0;
}
