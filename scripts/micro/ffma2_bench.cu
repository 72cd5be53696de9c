// Microbenchmark (B200): is an issue-bound FP32 kernel helped by the packed FFMA2 / FADD2 / FMUL2 instructions of sm_100?
// Same number of floating-point operations three ways: scalar FFMA, packed FFMA2, and scalar FFMA mixed with the integer /
// select work that keeps the issue port busy in the solve kernel.   nvcc -arch=sm_100a -O3 ffma2_bench.cu -o ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_scalar(float* out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(float* out, float a, float b) {
  float2 x[4];
  for (int i = 0; i < 4; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
  float s = 0;
  for (int i = 0; i < 4; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 FP ops + 8 ALU ops (min/max: alu pipe) per iteration, scalar vs packed FP
__global__ void k_mixed_scalar(float* out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fminf(fmaf(x[i], a, b), 1e30f);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mixed_packed(float* out, float a, float b) {
  float2 x[4];
  for (int i = 0; i < 4; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = __ffma2_rn(x[i], a2, b2); x[i].x = fminf(x[i].x, 1e30f); x[i].y = fminf(x[i].y, 1e30f); }
  float s = 0;
  for (int i = 0; i < 4; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float run(K k, float* d, int blocks, int threads) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(d, 0.999f, 0.001f);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<<<blocks, threads>>>(d, 0.999f, 0.001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  const int blocks = 148 * 8, threads = 256;
  float* d;
  cudaMalloc(&d, blocks * threads * sizeof(float));
  const double flops = 2.0 * 8 * ITERS * (double)blocks * threads;
  float a = run(k_scalar, d, blocks, threads), b = run(k_packed, d, blocks, threads);
  float c = run(k_mixed_scalar, d, blocks, threads), e = run(k_mixed_packed, d, blocks, threads);
  printf("{\"scalar_ffma_ms\": %.4f, \"packed_ffma2_ms\": %.4f, \"scalar_tflops\": %.1f, \"packed_tflops\": %.1f, "
         "\"mixed_scalar_ms\": %.4f, \"mixed_packed_ms\": %.4f}\n", a, b, flops / a / 1e9, flops / b / 1e9, c, e);
  return 0;
}
