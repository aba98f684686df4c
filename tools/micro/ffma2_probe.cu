// Micro-probe: FP32 FMA throughput of scalar FFMA (3-register form) vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_scalar(float* out, float a, float b, int iters) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  float x = a + threadIdx.x * 1e-7f, y = b + threadIdx.x * 1e-9f;   // both operands in registers: 3-register FFMA
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_packed(float* out, float a, float b, int iters) {
  float2 acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  float2 x = make_float2(a + threadIdx.x * 1e-7f, a - threadIdx.x * 1e-7f), y = make_float2(b, b * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = __ffma2_rn(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 8, block = 256, iters = 8192;
  float* out;
  cudaMalloc(&out, sizeof(float) * grid * block);
  const double threads = (double)grid * block;
  float ms_s = time_ms([&] { k_scalar<8><<<grid, block>>>(out, 0.999f, 0.001f, iters); });
  float ms_p = time_ms([&] { k_packed<8><<<grid, block>>>(out, 0.999f, 0.001f, iters); });
  float ms_p4 = time_ms([&] { k_packed<4><<<grid, block>>>(out, 0.999f, 0.001f, iters); });
  const double fma_s = threads * 8 * iters, fma_p = threads * 16 * iters, fma_p4 = threads * 8 * iters;
  printf("SMs %d\n", sms);
  printf("scalar FFMA   ILP 8 : %.3f ms  %.1f TFLOP/s (2 flop/FMA)\n", ms_s, 2 * fma_s / ms_s / 1e9);
  printf("packed FFMA2  ILP 8 : %.3f ms  %.1f TFLOP/s\n", ms_p, 2 * fma_p / ms_p / 1e9);
  printf("packed FFMA2  ILP 4 : %.3f ms  %.1f TFLOP/s\n", ms_p4, 2 * fma_p4 / ms_p4 / 1e9);
  return 0;
}
