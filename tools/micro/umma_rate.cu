// Micro-probe: cycles per tcgen05.mma (M = 128, cta_group::1) on sm_100a, issued back to back by one thread, as the
// discriminator kernels issue them -- kind::tf32 (K = 8) and kind::f16 with bf16 operands (K = 16), A operand in shared
// memory (SS) or in tensor memory (TS), N = 128 or 256, accumulating into the same TMEM block for `run` consecutive MMAs.
// Operand contents are irrelevant (shared memory is zero-filled).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}
// c F32, a / b format `fmt` (0 f16, 1 bf16, 2 tf32), K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc(int fmt, int m, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <bool TF32, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t id, uint32_t acc) {
  if (TF32 && TS)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(id), "r"(acc) : "memory");
  else if (TF32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(id), "r"(acc) : "memory");
  else if (TS)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(id), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(id), "r"(acc) : "memory");
}

template <bool TF32, bool TS, int N>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int n_mma, int run) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t id = idesc(TF32 ? 2 : 1, 128, N);
    // one MMA reads 32 bytes of K: two 16-byte core-matrix columns `lbo` apart; 8-row groups `sbo` apart
    const uint32_t lbo = N * 16, sbo = 128;
    const uint32_t b0 = smem_u32(smem), a0 = smem_u32(smem + 32 * 1024);
    uint64_t bd[4], ad[4];
    uint32_t at[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                                     // four k-steps of a 32 KB stage, built once
      bd[j] = smem_desc(b0 + (uint32_t)j * 2 * lbo, lbo, sbo);
      ad[j] = smem_desc(a0 + (uint32_t)j * 2 * 2048, 2048, 128);
      at[j] = tmem + 384 + 8 * j;
    }
    const int nblk = 384 / N;
    int blk = 0;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 12) {                             // `run` = 12 or 1: accumulate flag pattern only
      const uint32_t d = tmem + (uint32_t)(blk * N);
      blk = blk + 1 == nblk ? 0 : blk + 1;
#pragma unroll
      for (int j = 0; j < 12; ++j) mma<TF32, TS>(d, at[j & 3], ad[j & 3], bd[j & 3], id, (run == 1 || j == 0) ? 0u : 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[2 * gridDim.x + blockIdx.x] = (long long)(g1 - g0);       // nanoseconds: SM clock = cycles / ns
    out[blockIdx.x * 2] = t1 - t0;          // cycles the issuing thread needed (blocked when the MMA queue is full)
    out[blockIdx.x * 2 + 1] = t2 - t0;      // until the last MMA has completed
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <bool TF32, bool TS, int N>
void run(const char* name, int grid, int n_mma, int run_len) {
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 3 * grid);
  auto k = probe<TF32, TS, N>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, 64 * 1024>>>(out, n_mma, run_len);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[3 * 148];
  cudaMemcpy(h, out, sizeof(long long) * 3 * grid, cudaMemcpyDeviceToHost);
  double issue = 0, total = 0, ns = 0;
  for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; ns += h[2 * grid + i]; }
  const double k_per = TF32 ? 8 : 16;
  const double cyc = total / grid / n_mma;
  printf("%-20s grid %3d run %2d: %6.1f cycles / MMA (issue loop %6.1f) = %5.0f MAC/clk/SM at %4.0f MHz = %5.0f TFLOP/s  %s\n", name,
         grid, run_len, cyc, issue / grid / n_mma, 128.0 * N * k_per / cyc, total / ns * 1e3,
         2.0 * 128.0 * N * k_per * n_mma * grid / (ns / grid) * 1e-3, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  for (int grid : {1, 148}) {
    run<true, true, 128>("tf32 K=8  TS N=128", grid, 4092, 12);
    run<true, false, 128>("tf32 K=8  SS N=128", grid, 4092, 12);
    run<true, true, 256>("tf32 K=8  TS N=256", grid, 4092, 12);
    run<true, false, 256>("tf32 K=8  SS N=256", grid, 4092, 12);
    run<false, true, 128>("bf16 K=16 TS N=128", grid, 4092, 12);
    run<false, false, 128>("bf16 K=16 SS N=128", grid, 4092, 12);
    run<false, true, 256>("bf16 K=16 TS N=256", grid, 4092, 12);
    run<true, true, 128>("tf32 K=8  TS N=128", grid, 4092, 1);
    run<true, true, 128>("tf32 K=8  TS N=128", grid, 40920, 12);       // 1.4 ms: long enough for the clock to settle
  }
  return 0;
}
