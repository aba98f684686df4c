// K4: GAIL / VAIL discriminator reward  r = -log(1 - sigmoid(D(s)) + 1e-8)
//   make_discrim_reward   imitation_lib/imitation/gail_TRPO.py:320-327   (discrim_output :315-318, VAIL vail_TRPO.py:18-21)
//   VariationalNet.forward imitation_lib/utils/networks.py:258-284 (reparameterize :21-24), DiscriminatorNetwork.forward
//   :208-234, FullyConnectedNetwork :94-158, Standardizer.forward :68-74; shapes examples/imitation_learning/utils.py:151-179
//   + confs.yaml:113-130:  VAIL 32 -relu-> 256 -relu-> 128 -> (mu, logvar)[128+128] -> z -> 1 ;  GAIL 32 -tanh-> 512 -tanh-> 256 -> 1.
//
// The only dense contraction on the hot path, so the only tcgen05 kernel.  One CTA (128 threads, 1 per SM) owns a
// tile of 128 samples: TMEM lane = sample, TMEM column = layer output, i.e. thread t of the CTA owns sample t in every
// epilogue and the final 128->1 / 256->1 head is a per-thread dot product.  Each layer is D[128 x N] += A[128 x K] B[N x K]^T
// on tcgen05.mma kind::tf32 (M=128, N<=256, K=8), streamed in K-chunks of 32:
//   * A chunk (activations of the previous layer): tcgen05.ld from TMEM -> bias + activation -> split -> st.shared in the
//     canonical K-major no-swizzle UMMA layout [k/4][row][4]  (16-byte rows contiguous: conflict-free stores);
//   * B chunk (weights): a pre-split image in exactly that layout, fetched by ONE cp.async.bulk (TMA, 1-D) per chunk
//     onto an mbarrier; two stages so that the copy and the next A chunk overlap the MMAs of the current one;
//   * fp32 fidelity: every product is evaluated as 3xTF32 (a_hi b_hi + a_lo b_hi + a_hi b_lo, hi = cvt.rna.tf32), fp32
//     accumulation in TMEM -- the north star's 1e-5 tolerance on rewards rules out plain TF32 (~1e-3).
#include <cstring>
#include <vector>

#include "om_common.cuh"

namespace om {

constexpr int DISC_IN = 32;        // observation size of the H1 discriminators
constexpr int KC = 32;             // K-chunk (elements) = 4 MMA k-steps of 8
constexpr int TILE = 128;          // samples per CTA tile = UMMA_M
constexpr int STAGE_A_BYTES = TILE * KC * 4 * 2;      // hi + lo
constexpr int STAGE_B_BYTES = 256 * KC * 4 * 2;       // up to N = 256 rows, hi + lo
constexpr int DISC_MAX_PAR = 512 + 256 + 256 + 256;   // biases + head weights staged in shared memory

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded spin: a mis-programmed descriptor or copy must surface as a launch failure, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row x 16-byte core matrices,
// `lbo` = byte stride between the two 16-byte K slices of one MMA, `sbo` = byte stride between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);      // version 1 (sm_100), base offset 0, layout type 0 = SWIZZLE_NONE
}
// cute::UMMA::InstrDescriptor: c F32 (1<<4), a/b TF32 (2<<7, 2<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------- weight image (built on the host at create time)
// Chunk order = issue order of the kernel: for each 256-wide block nb of layer 1: [L1(nb)] then N1blk/32 chunks of layer 2;
// then (VAIL) N2/32 chunks of the [mu; logvar] layer.  Each chunk: hi image [8][rows][4] floats, then lo image.
struct DiscShape {
  int kind, n1, n2, z;          // kind 0 VAIL, 1 GAIL
  __host__ __device__ int n1_blocks() const { return n1 / 256; }
  __host__ __device__ int chunks_per_tile() const { return n1_blocks() * (1 + 8) + (kind == 0 ? n2 / KC : 0); }
};

struct DiscArgs {
  DiscShape sh;
  const float* image;           // pre-split weight chunks
  const float* params;          // b1[n1], b2[n2], (VAIL: bmu|blv [2z]), head weights wd[z or n2], bd
  const float* s; const float* mean; const float* stdv; const float* eps;
  float* reward; float* d_out;
  int n, ld;
};

enum { ACT_RELU = 0, ACT_TANH = 1 };
template <int ACT> __device__ __forceinline__ float act(float x) { return ACT == ACT_RELU ? fmaxf(x, 0.f) : tanhf(x); }

// split 32 activations of this thread's sample into the hi / lo A-stage images
__device__ __forceinline__ void store_a_chunk(uint8_t* stage, int row, const float (&a)[32]) {
  float4* hi = reinterpret_cast<float4*>(stage) + row;                       // [kc][128 rows] float4
  float4* lo = reinterpret_cast<float4*>(stage + TILE * KC * 4) + row;
#pragma unroll
  for (int kc = 0; kc < 8; ++kc) {
    float4 h, l;
    h.x = tf32_rna(a[4 * kc]); h.y = tf32_rna(a[4 * kc + 1]); h.z = tf32_rna(a[4 * kc + 2]); h.w = tf32_rna(a[4 * kc + 3]);
    l.x = a[4 * kc] - h.x; l.y = a[4 * kc + 1] - h.y; l.z = a[4 * kc + 2] - h.z; l.w = a[4 * kc + 3] - h.w;
    hi[kc * TILE] = h;
    lo[kc * TILE] = l;
  }
}

template <int N1, int N2, bool VAIL>
__global__ void __launch_bounds__(128, 1) disc_reward_kernel(DiscArgs a) {
  constexpr int ACT = VAIL ? ACT_RELU : ACT_TANH;
  constexpr int NB1 = N1 / 256;                 // 256-wide column blocks of layer 1
  constexpr int Z = 128;
  constexpr int N3 = 2 * Z;
  constexpr int D1_COL = 0, D2_COL = 256, D3_COL = 0;
  constexpr int HEAD_K = VAIL ? Z : N2;
  static_assert(N1 % 256 == 0 && N2 % 32 == 0 && N2 <= 256, "unsupported discriminator shape");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stA[2] = {smem, smem + STAGE_A_BYTES};
  uint8_t* stB[2] = {smem + 2 * STAGE_A_BYTES, smem + 2 * STAGE_A_BYTES + STAGE_B_BYTES};
  float* par = reinterpret_cast<float*>(smem + 2 * STAGE_A_BYTES + 2 * STAGE_B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(par + DISC_MAX_PAR + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const float* b1 = par;
  const float* b2 = par + N1;
  const float* b3 = par + N1 + N2;                          // VAIL: bmu | blv
  const float* wd = par + N1 + N2 + (VAIL ? N3 : 0);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_full[2] = {smem_u32(bars), smem_u32(bars + 1)};
  const uint32_t bar_done[2] = {smem_u32(bars + 2), smem_u32(bars + 3)};

  constexpr int NPAR = N1 + N2 + (VAIL ? N3 : 0) + HEAD_K + 1;
  for (int i = tid; i < NPAR; i += 128) par[i] = a.params[i];
  if (tid == 0) {
    mbar_init(bar_full[0], 1); mbar_init(bar_full[1], 1); mbar_init(bar_done[0], 1); mbar_init(bar_done[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);     // this warp's 32 TMEM lanes
  const float bd = par[NPAR - 1];

  const int ntiles = (a.n + TILE - 1) / TILE;
  int g = 0;                     // global chunk counter of this CTA (stage = g & 1)
  // commit of global chunk h has completed  <=>  phase (h >> 1) of bar_done[h & 1] has completed
  auto wait_chunk = [&](int h) {
    if (h >= 0) {
      mbar_wait(bar_done[h & 1], (uint32_t)(h >> 1) & 1u);
      tc_fence_after();
    }
  };

  // One K-chunk: (1) the stage is free once chunk g-2 retired; (2) thread 0 starts the weight copy; (3) every thread
  // writes its row of the A chunk; (4) thread 0 issues 4 k-steps x 3 products and commits.
  auto run_chunk = [&](const float (&act_in)[32], const float* img, int rows, uint32_t d_col, bool first) {
    const int st = g & 1;
    wait_chunk(g - 2);
    const uint32_t bytes = (uint32_t)rows * KC * 4 * 2;
    if (tid == 0) {
      mbar_expect_tx(bar_full[st], bytes);
      bulk_g2s(smem_u32(stB[st]), img, bytes, bar_full[st]);
    }
    store_a_chunk(stA[st], tid, act_in);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_full[st], (uint32_t)(g >> 1) & 1u);
      const uint32_t a_hi = smem_u32(stA[st]), a_lo = a_hi + TILE * KC * 4;
      const uint32_t b_hi = smem_u32(stB[st]), b_lo = b_hi + (uint32_t)rows * KC * 4;
      const uint32_t a_lbo = TILE * 16, b_lbo = (uint32_t)rows * 16, sbo = 128;
      const uint32_t idesc = idesc_tf32(TILE, rows);
#pragma unroll
      for (int j = 0; j < KC / 8; ++j) {
        const uint64_t dah = smem_desc(a_hi + j * 2 * a_lbo, a_lbo, sbo), dal = smem_desc(a_lo + j * 2 * a_lbo, a_lbo, sbo);
        const uint64_t dbh = smem_desc(b_hi + j * 2 * b_lbo, b_lbo, sbo), dbl = smem_desc(b_lo + j * 2 * b_lbo, b_lbo, sbo);
        umma_tf32(tmem + d_col, dal, dbh, idesc, (first && j == 0) ? 0u : 1u);      // small terms first
        umma_tf32(tmem + d_col, dah, dbl, idesc, 1u);
        umma_tf32(tmem + d_col, dah, dbh, idesc, 1u);
      }
      umma_commit(bar_done[st]);
    }
    ++g;
  };

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * TILE + tid;
    const bool live = env < a.n;
    const float* img = a.image;
    // ---- standardised input row (Standardizer.forward networks.py:73-74 with a frozen snapshot)
    float x[DISC_IN];
#pragma unroll
    for (int k = 0; k < DISC_IN; ++k)
      x[k] = live ? (a.s[(size_t)k * a.ld + env] - __ldg(a.mean + k)) / __ldg(a.stdv + k) : 0.f;
#pragma unroll 1
    for (int nb = 0; nb < NB1; ++nb) {
      run_chunk(x, img, 256, D1_COL, true);                                        // layer 1, column block nb
      img += 256 * KC * 2;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {                                                // layer 2, K-chunks fed by this block
        if (c == 0) wait_chunk(g - 1);                                             // D1 block complete
        float h[32];
        tmem_ld32(lane_addr + D1_COL + c * KC, h);
#pragma unroll
        for (int i = 0; i < 32; ++i) h[i] = act<ACT>(h[i] + b1[nb * 256 + c * KC + i]);
        run_chunk(h, img, N2, D2_COL, nb == 0 && c == 0);
        img += N2 * KC * 2;
      }
    }
    float dval = 0.f;
    if (VAIL) {
#pragma unroll 1
      for (int c = 0; c < N2 / KC; ++c) {                                          // [mu; logvar] layer
        if (c == 0) wait_chunk(g - 1);                                             // D2 complete
        float h[32];
        tmem_ld32(lane_addr + D2_COL + c * KC, h);
#pragma unroll
        for (int i = 0; i < 32; ++i) h[i] = act<ACT>(h[i] + b2[c * KC + i]);
        run_chunk(h, img, N3, D3_COL, c == 0);
        img += N3 * KC * 2;
      }
      wait_chunk(g - 1);
      // z = mu + exp(logvar / 2) * eps (networks.py:21-24), d = wd . z + bd
#pragma unroll 1
      for (int c = 0; c < Z / 32; ++c) {
        float mu[32], lv[32];
        tmem_ld32(lane_addr + D3_COL + c * 32, mu);
        tmem_ld32(lane_addr + D3_COL + Z + c * 32, lv);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int j = c * 32 + i;
          const float e = (a.eps && live) ? a.eps[(size_t)j * a.ld + env] : 0.f;
          const float zz = fmaf(expf(0.5f * (lv[i] + b3[Z + j])), e, mu[i] + b3[j]);
          dval = fmaf(wd[j], zz, dval);
        }
      }
    } else {
      wait_chunk(g - 1);
#pragma unroll 1
      for (int c = 0; c < N2 / 32; ++c) {
        float h[32];
        tmem_ld32(lane_addr + D2_COL + c * 32, h);
#pragma unroll
        for (int i = 0; i < 32; ++i) dval = fmaf(wd[c * 32 + i], act<ACT>(h[i] + b2[c * 32 + i]), dval);
      }
    }
    dval += bd;
    if (live) {
      // 1 - sigmoid(d) evaluated as sigmoid(-d): no cancellation for large d (gail_TRPO.py:326-327)
      const float one_minus_p = 1.f / (1.f + expf(dval));
      a.reward[env] = -logf(one_minus_p + 1e-8f);
      if (a.d_out) a.d_out[env] = dval;
    }
    // the next tile's layer-1 MMA overwrites the columns read above: order the TMEM loads before it
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace om

using namespace om;

struct OmDisc {
  DiscShape sh;
  float* image = nullptr;
  float* params = nullptr;
};

static void split_tf32(float x, float* hi, float* lo) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;                 // cvt.rna.tf32.f32: nearest, ties away from zero
  std::memcpy(hi, &u, 4);
  *lo = x - *hi;
}

// one chunk: rows [r0, r0+rows) x columns [k0, k0+32) of a row-major [*, ldw] matrix -> hi image then lo image
static void append_chunk(std::vector<float>& img, const float* w, int ldw, int r0, int rows, int k0) {
  const size_t base = img.size();
  img.resize(base + (size_t)rows * KC * 2);
  float* hi = img.data() + base;
  float* lo = hi + (size_t)rows * KC;
  for (int kc = 0; kc < 8; ++kc)
    for (int r = 0; r < rows; ++r)
      for (int e = 0; e < 4; ++e)
        split_tf32(w[(size_t)(r0 + r) * ldw + k0 + kc * 4 + e], hi + ((size_t)kc * rows + r) * 4 + e, lo + ((size_t)kc * rows + r) * 4 + e);
}

extern "C" int om_disc_create(const OmDiscDesc* d, OmDisc** out) {
  OM_REQUIRE(d && out, "om_disc_create: null argument");
  OM_REQUIRE(d->n_in == DISC_IN, "om_disc_create: n_in must be %d (the H1 observation), got %d", DISC_IN, d->n_in);
  const bool vail = d->kind == 0;
  OM_REQUIRE(d->kind == 0 || d->kind == 1, "om_disc_create: kind must be 0 (VAIL) or 1 (GAIL)");
  OM_REQUIRE((vail && d->n_h1 == 256 && d->n_h2 == 128 && d->z_size == 128) || (!vail && d->n_h1 == 512 && d->n_h2 == 256),
             "om_disc_create: only the reference shapes are built (VAIL 32-256-128-z128-1, GAIL 32-512-256-1)");
  OM_REQUIRE(d->w1 && d->b1 && d->w2 && d->b2 && d->wd && d->bd, "om_disc_create: null weights");
  OM_REQUIRE(!vail || (d->wmu && d->bmu && d->wlv && d->blv), "om_disc_create: VAIL needs mu / logvar layers");
  const int n1 = d->n_h1, n2 = d->n_h2, z = vail ? d->z_size : 0;
  std::vector<float> img;
  for (int nb = 0; nb < n1 / 256; ++nb) {
    append_chunk(img, d->w1, DISC_IN, nb * 256, 256, 0);
    for (int c = 0; c < 8; ++c) append_chunk(img, d->w2, n1, 0, n2, nb * 256 + c * KC);
  }
  if (vail) {
    std::vector<float> w3((size_t)2 * z * n2);                 // [mu; logvar] stacked: 2z rows x n2
    std::memcpy(w3.data(), d->wmu, sizeof(float) * z * n2);
    std::memcpy(w3.data() + (size_t)z * n2, d->wlv, sizeof(float) * z * n2);
    for (int c = 0; c < n2 / KC; ++c) append_chunk(img, w3.data(), n2, 0, 2 * z, c * KC);
  }
  std::vector<float> par;
  par.insert(par.end(), d->b1, d->b1 + n1);
  par.insert(par.end(), d->b2, d->b2 + n2);
  if (vail) {
    par.insert(par.end(), d->bmu, d->bmu + z);
    par.insert(par.end(), d->blv, d->blv + z);
  }
  par.insert(par.end(), d->wd, d->wd + (vail ? z : n2));
  par.push_back(d->bd[0]);
  OmDisc* h = new OmDisc();
  h->sh = DiscShape{d->kind, n1, n2, z};
  cudaError_t e = cudaMalloc(&h->image, img.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->params, par.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(h->image, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->params, par.data(), par.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (h->image) cudaFree(h->image);
    if (h->params) cudaFree(h->params);
    delete h;
    return fail("om_disc_create: device upload failed: %s (no CPU path)", cudaGetErrorString(e));
  }
  *out = h;
  return 0;
}

extern "C" void om_disc_destroy(OmDisc* h) {
  if (!h) return;
  cudaFree(h->image);
  cudaFree(h->params);
  delete h;
}

extern "C" int om_disc_reward(const OmDisc* h, const float* s, const float* mean, const float* stdv, const float* eps, int n,
                              int ld, float* reward, float* d_out, void* stream) {
  OM_REQUIRE(h, "om_disc_reward: null discriminator");
  OM_REQUIRE(n >= 0 && ld >= n, "om_disc_reward: need 0 <= n <= ld");
  if (n == 0) return 0;
  OM_REQUIRE(s && mean && stdv && reward, "om_disc_reward: null argument");
  int dev = 0, sms = 0;
  OM_CUDA_OK(cudaGetDevice(&dev));
  OM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntiles = ceil_div(n, TILE);
  const int grid = ntiles < sms ? ntiles : sms;                  // persistent: one CTA per SM
  const size_t smem = 2 * STAGE_A_BYTES + 2 * STAGE_B_BYTES + (DISC_MAX_PAR + 4) * sizeof(float) + 4 * 8 + 16;
  DiscArgs a{h->sh, h->image, h->params, s, mean, stdv, eps, reward, d_out, n, ld};
  cudaStream_t st = (cudaStream_t)stream;
  if (h->sh.kind == 0) {
    OM_CUDA_OK(cudaFuncSetAttribute(disc_reward_kernel<256, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    disc_reward_kernel<256, 128, true><<<grid, 128, smem, st>>>(a);
  } else {
    OM_CUDA_OK(cudaFuncSetAttribute(disc_reward_kernel<512, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    disc_reward_kernel<512, 256, false><<<grid, 128, smem, st>>>(a);
  }
  OM_LAUNCHED();
  return 0;
}
