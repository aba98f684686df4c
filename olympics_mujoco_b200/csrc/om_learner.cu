// K5 (returns / GAE reverse scans over a time-major rollout buffer) and K6 (moment partial sums).
#include <cstdlib>

#include "om_common.cuh"

namespace om {

// ---------------------------------------------------------------- K5a: PPO discounted returns
// R_t = r_t + gamma * R_{t+1}, restarted with the path's bootstrap wherever a path ends
// (rl/algos/ppo.py:68-84, bootstrap :195-196).  One thread per env; the recurrence is a 1-FMA chain, all
// loads are independent of it (the compiler hoists UNROLL of them ahead), rows are 128-byte coalesced.
template <int UNROLL>
__global__ void __launch_bounds__(128) ppo_returns_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                          const uint8_t* __restrict__ path_end,
                                                          const float* __restrict__ v_next, const float* __restrict__ v_last,
                                                          float gamma, int T, int n, int ld, float* __restrict__ ret,
                                                          float* __restrict__ adv) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  float R = 0.f;
  int t = T - 1;
  while (t >= 0) {
    const int cnt = min(UNROLL, t + 1);
    float r[UNROLL], v[UNROLL], vn[UNROLL];
    uint8_t e[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (u < cnt) {
        const size_t idx = (size_t)(t - u) * ld + env;
        r[u] = rewards[idx];
        v[u] = values ? values[idx] : 0.f;
        e[u] = path_end ? path_end[idx] : 0;
        vn[u] = (v_next && e[u] == 2) ? v_next[idx] : 0.f;
      }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (u < cnt) {
        const int tt = t - u;
        if (tt == T - 1) R = (e[u] == 1) ? 0.f : ((e[u] == 2 && v_next) ? vn[u] : (v_last ? v_last[env] : 0.f));
        else if (e[u] == 1) R = 0.f;
        else if (e[u] == 2) R = vn[u];
        R = fmaf(gamma, R, r[u]);
        const size_t idx = (size_t)tt * ld + env;
        if (ret) ret[idx] = R;
        if (adv) adv[idx] = R - v[u];
      }
    t -= cnt;
  }
}

// ---------------------------------------------------------------- K5b: GAE(lambda)
// mushroom_rl compute_gae (call site imitation_lib/imitation/gail_TRPO.py:126-128).
template <int UNROLL>
__global__ void __launch_bounds__(128) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ v,
                                                  const float* __restrict__ v_next, const uint8_t* __restrict__ absorbing,
                                                  const uint8_t* __restrict__ last, float gamma, float lam, int T, int n,
                                                  int ld, float* __restrict__ adv, float* __restrict__ v_target) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  const float gl = gamma * lam;
  float A = 0.f;
  int t = T - 1;
  while (t >= 0) {
    const int cnt = min(UNROLL, t + 1);
    float r[UNROLL], vv[UNROLL], vn[UNROLL];
    uint8_t ab[UNROLL], la[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (u < cnt) {
        const size_t idx = (size_t)(t - u) * ld + env;
        r[u] = rewards[idx]; vv[u] = v[idx]; vn[u] = v_next[idx];
        ab[u] = absorbing ? absorbing[idx] : 0;
        la[u] = last ? last[idx] : 0;
      }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (u < cnt) {
        const int tt = t - u;
        const bool end = la[u] || tt == T - 1;
        // delta with the bootstrap dropped only for absorbing segment ends
        const float boot = (end && ab[u]) ? 0.f : gamma * vn[u];
        const float delta = (r[u] - vv[u]) + boot;
        A = end ? delta : fmaf(gl, A, delta);
        const size_t idx = (size_t)tt * ld + env;
        if (adv) adv[idx] = A;
        if (v_target) v_target[idx] = A + vv[u];
      }
    t -= cnt;
  }
}

// ---------------------------------------------------------------- K5, scan flavour (T <= 32 * L)
// Both recurrences are affine: X_t = a_t * X_{t+1} + b_t.  One CTA owns 32 envs (lanes) x all T steps: warp w
// holds time segment [w*L, (w+1)*L) in registers, composes it into (A, B) with X_first = A * X_in + B, the
// 32 segment summaries are exchanged through shared memory, every warp folds the summaries of the LATER
// segments into its carry-in and then emits its L outputs.  Every global access is a 128-byte row segment.
// This is the "warp-scan over the rollout buffer" of the north star, laid out so that lanes stay on the
// contiguous env axis; it makes a 4096-env x 500-step buffer 131072 threads instead of 4096.
struct AffineIn {
  // GAE: rewards, v, v_next, absorbing, last.  Returns: rewards, values, path_end, v_next, v_last.
  const float* r; const float* v; const float* vn; const uint8_t* f0; const uint8_t* f1; const float* v_last;
  float gamma, gl;
  int mode;   // 0 = GAE, 1 = PPO returns
};

template <int L>
__global__ void __launch_bounds__(1024) affine_scan_kernel(AffineIn in, int T, int n, int ld, float* __restrict__ out0,
                                                           float* __restrict__ out1) {
  __shared__ float sA[32][33], sB[32][33];
  pdl_wait();
  const int lane = threadIdx.x, w = threadIdx.y;
  const int env = blockIdx.x * 32 + lane;
  const bool live = env < n;
  const int t0 = w * L;
  // a_t is either 0 (a path / segment ends at t) or the constant g: one bit per step instead of a register.  Steps past
  // the end of the buffer get (a, b) = (0, 0): everything behind the last step is zero anyway.  For L = 32 the values
  // V(s_t) are re-read for the output instead of being kept (b[32] + vv[32] spilled 1.5 KB per thread).
  constexpr bool KEEP_V = L <= 16;
  const float g = in.mode == 0 ? in.gl : in.gamma;
  float b[L], vv[KEEP_V ? L : 1];
  uint32_t amask = 0;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const int t = t0 + i;
    b[i] = 0.f;
    if (KEEP_V) vv[i] = 0.f;
    if (live && t < T) {
      const size_t idx = (size_t)t * ld + env;
      const float r = in.r[idx];
      const bool endbuf = t == T - 1;
      if (in.mode == 0) {
        const float v = in.v[idx], vn = in.vn[idx];
        const bool ab = in.f0 && in.f0[idx], la = (in.f1 && in.f1[idx]) || endbuf;
        const float boot = (la && ab) ? 0.f : in.gamma * vn;
        if (!la) amask |= 1u << i;
        b[i] = (r - v) + boot;
        if (KEEP_V) vv[i] = v;
      } else {
        const uint8_t e = in.f0 ? in.f0[idx] : 0;
        if (KEEP_V) vv[i] = in.v ? in.v[idx] : 0.f;
        // loaded whether or not the flag needs it: a load that waits for the flag byte costs a second round trip per
        // step (long_scoreboard was 56 % of this kernel's stalls at T = 64)
        const float vnx = in.vn ? in.vn[idx] : 0.f;
        // R_t = gamma * Rin + r, where Rin is the bootstrap when a path ends at t
        if (endbuf) {
          const float boot = (e == 1) ? 0.f : ((e == 2 && in.vn) ? vnx : (in.v_last ? in.v_last[env] : 0.f));
          b[i] = fmaf(in.gamma, boot, r);
        } else if (e == 1) { b[i] = r; }
        else if (e == 2) { b[i] = fmaf(in.gamma, vnx, r); }
        else { amask |= 1u << i; b[i] = r; }
      }
    }
  }
  // compose the segment from its last element backwards: X_{t0} = A * X_in + B
  float A = 1.f, B = 0.f;
#pragma unroll
  for (int i = L - 1; i >= 0; --i) {
    const float a = ((amask >> i) & 1u) ? g : 0.f;
    B = fmaf(a, B, b[i]);
    A = a * A;
  }
  sA[w][lane] = A; sB[w][lane] = B;
  __syncthreads();
  // carry-in of this segment = value at the first step of segment w+1 = fold of segments 31 .. w+1
  float X = 0.f;
  for (int s = (int)blockDim.y - 1; s > w; --s) X = fmaf(sA[s][lane], X, sB[s][lane]);
#pragma unroll
  for (int i = L - 1; i >= 0; --i) {
    const float a = ((amask >> i) & 1u) ? g : 0.f;
    X = fmaf(a, X, b[i]);
    const int t = t0 + i;
    if (live && t < T) {
      const size_t idx = (size_t)t * ld + env;
      if (out0) out0[idx] = X;
      if (out1) {
        const float v = KEEP_V ? vv[i] : (in.v ? in.v[idx] : 0.f);
        out1[idx] = in.mode == 0 ? X + v : X - v;
      }
    }
  }
}

// (Measured and rejected for 256 < T <= 512: the time axis cut in two, a cluster of two 512-thread CTAs per 32 envs, the
// earlier half taking its carry-in out of the later half's shared memory behind one cluster barrier -- twice the CTAs,
// two per SM, all 148 SMs busy at 4096 envs, and no faster: 0.5003 against 0.4980 ms for the whole 4096 x 500 step.)
static int launch_affine_scan(const AffineIn& in, int T, int n, int ld, float* o0, float* o1, cudaStream_t st) {
  // L steps per warp, W = ceil(T / L) warps per CTA: short horizons take 8 steps per warp (T = 64: 8 warps, 7 segment
  // folds and 4x more CTAs than 32 warps of 2 steps), long ones fill the 32 warps
  const int l8 = ceil_div(T, 8) < 8 ? ceil_div(T, 8) : 8;
  const int L = ceil_div(T, 32) > l8 ? ceil_div(T, 32) : l8;
  const int Lp = L <= 1 ? 1 : L <= 2 ? 2 : L <= 4 ? 4 : L <= 8 ? 8 : L <= 16 ? 16 : 32;
  dim3 block(32, ceil_div(T, Lp)), grid(ceil_div(n, 32));
  if (L <= 1) launch_pdl(affine_scan_kernel<1>, grid, block, 0, st, in, T, n, ld, o0, o1);
  else if (L <= 2) launch_pdl(affine_scan_kernel<2>, grid, block, 0, st, in, T, n, ld, o0, o1);
  else if (L <= 4) launch_pdl(affine_scan_kernel<4>, grid, block, 0, st, in, T, n, ld, o0, o1);
  else if (L <= 8) launch_pdl(affine_scan_kernel<8>, grid, block, 0, st, in, T, n, ld, o0, o1);
  else if (L <= 16) launch_pdl(affine_scan_kernel<16>, grid, block, 0, st, in, T, n, ld, o0, o1);
  else launch_pdl(affine_scan_kernel<32>, grid, block, 0, st, in, T, n, ld, o0, o1);
  return 0;
}

// ---------------------------------------------------------------- K6: moment partial sums
// grid = (env chunks, C, row slices).  float64 accumulation: per-thread -> warp shuffle -> one atomicAdd per warp.
// VEC = 4: 16-byte loads, four rows in flight per thread (64 B per thread outstanding -- a scalar one-load-per-iteration
// loop left the 262 MB observation buffer at 1.9 TB/s); VEC = 1 serves unaligned or ragged (n % 4, ld % 4) buffers.
template <int VEC>
__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ x, int rows, int C, int n, int ld,
                                                      double* __restrict__ out) {
  pdl_wait();
  const int c = blockIdx.y;
  double s = 0.0, s2 = 0.0;
  const int nv = n / VEC;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nv) {
    constexpr int U = 4;
    const int gz = gridDim.z;
    int r = blockIdx.z;
    for (; r + (U - 1) * gz < rows; r += U * gz) {
      float v[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float* row = x + ((size_t)(r + u * gz) * C + c) * ld;
        if (VEC == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(row) + i);
          v[u][0] = q.x; v[u][1 % VEC] = q.y; v[u][2 % VEC] = q.z; v[u][3 % VEC] = q.w;
        } else {
          v[u][0] = __ldg(row + i);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const double d = v[u][k];
          s += d;
          s2 = fma(d, d, s2);
        }
    }
    for (; r < rows; r += gz) {
      const float* row = x + ((size_t)r * C + c) * ld;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const double d = __ldg(row + (size_t)i * VEC + k);
        s += d;
        s2 = fma(d, d, s2);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  // one atomic pair per CTA: with C = 1 (advantages) every warp of the grid would otherwise queue on the same two L2
  // addresses (8000 same-address double atomics made a 16 us kernel out of an 8 MB pass)
  __shared__ double sh[2][8];
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    if (t != 0.0) atomicAdd(out + threadIdx.x * C + c, t);
  }
  if (c == 0 && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0) atomicAdd(out + 2 * C, (double)rows * (double)n);
}

// mean / (std + eps) of a single-component moment buffer [sum, sumsq, count] -> stats[0..1]
__global__ void adv_stats_kernel(const double* __restrict__ mom, int unbiased, double eps, double* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  const double cnt = mom[2];
  const double mean = mom[0] / cnt;
  double var = mom[1] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  if (unbiased && cnt > 1.0) var = var * cnt / (cnt - 1.0);
  stats[0] = mean;
  stats[1] = sqrt(var) + eps;
}

// mean / denominator per component from [sum[C], sumsq[C], count] (om_moment_stats kinds; distributed.mean_std_from_moments)
__global__ void __launch_bounds__(128) moment_stats_kernel(const double* __restrict__ mom, int C, int kind,
                                                           double* __restrict__ mean, double* __restrict__ denom,
                                                           float* __restrict__ mean32, float* __restrict__ denom32) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = mom[2 * C];
  const double m = mom[c] / n;
  const double var = mom[C + c] / n - m * m;
  const double v0 = var > 0.0 ? var : 0.0;
  double d;
  if (kind == 0) d = sqrt(var > 1e-2 ? var : 1e-2);
  else if (kind == 1) d = sqrt(v0 + 1e-8);
  else if (kind == 2) d = sqrt(v0 * n / (n - 1.0)) + 1e-5;
  else d = sqrt(v0) + 1e-8;
  if (mean) mean[c] = m;
  if (denom) denom[c] = d;
  if (mean32) mean32[c] = (float)m;
  if (denom32) denom32[c] = (float)d;
}

__global__ void __launch_bounds__(256) normalize_kernel(const float* x, const double* __restrict__ stats, int rows, int n, int ld,
                                                        float* y) {            // y may be x (in place): no __restrict__
  pdl_wait();
  const double mean = stats[0], inv = 1.0 / stats[1];
  const int r = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t idx = (size_t)r * ld + i;
    y[idx] = (float)(((double)x[idx] - mean) * inv);
  }
}

// The same over a CONTIGUOUS buffer (ld == n, 16-byte aligned, element count a multiple of four): one grid-stride pass of
// 128-bit accesses instead of rows x n / 256 CTAs of one element per thread (8.7 -> 4 us for the 500 x 4096 advantages).
__global__ void __launch_bounds__(256) normalize_flat_kernel(const float4* x, const double* __restrict__ stats, size_t n4,
                                                             float4* y) {       // y may be x (in place): no __restrict__
  pdl_wait();
  const double mean = stats[0], inv = 1.0 / stats[1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    y[i] = make_float4((float)(((double)v.x - mean) * inv), (float)(((double)v.y - mean) * inv),
                       (float)(((double)v.z - mean) * inv), (float)(((double)v.w - mean) * inv));
  }
}

// K per-thread float64 partial sums -> one atomicAdd per CTA and statistic (256-thread CTAs): warp shuffle, then the eight
// warp leaders meet in shared memory.  Per-warp atomics would queue thousands of same-address updates in L2.
template <int K>
__device__ __forceinline__ void block_sum_atomic(double (&acc)[K], double* __restrict__ sums) {
  __shared__ double sh[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    if (t != 0.0) atomicAdd(&sums[threadIdx.x], t);
  }
}

// ---------------------------------------------------------------- N2: discriminator-fit statistics, expert minibatch
// One pass over the logits of [policy batch; expert batch]: per-sample fp32 terms, float64 sums (warp shuffle, one
// atomicAdd per warp and statistic).
__global__ void __launch_bounds__(256) disc_loss_kernel(const float* __restrict__ logit, const float* __restrict__ target,
                                                        const float* __restrict__ kl, int n_plcy, int n, float entcoeff,
                                                        double* __restrict__ sums, float* __restrict__ dlogit) {
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = logit[i];
    const bool demo = i >= n_plcy;
    const float t = target ? target[i] : (demo ? 1.f : 0.f);
    const float sp = log1pf(expf(-fabsf(x)));                    // log(1 + exp(-|x|))
    const float bce = fmaxf(x, 0.f) - x * t + sp;
    const float sig = x >= 0.f ? 1.f / (1.f + expf(-x)) : expf(x) / (1.f + expf(x));
    const float logsig = fminf(x, 0.f) - sp;                     // logsigmoid(x)
    const float ent = (1.f - sig) * x - logsig;
    acc[0] += bce; acc[1] += ent;
    if (kl) acc[2] += kl[i];
    if (demo) { acc[4] += x > 0.f ? 1.0 : 0.0; acc[6] += sig; }
    else { acc[3] += x < 0.f ? 1.0 : 0.0; acc[5] += sig; }
    if (dlogit) dlogit[i] = (sig - t) + entcoeff * x * sig * (1.f - sig);
  }
  block_sum_atomic<7>(acc, sums);
  if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&sums[7], (double)n_plcy); atomicAdd(&sums[8], (double)(n - n_plcy)); }
}

// ---------------------------------------------------------------- N3: PPO minibatch losses
//   PPO.update_policy  rl/algos/ppo.py:231-282: clipped surrogate, entropy penalty, value loss, approximate KL, clip
//   fraction and the mirror-symmetry loss (policy(obs) against mirror_action(policy(mirror(obs))), wrappers.py:54-55,
//   the signed permutation applied on the fly) in ONE pass over the minibatch: per-sample fp32 terms, float64 sums.
struct PpoLossArgs {
  const float* logp; const float* old_logp; const float* adv; const float* mask; const float* values; const float* returns;
  const float* entropy;       // [nu][ld] or null
  const float* act;           // [nu][ld] policy(obs) or null
  const float* act_mirror;    // [nu][ld] policy(mirror(obs)) BEFORE mirror_action
  OmMirrorSpec mir;           // action mirror (identity when numel == 0)
  int nu, n, ld;
  float clip, vf_coeff;
  double* sums; float* dlogp; float* dvalues;
};
__global__ void __launch_bounds__(256) ppo_loss_kernel(PpoLossArgs a) {
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const float inv_n = 1.0f / (float)a.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const float lr = a.logp[i] - a.old_logp[i];
    const float ratio = expf(lr);
    const float m = a.mask ? a.mask[i] : 1.f;
    const float A = a.adv[i] * m;
    const float lo = 1.f - a.clip, hi = 1.f + a.clip;
    const float cpi = ratio * A, clp = fminf(fmaxf(ratio, lo), hi) * A;
    acc[0] += fminf(cpi, clp);
    acc[3] += (ratio - 1.f) - lr;
    acc[5] += fabsf(ratio - 1.f) > a.clip ? 1.0 : 0.0;
    if (a.values) {
      const float d = a.returns[i] - a.values[i];
      acc[2] += d * d;
      if (a.dvalues) a.dvalues[i] = -2.f * a.vf_coeff * d * inv_n;
    }
    if (a.dlogp) {
      const bool inside = ratio >= lo && ratio <= hi;                    // clamp passes the gradient on [lo, hi]
      a.dlogp[i] = (inside || cpi < clp) ? -cpi * inv_n : 0.f;
    }
    if (a.entropy)
      for (int k = 0; k < a.nu; ++k) acc[1] += a.entropy[(size_t)k * a.ld + i] * m;
    if (a.act && a.act_mirror)
      for (int k = 0; k < a.nu; ++k) {
        // mirror_action: y[index[k]] = sign[k] * x[k]
        const int j = a.mir.numel ? a.mir.index[k] : k;
        const float s = a.mir.numel ? a.mir.sign[k] : 1.f;
        const float d = a.act[(size_t)j * a.ld + i] - s * a.act_mirror[(size_t)k * a.ld + i];
        acc[4] += d * d;
      }
  }
  block_sum_atomic<6>(acc, a.sums);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.sums[6], (double)a.n);
}

// keyed permutation of [0, n): balanced Feistel network on 2 * half bits + cycle walking (contract: oracle/learner.py)
OM_HD uint32_t om_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
OM_HD uint32_t om_feistel_perm(uint32_t v, uint32_t n, int half, U4 key) {
  const uint32_t mask = (1u << half) - 1u;
  do {
    uint32_t l = v >> half, r = v & mask;
    const uint32_t k[4] = {key.x, key.y, key.z, key.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t f = om_mix32(r ^ k[i]) & mask;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    v = (l << half) | r;
  } while (v >= n);
  return v;
}
constexpr uint32_t EXPERT_STREAM = 48;      // oracle/philox.py STREAM_EXPERT

__global__ void __launch_bounds__(128) expert_minibatch_kernel(const float* __restrict__ src, int n_src, int ld_src, int D,
                                                               uint64_t seed, uint32_t draw, int half, int batch,
                                                               float* __restrict__ out, float* __restrict__ out_next,
                                                               int32_t* __restrict__ idx_out, int ld_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const uint32_t epoch = (uint32_t)(b / n_src), i = (uint32_t)(b % n_src);
  const uint32_t idx = om_feistel_perm(i, (uint32_t)n_src, half, om_draw(seed, epoch, draw, EXPERT_STREAM));
  if (idx_out) idx_out[b] = (int32_t)idx;
  for (int c0 = 0; c0 < D; c0 += 8) {                   // eight gathers in flight, not one round trip per row
    float v[8], w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (c0 + u < D) {
        const float* row = src + (size_t)(c0 + u) * ld_src + idx;
        v[u] = row[0];
        if (out_next) w[u] = row[1];
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (c0 + u < D) {
        out[(size_t)(c0 + u) * ld_out + b] = v[u];
        if (out_next) out_next[(size_t)(c0 + u) * ld_out + b] = w[u];
      }
    }
  }
}

}  // namespace om

using namespace om;

extern "C" int om_ppo_returns(const float* rewards, const float* values, const uint8_t* path_end, const float* v_next,
                              const float* v_last, float gamma, int T, int n, int ld, float* ret, float* adv,
                              void* stream) {
  OM_REQUIRE(T >= 0 && n >= 0 && ld >= n, "om_ppo_returns: bad sizes");
  if (T == 0 || n == 0) return 0;
  OM_REQUIRE(rewards && (ret || adv), "om_ppo_returns: null argument");
  OM_REQUIRE(!adv || values, "om_ppo_returns: advantages need values");
  if (T <= 1024 && T >= 8 && (long long)n < 148LL * 2048 && !g_knobs.serial_scan) {
    AffineIn in{rewards, values, v_next, path_end, nullptr, v_last, gamma, 0.f, 1};
    launch_affine_scan(in, T, n, ld, ret, adv, (cudaStream_t)stream);
  } else {
    ppo_returns_kernel<8><<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(rewards, values, path_end, v_next, v_last,
                                                                               gamma, T, n, ld, ret, adv);
  }
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_gae(const float* rewards, const float* v, const float* v_next, const uint8_t* absorbing,
                      const uint8_t* last, float gamma, float lam, int T, int n, int ld, float* adv, float* v_target,
                      void* stream) {
  OM_REQUIRE(T >= 0 && n >= 0 && ld >= n, "om_gae: bad sizes");
  if (T == 0 || n == 0) return 0;
  OM_REQUIRE(rewards && v && v_next && (adv || v_target), "om_gae: null argument");
  if (T <= 1024 && T >= 8 && (long long)n < 148LL * 2048 && !g_knobs.serial_scan) {
    AffineIn in{rewards, v, v_next, absorbing, last, nullptr, gamma, gamma * lam, 0};
    launch_affine_scan(in, T, n, ld, adv, v_target, (cudaStream_t)stream);
  } else {
    gae_kernel<8><<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(rewards, v, v_next, absorbing, last, gamma, lam, T, n,
                                                                       ld, adv, v_target);
  }
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_moments(const float* x, int rows, int C, int n, int ld, double* out, void* stream) {
  OM_REQUIRE(rows >= 0 && C >= 1 && n >= 0 && ld >= n, "om_moments: bad sizes");
  if (rows == 0 || n == 0) return 0;
  OM_REQUIRE(x && out, "om_moments: null argument");
  const bool vec = (n % 4 == 0) && (ld % 4 == 0) && ((uintptr_t)x % 16 == 0);
  int gx = ceil_div(vec ? n / 4 : n, 256);
  if (gx < 1) gx = 1;
  OM_REQUIRE(gx <= 65535 * 16, "om_moments: n too large for one launch");
  // slice the rows over grid.z until the launch has ~8 CTAs per SM, keeping >= 4 rows per thread for the unrolled loop
  int gz = 1184 / (gx * C > 0 ? gx * C : 1);
  if (gz > rows / 4) gz = rows / 4;
  if (gz < 1) gz = 1;
  if (gz > 65535) gz = 65535;
  dim3 grid(gx, C, gz);
  if (vec) OM_CUDA_OK(launch_pdl(moments_kernel<4>, grid, dim3(256), 0, (cudaStream_t)stream, x, rows, C, n, ld, out));
  else OM_CUDA_OK(launch_pdl(moments_kernel<1>, grid, dim3(256), 0, (cudaStream_t)stream, x, rows, C, n, ld, out));
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_adv_stats(const double* moments, int unbiased, double eps, double* stats, void* stream) {
  OM_REQUIRE(moments && stats, "om_adv_stats: null argument");
  OM_CUDA_OK(launch_pdl(adv_stats_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, moments, unbiased, eps, stats));
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_moment_stats(const double* moments, int C, int kind, double* mean, double* denom, float* mean32,
                               float* denom32, void* stream) {
  OM_REQUIRE(moments && C >= 1 && kind >= 0 && kind <= 3, "om_moment_stats: bad argument (C=%d kind=%d)", C, kind);
  moment_stats_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(moments, C, kind, mean, denom, mean32, denom32);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_normalize(const float* x, const double* stats, int rows, int n, int ld, float* y, void* stream) {
  OM_REQUIRE(rows >= 0 && n >= 0 && ld >= n, "om_normalize: bad sizes");
  if (rows == 0 || n == 0) return 0;
  OM_REQUIRE(x && stats && y, "om_normalize: null argument");
  const size_t total = (size_t)rows * (size_t)n;
  if (ld == n && total % 4 == 0 && ((uintptr_t)x | (uintptr_t)y) % 16 == 0) {
    const size_t n4 = total / 4;
    const size_t want = (n4 + 255) / 256;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    OM_CUDA_OK(launch_pdl(normalize_flat_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
                          reinterpret_cast<const float4*>(x), stats, n4, reinterpret_cast<float4*>(y)));
    OM_LAUNCHED();
    return 0;
  }
  int gx = ceil_div(n, 256);
  if (gx > 592) gx = 592;
  dim3 grid(gx, rows);
  OM_CUDA_OK(launch_pdl(normalize_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, stats, rows, n, ld, y));
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_disc_loss_stats(const float* logit, const float* target, const float* kl, int n_plcy, int n_demo,
                                  float entcoeff, double* sums, float* dlogit, void* stream) {
  OM_REQUIRE(n_plcy >= 0 && n_demo >= 0, "om_disc_loss_stats: bad sizes");
  const int n = n_plcy + n_demo;
  if (n == 0) return 0;
  OM_REQUIRE(logit && sums, "om_disc_loss_stats: null argument");
  const int grid = ceil_div(n, 256) < 148 * 8 ? ceil_div(n, 256) : 148 * 8;
  disc_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logit, target, kl, n_plcy, n, entcoeff, sums, dlogit);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_expert_minibatch(const float* src, int n_src, int ld_src, int D, uint64_t seed, uint32_t draw, int batch,
                                   float* out, float* out_next, int32_t* idx_out, int ld_out, void* stream) {
  OM_REQUIRE(batch >= 0 && D >= 1 && n_src >= 1 && ld_out >= batch, "om_expert_minibatch: bad sizes");
  OM_REQUIRE(ld_src >= n_src + (out_next ? 1 : 0), "om_expert_minibatch: the dataset needs n_src%s rows (ld_src = %d)",
             out_next ? " + 1" : "", ld_src);
  if (batch == 0) return 0;
  OM_REQUIRE(src && out, "om_expert_minibatch: null argument");
  int bits = 2;
  while (bits < 32 && (1ull << bits) < (unsigned long long)n_src) ++bits;
  const int half = (bits + 1) / 2;
  expert_minibatch_kernel<<<ceil_div(batch, 128), 128, 0, (cudaStream_t)stream>>>(src, n_src, ld_src, D, seed, draw, half, batch,
                                                                                 out, out_next, idx_out, ld_out);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_ppo_loss_stats(const float* logp, const float* old_logp, const float* adv, const float* mask,
                                 const float* values, const float* returns, const float* entropy, const float* act,
                                 const float* act_mirror, const OmMirrorSpec* action_mirror, int nu, int n, int ld, float clip,
                                 float vf_coeff, double* sums, float* dlogp, float* dvalues, void* stream) {
  OM_REQUIRE(n >= 0 && ld >= n && nu >= 0 && nu <= 64, "om_ppo_loss_stats: bad sizes (n=%d ld=%d nu=%d)", n, ld, nu);
  if (n == 0) return 0;
  OM_REQUIRE(logp && old_logp && adv && sums, "om_ppo_loss_stats: null argument");
  OM_REQUIRE(!values == !returns, "om_ppo_loss_stats: values and returns come together");
  OM_REQUIRE(!act == !act_mirror, "om_ppo_loss_stats: act and act_mirror come together");
  PpoLossArgs a{logp, old_logp, adv, mask, values, returns, entropy, act, act_mirror, {}, nu, n, ld, clip, vf_coeff, sums, dlogp,
                dvalues};
  a.mir.numel = 0;
  if (action_mirror && act) {
    OM_REQUIRE(action_mirror->numel == nu, "om_ppo_loss_stats: the action mirror has %d entries, nu = %d", action_mirror->numel, nu);
    for (int k = 0; k < nu; ++k)
      OM_REQUIRE(action_mirror->index[k] >= 0 && action_mirror->index[k] < nu, "om_ppo_loss_stats: mirror index[%d] out of range", k);
    a.mir = *action_mirror;
  }
  const int grid = ceil_div(n, 256) < 148 * 8 ? ceil_div(n, 256) : 148 * 8;
  ppo_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  OM_LAUNCHED();
  return 0;
}
