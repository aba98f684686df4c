// K5 (returns / GAE reverse scans over a time-major rollout buffer) and K6 (moment partial sums).
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

// ---------------------------------------------------------------- K6: moment partial sums
// grid = (env chunks, C).  float64 accumulation: per-thread -> warp shuffle -> one atomicAdd per warp.
__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ x, int rows, int C, int n, int ld,
                                                      double* __restrict__ out) {
  const int c = blockIdx.y;
  double s = 0.0, s2 = 0.0;
  for (int r = 0; r < rows; ++r) {
    const float* row = x + ((size_t)r * C + c) * ld;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const double v = row[i];
      s += v;
      s2 = fma(v, v, s2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + c, s);
    atomicAdd(out + C + c, s2);
  }
  if (c == 0 && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(out + 2 * C, (double)rows * (double)n);
}

// mean / (std + eps) of a single-component moment buffer [sum, sumsq, count] -> stats[0..1]
__global__ void adv_stats_kernel(const double* __restrict__ mom, int unbiased, double eps, double* __restrict__ stats) {
  const double cnt = mom[2];
  const double mean = mom[0] / cnt;
  double var = mom[1] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  if (unbiased && cnt > 1.0) var = var * cnt / (cnt - 1.0);
  stats[0] = mean;
  stats[1] = sqrt(var) + eps;
}

__global__ void __launch_bounds__(256) normalize_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                        int rows, int n, int ld, float* __restrict__ y) {
  const double mean = stats[0], inv = 1.0 / stats[1];
  const int r = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t idx = (size_t)r * ld + i;
    y[idx] = (float)(((double)x[idx] - mean) * inv);
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
  ppo_returns_kernel<8><<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(rewards, values, path_end, v_next, v_last, gamma,
                                                                             T, n, ld, ret, adv);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_gae(const float* rewards, const float* v, const float* v_next, const uint8_t* absorbing,
                      const uint8_t* last, float gamma, float lam, int T, int n, int ld, float* adv, float* v_target,
                      void* stream) {
  OM_REQUIRE(T >= 0 && n >= 0 && ld >= n, "om_gae: bad sizes");
  if (T == 0 || n == 0) return 0;
  OM_REQUIRE(rewards && v && v_next && (adv || v_target), "om_gae: null argument");
  gae_kernel<8><<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(rewards, v, v_next, absorbing, last, gamma, lam, T, n, ld,
                                                                     adv, v_target);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_moments(const float* x, int rows, int C, int n, int ld, double* out, void* stream) {
  OM_REQUIRE(rows >= 0 && C >= 1 && n >= 0 && ld >= n, "om_moments: bad sizes");
  if (rows == 0 || n == 0) return 0;
  OM_REQUIRE(x && out, "om_moments: null argument");
  int gx = ceil_div(n, 256 * 4);
  if (gx < 1) gx = 1;
  if (gx > 1184) gx = 1184;
  dim3 grid(gx, C);
  moments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, C, n, ld, out);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_adv_stats(const double* moments, int unbiased, double eps, double* stats, void* stream) {
  OM_REQUIRE(moments && stats, "om_adv_stats: null argument");
  adv_stats_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(moments, unbiased, eps, stats);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_normalize(const float* x, const double* stats, int rows, int n, int ld, float* y, void* stream) {
  OM_REQUIRE(rows >= 0 && n >= 0 && ld >= n, "om_normalize: bad sizes");
  if (rows == 0 || n == 0) return 0;
  OM_REQUIRE(x && stats && y, "om_normalize: null argument");
  int gx = ceil_div(n, 256);
  if (gx > 592) gx = 592;
  dim3 grid(gx, rows);
  normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, stats, rows, n, ld, y);
  OM_LAUNCHED();
  return 0;
}
