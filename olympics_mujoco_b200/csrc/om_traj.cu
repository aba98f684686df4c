// K3: reference-trajectory table + per-env integer state, and the fused H1 playback kernel.
#include <cstdlib>
#include <vector>

#include "om_common.cuh"
#include "om_sinks.cuh"
#include "om_traj.cuh"
#include "gen/fk_unitree_h1.cuh"
#include "gen/h1_perm.h"

namespace om {

// ---------------------------------------------------------------- per-step API kernels
__global__ void __launch_bounds__(128) traj_reset_kernel(TrajDev t, uint64_t seed, uint32_t env_id0,
                                                         const uint8_t* __restrict__ mask,
                                                         const int32_t* __restrict__ forced_traj,
                                                         const int32_t* __restrict__ forced_step,
                                                         int32_t* traj_no, int32_t* step_no, uint32_t* reset_count,
                                                         double* xy_off, float* sample, int n, int ld) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  if (mask && !mask[env]) return;
  int tr, st;
  uint32_t rc = reset_count[env];
  traj_draw(t, seed, env_id0 + env, rc, tr, st);
  if (forced_traj && forced_traj[env] >= 0) tr = min(forced_traj[env], t.n_traj - 1);
  if (forced_step && forced_step[env] >= 0) st = min(forced_step[env], t.T - 1);
  reset_count[env] = rc + 1;
  traj_no[env] = tr;
  step_no[env] = st;
  const double ox = t.xy[((size_t)tr * t.T + st) * 2], oy = t.xy[((size_t)tr * t.T + st) * 2 + 1];
  xy_off[env] = ox;
  xy_off[(size_t)ld + env] = oy;
  if (sample) traj_write_sample(t, tr, st, ox, oy, sample, ld, env);
}

__global__ void __launch_bounds__(128) traj_current_kernel(TrajDev t, const int32_t* __restrict__ traj_no,
                                                           const int32_t* __restrict__ step_no,
                                                           const double* __restrict__ xy_off, float* sample, int n,
                                                           int ld) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  traj_write_sample(t, traj_no[env], step_no[env], xy_off[env], xy_off[(size_t)ld + env], sample, ld, env);
}

__global__ void __launch_bounds__(128) traj_next_kernel(TrajDev t, uint64_t seed, uint32_t env_id0, int auto_reset,
                                                        int32_t* traj_no, int32_t* step_no, uint32_t* reset_count,
                                                        double* xy_off, float* sample, uint8_t* wrapped, int n, int ld) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  int tr = traj_no[env], st = step_no[env] + 1;
  double ox = xy_off[env], oy = xy_off[(size_t)ld + env];
  const bool wrap = st >= t.T;
  if (wrap && !auto_reset) {            // get_next_sample returns None: the index parks at T, sample untouched
    step_no[env] = t.T;
    if (wrapped) wrapped[env] = 1;
    return;
  }
  if (wrap) {
    uint32_t rc = reset_count[env];
    traj_draw(t, seed, env_id0 + env, rc, tr, st);
    reset_count[env] = rc + 1;
    ox = t.xy[((size_t)tr * t.T + st) * 2];
    oy = t.xy[((size_t)tr * t.T + st) * 2 + 1];
    xy_off[env] = ox;
    xy_off[(size_t)ld + env] = oy;
    traj_no[env] = tr;
  }
  step_no[env] = st;
  if (wrapped) wrapped[env] = wrap ? 1 : 0;
  if (sample) traj_write_sample(t, tr, st, ox, oy, sample, ld, env);
}

// ---------------------------------------------------------------- fused playback (sequential in time)
// One thread per env walks its episode step by step (loco_env_base.py:511-552).  Used when the episode is
// short or the caller asks for it; the time-parallel kernel below is the throughput path.
struct PlayArgs {
  TrajDev t;
  uint64_t seed;
  uint32_t env_id0;
  double dt;
  float target;
  int n_steps, end_reset, n, ld;      // (no use_absorbing: the playback loops call _has_fallen whatever the MDP setting,
                                      //  loco_env_base.py:422,541)
  int forced;         // 1 = play_trajectory (the model is forced to each sample), 0 = play_trajectory_from_velocity
  OmPlayState s;      // live carried state (written by the thread that runs the last step)
  OmPlayState snap;   // episode-start snapshot read by every chunk of the time-parallel kernel
  OmPlayOut o;
  double* obs_moments;   // = o.obs_moments (S1 fused into the playback), or null
};

OM_HD bool h1_fallen_f(float y, float tilt, float lst, float rot) {
  const double PI = 3.141592653589793;
  const double dy = y, dt = tilt, dl = lst, dr = rot;
  return (dy < -0.3) || (dy > 0.1) || (dt < (-PI / 4.5)) || (dt > (PI / 12)) || (dl < -PI / 12) || (dl > PI / 8) ||
         (dr < (-PI / 8)) || (dr > (PI / 8));
}

// One playback step given the sim state; shared by both playback kernels.
//   qs[17] (spec order, double) , dq[17] (spec order) -> FK outputs at time slot `slot`
// STREAM: streaming stores -- the time-parallel kernel (few envs, many steps: +2.5 % there); with 10^6 envs in the
// one-thread-per-env kernel they measured 1.5 % slower.
template <bool STREAM>
OM_HD void play_fk(const double (&qs)[17], const float (&dq)[17], const OmPlayOut& o, size_t slot, size_t ld, size_t env) {
  float q[17], qd[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) {
    q[OM_H1_PERM[k]] = (float)qs[k];
    qd[OM_H1_PERM[k]] = dq[k];
  }
  SoaSink<false, STREAM> S{o.xpos ? o.xpos + slot * 63 * ld : nullptr, o.xquat ? o.xquat + slot * 84 * ld : nullptr,
                   o.site_xpos ? o.site_xpos + slot * 3 * ld : nullptr, nullptr,
                   o.cvel ? o.cvel + slot * 126 * ld : nullptr, nullptr, ld, env};
  om_fk_unitree_h1(q, qd, S);
}

// obs / fallen / reward / integer state of one step from the freshly gathered sample row
constexpr int PLAY_BLOCK = 128;
template <bool STREAM>
OM_HD void play_emit(const float (&samp)[36], float prev_x_vel, const PlayArgs& a, size_t slot, size_t env, int tr, int st) {
  const size_t ld = a.ld;
  if (a.o.obs) {
    float* ob = a.o.obs + slot * 32 * ld + env;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
#ifdef __CUDA_ARCH__
      if (STREAM) __stcs(ob + k * ld, samp[k + 2]);   // written once, not read back by this call (the moments come from the prefix sums)
      else ob[k * ld] = samp[k + 2];
#else
      ob[k * ld] = samp[k + 2];
#endif
    }
  }
  if (a.o.fallen) a.o.fallen[slot * ld + env] = h1_fallen_f(samp[2], samp[3], samp[4], samp[5]) ? 1 : 0;
  if (a.o.reward) {
    const float d = prev_x_vel - a.target;
    a.o.reward[slot * ld + env] = expf(-(d * d));
  }
  if (a.o.traj_no_t) a.o.traj_no_t[slot * ld + env] = tr;
  if (a.o.step_no_t) a.o.step_no_t[slot * ld + env] = st;
}

// S1 of the playback WITHOUT touching the observation buffer: every observation a playback call emits is a row of the
// (L2-resident) trajectory table, and an env walks the table in contiguous runs -- rows st0+1 .. T-1 of its trajectory, then
// after each wrap reset rows st' .. of the drawn one.  With float64 prefix sums of the rows and of their squares
// (TrajDev::psum, 1 MB for 4 x 500 samples) the moment sums of a whole call are a few prefix differences per env and
// channel: one small kernel instead of a second pass over the [T][32][n] observation buffer (262 MB re-read per 4096 x 500 rollout; an
// in-kernel accumulation through shared memory cost the playback kernel 75 us of LSU bandwidth).  Layout of the result:
// om_moments' [sum[32], sumsq[32], count] (Standardizer networks.py:76-81, RunningMeanStd normalize.py:190-208).
constexpr int PM_ENVS = 4;     // envs per warp
__global__ void __launch_bounds__(256) play_moments_kernel(PlayArgs a, OmPlayState start, int start_reset) {
  // a warp takes PM_ENVS envs one after the other with its LANES ON THE 64 STATISTIC ROWS (lane l: rows l and l + 32), so
  // that every prefix-table access is one coalesced 256-byte read and no cross-lane reduction is needed; the 8 warps of
  // the CTA meet in shared memory: 64 atomics per 32 envs.  (One thread per env with 64 accumulators each made every load
  // a 32-way scattered access: 34 us for 4096 envs; per-warp atomics queued 2048 same-line float64 updates: 12 us.)
  __shared__ double part[8][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int env0 = (blockIdx.x * 8 + w) * PM_ENVS;
  const int T = a.t.T;
  int tr_[PM_ENVS], st_[PM_ENVS];
  uint32_t rc_[PM_ENVS];
#pragma unroll
  for (int u = 0; u < PM_ENVS; ++u) {                  // the state loads of all envs first (independent round trips)
    const int env = env0 + u;
    const bool live = env < a.n;
    tr_[u] = live ? start.traj_no[env] : 0;
    st_[u] = live ? start.step_no[env] : 0;
    rc_[u] = live ? start.reset_count[env] : 0u;
  }
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int u = 0; u < PM_ENVS; ++u) {
    const int env = env0 + u;
    if (env >= a.n) break;
    int tr = tr_[u], st = st_[u];
    uint32_t rc = rc_[u];
    if (start_reset) { traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st); ++rc; }
    int left = a.n_steps;
    int lo = st + 1;                                   // the start row itself was emitted by the previous step / by reset()
    while (left > 0) {
      if (lo >= T) {                                   // wrap -> reset: the drawn sample itself is the next observation
        traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st);
        ++rc;
        lo = st;
      }
      const int hi = min(T, lo + left);                // rows lo .. hi-1
      const double* p_lo = a.t.psum + ((size_t)tr * (T + 1) + lo) * 64 + lane;
      const double* p_hi = a.t.psum + ((size_t)tr * (T + 1) + hi) * 64 + lane;
      const double h0 = __ldg(p_hi), h1 = __ldg(p_hi + 32), l0 = __ldg(p_lo), l1 = __ldg(p_lo + 32);
      s0 += h0 - l0;
      s1 += h1 - l1;
      left -= hi - lo;
      lo = hi;
    }
  }
  part[w][lane] = s0;
  part[w][lane + 32] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += part[k][threadIdx.x];
    if (v != 0.0) atomicAdd(a.obs_moments + threadIdx.x, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.obs_moments + 64, (double)a.n * (double)a.n_steps);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) play_h1_seq_kernel(PlayArgs a) {
  const int env = blockIdx.x * BLOCK + threadIdx.x;
  if (env >= a.n) return;
  const size_t ld = a.ld, e = env;
  int tr = a.s.traj_no[e], st = a.s.step_no[e];
  uint32_t rc = a.s.reset_count[e];
  double ox = a.s.xy_off[e], oy = a.s.xy_off[ld + e];
  double cq[17];
  float dq[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) {
    cq[k] = a.s.curr_qpos ? a.s.curr_qpos[k * ld + e] : 0.0;
    dq[k] = a.s.pending[(17 + k) * ld + e];
  }
  float pxv = a.s.prev_x_vel[e];
  float samp[36];
  bool have_samp = false;
  if (a.forced) {
#pragma unroll
    for (int k = 0; k < 17; ++k) cq[k] = a.s.pending[k * ld + e];                    // play_trajectory: q = sample (:408)
  }
  for (int s = 0; s < a.n_steps; ++s) {
    if (!a.forced) {
#pragma unroll
      for (int k = 0; k < 17; ++k) cq[k] = fma(a.dt, (double)dq[k], cq[k]);          // :515-519
    }
    play_fk<false>(cq, dq, a.o, (size_t)s, ld, e);                                          // :521-525 / :408-410
    ++st;                                                                            // :532
    const bool wrap = st >= a.t.T;
    if (wrap) {                                                                      // :534-537
      traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st);
      ++rc;
      ox = a.t.xy[((size_t)tr * a.t.T + st) * 2];
      oy = a.t.xy[((size_t)tr * a.t.T + st) * 2 + 1];
    }
    traj_load_row(a.t, tr, st, samp);
    have_samp = true;
    if (a.forced) {                   // the next step is forced to this sample (x, y re-centred in double, then fp32)
      cq[0] = (float)(a.t.xy[((size_t)tr * a.t.T + st) * 2] - ox);
      cq[1] = (float)(a.t.xy[((size_t)tr * a.t.T + st) * 2 + 1] - oy);
#pragma unroll
      for (int k = 2; k < 17; ++k) cq[k] = samp[k];
    } else if (wrap) {
      cq[0] = 0.0; cq[1] = 0.0;       // x - x_off and y - y_off are exactly zero at the reset sample
#pragma unroll
      for (int k = 2; k < 17; ++k) cq[k] = samp[k];
    }
#pragma unroll
    for (int k = 0; k < 17; ++k) dq[k] = samp[17 + k];
    play_emit<false>(samp, pxv, a, (size_t)s, e, tr, st);                                   // :539-541
    pxv = samp[17];
  }
  // write the loop's `sample` variable back (x, y re-centred) before the end-of-episode reset
  if (have_samp) {
    a.s.pending[e] = (float)(a.t.xy[((size_t)tr * a.t.T + st) * 2] - ox);
    a.s.pending[ld + e] = (float)(a.t.xy[((size_t)tr * a.t.T + st) * 2 + 1] - oy);
#pragma unroll
    for (int k = 2; k < 34; ++k) a.s.pending[k * ld + e] = samp[k];
  }
  if (a.end_reset) {                                                                 // :555-557
    traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st);
    ++rc;
    ox = a.t.xy[((size_t)tr * a.t.T + st) * 2];
    oy = a.t.xy[((size_t)tr * a.t.T + st) * 2 + 1];
    traj_load_row(a.t, tr, st, samp);
    cq[0] = 0.0; cq[1] = 0.0;
#pragma unroll
    for (int k = 2; k < 17; ++k) cq[k] = samp[k];
    pxv = samp[17];
  }
  a.s.traj_no[e] = tr;
  a.s.step_no[e] = st;
  a.s.reset_count[e] = rc;
  a.s.xy_off[e] = ox;
  a.s.xy_off[ld + e] = oy;
#pragma unroll
  for (int k = 0; k < 17; ++k) if (a.s.curr_qpos) a.s.curr_qpos[k * ld + e] = cq[k];
  a.s.prev_x_vel[e] = pxv;
}

// ---------------------------------------------------------------- fused playback, time-parallel
// reset() at the START of a playback call (loco_env_base.py:481 / :377: every call begins with self.reset();
// sample = get_current_sample(); curr_qpos = sample[:len_qpos]) for one env: the draw, then the carried state is written
// to `dst` (the live state for the sequential kernel, the episode-start snapshot for the time-parallel one).
OM_HD void play_start_reset(const PlayArgs& a, const OmPlayState& src, const OmPlayState& dst, int env) {
  const size_t ld = a.ld, e = env;
  int tr, st;
  const uint32_t rc = src.reset_count[e];
  traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st);
  float samp[36];
  traj_load_row(a.t, tr, st, samp);
  dst.traj_no[e] = tr;
  dst.step_no[e] = st;
  dst.reset_count[e] = rc + 1;
  dst.xy_off[e] = a.t.xy[((size_t)tr * a.t.T + st) * 2];
  dst.xy_off[ld + e] = a.t.xy[((size_t)tr * a.t.T + st) * 2 + 1];
  samp[0] = 0.f; samp[1] = 0.f;                  // x - x_off and y - y_off are exactly zero at the reset sample
#pragma unroll
  for (int k = 0; k < 34; ++k) dst.pending[k * ld + e] = samp[k];
  if (dst.curr_qpos) {
#pragma unroll
    for (int k = 0; k < 17; ++k) dst.curr_qpos[k * ld + e] = (double)samp[k];
  }
  dst.prev_x_vel[e] = samp[17];
}

__global__ void __launch_bounds__(256) play_start_reset_kernel(PlayArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.n) play_start_reset(a, a.s, a.s, e);
}

template <bool RESET>
__global__ void __launch_bounds__(256) play_snapshot_kernel(PlayArgs a, OmPlayState live, OmPlayState snap, int n, int ld) {
  pdl_trigger();                                 // a few CTAs: the playback kernel's first wave may become resident now
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (RESET) {                                   // the call starts with reset(): the snapshot IS the reset state
    play_start_reset(a, live, snap, e);
    return;
  }
  // all loads first (the pointers may alias as far as the compiler knows: a load-store-load chain costs 34 round trips)
  double cq[17], xo[2];
  float pd[34];
#pragma unroll
  for (int k = 0; k < 17; ++k) cq[k] = live.curr_qpos ? live.curr_qpos[(size_t)k * ld + e] : 0.0;
#pragma unroll
  for (int k = 0; k < 34; ++k) pd[k] = live.pending[(size_t)k * ld + e];
  xo[0] = live.xy_off[e];
  xo[1] = live.xy_off[(size_t)ld + e];
  const int32_t tr = live.traj_no[e], st = live.step_no[e];
  const uint32_t rc = live.reset_count[e];
  const float pxv = live.prev_x_vel[e];
  snap.traj_no[e] = tr;
  snap.step_no[e] = st;
  snap.reset_count[e] = rc;
  snap.xy_off[e] = xo[0];
  snap.xy_off[(size_t)ld + e] = xo[1];
  snap.prev_x_vel[e] = pxv;
#pragma unroll
  for (int k = 0; k < 17; ++k) snap.curr_qpos[(size_t)k * ld + e] = cq[k];
#pragma unroll
  for (int k = 0; k < 34; ++k) snap.pending[(size_t)k * ld + e] = pd[k];
}

// The per-env recurrences of the playback loop are (a) the integer trajectory index, which only changes
// non-trivially at wrap resets whose draws depend on (seed, env, reset_count) alone, and (b) the Euler sum
// q_j = q_0 + dt * sum(dq), which telescopes through the float64 prefix table cdq.  A thread can therefore
// reconstruct the loop state in front of ANY step j0 in O(#resets) and then walk `chunk` steps exactly like
// the sequential kernel.  grid = (env tiles, time chunks): 4096 envs x 500 steps become ~300k threads
// instead of 4096, which is what lets a 4096-env rollout fill the 148 SMs.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) play_h1_tp_kernel(PlayArgs a, int chunk) {
  pdl_wait();                                    // launched while the snapshot kernel drains
  const int env = blockIdx.x * BLOCK + threadIdx.x;
  if (env >= a.n) return;
  const int j0 = blockIdx.y * chunk, j1 = min(j0 + chunk, a.n_steps);
  const size_t ld = a.ld, e = env;
  const int T = a.t.T;
  int tr = a.snap.traj_no[e], st = a.snap.step_no[e];
  uint32_t rc = a.snap.reset_count[e];
  // ---- replay the wrap resets that happen before step j0
  int seg_start = 0;
  bool first = true;
  while (seg_start + (T - st) <= j0) {
    seg_start += T - st;
    traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st);
    ++rc;
    first = false;
  }
  const int m = j0 - seg_start;                  // steps already taken inside this segment
  double ox, oy;
  double cq[17];
  float dq[17];
  float pxv;
  float samp[36];
  const double* c_hi = a.t.cdq + ((size_t)tr * (T + 1) + st + m) * 17;
  if (a.forced) {
    // play_trajectory: the sim state of step j0 IS the pending sample -- the carried one when nothing happened yet,
    // else row (tr, st + m) re-centred with the offsets of the reset that opened this segment
    if (first) { ox = a.snap.xy_off[e]; oy = a.snap.xy_off[ld + e]; }
    else { ox = a.t.xy[((size_t)tr * T + st) * 2]; oy = a.t.xy[((size_t)tr * T + st) * 2 + 1]; }
    if (first && m == 0) {
#pragma unroll
      for (int k = 0; k < 17; ++k) { cq[k] = a.snap.pending[k * ld + e]; dq[k] = a.snap.pending[(17 + k) * ld + e]; }
      pxv = a.snap.prev_x_vel[e];
    }
  } else if (first) {
    ox = a.snap.xy_off[e]; oy = a.snap.xy_off[ld + e];
    const double* c_lo = a.t.cdq + ((size_t)tr * (T + 1) + st + 1) * 17;
#pragma unroll
    for (int k = 0; k < 17; ++k) {
      const double q0 = a.snap.curr_qpos[k * ld + e];
      const float p0 = a.snap.pending[(17 + k) * ld + e];
      cq[k] = (m == 0) ? q0 : fma(a.dt, (double)p0 + (c_hi[k] - c_lo[k]), q0);
      dq[k] = p0;
    }
    pxv = a.snap.prev_x_vel[e];
  } else {
    ox = a.t.xy[((size_t)tr * T + st) * 2];
    oy = a.t.xy[((size_t)tr * T + st) * 2 + 1];
    traj_load_row(a.t, tr, st, samp);
    const double* c_lo = a.t.cdq + ((size_t)tr * (T + 1) + st) * 17;
    cq[0] = fma(a.dt, c_hi[0] - c_lo[0], 0.0);
    cq[1] = fma(a.dt, c_hi[1] - c_lo[1], 0.0);
#pragma unroll
    for (int k = 2; k < 17; ++k) cq[k] = fma(a.dt, c_hi[k] - c_lo[k], (double)samp[k]);
  }
  st += m;
  if (!first || m > 0) {                         // the pending sample is row (tr, st)
    traj_load_row(a.t, tr, st, samp);
#pragma unroll
    for (int k = 0; k < 17; ++k) dq[k] = samp[17 + k];
    pxv = samp[17];
    if (a.forced) {
      cq[0] = (float)(a.t.xy[((size_t)tr * T + st) * 2] - ox);
      cq[1] = (float)(a.t.xy[((size_t)tr * T + st) * 2 + 1] - oy);
#pragma unroll
      for (int k = 2; k < 17; ++k) cq[k] = samp[k];
    }
  }
  // ---- walk the chunk (identical to the sequential kernel's loop body)
  for (int s = j0; s < j1; ++s) {
    if (!a.forced) {
#pragma unroll
      for (int k = 0; k < 17; ++k) cq[k] = fma(a.dt, (double)dq[k], cq[k]);
    }
    play_fk<true>(cq, dq, a.o, (size_t)s, ld, e);
    ++st;
    const bool wrap = st >= T;
    if (wrap) {
      traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st);
      ++rc;
      ox = a.t.xy[((size_t)tr * T + st) * 2];
      oy = a.t.xy[((size_t)tr * T + st) * 2 + 1];
    }
    traj_load_row(a.t, tr, st, samp);
    if (a.forced) {
      cq[0] = (float)(a.t.xy[((size_t)tr * T + st) * 2] - ox);
      cq[1] = (float)(a.t.xy[((size_t)tr * T + st) * 2 + 1] - oy);
#pragma unroll
      for (int k = 2; k < 17; ++k) cq[k] = samp[k];
    } else if (wrap) {
      cq[0] = 0.0; cq[1] = 0.0;
#pragma unroll
      for (int k = 2; k < 17; ++k) cq[k] = samp[k];
    }
#pragma unroll
    for (int k = 0; k < 17; ++k) dq[k] = samp[17 + k];
    play_emit<true>(samp, pxv, a, (size_t)s, e, tr, st);
    pxv = samp[17];
  }
  if (j1 != a.n_steps) return;
  // ---- the thread that ran the last step owns the carried state
  a.s.pending[e] = (float)(a.t.xy[((size_t)tr * T + st) * 2] - ox);
  a.s.pending[ld + e] = (float)(a.t.xy[((size_t)tr * T + st) * 2 + 1] - oy);
#pragma unroll
  for (int k = 2; k < 34; ++k) a.s.pending[k * ld + e] = samp[k];
  if (a.end_reset) {
    traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st);
    ++rc;
    ox = a.t.xy[((size_t)tr * T + st) * 2];
    oy = a.t.xy[((size_t)tr * T + st) * 2 + 1];
    traj_load_row(a.t, tr, st, samp);
    cq[0] = 0.0; cq[1] = 0.0;
#pragma unroll
    for (int k = 2; k < 17; ++k) cq[k] = samp[k];
    pxv = samp[17];
  }
  a.s.traj_no[e] = tr;
  a.s.step_no[e] = st;
  a.s.reset_count[e] = rc;
  a.s.xy_off[e] = ox;
  a.s.xy_off[ld + e] = oy;
#pragma unroll
  for (int k = 0; k < 17; ++k) if (a.s.curr_qpos) a.s.curr_qpos[k * ld + e] = cq[k];
  a.s.prev_x_vel[e] = pxv;
}

// ---------------------------------------------------------------- fused live step
// One env step of a trajectory-driven rollout in ONE kernel (was traj_next -> set_sim_state -> h1_step, the sample and
// qpos / qvel round-tripping through HBM between three launches): advance the trajectory index (wrap -> reset,
// loco_env_base.py:534-537), gather the sample row (get_next_sample, trajectory.py:389-401), scatter it into qpos / qvel
// (set_sim_state :659-684; only written when the caller wants the MjData mirrors), K1 on it, observation
// (_create_observation :737-767), has_fallen, TargetVelocityReward on the PREVIOUS observation (mushroom MuJoCo.step:
// reward(self._obs, action, cur_obs, absorbing); self._obs = cur_obs).
struct LiveArgs {
  TrajDev t;
  uint64_t seed;
  uint32_t env_id0;
  float target;
  int use_absorbing, n, ld;
  int32_t* traj_no; int32_t* step_no; uint32_t* reset_count; double* xy_off; float* prev_x_vel;
  float* qpos; float* qvel;
  float* xpos; float* xquat; float* site_xpos; float* cvel;
  float* obs; float* reward; uint8_t* absorbing; uint8_t* wrapped;
};

// STREAM: streaming stores for the FK outputs (131 072 envs: 34.9 -> 31.8 us = 0.97 of HBM; 1 048 576 envs: 0.261 -> 0.266 ms,
// so the host picks by batch size)
template <int BLOCK, bool STREAM>
__global__ void __launch_bounds__(BLOCK) h1_live_step_kernel(LiveArgs a) {
  pdl_wait();                                    // consecutive live steps: this launch overlaps the previous step's drain
  const int env = blockIdx.x * BLOCK + threadIdx.x;
  if (env >= a.n) return;
  const size_t ld = a.ld, e = env;
  const int T = a.t.T;
  int tr = a.traj_no[e], st = a.step_no[e] + 1;
  double ox = a.xy_off[e], oy = a.xy_off[ld + e];
  const float pxv = a.prev_x_vel[e];
  const bool wrap = st >= T;
  if (wrap) {
    const uint32_t rc = a.reset_count[e];
    traj_draw(a.t, a.seed, a.env_id0 + env, rc, tr, st);
    a.reset_count[e] = rc + 1;
    ox = a.t.xy[((size_t)tr * T + st) * 2];
    oy = a.t.xy[((size_t)tr * T + st) * 2 + 1];
    a.xy_off[e] = ox;
    a.xy_off[ld + e] = oy;
    a.traj_no[e] = tr;
  }
  a.step_no[e] = st;
  if (a.wrapped) a.wrapped[e] = wrap ? 1 : 0;
  float samp[36];
  traj_load_row(a.t, tr, st, samp);
  samp[0] = (float)(a.t.xy[((size_t)tr * T + st) * 2] - ox);
  samp[1] = (float)(a.t.xy[((size_t)tr * T + st) * 2 + 1] - oy);
  float q[17], qd[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) {
    q[OM_H1_PERM[k]] = samp[k];
    qd[OM_H1_PERM[k]] = samp[17 + k];
  }
  if (a.qpos) {
#pragma unroll
    for (int k = 0; k < 17; ++k) a.qpos[k * ld + e] = q[k];
  }
  if (a.qvel) {
#pragma unroll
    for (int k = 0; k < 17; ++k) a.qvel[k * ld + e] = qd[k];
  }
  if (a.obs) {
#pragma unroll
    for (int k = 0; k < 32; ++k) a.obs[k * ld + e] = samp[k + 2];
  }
  if (a.absorbing) a.absorbing[e] = (a.use_absorbing && h1_fallen_f(samp[2], samp[3], samp[4], samp[5])) ? 1 : 0;
  if (a.reward) {
    const float d = pxv - a.target;
    a.reward[e] = expf(-(d * d));
  }
  a.prev_x_vel[e] = samp[17];                     // dq_pelvis_tx: emitted row 15 of UnitreeH1's own spec (checked on the host)
  if (a.xpos || a.xquat || a.site_xpos || a.cvel) {
    SoaSink<false, STREAM> S{a.xpos, a.xquat, a.site_xpos, nullptr, a.cvel, nullptr, ld, e};
    om_fk_unitree_h1(q, qd, S);
  }
}

}  // namespace om

using namespace om;

struct OmTraj {
  TrajDev d;
  // Scratch of the time-parallel playback kernel (episode-start snapshot of the carried state), grown on demand
  // and kept: a stream-ordered allocation per call cost more than the kernel whenever the pool had been trimmed.
  // One playback call per handle may be in flight at a time.
  mutable char* scratch = nullptr;
  mutable size_t scratch_bytes = 0;
  // fork / join of the observation-moments kernel beside the time-parallel playback kernel (both only READ the snapshot)
  mutable cudaStream_t side = nullptr;
  mutable cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

extern "C" int om_traj_create(const double* table, int K, int n_traj, int T, OmTraj** out) {
  OM_REQUIRE(table && out, "om_traj_create: null argument");
  OM_REQUIRE(K >= 2 && K <= 64 && n_traj >= 1 && T >= 1, "om_traj_create: bad shape K=%d n_traj=%d T=%d", K, n_traj, T);
  const int kpad = (K + 3) / 4 * 4;
  std::vector<float> rows((size_t)n_traj * T * kpad, 0.f);
  std::vector<double> xy((size_t)n_traj * T * 2);
  for (int k = 0; k < K; ++k)
    for (int tr = 0; tr < n_traj; ++tr)
      for (int s = 0; s < T; ++s) {
        const double v = table[((size_t)k * n_traj + tr) * T + s];
        rows[((size_t)tr * T + s) * kpad + k] = (float)v;
        if (k < 2) xy[((size_t)tr * T + s) * 2 + k] = v;
      }
  // exclusive prefix sums of the (fp32-rounded, as stored) velocity channels, accumulated in float64
  const int nq = K / 2;
  std::vector<double> cdq((size_t)n_traj * (T + 1) * nq, 0.0);
  for (int tr = 0; tr < n_traj; ++tr)
    for (int s = 0; s < T; ++s)
      for (int k = 0; k < nq; ++k)
        cdq[((size_t)tr * (T + 1) + s + 1) * nq + k] =
            cdq[((size_t)tr * (T + 1) + s) * nq + k] + (double)rows[((size_t)tr * T + s) * kpad + nq + k];
  // exclusive prefix sums of the emitted observation rows (channels 2 .. K-1, fp32-rounded as stored) and of their squares,
  // float64: psum[tr][i][c] = sum_{j<i} row_j[c+2], psum[tr][i][32+c] = sum_{j<i} row_j[c+2]^2  (34-key tables only)
  std::vector<double> psum;
  if (K == 34) {
    psum.assign((size_t)n_traj * (T + 1) * 64, 0.0);
    for (int tr = 0; tr < n_traj; ++tr)
      for (int s = 0; s < T; ++s)
        for (int c = 0; c < 32; ++c) {
          const double v = (double)rows[((size_t)tr * T + s) * kpad + c + 2];
          const size_t at = ((size_t)tr * (T + 1) + s) * 64;
          psum[at + 64 + c] = psum[at + c] + v;
          psum[at + 64 + 32 + c] = psum[at + 32 + c] + v * v;
        }
  }
  OmTraj* t = new OmTraj();
  t->d.K = K; t->d.kpad = kpad; t->d.n_traj = n_traj; t->d.T = T;
  float* drows = nullptr;
  double* dxy = nullptr;
  double* dcdq = nullptr;
  double* dpsum = nullptr;
  cudaError_t e = cudaMalloc(&drows, rows.size() * sizeof(float));
  if (e == cudaSuccess && !psum.empty()) e = cudaMalloc(&dpsum, psum.size() * sizeof(double));
  if (e == cudaSuccess && !psum.empty()) e = cudaMemcpy(dpsum, psum.data(), psum.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&dxy, xy.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dcdq, cdq.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(drows, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dxy, xy.data(), xy.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dcdq, cdq.data(), cdq.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (drows) cudaFree(drows);
    if (dxy) cudaFree(dxy);
    if (dcdq) cudaFree(dcdq);
    if (dpsum) cudaFree(dpsum);
    delete t;
    return fail("om_traj_create: device upload failed: %s (no CPU path)", cudaGetErrorString(e));
  }
  t->d.rows = drows;
  t->d.xy = dxy;
  t->d.cdq = dcdq;
  t->d.psum = dpsum;
  *out = t;
  return 0;
}

extern "C" void om_traj_destroy(OmTraj* t) {
  if (!t) return;
  cudaFree((void*)t->d.rows);
  cudaFree((void*)t->d.xy);
  cudaFree((void*)t->d.cdq);
  if (t->d.psum) cudaFree((void*)t->d.psum);
  if (t->scratch) cudaFree(t->scratch);
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  if (t->ev_join) cudaEventDestroy(t->ev_join);
  if (t->side) cudaStreamDestroy(t->side);
  delete t;
}

extern "C" int om_traj_reset(const OmTraj* t, uint64_t seed, uint32_t env_id0, const uint8_t* mask,
                             const int32_t* forced_traj, const int32_t* forced_step, int32_t* traj_no, int32_t* step_no,
                             uint32_t* reset_count, double* xy_off, float* sample, int n, int ld, void* stream) {
  OM_REQUIRE(t, "om_traj_reset: null table");
  OM_REQUIRE(n >= 0 && ld >= n, "om_traj_reset: need 0 <= n <= ld");
  if (n == 0) return 0;
  OM_REQUIRE(traj_no && step_no && reset_count && xy_off, "om_traj_reset: null state");
  traj_reset_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(t->d, seed, env_id0, mask, forced_traj, forced_step,
                                                                         traj_no, step_no, reset_count, xy_off, sample, n, ld);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_traj_current(const OmTraj* t, const int32_t* traj_no, const int32_t* step_no, const double* xy_off,
                               float* sample, int n, int ld, void* stream) {
  OM_REQUIRE(t, "om_traj_current: null table");
  OM_REQUIRE(n >= 0 && ld >= n, "om_traj_current: need 0 <= n <= ld");
  if (n == 0) return 0;
  OM_REQUIRE(traj_no && step_no && xy_off && sample, "om_traj_current: null argument");
  traj_current_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(t->d, traj_no, step_no, xy_off, sample, n, ld);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_traj_next(const OmTraj* t, uint64_t seed, uint32_t env_id0, int auto_reset, int32_t* traj_no,
                            int32_t* step_no, uint32_t* reset_count, double* xy_off, float* sample, uint8_t* wrapped,
                            int n, int ld, void* stream) {
  OM_REQUIRE(t, "om_traj_next: null table");
  OM_REQUIRE(n >= 0 && ld >= n, "om_traj_next: need 0 <= n <= ld");
  if (n == 0) return 0;
  OM_REQUIRE(traj_no && step_no && reset_count && xy_off, "om_traj_next: null state");
  traj_next_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(t->d, seed, env_id0, auto_reset, traj_no, step_no, reset_count,
                                                                        xy_off, sample, wrapped, n, ld);
  OM_LAUNCHED();
  return 0;
}

static int play_impl(const OmModel* m, const OmH1Spec* spec, const OmTraj* t, uint64_t seed, uint32_t env_id0, double dt,
                     int n_steps, int end_episode_reset, const OmPlayState* state, const OmPlayOut* out, int n, int ld,
                     void* stream, int forced) {
  OM_REQUIRE(m && spec && t && state && out, "om_h1_play_from_velocity: null argument");
  OM_REQUIRE(n >= 0 && ld >= n && n_steps >= 0, "om_h1_play_from_velocity: bad sizes");
  OM_REQUIRE(m->specialised == SPEC_H1, "om_h1_play_from_velocity: model is not the UnitreeH1 (arms disabled) model");
  OM_REQUIRE(spec->n_obs_q == 17 && t->d.K == 34, "om_h1_play_from_velocity: expects the 34-key H1 trajectory");
  for (int k = 0; k < 17; ++k)
    OM_REQUIRE(spec->obs_perm[k] == OM_H1_PERM_HOST[k], "om_h1_play_from_velocity: observation spec differs from UnitreeH1's");
  if (n == 0) return 0;
  OM_REQUIRE(state->traj_no && state->step_no && state->reset_count && state->xy_off && (forced || state->curr_qpos) &&
                 state->pending && state->prev_x_vel, "om_h1_play_from_velocity: incomplete state");
  PlayArgs a;
  a.t = t->d; a.seed = seed; a.env_id0 = env_id0; a.dt = dt; a.target = spec->target_velocity;
  a.n_steps = n_steps; a.end_reset = end_episode_reset & 1;
  const bool start_reset = (end_episode_reset & 2) != 0;          // OM_PLAY_START_RESET
  a.n = n; a.ld = ld; a.s = *state; a.snap = *state; a.o = *out; a.forced = forced;
  a.obs_moments = out->obs_moments;
  constexpr int BLOCK = PLAY_BLOCK;
  const bool mom = a.obs_moments != nullptr;
  OM_REQUIRE(!mom || t->d.psum, "om_h1_play_from_velocity: this trajectory handle has no prefix tables");
  // time-parallel when the env count alone cannot fill the machine (148 SMs x 2048 threads)
  const long long target_threads = 148LL * 2048;
  const long long chunks_wanted = target_threads / n > 0 ? target_threads / n : 1;
  int chunk = ceil_div(n_steps, chunks_wanted);
  if (g_knobs.play_chunk > 0) chunk = g_knobs.play_chunk;            // tuning / test hook (om_debug_set)
  if (chunk < 1) chunk = 1;
  if (chunk >= n_steps) {
    if (mom) {                                      // reads the START state: before anything below touches it
      play_moments_kernel<<<ceil_div(n, 8 * PM_ENVS), 256, 0, (cudaStream_t)stream>>>(a, a.s, start_reset ? 1 : 0);
      OM_LAUNCHED();
    }
    if (start_reset) {
      play_start_reset_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a);
      OM_LAUNCHED();
    }
    play_h1_seq_kernel<BLOCK><<<ceil_div(n, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(a);
  } else {
    // the carried state is read by every chunk and overwritten by the last one: snapshot the episode-start
    // state (stream-ordered scratch) so that no thread can observe another's end-of-episode write
    cudaStream_t st = (cudaStream_t)stream;
    const size_t L = (size_t)ld;
    const size_t bytes = L * (4 + 4 + 4 + 16 + 4 + 17 * 8 + 34 * 4) + 256;
    if (t->scratch_bytes < bytes) {
      if (t->scratch) OM_CUDA_OK(cudaFree(t->scratch));       // synchronises: no earlier call still reads it
      t->scratch = nullptr;
      t->scratch_bytes = 0;
      OM_CUDA_OK(cudaMalloc((void**)&t->scratch, bytes));
      t->scratch_bytes = bytes;
    }
    char* p = t->scratch;
    a.snap.curr_qpos = (double*)p; p += L * 17 * 8;
    a.snap.xy_off = (double*)p; p += L * 16;
    a.snap.pending = (float*)p; p += L * 34 * 4;
    a.snap.prev_x_vel = (float*)p; p += L * 4;
    a.snap.traj_no = (int32_t*)p; p += L * 4;
    a.snap.step_no = (int32_t*)p; p += L * 4;
    a.snap.reset_count = (uint32_t*)p;
    if (start_reset) play_snapshot_kernel<true><<<ceil_div(n, 256), 256, 0, st>>>(a, a.s, a.snap, n, ld);
    else play_snapshot_kernel<false><<<ceil_div(n, 256), 256, 0, st>>>(a, a.s, a.snap, n, ld);
    OM_LAUNCHED();
    dim3 grid(ceil_div(n, BLOCK), ceil_div(n_steps, chunk));
    if (mom) {
      // The moments kernel (L2-bound: prefix-table reads) runs BESIDE the playback kernel (DRAM-bound) on a side stream:
      // both only read the snapshot, which now holds the state after the call's start reset.  Plain fork / join with two
      // events, legal under stream capture.
      if (!t->side) {
        OM_CUDA_OK(cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking));
        OM_CUDA_OK(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
        OM_CUDA_OK(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
      }
      OM_CUDA_OK(cudaEventRecord(t->ev_fork, st));
      OM_CUDA_OK(cudaStreamWaitEvent(t->side, t->ev_fork, 0));
      play_moments_kernel<<<ceil_div(n, 8 * PM_ENVS), 256, 0, t->side>>>(a, a.snap, 0);
      OM_LAUNCHED();
      OM_CUDA_OK(cudaEventRecord(t->ev_join, t->side));
    }
    OM_CUDA_OK(launch_pdl(play_h1_tp_kernel<BLOCK>, grid, dim3(BLOCK), 0, st, a, chunk));
    OM_LAUNCHED();
    if (mom) OM_CUDA_OK(cudaStreamWaitEvent(st, t->ev_join, 0));
    return 0;
  }
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_h1_play_from_velocity(const OmModel* m, const OmH1Spec* spec, const OmTraj* t, uint64_t seed,
                                        uint32_t env_id0, double dt, int n_steps, int end_episode_reset,
                                        const OmPlayState* state, const OmPlayOut* out, int n, int ld, void* stream) {
  return play_impl(m, spec, t, seed, env_id0, dt, n_steps, end_episode_reset, state, out, n, ld, stream, 0);
}

extern "C" int om_h1_play_trajectory(const OmModel* m, const OmH1Spec* spec, const OmTraj* t, uint64_t seed, uint32_t env_id0,
                                     int n_steps, int end_episode_reset, const OmPlayState* state, const OmPlayOut* out, int n,
                                     int ld, void* stream) {
  return play_impl(m, spec, t, seed, env_id0, 0.0, n_steps, end_episode_reset, state, out, n, ld, stream, 1);
}

extern "C" int om_h1_live_step(const OmModel* m, const OmH1Spec* spec, const OmTraj* t, uint64_t seed, uint32_t env_id0,
                               const OmLiveState* state, const OmLiveOut* out, int n, int ld, void* stream) {
  OM_REQUIRE(m && spec && t && state && out, "om_h1_live_step: null argument");
  OM_REQUIRE(n >= 0 && ld >= n, "om_h1_live_step: need 0 <= n <= ld (n=%d ld=%d)", n, ld);
  OM_REQUIRE(m->specialised == SPEC_H1, "om_h1_live_step: model is not the UnitreeH1 (arms disabled) model");
  OM_REQUIRE(spec->n_obs_q == 17 && t->d.K == 34, "om_h1_live_step: expects the 34-key H1 trajectory");
  OM_REQUIRE(spec->x_vel_idx == 15, "om_h1_live_step: x_vel_idx must be 15 (dq_pelvis_tx of UnitreeH1's own spec)");
  for (int k = 0; k < 17; ++k)
    OM_REQUIRE(spec->obs_perm[k] == OM_H1_PERM_HOST[k], "om_h1_live_step: observation spec differs from UnitreeH1's");
  if (n == 0) return 0;
  OM_REQUIRE(state->traj_no && state->step_no && state->reset_count && state->xy_off && state->prev_x_vel,
             "om_h1_live_step: incomplete state");
  LiveArgs a;
  a.t = t->d; a.seed = seed; a.env_id0 = env_id0; a.target = spec->target_velocity;
  a.use_absorbing = spec->use_absorbing_states; a.n = n; a.ld = ld;
  a.traj_no = state->traj_no; a.step_no = state->step_no; a.reset_count = state->reset_count; a.xy_off = state->xy_off;
  a.prev_x_vel = state->prev_x_vel;
  a.qpos = out->qpos; a.qvel = out->qvel; a.xpos = out->xpos; a.xquat = out->xquat; a.site_xpos = out->site_xpos;
  a.cvel = out->cvel; a.obs = out->obs; a.reward = out->reward; a.absorbing = out->absorbing; a.wrapped = out->wrapped;
  constexpr int BLOCK = 128;
  if (n <= 524288) OM_CUDA_OK(launch_pdl(h1_live_step_kernel<BLOCK, true>, dim3(ceil_div(n, BLOCK)), dim3(BLOCK), 0, (cudaStream_t)stream, a));
  else OM_CUDA_OK(launch_pdl(h1_live_step_kernel<BLOCK, false>, dim3(ceil_div(n, BLOCK)), dim3(BLOCK), 0, (cudaStream_t)stream, a));
  OM_LAUNCHED();
  return 0;
}
