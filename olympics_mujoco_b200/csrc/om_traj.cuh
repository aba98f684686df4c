// Device view of a reference-trajectory table (K3) and the gather / reset helpers shared by kernels.
#pragma once
#include "om_common.cuh"

namespace om {

struct TrajDev {
  const float* rows;   // [n_traj][T][kpad] sample-major rows (one 16-byte aligned row per sample)
  const double* xy;    // [n_traj][T][2]   channels 0,1 (root x, y) in float64
  const double* cdq;   // [n_traj][T+1][K/2] exclusive prefix sums of the velocity channels (float64):
                       //   cdq[tr][i][k] = sum_{j<i} rows[tr][j][K/2+k]  (time-parallel playback)
  const double* psum;  // [n_traj][T+1][64] exclusive prefix sums of the observation rows (channels 2..33) and of their
                       //   squares, float64 (S1 of a playback call as prefix differences); null unless K == 34
  int K, kpad, n_traj, T;
};

// reset_trajectory draws (trajectory.py:304,311) under the Philox contract
OM_HD void traj_draw(const TrajDev& t, uint64_t seed, uint32_t env_id, uint32_t reset_count, int& traj_no, int& step_no) {
  const U4 w = om_draw(seed, env_id, reset_count, 0u);
  traj_no = to_int(w.x, t.n_traj);
  step_no = to_int(w.y, t.T);
}

// one sample row into registers (kpad <= 36 for the H1 table); vectorised 16-byte gathers
OM_HD void traj_load_row(const TrajDev& t, int traj_no, int step_no, float (&samp)[36]) {
  const float4* r = reinterpret_cast<const float4*>(t.rows + ((size_t)traj_no * t.T + step_no) * t.kpad);
#pragma unroll
  for (int v = 0; v < 9; ++v) {
    const float4 x = __ldg(r + v);
    samp[4 * v] = x.x; samp[4 * v + 1] = x.y; samp[4 * v + 2] = x.z; samp[4 * v + 3] = x.w;
  }
}

// get_current_sample (trajectory.py:381-387) for one env into a [K][ld] SoA array
OM_HD void traj_write_sample(const TrajDev& t, int traj_no, int step_no, double ox, double oy, float* sample, int ld, int env) {
  const size_t row = (size_t)traj_no * t.T + step_no;
  sample[env] = (float)(t.xy[row * 2] - ox);
  sample[(size_t)ld + env] = (float)(t.xy[row * 2 + 1] - oy);
  const float* r = t.rows + row * t.kpad;
  for (int k = 2; k < t.K; ++k) sample[(size_t)k * ld + env] = __ldg(r + k);
}

}  // namespace om
