// N1 (the step immediately BEFORE the path): action de-normalisation and the PD law that feed mj_step.
//   LocoEnvBase._preprocess_action   loco_env_base.py:1050-1069   ctrl = action * delta + mean
//   JVRC.step / do_simulation        environments/robot.py:88-115  target = action + motor_offset;  ctrl = tau / gear
//   MujocoRobotInterface.step_pd     interfaces/mujoco_robot_interface.py:425-443  tau = kp (p - q) + kv (v - dq)
// Element-wise SoA passes (one thread per env, nu <= 32 actuators unrolled at run time); HBM-bound, 12 B/actuator.
#include "om_common.cuh"

namespace om {

__global__ void __launch_bounds__(256) action_affine_kernel(OmActionSpec sp, const float* __restrict__ action, int n, int ld,
                                                            float* __restrict__ ctrl) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  for (int a = 0; a < sp.nu; ++a) ctrl[(size_t)a * ld + e] = fmaf(action[(size_t)a * ld + e], sp.delta[a], sp.mean[a]);
}

__global__ void __launch_bounds__(256) pd_torque_kernel(OmPdSpec sp, const float* __restrict__ target,
                                                        const float* __restrict__ vel_target, const float* __restrict__ qpos,
                                                        const float* __restrict__ qvel, int add_offset, int n, int ld,
                                                        float* __restrict__ ctrl) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  for (int a = 0; a < sp.nu; ++a) {
    const float p = target[(size_t)a * ld + e] + (add_offset ? sp.offset[a] : 0.f);
    const float v = vel_target ? vel_target[(size_t)a * ld + e] : 0.f;
    const float perr = p - qpos[(size_t)sp.qposadr[a] * ld + e];
    const float verr = v - qvel[(size_t)sp.dofadr[a] * ld + e];
    ctrl[(size_t)a * ld + e] = (sp.kp[a] * perr + sp.kd[a] * verr) / sp.gear[a];
  }
}

}  // namespace om

using namespace om;

extern "C" int om_action_affine(const OmActionSpec* spec, const float* action, int n, int ld, float* ctrl, void* stream) {
  OM_REQUIRE(spec && spec->nu >= 0 && spec->nu <= 32, "om_action_affine: need 0 <= nu <= 32");
  OM_REQUIRE(n >= 0 && ld >= n, "om_action_affine: need 0 <= n <= ld");
  if (n == 0 || spec->nu == 0) return 0;
  OM_REQUIRE(action && ctrl, "om_action_affine: null argument");
  action_affine_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(*spec, action, n, ld, ctrl);
  OM_LAUNCHED();
  return 0;
}

extern "C" int om_pd_torque(const OmPdSpec* spec, const float* target, const float* vel_target, const float* qpos,
                            const float* qvel, int add_offset, int n, int ld, float* ctrl, void* stream) {
  OM_REQUIRE(spec && spec->nu >= 0 && spec->nu <= 32, "om_pd_torque: need 0 <= nu <= 32");
  OM_REQUIRE(n >= 0 && ld >= n, "om_pd_torque: need 0 <= n <= ld");
  if (n == 0 || spec->nu == 0) return 0;
  OM_REQUIRE(target && qpos && qvel && ctrl, "om_pd_torque: null argument");
  for (int a = 0; a < spec->nu; ++a)
    OM_REQUIRE(spec->gear[a] != 0.f && spec->qposadr[a] >= 0 && spec->dofadr[a] >= 0, "om_pd_torque: bad actuator %d", a);
  pd_torque_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(*spec, target, vel_target, qpos, qvel, add_offset, n, ld, ctrl);
  OM_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------- N3: mirror-symmetry transforms
//   SymmetricEnv.mirror_observation / mirror_action / mirror_clock_observation   rl/envs/wrappers.py:51-72
//   _get_symmetry_matrix :75-82: mat[i, |m_i|] = sign(m_i)  =>  (x @ mat)[|m_i|] = sign(m_i) x[i]: a signed permutation,
//   applied here as a scatter over SoA rows; clock rows c become sin(arcsin(y_c) + pi) = -y_c (:67-69).
namespace om {
__global__ void __launch_bounds__(256) mirror_kernel(OmMirrorSpec sp, const float* __restrict__ x, int n, int ld,
                                                     float* __restrict__ y) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
#pragma unroll 8
  for (int i = 0; i < sp.numel; ++i) {                  // (the loads do not depend on the stores: x and y never alias)
    const float v = x[(size_t)i * ld + e] * sp.sign[i];
    y[(size_t)sp.index[i] * ld + e] = sp.negate[sp.index[i]] ? -v : v;
  }
}
}  // namespace om

extern "C" int om_mirror(const OmMirrorSpec* spec, const float* x, int n, int ld, float* y, void* stream) {
  OM_REQUIRE(spec && spec->numel >= 0 && spec->numel <= 64, "om_mirror: need 0 <= numel <= 64");
  OM_REQUIRE(n >= 0 && ld >= n, "om_mirror: need 0 <= n <= ld");
  if (n == 0 || spec->numel == 0) return 0;
  OM_REQUIRE(x && y && x != y, "om_mirror: null or aliased argument (the scatter is not in-place safe)");
  uint64_t seen = 0;
  for (int i = 0; i < spec->numel; ++i) {
    OM_REQUIRE(spec->index[i] >= 0 && spec->index[i] < spec->numel, "om_mirror: index[%d] out of range", i);
    seen |= 1ull << spec->index[i];
  }
  OM_REQUIRE(seen == (spec->numel == 64 ? ~0ull : (1ull << spec->numel) - 1), "om_mirror: indices are not a permutation");
  om::mirror_kernel<<<om::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(*spec, x, n, ld, y);
  OM_LAUNCHED();
  return 0;
}
