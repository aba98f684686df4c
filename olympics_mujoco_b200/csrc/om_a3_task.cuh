// K2 (A3 flavour): per-env device functions of the StickFigureA3 RL step tail -- everything
// StickFigureA3.step does after robot.step() (= mj_step, outside this path):
//   WalkingTask.step          tasks/walking_task.py:246-293  (+ update_target_steps :228-244, update_goal_steps :184-225)
//   WalkingTask.calc_reward   :74-110, step_reward :56-72; terms tasks/rewards.py:27-40, :65-83, :85-102, :121-126
//   WalkingTask.done          :298-319
//   StickFigureA3.get_obs     real_humanoid_robots/StickFigureA3.py:144-178
//   reset_model + task.reset  StickFigureA3.py:205-235, walking_task.py:321-397 (:137-182, :113-135)
// The body/site quantities the task reads through MujocoRobotInterface (interfaces/mujoco_robot_interface.py
// :299-346) are captured straight out of the generated FK by A3Sink, in registers; nothing the task does not
// read is computed (the compiler removes the arms and the COM pass: foot velocities are taken about the root
// origin, v(xpos) = v_P + w x (xpos - P), which is what mj_objectVelocity(mjOBJ_XBODY) evaluates to).
// Host-compilable (OM_HD pre-defined) so tests/host/a3_host_harness.cpp can check it without a GPU.
#pragma once
#include "om_math.cuh"
#include "gen/a3_ids.h"
#include "gen/fk_stick_figure_a3.cuh"
#include "gen/fk_pos_stick_figure_a3.cuh"
#include "gen/fk_pos_f64_stick_figure_a3.cuh"

namespace om {

constexpr int A3_NQ = 25, A3_NV = 24, A3_NOBS = 41, A3_MAX_STEPS = 20, A3_NINT = 7, A3_LUT_COLS = 6, A3_NU = 60;
enum { A3I_PHASE = 0, A3I_T1, A3I_T2, A3I_FRAMES, A3I_MODE, A3I_SEQLEN, A3I_REACHED };
enum { A3_STANDING = 0, A3_FORWARD = 1 };
constexpr uint32_t A3_RESET_STREAM = 16;      // oracle/philox.py STREAM_A3_RESET

struct A3TaskConst {
  int period, delay_frames;
  float fmax, vmax;                   // rewards.py:66 (mass*9.8*0.5), :87 (0.2)
  float inv_fmax, inv_vmax;           // their fp32 reciprocals (x / xmax as one multiply)
  double target_radius;               // walking_task.py:333
  float near_d2;                      // smallest fp32 d2 with (double)sqrtf(d2) >= target_radius (a3_near_d2)
  float near_lo2, near_hi2;           // (radius -+ A3_BAND)^2 rounded outwards: outside [lo2, hi2] the fp32 test decides
  double goal_height_ref, deadzone;   // StickFigureA3.py:110; rewards.py:36 (0.01 + 0.05*goal_speed_ref)
  const float* lut;                   // [period][6]: r_frc, r_vel, l_frc, l_vel clocks, sin/cos(2 pi phase/period)
};

struct A3Feat {
  V3 root_p, head_p, lfoot_p, rfoot_p, lsite, rsite;
  Q4 root_q;
  V3 lw, lv, rw, rv;                  // foot body spatial velocity [w; v] about the root origin
};

struct NullFkSink {
  static constexpr bool want_site_xmat = false;
  OM_HD void xpos(int, float, float, float) const {}
  OM_HD void xquat(int, float, float, float, float) const {}
  OM_HD void site_xpos(int, float, float, float) const {}
  OM_HD void site_xmat(int, float, float, float, float, float, float, float, float, float) const {}
  OM_HD void cvel(int, float, float, float, float, float, float) const {}
  OM_HD void com(float, float, float) const {}
  OM_HD void vel_p(int, float, float, float, float, float, float) const {}
};

// Captures what the task reads; forwards everything to Inner (NullFkSink, or an HBM sink when the caller also
// wants the MjData fields).  Body / site indices are literals at every call site, so the branches fold.
template <class Inner>
struct A3Sink {
  static constexpr bool want_site_xmat = Inner::want_site_xmat;
  Inner inner;
  A3Feat f;
  OM_HD void xpos(int b, float x, float y, float z) {
    inner.xpos(b, x, y, z);
    if (b == OM_A3_ROOT) f.root_p = V3{x, y, z};
    else if (b == OM_A3_HEAD) f.head_p = V3{x, y, z};
    else if (b == OM_A3_LFOOT) f.lfoot_p = V3{x, y, z};
    else if (b == OM_A3_RFOOT) f.rfoot_p = V3{x, y, z};
  }
  OM_HD void xquat(int b, float w, float x, float y, float z) {
    inner.xquat(b, w, x, y, z);
    if (b == OM_A3_ROOT) f.root_q = Q4{w, x, y, z};
  }
  OM_HD void site_xpos(int s, float x, float y, float z) {
    inner.site_xpos(s, x, y, z);
    if (s == OM_A3_LSITE) f.lsite = V3{x, y, z};
    else if (s == OM_A3_RSITE) f.rsite = V3{x, y, z};
  }
  OM_HD void site_xmat(int s, float a, float b, float c, float d, float e, float g, float h, float i, float j) {
    inner.site_xmat(s, a, b, c, d, e, g, h, i, j);
  }
  OM_HD void cvel(int b, float wx, float wy, float wz, float vx, float vy, float vz) { inner.cvel(b, wx, wy, wz, vx, vy, vz); }
  OM_HD void com(float x, float y, float z) { inner.com(x, y, z); }
  OM_HD void vel_p(int b, float wx, float wy, float wz, float vx, float vy, float vz) {
    if (b == OM_A3_LFOOT) { f.lw = V3{wx, wy, wz}; f.lv = V3{vx, vy, vz}; }
    else if (b == OM_A3_RFOOT) { f.rw = V3{wx, wy, wz}; f.rv = V3{vx, vy, vz}; }
  }
};

// ---------------------------------------------------------------- exact threshold decisions (A10 target_reached, A12 done)
// The reference takes both decisions in float64 on MuJoCo's float64 site positions (walking_task.py:266-283, :298-319).
// Here the site positions come out of an fp32 chain (error ~1e-6 m), so a decision whose fp32 margin is below A3_BAND
// (1e-5 m, ten times the chain's error) is RE-EVALUATED from a float64 forward pass of the same fp32 inputs
// (gen/fk_pos_f64_*.cuh, generated from the same tables): the flags are then functions of the float64 arithmetic alone,
// i.e. bit-identical to the float64 oracle on the same (fp32-representable) qpos / contact / sequence inputs.  About one
// env-step in 10^4 takes the slow path; it is one out-of-line call.
constexpr double A3_BAND = 1e-5;
struct A3SiteSinkF64 {
  static constexpr bool want_site_xmat = false;
  double ls[3], rs[3];
  OM_HD void xpos(int, double, double, double) {}
  OM_HD void xquat(int, double, double, double, double) {}
  OM_HD void site_xpos(int s, double x, double y, double z) {
    if (s == OM_A3_LSITE) { ls[0] = x; ls[1] = y; ls[2] = z; }
    else if (s == OM_A3_RSITE) { rs[0] = x; rs[1] = y; rs[2] = z; }
  }
  OM_HD void vel_p(int, double, double, double, double, double, double) {}
};
struct A3SitesF64 { double ls[3], rs[3]; };
// q: component k of the env-step's qpos at q[k * ld] (the kernels pass the global SoA pointer: the slow path re-reads its 25
// inputs instead of keeping the caller's register copy alive and addressable)
OM_NOINLINE A3SitesF64 a3_sites_f64(const float* q, size_t ld) {
  double qd[A3_NQ], v0[A3_NV];
#pragma unroll
  for (int k = 0; k < A3_NQ; ++k) qd[k] = (double)q[k * ld];
#pragma unroll
  for (int k = 0; k < A3_NV; ++k) v0[k] = 0.0;
  A3SiteSinkF64 S{};
  om_fk_pos_f64_stick_figure_a3(qd, v0, S);
  A3SitesF64 r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.ls[k] = S.ls[k]; r.rs[k] = S.rs[k]; }
  return r;
}
// Where the slow path finds the env-step's qpos -- or, with `defer` (the (env, t)-parallel replay kernel, whose hot
// path must stay free of an out-of-line call and of its register saves), only a note that this env-step needs one: the
// fp32 decision is returned for now, the env-step's byte is flagged and the sequential pass re-takes it in float64.
struct A3Exact {
  const float* q;
  size_t ld;
  bool defer;
  bool unsure;
  OM_HD A3SitesF64 get() const { return a3_sites_f64(q, ld); }
};
// done (:298-319): root z (= qpos[2], exact) - lowest foot-site z < 0.6
OM_HD bool a3_done_height(float root_z, float lz, float rz, A3Exact& ex) {
  const double h = (double)root_z - (double)fminf(lz, rz);
  if (fabs(h - 0.6) > A3_BAND) return h < 0.6;
  if (ex.defer) { ex.unsure = true; return h < 0.6; }
  const A3SitesF64 s = ex.get();
  return (double)root_z - fmin(s.ls[2], s.rs[2]) < 0.6;
}
// "a foot is within target_radius of p" (:266-269): np.linalg.norm(foot - target) < radius, either foot
OM_HD bool a3_near_exact(const A3TaskConst& C, V3 lsite, V3 rsite, V3 p, A3Exact& ex) {
  const V3 a = lsite - p, b = rsite - p;
  const float dl2 = dot(a, a), dr2 = dot(b, b);
  if (dl2 < C.near_lo2 || dr2 < C.near_lo2) return true;          // certainly inside
  if (dl2 > C.near_hi2 && dr2 > C.near_hi2) return false;         // certainly outside
  if (ex.defer) { ex.unsure = true; return dl2 < C.near_d2 || dr2 < C.near_d2; }
  const A3SitesF64 s = ex.get();
  const double px = (double)p.x, py = (double)p.y, pz = (double)p.z;
  const double lx = s.ls[0] - px, ly = s.ls[1] - py, lz = s.ls[2] - pz;
  const double rx = s.rs[0] - px, ry = s.rs[1] - py, rz = s.rs[2] - pz;
  return sqrt(lx * lx + ly * ly + lz * lz) < C.target_radius || sqrt(rx * rx + ry * ry + rz * rz) < C.target_radius;
}
inline void a3_near_band(double radius, float* lo2, float* hi2) {     // host: task creation, test harness
  float lo = (float)((radius - A3_BAND) * (radius - A3_BAND)), hi = (float)((radius + A3_BAND) * (radius + A3_BAND));
  while ((double)lo > (radius - A3_BAND) * (radius - A3_BAND)) lo = nextafterf(lo, 0.f);
  while ((double)hi < (radius + A3_BAND) * (radius + A3_BAND)) hi = nextafterf(hi, INFINITY);
  *lo2 = radius > A3_BAND ? lo : 0.f;
  *hi2 = hi;
}

struct A3TaskRegs { int phase, t1, t2, frames, mode, seq_len, reached; };

OM_HD float norm3(V3 a) { return sqrtf(dot(a, a)); }

// What WalkingTask.step / calc_reward / done read per env-step, reduced to 17 floats
struct A3TaskIn {
  V3 root_p; Q4 root_q; float head_x, head_y; V3 lsite, rsite; float lvel_n, rvel_n;
};
OM_HD A3TaskIn a3_task_in(const A3Feat& f) {
  A3TaskIn t;
  t.root_p = f.root_p; t.root_q = f.root_q; t.head_x = f.head_p.x; t.head_y = f.head_p.y;
  t.lsite = f.lsite; t.rsite = f.rsite;
  t.lvel_n = norm3(f.lv + cross(f.lw, f.lfoot_p - f.root_p));     // mj_objectVelocity(XBODY), linear part, at xpos
  t.rvel_n = norm3(f.rv + cross(f.rw, f.rfoot_p - f.root_p));
  return t;
}
// sequence[t1], sequence[t2] kept in registers: they change only when a target is reached
struct A3Targets { V3 p1; float th1; V3 p2; float th2; };
template <class Seq>
OM_HD A3Targets a3_targets_load(const A3TaskRegs& s, const Seq& seq) {
  return A3Targets{V3{seq(s.t1, 0), seq(s.t1, 1), seq(s.t1, 2)}, seq(s.t1, 3), V3{seq(s.t2, 0), seq(s.t2, 1), seq(s.t2, 2)},
                   seq(s.t2, 3)};
}

// transforms3d quaternions.quat2mat (scale-invariant: s = 2/|q|^2), rows 0..2
OM_HD void tf3_quat2mat(Q4 q, float (&m)[9]) {
  const float nq = fmaf(q.w, q.w, fmaf(q.x, q.x, fmaf(q.y, q.y, q.z * q.z)));
  if (nq < 2.220446e-16f) {            // transforms3d: Nq < float64 eps -> identity
    m[0] = 1.f; m[1] = 0.f; m[2] = 0.f; m[3] = 0.f; m[4] = 1.f; m[5] = 0.f; m[6] = 0.f; m[7] = 0.f; m[8] = 1.f;
    return;
  }
  const float s = 2.0f * om_rcp(nq);
  const float X = q.x * s, Y = q.y * s, Z = q.z * s;
  const float wX = q.w * X, wY = q.w * Y, wZ = q.w * Z, xX = q.x * X, xY = q.x * Y, xZ = q.x * Z;
  const float yY = q.y * Y, yZ = q.y * Z, zZ = q.z * Z;
  m[0] = 1.0f - (yY + zZ); m[1] = xY - wZ; m[2] = xZ + wY;
  m[3] = xY + wZ; m[4] = 1.0f - (xX + zZ); m[5] = yZ - wX;
  m[6] = xZ - wY; m[7] = yZ + wX; m[8] = 1.0f - (xX + yY);
}

constexpr float A3_EPS4 = 8.8817842e-16f;     // transforms3d euler._EPS4 (4 * float64 eps)

// StickFigureA3.get_obs rows 0..3: euler2quat(roll, pitch, 0) of quat2euler(qpos[3:7]) (axes 'sxyz'), literally
OM_NOINLINE Q4 a3_root_orient_trig(float qw, float qx, float qy, float qz) {
  float m[9];
  tf3_quat2mat(Q4{qw, qx, qy, qz}, m);
  const float cy = sqrtf(fmaf(m[0], m[0], m[3] * m[3]));
  const float roll = cy > A3_EPS4 ? atan2f(m[7], m[8]) : atan2f(-m[5], m[4]);
  const float pitch = atan2f(-m[6], cy);
  float si, ci, sj, cj;
  sincosf(0.5f * roll, &si, &ci);
  sincosf(0.5f * pitch, &sj, &cj);
  return Q4{cj * ci, cj * si, sj * ci, -(sj * si)};
}
// The same quantity without inverse trigonometry: q = qz(yaw) qy(pitch) qx(roll) up to sign, so the observation is
// conj(qz(yaw)) q with yaw = atan2(M10, M00); cos/sin of yaw/2 follow from (M00, M10)/cy by the half-angle formulas,
// and the sign is the one that makes w = cos(pitch/2) cos(roll/2) non-negative.  Near gimbal lock (cy -> 0) the yaw
// split is ill-conditioned and the literal path is taken.
OM_HD void a3_root_orient(float qw, float qx, float qy, float qz, float* o) {
  const float nq = fmaf(qw, qw, fmaf(qx, qx, fmaf(qy, qy, qz * qz)));
  const float s2 = 2.0f * om_rcp(nq);
  const float m00 = 1.0f - (qy * qy + qz * qz) * s2, m10 = (qx * qy + qw * qz) * s2;
  const float cy2 = fmaf(m00, m00, m10 * m10);
  if (!(cy2 > 1e-4f) || !(nq > 1e-12f)) {
    const Q4 r = a3_root_orient_trig(qw, qx, qy, qz);
    o[0] = r.w; o[1] = r.x; o[2] = r.y; o[3] = r.z;
    return;
  }
  const float icy = rsqrtf(cy2), c = m00 * icy, sn = m10 * icy;
  const float h = sqrtf(0.5f * (1.0f + fabsf(c)));            // the larger of |cos|, |sin| of yaw/2
  const float g = 0.5f * sn * om_rcp(h);
  const float chz = c >= 0.f ? h : fabsf(g), shz = c >= 0.f ? g : copysignf(h, sn);
  const float inv = rsqrtf(nq);
  float w = fmaf(chz, qw, shz * qz) * inv, x = fmaf(chz, qx, shz * qy) * inv;
  float y = fmaf(chz, qy, -shz * qx) * inv, z = fmaf(chz, qz, -shz * qw) * inv;
  const bool neg = w < 0.f || (w == 0.f && x < 0.f);
  o[0] = neg ? -w : w; o[1] = neg ? -x : x; o[2] = neg ? -y : y; o[3] = neg ? -z : z;
}

// StickFigureA3.get_obs rows 0..30 (everything that does not depend on the task state)
OM_HD void a3_obs_robot(const float (&q)[A3_NQ], const float (&qd)[A3_NV], float (&obs)[A3_NOBS]) {
  a3_root_orient(q[3], q[4], q[5], q[6], obs);
#pragma unroll
  for (int k = 0; k < 3; ++k) obs[4 + k] = qd[3 + k];
#pragma unroll
  for (int k = 0; k < 12; ++k) obs[7 + k] = q[7 + k];                 // actuator_length / gear, gear = 1
#pragma unroll
  for (int k = 0; k < 12; ++k) obs[19 + k] = qd[6 + k];
}

// WalkingTask.step + calc_reward + done for one env.  Seq: float operator()(int step, int component); `tc` caches
// sequence[t1] / sequence[t2] across steps.  Fills obs[31..40], the six weighted terms, their sum and done.
template <class Seq>
OM_HD void a3_task_step(const A3TaskConst& C, A3Exact ex, const A3TaskIn& f, A3TaskRegs& s, A3Targets& tc,
                        const Seq& seq, float l_grf, float r_grf, float min_z, bool foot_contact, bool bad_collision,
                        float (&obs)[A3_NOBS], float (&terms)[6], float& total, bool& done) {
  s.phase += 1;                                                      // :248-250
  if (s.phase >= C.period) s.phase = 0;
  float dl = norm3(f.lsite - tc.p1), dr = norm3(f.rsite - tc.p1);    // :266-283
  if (a3_near_exact(C, f.lsite, f.rsite, tc.p1, ex)) {
    s.reached = 1;
    s.frames += 1;
  } else {
    s.reached = 0;
    s.frames = 0;
  }
  if (s.reached && s.frames >= C.delay_frames) {                     // :286-289, update_target_steps :228-244
    const int old_t2 = s.t2;
    s.t1 = s.t2;
    s.t2 += 1;
    if (s.t2 == s.seq_len) s.t2 = s.seq_len - 1;
    s.reached = 0;
    s.frames = 0;
    tc.p1 = tc.p2;
    tc.th1 = tc.th2;
    if (s.t2 != old_t2) { tc.p2 = V3{seq(s.t2, 0), seq(s.t2, 1), seq(s.t2, 2)}; tc.th2 = seq(s.t2, 3); }
    dl = norm3(f.lsite - tc.p1);
    dr = norm3(f.rsite - tc.p1);
  }

  // ---- update_goal_steps :184-225: inv([R p; 0 1]) [Rz(theta) t; 0 1]
  float R[9];
  tf3_quat2mat(f.root_q, R);
  const float* lrow = C.lut + (size_t)s.phase * A3_LUT_COLS;
  obs[31] = lrow[4];
  obs[32] = lrow[5];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const V3 d = (i == 0 ? tc.p1 : tc.p2) - f.root_p;
    float st, ct;
    om_sincos(i == 0 ? tc.th1 : tc.th2, &st, &ct);
    const float a = fmaf(R[0], ct, R[3] * st), b = fmaf(R[1], ct, R[4] * st);   // column 0 of R^T Rz
    const float cy = sqrtf(fmaf(a, a, b * b));
    const bool walk = s.mode != A3_STANDING;
    obs[33 + i] = walk ? fmaf(R[0], d.x, fmaf(R[3], d.y, R[6] * d.z)) : 0.f;
    obs[35 + i] = walk ? fmaf(R[1], d.x, fmaf(R[4], d.y, R[7] * d.z)) : 0.f;
    obs[37 + i] = walk ? fmaf(R[2], d.x, fmaf(R[5], d.y, R[8] * d.z)) : 0.f;
    obs[39 + i] = (walk && cy > A3_EPS4) ? om_atan2(b, a) : 0.f;
  }

  // ---- calc_reward :74-110
  float r_frc_c = 1.f, r_vel_c = -1.f, l_frc_c = 1.f, l_vel_c = -1.f;                  // STANDING :83-91
  if (s.mode != A3_STANDING) { r_frc_c = lrow[0]; r_vel_c = lrow[1]; l_frc_c = lrow[2]; l_vel_c = lrow[3]; }
  const float PI4 = 0.78539816339744831f;
  const float nl = fminf(l_grf, C.fmax) * C.inv_fmax * 2.f - 1.f, nr = fminf(r_grf, C.fmax) * C.inv_fmax * 2.f - 1.f;
  const float frc = (om_tan_q(PI4 * l_frc_c * nl) + om_tan_q(PI4 * r_frc_c * nr)) * 0.5f;      // rewards.py:65-83
  const float vl = fminf(f.lvel_n, C.vmax) * C.inv_vmax * 2.f - 1.f, vr = fminf(f.rvel_n, C.vmax) * C.inv_vmax * 2.f - 1.f;
  const float vel = (om_tan_q(PI4 * l_vel_c * vl) + om_tan_q(PI4 * r_vel_c * vr)) * 0.5f;      // rewards.py:85-102
  float sh, ch;
  om_sincos(0.5f * tc.th1, &sh, &ch);                                                    // euler2quat(0, 0, theta)
  const float inner = fmaf(ch, f.root_q.w, sh * f.root_q.z);
  const float orient = expf(-10.f * (1.f - inner * inner));                            // rewards.py:121-126
  // rewards.py:27-40; the dead-zone test is evaluated in double on the fp32 inputs (root z is qpos[2] itself)
  double err = fabs((double)f.root_p.z - (foot_contact ? (double)min_z : 0.0) - C.goal_height_ref);
  if (err < C.deadzone) err = 0.0;
  const float errf = (float)err;
  const float height = expf(-40.f * errf * errf);
  const float fd = fminf(dl, dr);                                                      // :56-72
  const float hit = s.reached ? expf(-fd / 0.25f) : 0.f;
  const float mx = (tc.p1.x + tc.p2.x) * 0.5f - f.root_p.x, my = (tc.p1.y + tc.p2.y) * 0.5f - f.root_p.y;
  const float progress = expf(-sqrtf(fmaf(mx, mx, my * my)) * 0.5f);
  const float step_r = fmaf(0.8f, hit, 0.2f * progress);
  const float hx = f.head_x - f.root_p.x, hy = f.head_y - f.root_p.y;
  const float upper = expf(-10.f * fmaf(hx, hx, hy * hy));
  terms[0] = 0.150f * frc; terms[1] = 0.150f * vel; terms[2] = 0.050f * orient;
  terms[3] = 0.050f * height; terms[4] = 0.450f * step_r; terms[5] = 0.050f * upper;
  total = ((((terms[0] + terms[1]) + terms[2]) + terms[3]) + terms[4]) + terms[5];      // StickFigureA3.py:192
  // ---- done :298-319 (float64 decision, see a3_done_height)
  done = a3_done_height(f.root_p.z, f.lsite.z, f.rsite.z, ex) || bad_collision;
}

// ---------------------------------------------------------------- time-parallel split of the step tail
// Everything in WalkingTask.step / calc_reward / done that does NOT depend on the footstep-target state (t1, t2,
// target_reached) depends on (env, t) alone -- including the phase clock, because phase(t) = (phase0 + t + 1) mod period
// and the mode is fixed for an episode.  The (env, t)-parallel pass evaluates those pieces (four of the six reward
// terms with their tanf/expf, done, the clock rows of the observation) and leaves a 16-float record; the sequential
// pass keeps only the target state machine, reduced to integer work on per-candidate "target near" bits; a second
// (env, t)-parallel pass finishes the goal steps, the orientation and step terms.
constexpr int A3_NREC = 16;
struct A3Rec {
  V3 root_p; Q4 root_q; V3 lsite, rsite;
  float t01, t3, t5;                  // terms[0] + terms[1], terms[3], terms[5] (summed later in the reference's order)
};
// ld as a 32-bit row stride: b + k * ld is one widening multiply-add per row (see om_a3.cu: row())
OM_HD void a3_rec_store(const A3Rec& r, float* b, unsigned ld) {
  b[0] = r.root_p.x; b[ld] = r.root_p.y; b[2u * ld] = r.root_p.z;
  b[3u * ld] = r.root_q.w; b[4u * ld] = r.root_q.x; b[5u * ld] = r.root_q.y; b[6u * ld] = r.root_q.z;
  b[7u * ld] = r.lsite.x; b[8u * ld] = r.lsite.y; b[9u * ld] = r.lsite.z;
  b[10u * ld] = r.rsite.x; b[11u * ld] = r.rsite.y; b[12u * ld] = r.rsite.z;
  b[13u * ld] = r.t01; b[14u * ld] = r.t3; b[15u * ld] = r.t5;
}
// (read once: streaming loads on the device, the line may leave L2 first)
OM_HD float a3_ld_once(const float* p) {
#ifdef __CUDA_ARCH__
  return __ldcs(p);
#else
  return *p;
#endif
}
OM_HD A3Rec a3_rec_load(const float* b, unsigned ld) {
  A3Rec r;
  r.root_p = V3{a3_ld_once(b), a3_ld_once(b + ld), a3_ld_once(b + 2u * ld)};
  r.root_q = Q4{a3_ld_once(b + 3u * ld), a3_ld_once(b + 4u * ld), a3_ld_once(b + 5u * ld), a3_ld_once(b + 6u * ld)};
  r.lsite = V3{a3_ld_once(b + 7u * ld), a3_ld_once(b + 8u * ld), a3_ld_once(b + 9u * ld)};
  r.rsite = V3{a3_ld_once(b + 10u * ld), a3_ld_once(b + 11u * ld), a3_ld_once(b + 12u * ld)};
  r.t01 = a3_ld_once(b + 13u * ld); r.t3 = a3_ld_once(b + 14u * ld); r.t5 = a3_ld_once(b + 15u * ld);
  return r;
}

// (env, t)-parallel part.  `phase` is the phase AFTER this step's increment.  Writes terms[0,1,3,5], obs[31,32], done.
OM_HD A3Rec a3_task_pre(const A3TaskConst& C, const A3TaskIn& f, int phase, int mode, float l_grf, float r_grf, float min_z,
                        bool foot_contact, bool bad_collision, float (&terms)[6], float& clock_sin, float& clock_cos,
                        bool& done, A3Exact& ex) {
  const float* lrow = C.lut + (size_t)phase * A3_LUT_COLS;
  clock_sin = lrow[4];
  clock_cos = lrow[5];
  float r_frc_c = 1.f, r_vel_c = -1.f, l_frc_c = 1.f, l_vel_c = -1.f;                  // STANDING :83-91
  if (mode != A3_STANDING) { r_frc_c = lrow[0]; r_vel_c = lrow[1]; l_frc_c = lrow[2]; l_vel_c = lrow[3]; }
  const float PI4 = 0.78539816339744831f;
  const float nl = fminf(l_grf, C.fmax) * C.inv_fmax * 2.f - 1.f, nr = fminf(r_grf, C.fmax) * C.inv_fmax * 2.f - 1.f;
  const float frc = (om_tan_q(PI4 * l_frc_c * nl) + om_tan_q(PI4 * r_frc_c * nr)) * 0.5f;      // rewards.py:65-83
  const float vl = fminf(f.lvel_n, C.vmax) * C.inv_vmax * 2.f - 1.f, vr = fminf(f.rvel_n, C.vmax) * C.inv_vmax * 2.f - 1.f;
  const float vel = (om_tan_q(PI4 * l_vel_c * vl) + om_tan_q(PI4 * r_vel_c * vr)) * 0.5f;      // rewards.py:85-102
  double err = fabs((double)f.root_p.z - (foot_contact ? (double)min_z : 0.0) - C.goal_height_ref);   // rewards.py:27-40
  if (err < C.deadzone) err = 0.0;
  const float errf = (float)err;
  const float height = expf(-40.f * errf * errf);
  const float hx = f.head_x - f.root_p.x, hy = f.head_y - f.root_p.y;
  const float upper = expf(-10.f * fmaf(hx, hx, hy * hy));
  terms[0] = 0.150f * frc; terms[1] = 0.150f * vel; terms[3] = 0.050f * height; terms[5] = 0.050f * upper;
  done = a3_done_height(f.root_p.z, f.lsite.z, f.rsite.z, ex) || bad_collision;                       // :298-319
  A3Rec r;
  r.root_p = f.root_p; r.root_q = f.root_q; r.lsite = f.lsite; r.rsite = f.rsite;
  r.t01 = terms[0] + terms[1]; r.t3 = terms[3]; r.t5 = terms[5];
  return r;
}

// sin / cos of the cached targets' headings (they change only when a target is reached)
struct A3TargetTrig { float s1, c1, s2, c2, sh, ch; };
OM_HD A3TargetTrig a3_target_trig(const A3Targets& tc) {
  A3TargetTrig g;
  om_sincos(tc.th1, &g.s1, &g.c1);
  om_sincos(tc.th2, &g.s2, &g.c2);
  om_sincos(0.5f * tc.th1, &g.sh, &g.ch);
  return g;
}

// The target state machine (:266-289, update_target_steps :228-244) is the only recurrence over time.  Within one call
// the target index t1 can only move along a chain fixed by the call's start state: candidate 0 = t1_0, candidate 1 =
// t2_0, candidate j = min(t2_0 + j - 1, len - 1); t2 is always the next candidate.  The (env, t)-parallel pass
// therefore evaluates "a foot is within target_radius of candidate j" for the first few candidates (one bit each),
// and the recurrence shrinks to integer work on those bits: no geometry on the sequential path.
constexpr int A3_MAX_CAND = 6;     // candidate bits 0..5 of the per-step byte; bit 6 flags a decision to re-take in float64,
                                   // bit 7 marks the byte as rewritten by the sequential pass (om_a3.cu: A3Scratch::step)
OM_HD int a3_cand(int j, int t1_0, int t2_0, int seq_len) {
  if (j == 0) return t1_0;
  const int k = t2_0 + j - 1;
  return k < seq_len - 1 ? k : seq_len - 1;
}
// "(double)sqrtf(d2) < target_radius" as a threshold on d2 itself: sqrtf is correctly rounded, hence monotone, so the
// two tests agree on every float.  Host-side helper (task creation, test harness).
inline float a3_near_d2(double radius) {
  float t = (float)(radius * radius);
  while ((double)sqrtf(t) < radius) t = nextafterf(t, INFINITY);
  while (t > 0.f && (double)sqrtf(nextafterf(t, 0.f)) >= radius) t = nextafterf(t, 0.f);
  return t;
}
template <class Seq>
OM_HD uint32_t a3_near_bits(const A3TaskConst& C, V3 lsite, V3 rsite, int ncand, int t1_0, int t2_0, int seq_len, const Seq& seq,
                            A3Exact& ex) {
  uint32_t bits = 0;
#pragma unroll 1
  for (int j = 0; j < ncand; ++j) {
    const int k = a3_cand(j, t1_0, t2_0, seq_len);
    const V3 p{seq(k, 0), seq(k, 1), seq(k, 2)};
    if (a3_near_exact(C, lsite, rsite, p, ex)) bits |= 1u << j;                 // float64 decision near the radius
  }
  return bits;
}
// Candidates that can be consulted at step t of a call: advance k needs max(delay - frames0, 1) + (k - 1) max(delay, 1)
// steps, and the new target's bit is first read on the step after the advance.
OM_HD int a3_cand_needed(int t, int frames0, int delay_frames, int ncand) {
  const int dm = delay_frames > 1 ? delay_frames : 1;
  int reach = delay_frames - frames0 > 1 ? delay_frames - frames0 : 1;     // first step that can read candidate 1
  int n = 1;
  while (n < ncand && t >= reach) { ++n; reach += dm; }                    // (no integer division on the hot path)
  return n;
}
struct A3Walk { int j, frames, reached; };       // j = number of target advances since the start of the call
OM_HD void a3_walk_step(const A3TaskConst& C, uint32_t bits, A3Walk& w) {
  if ((bits >> w.j) & 1u) {
    w.reached = 1;
    w.frames += 1;
  } else {
    w.reached = 0;
    w.frames = 0;
  }
  if (w.reached && w.frames >= C.delay_frames) {
    w.j += 1;
    w.reached = 0;
    w.frames = 0;
  }
}
// longest call the candidate bits cover: every advance needs max(delay_frames, 1) steps, the first may come at once
constexpr int a3_max_steps_per_call(int delay_frames) { return (A3_MAX_CAND - 1) * (delay_frames > 1 ? delay_frames : 1); }
inline int a3_num_cand_host(int n_steps, int delay_frames) {
  const int d = delay_frames > 1 ? delay_frames : 1;
  const int nc = 1 + (n_steps + d - 1) / d;
  return nc < A3_MAX_CAND ? nc : A3_MAX_CAND;
}

// (env, t)-parallel again: goal steps (update_goal_steps :184-225), orientation and step terms, total, given the
// state the machine was in after this step.  Writes goal[8] (obs rows 33..40), terms[2], terms[4], total.
// `tg`: sin / cos of the two targets' headings (a3_target_trig), computed here or read from the per-candidate table the
// walk pass leaves (same function of the same angle: identical bits)
template <class Seq>
OM_HD void a3_task_post(const A3TaskConst& C, const A3Rec& f, int mode, int t1, int t2, bool reached, const Seq& seq,
                        const A3TargetTrig* tg_in, float (&goal)[8], float& t2_orient, float& t4_step, float& total) {
  A3TaskRegs s{};
  s.t1 = t1; s.t2 = t2;
  const A3Targets tc = a3_targets_load(s, seq);
  const A3TargetTrig tg = tg_in ? *tg_in : a3_target_trig(tc);
  const float dl = norm3(f.lsite - tc.p1), dr = norm3(f.rsite - tc.p1);
  if (mode != A3_STANDING) {
    float R[9];
    tf3_quat2mat(f.root_q, R);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const V3 d = (i == 0 ? tc.p1 : tc.p2) - f.root_p;
      const float st = i == 0 ? tg.s1 : tg.s2, ct = i == 0 ? tg.c1 : tg.c2;
      const float a = fmaf(R[0], ct, R[3] * st), b = fmaf(R[1], ct, R[4] * st);
      const float cy = sqrtf(fmaf(a, a, b * b));
      goal[0 + i] = fmaf(R[0], d.x, fmaf(R[3], d.y, R[6] * d.z));
      goal[2 + i] = fmaf(R[1], d.x, fmaf(R[4], d.y, R[7] * d.z));
      goal[4 + i] = fmaf(R[2], d.x, fmaf(R[5], d.y, R[8] * d.z));
      goal[6 + i] = cy > A3_EPS4 ? om_atan2(b, a) : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) goal[i] = 0.f;
  }
  const float inner = fmaf(tg.ch, f.root_q.w, tg.sh * f.root_q.z);
  t2_orient = 0.050f * expf(-10.f * (1.f - inner * inner));                            // rewards.py:121-126
  const float fd = fminf(dl, dr);                                                      // walking_task.py:56-72
  const float hit = reached ? expf(-fd / 0.25f) : 0.f;
  const float mx = (tc.p1.x + tc.p2.x) * 0.5f - f.root_p.x, my = (tc.p1.y + tc.p2.y) * 0.5f - f.root_p.y;
  const float progress = expf(-sqrtf(fmaf(mx, mx, my * my)) * 0.5f);
  t4_step = 0.450f * fmaf(0.8f, hit, 0.2f * progress);
  total = (((f.t01 + t2_orient) + f.t3) + t4_step) + f.t5;                             // StickFigureA3.py:192, dict order
}

// ---------------------------------------------------------------- reset (A13)
OM_HD void a3_reset_uniforms(uint64_t seed, uint32_t env_id, uint32_t reset_count, float (&u)[A3_NU]) {
#pragma unroll
  for (int s = 0; s < A3_NU / 4; ++s) {
    const U4 w = om_draw(seed, env_id, reset_count, A3_RESET_STREAM + s);
    u[4 * s] = to_unit(w.x); u[4 * s + 1] = to_unit(w.y); u[4 * s + 2] = to_unit(w.z); u[4 * s + 3] = to_unit(w.w);
  }
}

// reset_model (StickFigureA3.py:205-235): draw order documented in oracle/a3.py reset()
OM_HD void a3_reset_qpos_qvel(const float* init_qpos, const float (&u)[A3_NU], float (&q)[A3_NQ], float (&qd)[A3_NV]) {
  const float c = 0.02f;
#pragma unroll
  for (int k = 0; k < A3_NQ; ++k) q[k] = fmaf(fmaf(u[k], 2.f, -1.f), c, init_qpos[k]);
#pragma unroll
  for (int k = 0; k < A3_NV; ++k) qd[k] = fmaf(u[25 + k], 2.f, -1.f) * c;
  q[0] = fmaf(u[49], 2.f, -1.f);
  q[1] = fmaf(u[50], 2.f, -1.f);
  q[2] = 1.34f;
  const float pitch = fmaf(u[51], 10.f, -5.f) * 0.017453292519943295f;
  const float yaw = fmaf(u[52], 2.f, -1.f) * 3.14159265358979323846f;
  float sj, cj, sk, ck;
  om_sincos(0.5f * pitch, &sj, &cj);
  om_sincos(0.5f * yaw, &sk, &ck);
  q[3] = cj * ck; q[4] = -(sj * sk); q[5] = sj * ck; q[6] = cj * sk;     // euler2quat(0, pitch, yaw)
}

// WalkingTask.reset (:321-397) after set_state; SeqOut: void operator()(int step, int component, float value)
template <class SeqOut>
OM_HD void a3_task_reset(const A3TaskConst& C, const A3Feat& f, const float (&u)[A3_NU], float step_h, A3TaskRegs& s,
                         const SeqOut& seq_out) {
  s.phase = (double)u[53] < 0.5 ? 0 : C.period / 2;                   // :354
  s.mode = (double)u[54] < 0.2 ? A3_STANDING : A3_FORWARD;            // :362-364
  int num_steps = A3_MAX_STEPS;
  float step_height = 0.f;
  if (s.mode == A3_STANDING) num_steps = 1;
  else step_height = (double)u[55] < 0.5 ? -step_h : step_h;          // :377-379
  const float first_y = fmaf(0.01f, u[56], 0.095f);                   // generate_step_sequence :137-182
  const bool half = s.phase == C.period / 2;
  const int cc = (double)u[57] < 0.5 ? 2 : 3;
  float R[9];
  tf3_quat2mat(f.root_q, R);                                          // transform_sequence :113-135
  const float cyy = sqrtf(fmaf(R[0], R[0], R[3] * R[3]));
  const float yaw = cyy > A3_EPS4 ? om_atan2(R[3], R[0]) : 0.f;
  float sy, cy;
  om_sincos(yaw, &sy, &cy);
  const float mx = (f.lfoot_p.x + f.rfoot_p.x) * 0.5f, my = (f.lfoot_p.y + f.rfoot_p.y) * 0.5f;
  float x = 0.f, z = 0.f, y = half ? -0.15f : 0.15f;
  for (int i = 0; i < A3_MAX_STEPS; ++i) {
    float sx, syy, sz;
    if (i == 0) { sx = 0.f; syy = half ? -first_y : first_y; sz = 0.f; }
    else {
      x += 0.3f;
      y = -y;
      if (i > cc) z += step_height;
      sx = x; syy = y; sz = z;
    }
    const bool live = i < num_steps;
    seq_out(i, 0, live ? mx + sx * cy - syy * sy : 0.f);
    seq_out(i, 1, live ? my + sx * sy + syy * cy : 0.f);
    seq_out(i, 2, live ? sz : 0.f);
    seq_out(i, 3, live ? yaw : 0.f);
  }
  s.seq_len = num_steps;
  s.t1 = 0;                                                           // update_target_steps from t1 = t2 = 0
  s.t2 = 1;
  if (s.t2 == s.seq_len) s.t2 = s.seq_len - 1;
  s.frames = 0;
  s.reached = 0;
}

}  // namespace om
