// TEST INFRASTRUCTURE: compiles the per-env device functions of the A3 step tail (csrc/om_a3_task.cuh, fp32) as
// plain host C++ so that they can be checked against the reference-generated golden fixture in a container
// without a GPU.  The loops below mirror a3_task_kernel / a3_reset_kernel in csrc/om_a3.cu for ONE env.
// Never shipped, never used by the product (which has no CPU path).
#include <cmath>
#include <cstring>
#define OM_HD inline
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
#include "om_a3_task.cuh"

using namespace om;

static A3TaskConst make_const(const float* lut6, int period, int delay, double radius, double gh, double dz, float fmax) {
  A3TaskConst C;
  C.period = period; C.delay_frames = delay; C.fmax = fmax; C.vmax = 0.2f; C.inv_fmax = 1.0f / fmax; C.inv_vmax = 1.0f / 0.2f;
  C.target_radius = radius; C.near_d2 = a3_near_d2(radius); C.goal_height_ref = gh; C.deadzone = dz; C.lut = lut6;
  a3_near_band(radius, &C.near_lo2, &C.near_hi2);
  return C;
}

struct SeqHost {
  const float* s;
  float operator()(int t, int c) const { return s[t * 4 + c]; }
};
struct SeqHostOut {
  float* s;
  void operator()(int t, int c, float v) const { s[t * 4 + c] = v; }
};

extern "C" void host_a3_rollout(const float* lut6, int period, int delay, double radius, double gh, double dz, float fmax,
                                const float* qpos, const float* qvel, const float* contact, int T, int* ints, float* seq,
                                float* obs, float* terms, float* reward, unsigned char* done) {
  const A3TaskConst C = make_const(lut6, period, delay, radius, gh, dz, fmax);
  A3TaskRegs s{ints[0], ints[1], ints[2], ints[3], ints[4], ints[5], ints[6]};
  A3Targets tc = a3_targets_load(s, SeqHost{seq});
  for (int t = 0; t < T; ++t) {
    float q[A3_NQ], qd[A3_NV], o[A3_NOBS], tr[6], total;
    bool d;
    std::memcpy(q, qpos + t * A3_NQ, sizeof q);
    std::memcpy(qd, qvel + t * A3_NV, sizeof qd);
    const float* c = contact + t * 4;
    a3_obs_robot(q, qd, o);
    A3Sink<NullFkSink> S{};
    om_fk_pos_stick_figure_a3(q, qd, S);
    const int fl = (int)c[3];
    a3_task_step(C, A3Exact{qpos + t * A3_NQ, 1, false, false}, a3_task_in(S.f), s, tc, SeqHost{seq}, c[0], c[1], c[2], (fl & 1) != 0, (fl & 2) != 0, o, tr, total, d);
    std::memcpy(obs + t * A3_NOBS, o, sizeof o);
    std::memcpy(terms + t * 6, tr, sizeof tr);
    reward[t] = total;
    done[t] = d ? 1 : 0;
  }
  int out[7] = {s.phase, s.t1, s.t2, s.frames, s.mode, s.seq_len, s.reached};
  std::memcpy(ints, out, sizeof out);
}

// the time-parallel split (a3_feat_kernel with its state-machine tail + a3_post_kernel of csrc/om_a3.cu) for one env
extern "C" void host_a3_rollout_split(const float* lut6, int period, int delay, double radius, double gh, double dz, float fmax,
                                      const float* qpos, const float* qvel, const float* contact, int T, int* ints, float* seq,
                                      float* obs, float* terms, float* reward, unsigned char* done) {
  const A3TaskConst C = make_const(lut6, period, delay, radius, gh, dz, fmax);
  float* rec = new float[(size_t)T * A3_NREC];
  for (int t = 0; t < T; ++t) {                                   // pass 1: any order
    float q[A3_NQ], qd[A3_NV], o[A3_NOBS], tr[6];
    bool d;
    std::memcpy(q, qpos + t * A3_NQ, sizeof q);
    std::memcpy(qd, qvel + t * A3_NV, sizeof qd);
    const float* c = contact + t * 4;
    a3_obs_robot(q, qd, o);
    A3Sink<NullFkSink> S{};
    om_fk_pos_stick_figure_a3(q, qd, S);
    const int fl = (int)c[3];
    const int phase = (ints[0] + t + 1) % period;
    A3Exact ex{qpos + t * A3_NQ, 1, false, false};
    const A3Rec r = a3_task_pre(C, a3_task_in(S.f), phase, ints[4], c[0], c[1], c[2], (fl & 1) != 0, (fl & 2) != 0, tr, o[31],
                                o[32], d, ex);
    a3_rec_store(r, rec + (size_t)t * A3_NREC, 1);
    std::memcpy(obs + t * A3_NOBS, o, 33 * sizeof(float));
    terms[t * 6 + 0] = tr[0]; terms[t * 6 + 1] = tr[1]; terms[t * 6 + 3] = tr[3]; terms[t * 6 + 5] = tr[5];
    done[t] = d ? 1 : 0;
  }
  // passes B + C: integer walk over the candidate bits, then the state-dependent outputs.  Calls longer than the bits
  // cover are cut into sub-calls exactly like om_a3_task_step does.
  uint8_t* near = new uint8_t[T];
  int st[7] = {ints[0], ints[1], ints[2], ints[3], ints[4], ints[5], ints[6]};
  for (int c0 = 0; c0 < T; c0 += a3_max_steps_per_call(delay)) {
    const int len = T - c0 < a3_max_steps_per_call(delay) ? T - c0 : a3_max_steps_per_call(delay);
    const int nc = a3_num_cand_host(len, delay);
    const int t1_0 = st[1], t2_0 = st[2], sl = st[5];
    for (int t = c0; t < c0 + len; ++t) {
      const A3Rec r = a3_rec_load(rec + (size_t)t * A3_NREC, 1);
      A3Exact ex{qpos + t * A3_NQ, 1, false, false};
      near[t] = (uint8_t)a3_near_bits(C, r.lsite, r.rsite, a3_cand_needed(t - c0, st[3], delay, nc), t1_0, t2_0, sl, SeqHost{seq}, ex);
    }
    A3Walk w{0, st[3], st[6]};
    for (int t = c0; t < c0 + len; ++t) {
      a3_walk_step(C, near[t], w);
      float goal[8], t2, t4, total;
      a3_task_post(C, a3_rec_load(rec + (size_t)t * A3_NREC, 1), st[4], a3_cand(w.j, t1_0, t2_0, sl),
                   a3_cand(w.j + 1, t1_0, t2_0, sl), w.reached != 0, SeqHost{seq}, nullptr, goal, t2, t4, total);
      std::memcpy(obs + t * A3_NOBS + 33, goal, sizeof goal);
      terms[t * 6 + 2] = t2; terms[t * 6 + 4] = t4;
      reward[t] = total;
    }
    st[1] = a3_cand(w.j, t1_0, t2_0, sl); st[2] = a3_cand(w.j + 1, t1_0, t2_0, sl); st[3] = w.frames; st[6] = w.reached;
  }
  delete[] near;
  delete[] rec;
  int out[7] = {(ints[0] + T) % period, st[1], st[2], st[3], st[4], st[5], st[6]};
  std::memcpy(ints, out, sizeof out);
}

extern "C" void host_a3_reset(const float* lut6, int period, int delay, double radius, double gh, double dz, float fmax,
                              const float* init_qpos, unsigned long long seed, unsigned env, unsigned rc, float step_h,
                              float* qpos, float* qvel, int* ints, float* seq, float* obs) {
  const A3TaskConst C = make_const(lut6, period, delay, radius, gh, dz, fmax);
  float u[A3_NU], q[A3_NQ], qd[A3_NV];
  a3_reset_uniforms(seed, env, rc, u);
  a3_reset_qpos_qvel(init_qpos, u, q, qd);
  A3Sink<NullFkSink> S{};
  om_fk_pos_stick_figure_a3(q, qd, S);
  A3TaskRegs s;
  a3_task_reset(C, S.f, u, step_h, s, SeqHostOut{seq});
  std::memcpy(qpos, q, sizeof q);
  std::memcpy(qvel, qd, sizeof qd);
  int out[7] = {s.phase, s.t1, s.t2, s.frames, s.mode, s.seq_len, s.reached};
  std::memcpy(ints, out, sizeof out);
  float o[A3_NOBS];
  a3_obs_robot(q, qd, o);
  o[31] = lut6[s.phase * A3_LUT_COLS + 4];
  o[32] = lut6[s.phase * A3_LUT_COLS + 5];
  for (int k = 33; k < A3_NOBS; ++k) o[k] = 0.f;
  std::memcpy(obs, o, sizeof o);
}

// Property check of the candidate pruning (a3_cand_needed): walking the state machine over the FULL candidate bits and
// over bits masked to the candidates declared reachable at each step must give the same (advances, frames, reached) at
// every step.  Returns the first step at which they differ, or -1.
extern "C" int host_a3_walk_pruning_check(int delay, int frames0, int reached0, int ncand, int T, const unsigned char* bits) {
  A3TaskConst C{};
  C.delay_frames = delay;
  A3Walk full{0, frames0, reached0}, pruned{0, frames0, reached0};
  for (int t = 0; t < T; ++t) {
    const int nc = a3_cand_needed(t, frames0, delay, ncand);
    const unsigned masked = bits[t] & ((1u << nc) - 1u);
    a3_walk_step(C, bits[t], full);
    a3_walk_step(C, masked, pruned);
    if (full.j != pruned.j || full.frames != pruned.frames || full.reached != pruned.reached) return t;
    if (full.j >= ncand) return -1;            // beyond the call's candidate budget (the ABI cuts calls before this)
  }
  return -1;
}

// float64 site positions of the exact-decision slow path (gen/fk_pos_f64_*.cuh) for the oracle comparison
extern "C" void host_a3_sites_f64(const float* qpos, int n, double* lsite, double* rsite) {
  for (int i = 0; i < n; ++i) {
    const A3SitesF64 s = a3_sites_f64(qpos + (size_t)i * A3_NQ, 1);
    for (int k = 0; k < 3; ++k) { lsite[3 * i + k] = s.ls[k]; rsite[3 * i + k] = s.rs[k]; }
  }
}

// closed-form root roll / pitch quaternion (a3_root_orient) next to the literal quat2euler -> euler2quat path
extern "C" void host_a3_root_orient(const float* q, int n, float* closed, float* literal) {
  for (int i = 0; i < n; ++i) {
    a3_root_orient(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3], closed + 4 * i);
    const Q4 r = a3_root_orient_trig(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
    literal[4 * i] = r.w; literal[4 * i + 1] = r.x; literal[4 * i + 2] = r.y; literal[4 * i + 3] = r.z;
  }
}

// bounded-range trigonometry of om_math.cuh
extern "C" void host_trig(const float* x, int n, float* s, float* c, float* t) {
  for (int i = 0; i < n; ++i) {
    om_sincos(x[i], s + i, c + i);
    t[i] = om_tan_q(x[i]);
  }
}
extern "C" void host_atan2(const float* y, const float* x, int n, float* r) {
  for (int i = 0; i < n; ++i) r[i] = om_atan2(y[i], x[i]);
}
