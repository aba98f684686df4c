// TEST INFRASTRUCTURE: compiles the GENERATED fk device functions as plain host C++ so that the code
// generator can be checked against the float64 oracle in a container without a GPU.  Never shipped,
// never used by the product (which has no CPU path).
#include <cmath>
#include <cstring>
#define OM_HD inline
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
#include "om_math.cuh"
#include "fk_unitree_h1.cuh"
#include "fk_stick_figure_a3.cuh"
#include "fk_pos_stick_figure_a3.cuh"
#include "fk_unitree_h1_parts.cuh"
#include "fk_stick_figure_a3_parts.cuh"      // test-only: the partitioner on a free-joint-rooted tree

struct HostSink {
  static constexpr bool want_site_xmat = true;
  float *xp, *xq, *sp, *sm, *cv, *cm;
  float* vp = nullptr;
  void xpos(int b, float x, float y, float z) { xp[b*3]=x; xp[b*3+1]=y; xp[b*3+2]=z; }
  void xquat(int b, float w, float x, float y, float z) { xq[b*4]=w; xq[b*4+1]=x; xq[b*4+2]=y; xq[b*4+3]=z; }
  void site_xpos(int s, float x, float y, float z) { sp[s*3]=x; sp[s*3+1]=y; sp[s*3+2]=z; }
  void site_xmat(int s, float a, float b, float c, float d, float e, float f, float g, float h, float i) {
    float m[9] = {a,b,c,d,e,f,g,h,i}; std::memcpy(sm + 9*s, m, sizeof m); }
  void cvel(int b, float wx, float wy, float wz, float vx, float vy, float vz) {
    float v[6] = {wx,wy,wz,vx,vy,vz}; std::memcpy(cv + 6*b, v, sizeof v); }
  void com(float x, float y, float z) { cm[0]=x; cm[1]=y; cm[2]=z; }
  void vel_p(int b, float wx, float wy, float wz, float vx, float vy, float vz) {
    if (vp) { float v[6] = {wx,wy,wz,vx,vy,vz}; std::memcpy(vp + 6*b, v, sizeof v); } }
};

extern "C" void host_fk_h1(const float* q, const float* qd, int n, float* xp, float* xq, float* sp, float* sm,
                           float* cv, float* cm) {
  for (int e = 0; e < n; ++e) {
    float qq[17], dd[17];
    std::memcpy(qq, q + 17*e, sizeof qq); std::memcpy(dd, qd + 17*e, sizeof dd);
    HostSink S{xp + 63*e, xq + 84*e, sp + 3*e, sm + 9*e, cv + 126*e, cm + 3*e};
    om_fk_unitree_h1(qq, dd, S);
  }
}
extern "C" void host_fk_a3(const float* q, const float* qd, int n, float* xp, float* xq, float* sp, float* sm,
                           float* cv, float* cm) {
  for (int e = 0; e < n; ++e) {
    float qq[25], dd[24];
    std::memcpy(qq, q + 25*e, sizeof qq); std::memcpy(dd, qd + 24*e, sizeof dd);
    HostSink S{xp + 51*e, xq + 68*e, sp + 6*e, sm + 18*e, cv + 102*e, cm + 3*e};
    om_fk_stick_figure_a3(qq, dd, S);
  }
}

// the matrix-chain variant (codegen.generate_fk_pos) next to the quaternion one: xpos, site_xpos, root xquat and the
// spatial velocities about the root origin of both
extern "C" void host_fk_a3_pos(const float* q, const float* qd, int n, float* xp, float* xq, float* sp, float* vp,
                               float* vp_quat) {
  static float dump[17 * 9 + 64];
  for (int e = 0; e < n; ++e) {
    float qq[25], dd[24];
    std::memcpy(qq, q + 25*e, sizeof qq); std::memcpy(dd, qd + 24*e, sizeof dd);
    HostSink S{xp + 51*e, xq + 68*e, sp + 6*e, dump, dump, dump};
    S.vp = vp + 102*e;
    om_fk_pos_stick_figure_a3(qq, dd, S);
    static float xp2[51], xq2[68], sp2[6], sm2[18], cv2[102], cm2[3];
    HostSink S2{xp2, xq2, sp2, sm2, cv2, cm2};
    S2.vp = vp_quat + 102*e;
    om_fk_stick_figure_a3(qq, dd, S2);
  }
}

// the three part functions of the split H1 kernel (csrc/om_fk.cu: h1_step_split_kernel): pass 1 collects the partial
// centre-of-mass sums, pass 2 hands every part the combined centre of mass, like the shared-memory exchange does
struct HostExchange {
  float (*sum)[3]; int part; bool second;
  void com_exchange(float sx, float sy, float sz, float inv_mass, float& cx, float& cy, float& cz) const {
    if (!second) { sum[part][0] = sx; sum[part][1] = sy; sum[part][2] = sz; cx = cy = cz = 0.f; return; }
    cx = ((sum[0][0] + sum[1][0]) + sum[2][0]) * inv_mass;
    cy = ((sum[0][1] + sum[1][1]) + sum[2][1]) * inv_mass;
    cz = ((sum[0][2] + sum[1][2]) + sum[2][2]) * inv_mass;
  }
};
extern "C" void host_fk_h1_parts(const float* q, const float* qd, int n, float* xp, float* xq, float* sp, float* sm,
                                 float* cv, float* cm) {
  for (int e = 0; e < n; ++e) {
    float qq[17], dd[17], sum[3][3];
    std::memcpy(qq, q + 17*e, sizeof qq); std::memcpy(dd, qd + 17*e, sizeof dd);
    HostSink S{xp + 63*e, xq + 84*e, sp + 3*e, sm + 9*e, cv + 126*e, cm + 3*e};
    for (int pass = 0; pass < 2; ++pass) {
      HostExchange X0{sum, 0, pass == 1}, X1{sum, 1, pass == 1}, X2{sum, 2, pass == 1};
      om_fk_unitree_h1_part0(qq, dd, S, X0);
      om_fk_unitree_h1_part1(qq, dd, S, X1);
      om_fk_unitree_h1_part2(qq, dd, S, X2);
    }
  }
}

// the same two-pass emulation for a three-way split of the StickFigureA3 tree (not used by any kernel: it checks the
// part generator on a model whose root is a free joint and whose subtrees are unbalanced)
extern "C" void host_fk_a3_parts(const float* q, const float* qd, int n, float* xp, float* xq, float* sp, float* sm,
                                 float* cv, float* cm) {
  for (int e = 0; e < n; ++e) {
    float qq[25], dd[24], sum[3][3];
    std::memcpy(qq, q + 25*e, sizeof qq); std::memcpy(dd, qd + 24*e, sizeof dd);
    HostSink S{xp + 51*e, xq + 68*e, sp + 6*e, sm + 18*e, cv + 102*e, cm + 3*e};
    for (int pass = 0; pass < 2; ++pass) {
      HostExchange X0{sum, 0, pass == 1}, X1{sum, 1, pass == 1}, X2{sum, 2, pass == 1};
      om_fk_stick_figure_a3_part0(qq, dd, S, X0);
      om_fk_stick_figure_a3_part1(qq, dd, S, X1);
      om_fk_stick_figure_a3_part2(qq, dd, S, X2);
    }
  }
}
