// Output sinks for the generated FK functions (gen/fk_*.cuh).  The generated code calls
// S.xpos(b, ...) etc. with literal indices; a sink decides where (and whether) a value is stored.
#pragma once
#include "om_common.cuh"

namespace om {

// Structure-of-arrays HBM sink: component c of env e at base[c*ld + e]; null arrays are skipped.
template <bool SITE_XMAT>
struct SoaSink {
  static constexpr bool want_site_xmat = SITE_XMAT;
  float* __restrict__ xp; float* __restrict__ xq; float* __restrict__ sp; float* __restrict__ sm;
  float* __restrict__ cv; float* __restrict__ cm;
  size_t ld, env;
  OM_HD void xpos(int b, float x, float y, float z) const {
    if (xp) { xp[(3 * b) * ld + env] = x; xp[(3 * b + 1) * ld + env] = y; xp[(3 * b + 2) * ld + env] = z; }
  }
  OM_HD void xquat(int b, float w, float x, float y, float z) const {
    if (xq) { xq[(4 * b) * ld + env] = w; xq[(4 * b + 1) * ld + env] = x; xq[(4 * b + 2) * ld + env] = y; xq[(4 * b + 3) * ld + env] = z; }
  }
  OM_HD void site_xpos(int s, float x, float y, float z) const {
    if (sp) { sp[(3 * s) * ld + env] = x; sp[(3 * s + 1) * ld + env] = y; sp[(3 * s + 2) * ld + env] = z; }
  }
  OM_HD void site_xmat(int s, float a, float b, float c, float d, float e, float f, float g, float h, float i) const {
    if (SITE_XMAT && sm) {
      float* o = sm + (9 * s) * ld + env;
      o[0] = a; o[ld] = b; o[2 * ld] = c; o[3 * ld] = d; o[4 * ld] = e; o[5 * ld] = f; o[6 * ld] = g; o[7 * ld] = h; o[8 * ld] = i;
    }
  }
  OM_HD void cvel(int b, float wx, float wy, float wz, float vx, float vy, float vz) const {
    if (cv) {
      float* o = cv + (6 * b) * ld + env;
      o[0] = wx; o[ld] = wy; o[2 * ld] = wz; o[3 * ld] = vx; o[4 * ld] = vy; o[5 * ld] = vz;
    }
  }
  OM_HD void com(float x, float y, float z) const {
    if (cm) { cm[env] = x; cm[ld + env] = y; cm[2 * ld + env] = z; }
  }
  OM_HD void vel_p(int, float, float, float, float, float, float) const {}
};

}  // namespace om
