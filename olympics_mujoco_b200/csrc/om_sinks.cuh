// Output sinks for the generated FK functions (gen/fk_*.cuh).  The generated code calls
// S.xpos(b, ...) etc. with literal indices; a sink decides where (and whether) a value is stored.
#pragma once
#include "om_common.cuh"

namespace om {

// Structure-of-arrays HBM sink: component c of env e at base[c*ld + e]; null arrays are skipped.  STREAM: the stores are
// streaming (st.global.cs, evict-first) -- for the multi-step kernels, whose outputs are written once and not read back by
// the same call, so that what IS re-read (trajectory table, prefix sums, rewards for the returns pass) keeps its L2 lines.
template <bool SITE_XMAT, bool STREAM = false>
struct SoaSink {
  static constexpr bool want_site_xmat = SITE_XMAT;
  float* __restrict__ xp; float* __restrict__ xq; float* __restrict__ sp; float* __restrict__ sm;
  float* __restrict__ cv; float* __restrict__ cm;
  size_t ld, env;
  static OM_HD void st(float* p, float v) {
#ifdef __CUDA_ARCH__
    if (STREAM) __stcs(p, v);
    else *p = v;
#else
    *p = v;
#endif
  }
  OM_HD void xpos(int b, float x, float y, float z) const {
    if (xp) { st(xp + (3 * b) * ld + env, x); st(xp + (3 * b + 1) * ld + env, y); st(xp + (3 * b + 2) * ld + env, z); }
  }
  OM_HD void xquat(int b, float w, float x, float y, float z) const {
    if (xq) {
      st(xq + (4 * b) * ld + env, w); st(xq + (4 * b + 1) * ld + env, x); st(xq + (4 * b + 2) * ld + env, y);
      st(xq + (4 * b + 3) * ld + env, z);
    }
  }
  OM_HD void site_xpos(int s, float x, float y, float z) const {
    if (sp) { st(sp + (3 * s) * ld + env, x); st(sp + (3 * s + 1) * ld + env, y); st(sp + (3 * s + 2) * ld + env, z); }
  }
  OM_HD void site_xmat(int s, float a, float b, float c, float d, float e, float f, float g, float h, float i) const {
    if (SITE_XMAT && sm) {
      float* o = sm + (9 * s) * ld + env;
      st(o, a); st(o + ld, b); st(o + 2 * ld, c); st(o + 3 * ld, d); st(o + 4 * ld, e); st(o + 5 * ld, f); st(o + 6 * ld, g);
      st(o + 7 * ld, h); st(o + 8 * ld, i);
    }
  }
  OM_HD void cvel(int b, float wx, float wy, float wz, float vx, float vy, float vz) const {
    if (cv) {
      float* o = cv + (6 * b) * ld + env;
      st(o, wx); st(o + ld, wy); st(o + 2 * ld, wz); st(o + 3 * ld, vx); st(o + 4 * ld, vy); st(o + 5 * ld, vz);
    }
  }
  OM_HD void com(float x, float y, float z) const {
    if (cm) { st(cm + env, x); st(cm + ld + env, y); st(cm + 2 * ld + env, z); }
  }
  OM_HD void vel_p(int, float, float, float, float, float, float) const {}
};

}  // namespace om
