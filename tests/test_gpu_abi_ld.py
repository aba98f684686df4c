"""C-ABI contract `ld >= n`: arrays may be padded (leading dimension larger than the number of envs).  Every entry
point is called through ctypes with ld = n + pad; results must equal the ld == n call and the padding columns, filled
with a sentinel, must come back untouched (compute-sanitizer is closed on this pool: this is the out-of-bounds check)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SENT = -12345.0


def _padded(x, ld):
    """[..., C, n] tensor -> same data inside a [..., C, ld] buffer whose extra columns hold the sentinel."""
    import torch
    buf = torch.full(x.shape[:-1] + (ld,), SENT if x.dtype.is_floating_point else 77, dtype=x.dtype, device=x.device)
    buf[..., :x.shape[-1]] = x
    return buf


def _check(buf, ref, n, name):
    import torch
    assert torch.equal(buf[..., :n], ref), f"{name}: padded call differs"
    pad = buf[..., n:]
    want = SENT if buf.dtype.is_floating_point else 77
    assert bool((pad == want).all()), f"{name}: wrote into the padding"


def test_padded_leading_dimension(h1_model, a3_model, h1_states):
    import torch
    from olympics_mujoco_b200 import _lib
    from olympics_mujoco_b200 import kernels as Kn
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    from oracle import a3 as OA
    from oracle import h1 as OH
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    keep = []                                            # padded temporaries must outlive the asynchronous launches

    def P(t):
        if t is None:
            return None
        keep.append(t)
        return C.c_void_p(t.data_ptr())
    n, ld = 193, 256
    qpos, qvel = (Kn.to_soa(a[:n]) for a in h1_states)
    dm = Kn.DeviceModel(h1_model)
    spec = Kn.make_h1_spec(OH.perm(h1_model), OH.x_vel_idx(h1_model))
    pxv = torch.randn(n, device="cuda")
    ref = Kn.h1_step(dm, spec, qpos, qvel, pxv)
    out = {k: _padded(torch.zeros_like(v), ld) for k, v in ref.items()}
    _lib.check(lib.om_h1_step(dm.handle, C.byref(spec), P(_padded(qpos, ld)), P(_padded(qvel, ld)), P(_padded(pxv, ld)), n, ld,
                              P(out["xpos"]), P(out["xquat"]), P(out["site_xpos"]), P(out["cvel"]), P(out["obs"]),
                              P(out["reward"]), P(out["absorbing"]), st))
    torch.cuda.synchronize()
    for k in ref:
        _check(out[k], ref[k], n, "om_h1_step." + k)
    # om_fk, generic kernel
    reff = Kn.fk(dm, qpos, qvel, force_generic=True)
    outf = {k: _padded(torch.zeros_like(v), ld) for k, v in reff.items()}
    _lib.check(lib.om_fk(dm.handle, P(_padded(qpos, ld)), P(_padded(qvel, ld)), n, ld, P(outf["xpos"]), P(outf["xquat"]),
                         P(outf["site_xpos"]), P(outf["site_xmat"]), P(outf["cvel"]), P(outf["subtree_com"]), 1, st))
    torch.cuda.synchronize()
    for k in reff:
        _check(outf[k], reff[k], n, "om_fk." + k)
    # A3 multi-step replay (time-parallel pair) with padded inputs, state and outputs
    T = 5
    task = Kn.A3Task(Kn.DeviceModel(a3_model), n, phase_clock_lut(), OA.init_qpos(), seed=2)
    q0, v0 = Kn.soa(25, n), Kn.soa(24, n)
    task.reset(q0, v0)
    g = torch.Generator(device="cuda").manual_seed(0)
    q = (q0[None] + 0.01 * torch.randn((T, 25, n), device="cuda", generator=g)).contiguous()
    v = torch.randn((T, 24, n), device="cuda", generator=g)
    con = torch.rand((T, 4, n), device="cuda", generator=g)
    ints0, seq0 = task.ints.clone(), task.sequence.clone()
    refa = task.step(q, v, con)
    ints_ref = task.ints.clone()
    ints_p, seq_p = _padded(ints0, ld), _padded(seq0, ld)
    outa = {k: _padded(torch.zeros_like(x), ld) for k, x in refa.items()}
    sa = _lib.OmA3State(ints=ints_p.data_ptr(), sequence=seq_p.data_ptr())
    oa = _lib.OmA3Out(obs=outa["obs"].data_ptr(), terms=outa["terms"].data_ptr(), reward=outa["reward"].data_ptr(),
                      done=outa["done"].data_ptr())
    _lib.check(lib.om_a3_task_step(task.dm.handle, task.handle, P(_padded(q, ld)), P(_padded(v, ld)), P(_padded(con, ld)), T,
                                   C.byref(sa), C.byref(oa), n, ld, st))
    torch.cuda.synchronize()
    for k in refa:
        _check(outa[k], refa[k], n, "om_a3_task_step." + k)
    _check(ints_p, ints_ref, n, "om_a3_task_step.ints")
    # GAE + moments on a padded rollout buffer
    Tt = 40
    r, vv, vn = (torch.randn((Tt, n), device="cuda", generator=g) for _ in range(3))
    ab = (torch.rand((Tt, n), device="cuda", generator=g) < 0.05).to(torch.uint8)
    la = (torch.rand((Tt, n), device="cuda", generator=g) < 0.1).to(torch.uint8)
    vt_ref, adv_ref = Kn.gae(r, vv, vn, ab, la, 0.99, 0.97)
    adv_p, vt_p = _padded(torch.zeros_like(r), ld), _padded(torch.zeros_like(r), ld)
    _lib.check(lib.om_gae(P(_padded(r, ld)), P(_padded(vv, ld)), P(_padded(vn, ld)), P(_padded(ab, ld)), P(_padded(la, ld)),
                          C.c_float(0.99), C.c_float(0.97), Tt, n, ld, P(adv_p), P(vt_p), st))
    mom_ref = Kn.moments_scalar(adv_ref)
    mom = torch.zeros(3, dtype=torch.float64, device="cuda")
    _lib.check(lib.om_moments(P(adv_p), Tt, 1, n, ld, P(mom), st))
    torch.cuda.synchronize()
    _check(adv_p, adv_ref, n, "om_gae.adv")
    _check(vt_p, vt_ref, n, "om_gae.v_target")
    assert torch.allclose(mom, mom_ref, rtol=1e-12), "om_moments must skip the padding columns"
    # discriminator with padded observations / noise / outputs
    from conftest import GOLDEN
    gd = np.load(GOLDEN / "discriminator_ref.npz")
    disc = Kn.Discriminator("vail", {k: gd["v_" + k] for k in ("w1", "b1", "w2", "b2", "wmu", "bmu", "wlv", "blv", "wd", "bd")})
    s = torch.randn((32, n), device="cuda", generator=g)
    eps = torch.randn((128, n), device="cuda", generator=g)
    mean, std = torch.zeros(32, device="cuda"), torch.ones(32, device="cuda")
    r_ref, d_ref = disc.reward(s, mean, std, eps=eps, want_d=True)
    r_p, d_p = _padded(torch.zeros_like(r_ref), ld), _padded(torch.zeros_like(d_ref), ld)
    _lib.check(lib.om_disc_reward(disc.handle, P(_padded(s, ld)), P(mean), P(std), P(_padded(eps, ld)), n, ld, P(r_p), P(d_p), st))
    torch.cuda.synchronize()
    _check(r_p, r_ref, n, "om_disc_reward.reward")
    _check(d_p, d_ref, n, "om_disc_reward.d")
