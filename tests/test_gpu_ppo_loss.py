"""GPU parity for the N3 losses: PPO.update_policy (rl/algos/ppo.py:231-282) as one pass over the minibatch, through the
C ABI, against the reference method's own outputs (tests/golden/ppo_loss_ref.npz) and the float64 oracle."""
import numpy as np
import pytest

from conftest import GOLDEN, assert_close

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device="cuda")


def test_ppo_update_losses_vs_reference_update_policy():
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from olympics_mujoco_b200.learners import ppo_update_losses
    z = np.load(GOLDEN / "ppo_loss_ref.npz")
    mir = Kn.make_mirror_spec(list(z["mirrored_acts"]))
    out, dlogp, dvalues = ppo_update_losses(_t(z["logp"].ravel()), _t(z["old_logp"].ravel()), _t(z["adv"].ravel()),
                                            _t(z["mask"].ravel()), _t(z["values"].ravel()), _t(z["returns"].ravel()),
                                            clip=float(z["clip"]), vf_coeff=float(z["vf_coeff"]), entropy=_t(z["entropy"].T),
                                            act=_t(z["det_actions"].T), act_mirror=_t(z["mirror_raw"].T), action_mirror=mir)
    torch.cuda.synchronize()
    for k in ("actor_loss", "entropy_penalty", "critic_loss", "approx_kl", "mirror_loss", "clip_fraction"):
        assert abs(out[k] - float(z[k])) < 1e-5 * max(1.0, abs(float(z[k]))), (k, out[k], float(z[k]))
    assert_close(dlogp.cpu().numpy(), z["dlogp"].ravel(), "d actor_loss / d logp vs torch autograd", rtol=1e-5, atol=1e-9)
    assert_close(dvalues.cpu().numpy(), z["dvalues"].ravel(), "d critic_loss / d values vs torch autograd", rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("n", [1, 33, 100003])
def test_ppo_loss_ragged_sizes_and_optional_terms(n):
    """Random batches of ragged sizes against the float64 oracle; without mask / value / entropy / mirror inputs the
    corresponding sums stay zero; a pre-mirrored act_mirror (no spec) gives the same mirror loss."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from oracle import learner as L
    rng = np.random.default_rng(n)
    nu = 12
    mirrored = [6, -7, -8, 9, -10, 11, 0.1, -1, -2, 3, -4, 5]
    old = rng.normal(-8, 2, n); logp = old + rng.normal(0, 0.15, n)
    adv, mask = rng.normal(0, 1, n), (rng.random(n) > 0.1).astype(np.float64)
    values, returns = rng.normal(0, 1, n), rng.normal(0, 1, n)
    ent, act, mraw = rng.normal(0.5, 0.1, (n, nu)), rng.normal(0, 1, (n, nu)), rng.normal(0, 1, (n, nu))
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    o = L.ppo_losses(f32(logp), f32(old), f32(adv), mask, f32(values), f32(returns), f32(ent), f32(act), f32(mraw), mirrored, 0.2, 0.5)
    sums, dlogp, dvalues = Kn.ppo_loss_stats(_t(logp), _t(old), _t(adv), 0.2, mask=_t(mask), values=_t(values), returns=_t(returns),
                                             vf_coeff=0.5, entropy=_t(ent.T), act=_t(act.T), act_mirror=_t(mraw.T),
                                             action_mirror=Kn.make_mirror_spec(mirrored))
    torch.cuda.synchronize()
    s = sums.cpu().numpy()
    assert s[6] == n
    got = dict(actor_loss=-s[0] / n, entropy_penalty=-s[1] / (n * nu), critic_loss=0.5 * s[2] / n, approx_kl=s[3] / n,
               mirror_loss=s[4] / (n * nu), clip_fraction=s[5] / n)
    for k, v in got.items():
        assert abs(v - o[k]) < 2e-5 * max(1.0, abs(o[k])), (k, v, o[k])
    assert_close(dlogp.cpu().numpy(), o["dlogp"], "dlogp", rtol=2e-5, atol=1e-9)
    assert_close(dvalues.cpu().numpy(), o["dvalues"], "dvalues", rtol=2e-5, atol=1e-9)
    pre = L.mirror(f32(mraw), mirrored)
    s2, _, _ = Kn.ppo_loss_stats(_t(logp), _t(old), _t(adv), 0.2, act=_t(act.T), act_mirror=_t(pre.T), want_grad=False)
    torch.cuda.synchronize()
    s2 = s2.cpu().numpy()
    assert abs(s2[4] - s[4]) < 1e-6 * max(1.0, s[4]) and s2[1] == 0 and s2[2] == 0
