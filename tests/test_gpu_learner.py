"""GPU parity: returns / GAE (K5) and moments / normalisation (K6) vs oracle/learner.py and the
reference's own PPOBuffer outputs (tests/golden/ppo_returns_ref.npz)."""
import numpy as np
import pytest

from conftest import GOLDEN, assert_close

pytestmark = pytest.mark.gpu


def _cu(a, dtype=None):
    import torch
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def test_ppo_returns_vs_reference_buffer():
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    z = np.load(GOLDEN / "ppo_returns_ref.npz")
    r, v = z["rewards"].astype(np.float32), z["values"].astype(np.float32)
    v_last = (~z["done_last"]) * z["v_boot"]
    ret, adv = Kn.ppo_returns(_cu(r), _cu(v), float(z["gamma"]), v_last=_cu(v_last, torch.float32))
    assert_close(ret.cpu().numpy(), z["returns"], "returns vs PPOBuffer.finish_path")
    # advantage normalisation ppo.py:335-336 (torch.std unbiased, eps 1e-5)
    mom = Kn.moments_scalar(adv)
    stats = Kn.adv_stats(mom, unbiased=True, eps=1e-5)
    an = Kn.normalize(adv, stats)
    assert_close(an.cpu().numpy(), z["adv_norm"], "normalised advantages", rtol=2e-5, atol=2e-5)
    # known-answer vector recorded in SURVEY.md
    kat, _ = Kn.ppo_returns(_cu(np.array([[1.0], [2.0], [3.0]], np.float32)), _cu(np.zeros((3, 1), np.float32)), 0.99,
                            v_last=_cu(np.array([10.0], np.float32)))
    assert_close(kat.cpu().numpy()[:, 0], [15.62329, 14.771, 12.9], "KAT")


@pytest.mark.parametrize("serial", [False, True])
@pytest.mark.parametrize("T,n", [(64, 1000), (1, 5), (37, 1), (500, 45), (257, 33), (512, 64), (400, 4100)])
def test_ppo_returns_segmented(T, n, serial, om_knob):
    from olympics_mujoco_b200 import kernels as Kn
    if serial:
        om_knob("serial_scan", int("1"))
    from oracle import learner as L
    rng = np.random.default_rng(T + n)
    r = rng.normal(0, 1, (T, n)).astype(np.float32)
    v = rng.normal(0, 1, (T, n)).astype(np.float32)
    vn = rng.normal(0, 1, (T, n)).astype(np.float32)
    done = rng.random((T, n)) < 0.1
    # oracle semantics: terminated paths bootstrap 0, the buffer end bootstraps with v_next[T-1]
    ref_ret, ref_adv = L.ppo_returns_segmented(r.astype(np.float64), v.astype(np.float64), done, vn.astype(np.float64), 0.99)
    ret, adv = Kn.ppo_returns(_cu(r), _cu(v), 0.99, path_end=_cu(done.astype(np.uint8)), v_last=_cu(vn[T - 1]))
    assert_close(ret.cpu().numpy(), ref_ret, "returns")
    assert_close(adv.cpu().numpy(), ref_adv, "advantages")


@pytest.mark.parametrize("serial", [False, True])                 # affine-scan kernel / one-thread-per-env kernel
@pytest.mark.parametrize("T,n", [(64, 513), (1000, 3), (500, 70), (1, 1), (1500, 33), (258, 31), (511, 97), (384, 4096)])
def test_gae_vs_mushroom_restatement(T, n, serial, om_knob):
    from olympics_mujoco_b200 import kernels as Kn
    if serial:
        om_knob("serial_scan", int("1"))
    from oracle import learner as L
    rng = np.random.default_rng(T * 7 + n)
    r = rng.normal(0, 1, (T, n)).astype(np.float32)
    v = rng.normal(0, 1, (T, n)).astype(np.float32)
    vn = rng.normal(0, 1, (T, n)).astype(np.float32)
    last = rng.random((T, n)) < 0.05
    absorbing = last & (rng.random((T, n)) < 0.5)
    vt, adv = Kn.gae(_cu(r), _cu(v), _cu(vn), _cu(absorbing.astype(np.uint8)), _cu(last.astype(np.uint8)), 0.99, 0.97)
    ref_vt, ref_adv = L.compute_gae_batched(v, vn, r, absorbing, last, 0.99, 0.97)
    assert_close(adv.cpu().numpy(), ref_adv, "adv")
    assert_close(vt.cpu().numpy(), ref_vt, "v_target")
    # flat (single-env dataset) form of compute_gae, column 0
    f_vt, f_adv = L.compute_gae(v[:, 0], vn[:, 0], r[:, 0], absorbing[:, 0], last[:, 0], 0.99, 0.97)
    assert_close(adv.cpu().numpy()[:, 0], f_adv, "adv flat")
    # gail_TRPO.py:128 normalisation (np.std population, eps 1e-8)
    stats = Kn.adv_stats(Kn.moments_scalar(adv), unbiased=False, eps=1e-8)
    an = Kn.normalize(adv, stats).cpu().numpy()
    assert_close(an, L.normalize_advantage_gail(adv.cpu().numpy()), "normalised adv", rtol=2e-5, atol=2e-5)


def test_moments_match_standardizer_and_running_mean_std():
    """K6 sums reproduce Standardizer.update_mean_std (networks.py:76-81) and, merged batch by batch, the
    reference's RunningMeanStd (normalize.py:190-208; golden fixture from the reference module)."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    z = np.load(GOLDEN / "standardizer_ref.npz")
    mom = torch.zeros(11, dtype=torch.float64, device="cuda")
    for k in ("x0", "x1", "x2"):
        Kn.moments(Kn.to_soa(z[k].astype(np.float32)), out=mom)
    m = mom.cpu().numpy()
    s, s2, cnt = m[:5] + 0.0, m[5:10] + 1e-2, m[10] + 1e-2             # Standardizer's initial sums
    mean = s / cnt
    std = np.sqrt(np.maximum(s2 / cnt - mean ** 2, 1e-2))
    assert_close(mean, z["mean"], "standardizer mean", rtol=1e-6, atol=1e-6)
    assert_close(std, z["std"], "standardizer std", rtol=1e-6, atol=1e-6)
    z = np.load(GOLDEN / "running_mean_std_ref.npz")
    x = np.concatenate([z["x0"], z["x1"], z["x2"]]).astype(np.float32)
    m = Kn.moments(Kn.to_soa(x)).cpu().numpy()
    cnt = m[8]
    mean = m[:4] / cnt
    var = m[4:8] / cnt - mean ** 2
    # RunningMeanStd starts from (mean 0, var 0, count 1e-4): merge the same prior analytically
    tot = cnt + 1e-4
    mean_r = mean * cnt / tot
    var_r = (var * cnt + mean ** 2 * 1e-4 * cnt / tot) / tot
    assert_close(mean_r, z["mean"], "rms mean", rtol=1e-5, atol=1e-6)
    assert_close(var_r, z["var"], "rms var", rtol=1e-5, atol=1e-6)


def test_moments_rollout_buffer_large():
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((16, 32, 20000), device="cuda", generator=g) * 3 + 1
    m = Kn.moments(x).cpu().numpy()
    xd = x.double()
    assert_close(m[:32], xd.sum(dim=(0, 2)).cpu().numpy(), "sum", rtol=1e-9, atol=1e-6)
    assert_close(m[32:64], (xd * xd).sum(dim=(0, 2)).cpu().numpy(), "sumsq", rtol=1e-9, atol=1e-6)
    assert m[64] == 16 * 20000
