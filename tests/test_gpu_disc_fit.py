"""GPU parity for the N2 row (data side of GAIL._fit_discriminator, gail_TRPO.py:167-220): expert minibatch gather,
forward pass with logit + KL outputs, loss statistics and the VDB beta update, through the C ABI, against the float64
oracle (oracle/learner.py, pinned to the reference's loss classes by tests/golden/disc_loss_ref.npz)."""
import numpy as np
import pytest

from conftest import GOLDEN, assert_close

pytestmark = pytest.mark.gpu


def _t(a, dtype=np.float32):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype), device="cuda")


@pytest.mark.parametrize("n_src,batch", [(1, 5), (1999, 512), (1999, 5000), (4096, 4096), (70001, 65536)])
def test_expert_minibatch_indices_and_gather_exact(n_src, batch):
    """Indices are bit-exact against the contract; gathered states / next_states are exact copies of the rows."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from oracle import learner as L
    rng = np.random.default_rng(n_src)
    full = rng.normal(0, 1, (n_src + 1, 32)).astype(np.float32)
    src = Kn.to_soa(full)
    for draw in (0, 3):
        s, nxt, idx = Kn.expert_minibatch(src, n_src, seed=11, draw=draw, batch=batch, want_next=True, want_idx=True)
        torch.cuda.synchronize()
        want = L.expert_indices(11, draw, batch, n_src)
        assert np.array_equal(idx.cpu().numpy(), want)
        assert np.array_equal(s.cpu().numpy().T, full[want])
        assert np.array_equal(nxt.cpu().numpy().T, full[want + 1])


def test_expert_dataset_from_create_dataset_layout():
    """ExpertDataset stores states + the last next_state so that next_states[i] is row i + 1 (trajectory.py:170-171)."""
    import torch
    from olympics_mujoco_b200.learners import ExpertDataset
    rng = np.random.default_rng(0)
    states = rng.normal(0, 1, (301, 34))
    ds = dict(states=states[:-1], next_states=states[1:], absorbing=np.zeros(300), last=np.zeros(301))
    mask = np.arange(32)
    ex = ExpertDataset(ds, state_mask=mask, seed=5)
    s, nxt, idx = ex.sample(300, want_next=True, want_idx=True)
    torch.cuda.synchronize()
    i = idx.cpu().numpy()
    assert sorted(i) == list(range(300)) and ex.draws == 1
    assert np.array_equal(s.cpu().numpy().T, states[i][:, mask].astype(np.float32))
    assert np.array_equal(nxt.cpu().numpy().T, states[i + 1][:, mask].astype(np.float32))


@pytest.mark.parametrize("noisy", [False, True])
def test_loss_statistics_vs_reference_losses(noisy):
    """om_disc_loss_stats on the reference fixture's logits: GAIL loss, its gradient, entropy, accuracies."""
    import torch
    from olympics_mujoco_b200.learners import GailDiscriminatorLoss, VDBLoss
    from oracle import learner as L
    z = np.load(GOLDEN / "disc_loss_ref.npz")
    logits = z["logits"].ravel()
    n = logits.size // 2
    target = z["tnoisy"].ravel() if noisy else None
    loss, grad, st = GailDiscriminatorLoss(float(z["entcoeff"]))(_t(logits), n, target=None if target is None else _t(target))
    torch.cuda.synchronize()
    ref = float(z["gail_loss_noisy"] if noisy else z["gail_loss01"])
    assert abs(loss - ref) < 1e-5 * max(1.0, abs(ref))
    want_loss, want_grad = L.gail_disc_loss(logits, z["tnoisy"] if noisy else z["t01"], float(z["entcoeff"]))
    assert abs(loss - want_loss) < 1e-6
    assert_close(grad.cpu().numpy(), want_grad, "d loss / d logit", rtol=1e-5, atol=1e-9)
    if not noisy:
        assert_close(grad.cpu().numpy(), z["gail_grad01"].ravel(), "gradient vs torch autograd", rtol=2e-5, atol=1e-9)
    sig = 1 / (1 + np.exp(-logits.astype(np.float64)))
    assert st["accuracy_gen"] == np.mean(sig[:n] < 0.5) and st["accuracy_exp"] == np.mean(sig[n:] > 0.5)
    assert abs(st["d_out_gen"] - sig[:n].mean()) < 1e-6 and abs(st["d_out_exp"] - sig[n:].mean()) < 1e-6
    assert abs(st["bernoulli_ent"] - z["ent"].mean()) < 1e-6
    # VDB: three fits move beta exactly like the reference's VDBLoss
    vl = VDBLoss(float(z["info_constraint"]), float(z["lr_beta"]))
    for i in range(3):
        v, _, vst = vl(_t(logits), n, kl=_t(z["kl"]))
        assert abs(v - z["vdb_losses"][i]) < 1e-5 * abs(z["vdb_losses"][i])
        assert abs(vl.beta - z["betas"][i + 1]) < 1e-7


@pytest.mark.parametrize("kernel", ["a_in_tmem_one_cta", "a_in_tmem_rotating", "a_in_tmem_split_release", "alternating_producers", "a_in_tmem_two_ctas", "two_ctas_per_sm", "one_cta_per_sm", "one_producer_group"])
def test_vail_forward_logit_and_kl_vs_oracle(kernel, om_knob):
    """The fit's forward pass (logit + per-sample KL) out of the tcgen05 kernels against the float64 oracle and the
    reference network's own mu / logvar (discriminator_ref.npz); ragged sample counts."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from oracle import learner as L
    om_knob("disc_vail2", {"a_in_tmem_one_cta": 4, "a_in_tmem_rotating": 5, "a_in_tmem_split_release": 6, "alternating_producers": 7, "a_in_tmem_two_ctas": 3, "two_ctas_per_sm": 1}.get(kernel, 0))
    if kernel == "one_producer_group":
        om_knob("disc_pg2", int("0"))
    g = np.load(GOLDEN / "discriminator_ref.npz")
    p = {k: g["v_" + k] for k in ("w1", "b1", "w2", "b2", "wmu", "bmu", "wlv", "blv", "wd", "bd")}
    disc = Kn.Discriminator("vail", p)
    mean, std = g["vail_mean"].astype(np.float32), g["vail_std"].astype(np.float32)
    for n in (256, 1, 131):
        s, eps = g["s"][:n], g["eps"][:n]
        fw = disc.forward(_t(s.T), _t(mean), _t(std), eps=_t(eps.T), want_reward=True)
        torch.cuda.synchronize()
        d64, mu, lv = L.vail_forward({k: v.astype(np.float64) for k, v in p.items()}, s, eps, mean.astype(np.float64),
                                     std.astype(np.float64))
        assert_close(fw["logit"].cpu().numpy(), d64, "logit")
        assert_close(fw["kl"].cpu().numpy(), L.vdb_kl(mu, lv), "kl vs float64 oracle")
        assert_close(fw["kl"].cpu().numpy(), L.vdb_kl(g["vail_mu"][:n], g["vail_logvar"][:n]), "kl vs the reference net's mu / logvar",
                     rtol=2e-5, atol=2e-5)
        assert_close(fw["reward"].cpu().numpy(), -np.log(1 - 1 / (1 + np.exp(-d64)) + 1e-8), "reward")


def test_fit_batch_end_to_end_vs_oracle():
    """DiscriminatorReward.fit_batch = one epoch of _fit_discriminator without the weight update: expert sample,
    standardiser update on [policy; expert], forward, VDB loss -- restated step by step with the oracle."""
    import torch
    from olympics_mujoco_b200.learners import DiscriminatorReward, ExpertDataset, VDBLoss
    from oracle import learner as L
    g = np.load(GOLDEN / "discriminator_ref.npz")
    p = {k: g["v_" + k] for k in ("w1", "b1", "w2", "b2", "wmu", "bmu", "wlv", "blv", "wd", "bd")}
    rng = np.random.default_rng(2)
    expert_states = (rng.normal(0, 1, (501, 32)) * 1.5 + 0.3).astype(np.float32)
    ds = dict(states=expert_states[:-1], next_states=expert_states[1:])
    ex = ExpertDataset(ds, seed=4)
    dr = DiscriminatorReward("vail", p, seed=0)
    n = 192
    plcy = g["s"][:n]
    eps = rng.normal(0, 1, (2 * n, 128)).astype(np.float32)
    vl = VDBLoss(0.5, 1e-5)
    value, st, batch = dr.fit_batch(_t(plcy.T), ex, vl, eps=_t(eps.T))
    torch.cuda.synchronize()
    idx = L.expert_indices(4, 0, n, 500)
    inputs = np.concatenate([plcy, expert_states[idx]])
    assert np.array_equal(batch["inputs"].cpu().numpy().T, inputs)
    # the reference folds the batch in TWICE before it is standardised: the explicit update_mean_std (gail_TRPO.py:200-201)
    # and the forward inside D.fit (networks.py:66-71) -- one Standardizer object serves both
    sd = L.Standardizer()
    sd.update_mean_std(inputs)
    sd.forward(inputs)
    mean32, std32 = sd.mean.astype(np.float32).astype(np.float64), sd.std.astype(np.float32).astype(np.float64)
    d64, mu, lv = L.vail_forward({k: v.astype(np.float64) for k, v in p.items()}, inputs, eps, mean32, std32)
    assert_close(batch["logit"].cpu().numpy(), d64, "logit")
    t = np.concatenate([np.zeros(n), np.ones(n)])
    want, beta = L.vdb_loss(d64, L.vdb_kl(mu, lv), t, 0.1, 0.5, 1e-5)
    assert abs(value - want) < 1e-5 * abs(want) and abs(vl.beta - beta) < 1e-9
    sig = 1 / (1 + np.exp(-d64))
    assert_close(batch["dlogit"].cpu().numpy(), (sig - t) / (2 * n), "bce gradient", rtol=1e-4, atol=1e-8)
    # running sums after one fit epoch (with the logging forwards of an attached summary writer) plus one reward call,
    # against the reference Standardizer restatement replaying the reference's call sequence
    dr2 = DiscriminatorReward("vail", p, seed=0)
    ex2 = ExpertDataset(ds, seed=4)
    dr2.fit_batch(_t(plcy.T), ex2, VDBLoss(0.5, 1e-5), eps=_t(eps.T), logging_forwards=True)
    reward_states = g["s"][40:168]
    dr2.make_discrim_reward(_t(reward_states.T), eps=_t(eps[:128].T))
    torch.cuda.synchronize()
    sd2 = L.Standardizer()
    sd2.update_mean_std(inputs)                      # gail_TRPO.py:200-201
    sd2.forward(inputs)                              # D.fit
    for x in (inputs, inputs[n:], inputs[:n], inputs, inputs[n:], inputs[:n]):
        sd2.forward(x)                               # _discriminator_logging :222-258 (full, demo, policy, full, demo, policy)
    sd2.forward(reward_states)                       # make_discrim_reward :320-327
    run = dr2.standardizer.running.cpu().numpy()
    assert_close(run[:32], sd2._sum, "running sum", rtol=1e-12, atol=1e-9)
    assert_close(run[32:64], sd2._sumsq, "running sum of squares", rtol=1e-12, atol=1e-9)
    assert abs(run[64] - float(sd2._count[0])) < 1e-9
    assert_close(dr2.standardizer.mean.cpu().numpy(), sd2.mean, "mean", rtol=1e-12, atol=1e-12)
    assert_close(dr2.standardizer.std.cpu().numpy(), sd2.std, "std", rtol=1e-12, atol=1e-12)
