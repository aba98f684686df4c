"""CPU: pin the oracle (and the host-side restatements) against fixtures produced by the reference's own
Python modules (tools/gen_golden.py) and against published known-answer vectors."""
import numpy as np

from conftest import GOLDEN, assert_close
from oracle import h1 as OH
from oracle import learner as L
from oracle import philox
from oracle.trajectory import TrajectoryState


def test_philox_random123_known_answers():
    u = lambda *v: [np.uint32(x) for x in v]
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        out = philox.philox4x32_10(u(*ctr), u(*key))
        assert tuple(int(x) for x in out) == exp


def test_philox_int_mapping_range_and_vectorisation():
    w = philox.draw(7, np.arange(1000, dtype=np.uint32), np.uint32(3))
    a = philox.to_int(w[0], 4)
    b = philox.to_int(w[1], 500)
    assert a.min() == 0 and a.max() == 3 and b.min() >= 0 and b.max() <= 499
    single = philox.draw(7, np.uint32(17), np.uint32(3))
    assert int(single[0]) == int(w[0][17])
    un = philox.to_unit(w[2])
    assert un.min() >= 0 and un.max() < 1 and np.all(un == un.astype(np.float32))


def test_trajectory_oracle_matches_reference_class():
    """Scripted reset / current / next sequence recorded from the reference's Trajectory object."""
    z = np.load(GOLDEN / "trajectory_ref.npz")
    tr = TrajectoryState(z["table"])
    for (op, traj_no, sub), sample, state in zip(z["script"], z["samples"], z["state"]):
        if op == 0:
            s = tr.reset_trajectory(substep_no=int(sub), traj_no=int(traj_no))
        elif op == 1:
            s = tr.get_current_sample()
        else:
            s = tr.get_next_sample()
        assert (tr.traj_no, tr.step_no) == tuple(state)
        if s is None:
            assert np.isnan(sample).all()
        else:
            assert np.array_equal(s, sample)


def test_h1_recorded_rollouts_reward_and_fallen():
    """saved_npz rollouts: key order = observation-spec order, non-terminal, reward statistics recorded in
    SURVEY.md section 8c and per-sample values from the reference's TargetVelocityReward."""
    from olympics_mujoco_b200 import mjcf
    model = mjcf.load_builtin("unitree_h1")
    z = np.load(GOLDEN / "saved_rollouts_ref.npz")
    assert list(z["keys"]) == OH.keys(model)
    assert OH.x_vel_idx(model) == 15
    assert list(OH.perm(model)) == [0, 1, 2, 3, 4, 5, 16, 13, 12, 11, 14, 15, 8, 7, 6, 9, 10]
    for name in ("vail_unprocessed_0", "gail_unprocessed_0", "vail_processed_0", "gail_processed_0"):
        obs = OH.create_observation(z[name])
        assert not OH.has_fallen(obs).any()
        r = OH.target_velocity_reward(obs, 15)
        assert np.array_equal(r, z[name + "_reward"])
    r = OH.target_velocity_reward(OH.create_observation(z["vail_unprocessed_0"]), 15)
    assert_close([r.min(), r.max(), r.mean()], [0.946320, 0.999996, 0.988174], "vail stats", rtol=0, atol=1e-6)
    r = OH.target_velocity_reward(OH.create_observation(z["gail_unprocessed_0"]), 15)
    assert_close([r.min(), r.max(), r.mean()], [0.941353, 0.999999, 0.990539], "gail stats", rtol=0, atol=1e-6)


def test_ppo_returns_oracle_vs_reference_buffer():
    z = np.load(GOLDEN / "ppo_returns_ref.npz")
    ret, adv = L.ppo_returns(z["rewards"], z["values"], z["done_last"], z["v_boot"], float(z["gamma"]))
    assert_close(ret, z["returns"], "returns", rtol=1e-12, atol=1e-12)
    # torch.Tensor() made the reference's advantages fp32: compare at fp32 resolution
    assert_close(L.normalize_advantage_ppo(adv), z["adv_norm"], "adv", rtol=1e-5, atol=1e-5)
    assert_close(z["kat_returns"], [15.62329, 14.771, 12.9], "KAT", rtol=1e-9, atol=1e-9)
    # segmented form reduces to the single-path form
    done = np.zeros(z["rewards"].shape, bool)
    done[-1] = z["done_last"]
    vn = np.zeros(z["rewards"].shape)
    vn[-1] = z["v_boot"]
    ret2, _ = L.ppo_returns_segmented(z["rewards"], z["values"], done, vn, float(z["gamma"]))
    assert_close(ret2, ret, "segmented", rtol=1e-12, atol=1e-12)


def test_gae_batched_equals_flat_restatement():
    rng = np.random.default_rng(0)
    T, n = 50, 7
    r, v, vn = rng.normal(0, 1, (3, T, n))
    last = rng.random((T, n)) < 0.1
    absorbing = last & (rng.random((T, n)) < 0.5)
    vt, adv = L.compute_gae_batched(v, vn, r, absorbing, last, 0.99, 0.97)
    for e in range(n):
        f_vt, f_adv = L.compute_gae(v[:, e], vn[:, e], r[:, e], absorbing[:, e], last[:, e], 0.99, 0.97)
        assert_close(adv[:, e], f_adv, "adv", rtol=1e-12, atol=1e-12)
        assert_close(vt[:, e], f_vt, "vt", rtol=1e-12, atol=1e-12)
    # hand-computed 3-step case: gamma = 0.5, lam = 0.5, second step ends a non-absorbing episode
    vt, adv = L.compute_gae([1.0, 2.0, 3.0], [2.0, 3.0, 4.0], [1.0, 1.0, 1.0], [False, False, True],
                            [False, True, True], 0.5, 0.5)
    assert_close(adv, [1 + 1 - 1 + 0.25 * 0.5, 1 + 1.5 - 2, 1 - 3], "hand", rtol=1e-12, atol=1e-12)


def test_standardizer_and_running_mean_std_oracles():
    z = np.load(GOLDEN / "standardizer_ref.npz")
    s = L.Standardizer()
    for k in ("x0", "x1", "x2"):
        s.update_mean_std(z[k])
    assert_close(s.mean, z["mean"], "mean", rtol=1e-12, atol=1e-12)
    assert_close(s.std, z["std"], "std", rtol=1e-12, atol=1e-12)
    z = np.load(GOLDEN / "running_mean_std_ref.npz")
    rms = L.RunningMeanStd(epsilon=1e-4, shape=(4,))
    for k in ("x0", "x1", "x2"):
        rms.update(z[k])
    assert_close(rms.mean, z["mean"], "rms mean", rtol=1e-12, atol=1e-12)
    assert_close(rms.var, z["var"], "rms var", rtol=1e-12, atol=1e-12)
    # the reference's own self-check property (rl/envs/normalize.py:210-225)
    rng = np.random.default_rng(1)
    xs = [rng.normal(size=(k, 2)) for k in (3, 4, 5)]
    rms = L.RunningMeanStd(epsilon=0.0, shape=(2,))
    for x in xs:
        rms.update(x)
    x = np.concatenate(xs)
    assert np.allclose([x.mean(axis=0), x.var(axis=0)], [rms.mean, rms.var])


def test_discriminator_oracle_vs_reference_networks():
    z = np.load(GOLDEN / "discriminator_ref.npz")
    assert int(z["n_par_vail"]) == 74497 and int(z["n_par_gail"]) == 148481
    pv = {k[2:]: z[k] for k in z.files if k.startswith("v_")}
    d, mu, logvar = L.vail_forward(pv, z["s"], z["eps"], z["vail_mean"], z["vail_std"])
    assert_close(mu, z["vail_mu"], "mu", rtol=1e-4, atol=1e-5)            # reference ran in fp32
    assert_close(logvar, z["vail_logvar"], "logvar", rtol=1e-4, atol=1e-5)
    assert_close(d, z["vail_d"][:, 0], "d", rtol=1e-4, atol=1e-5)
    assert_close(L.discrim_reward(d), z["vail_reward"], "vail reward", rtol=1e-4, atol=1e-5)
    pg = {k[2:]: z[k] for k in z.files if k.startswith("g_")}
    dg = L.gail_forward(pg, z["s"], z["gail_mean"], z["gail_std"])
    assert_close(dg, z["gail_d"][:, 0], "gail d", rtol=1e-4, atol=1e-5)
    assert_close(L.discrim_reward(dg), z["gail_reward"], "gail reward", rtol=1e-4, atol=1e-5)


def test_mirror_oracle_vs_reference_symmetric_env():
    """oracle.learner.mirror against the outputs of the reference's own SymmetricEnv (tests/golden/mirror_ref.npz,
    generated by tools/gen_golden.py:gen_mirror from rl/envs/wrappers.py)."""
    from oracle import learner as L
    g = np.load(GOLDEN / "mirror_ref.npz")
    np.testing.assert_allclose(L.mirror(g["obs"], g["mirrored_obs"]), g["mirror_obs"], rtol=0, atol=0)
    np.testing.assert_allclose(L.mirror(g["act"], g["mirrored_acts"]), g["mirror_act"], rtol=0, atol=0)
    np.testing.assert_allclose(L.mirror(g["obs"], g["mirrored_obs"], g["clock_inds"]), g["mirror_clock_obs"], rtol=0, atol=2e-7)
    m = L.symmetry_matrix(g["mirrored_obs"])
    assert np.array_equal(m @ m, np.eye(41))                       # the A3 observation mirror is an involution


def test_disc_fit_loss_oracle_vs_reference_losses():
    """N2: oracle/learner.py's GAIL / VDB loss restatements against the reference's own GailDiscriminatorLoss and
    VDBLoss (imitation_lib/utils/math.py, executed by tools/gen_golden.py:gen_disc_loss): value, autograd gradient,
    per-sample entropy and KL, three consecutive beta updates."""
    from oracle import learner as L
    z = np.load(GOLDEN / "disc_loss_ref.npz")
    loss, grad = L.gail_disc_loss(z["logits"], z["t01"], float(z["entcoeff"]))
    assert abs(loss - float(z["gail_loss01"])) < 1e-6
    assert_close(grad, z["gail_grad01"].ravel(), "d loss / d logit vs torch autograd", rtol=1e-5, atol=1e-9)
    assert abs(L.gail_disc_loss(z["logits"], z["tnoisy"], float(z["entcoeff"]))[0] - float(z["gail_loss_noisy"])) < 1e-6
    assert_close(L.logit_bernoulli_entropy(z["logits"]), z["ent"], "bernoulli entropy", rtol=1e-5, atol=1e-6)
    kl = L.vdb_kl(z["mu"], z["logvar"])
    assert_close(kl, z["kl"], "kl", rtol=1e-5, atol=1e-5)
    beta = float(z["betas"][0])
    for i in range(3):
        loss, beta = L.vdb_loss(z["logits"], kl, z["t01"], beta, float(z["info_constraint"]), float(z["lr_beta"]))
        assert abs(loss - z["vdb_losses"][i]) < 1e-5 * abs(z["vdb_losses"][i])
        assert abs(beta - z["betas"][i + 1]) < 1e-7


def test_expert_index_contract_is_a_permutation_per_epoch():
    """N2: every epoch of the expert-minibatch contract is a sample without replacement (what taking the first batch of
    mushroom_rl's shuffled minibatch_generator gives), draws differ, and the stream is reproducible."""
    from oracle import learner as L
    for n_src in (1, 2, 5, 1999, 4096, 70001):
        idx = L.expert_indices(3, 0, 2 * n_src + 3, n_src)
        assert sorted(idx[:n_src]) == list(range(n_src)) and sorted(idx[n_src:2 * n_src]) == list(range(n_src))
    a, b = L.expert_indices(3, 0, 512, 1999), L.expert_indices(3, 1, 512, 1999)
    assert not np.array_equal(a, b) and np.array_equal(a, L.expert_indices(3, 0, 512, 1999))
    # roughly uniform first positions over many draws
    first = np.array([L.expert_indices(9, d, 1, 10)[0] for d in range(2000)])
    assert np.bincount(first, minlength=10).min() > 120


def test_ppo_loss_oracle_vs_reference_update_policy():
    """N3: oracle/learner.py ppo_losses against the six values PPO.update_policy returned (rl/algos/ppo.py:231-282, run by
    tools/gen_golden.py:gen_ppo_loss with the reference's SymmetricEnv mirror functions) and torch autograd's gradients."""
    from oracle import learner as L
    z = np.load(GOLDEN / "ppo_loss_ref.npz")
    o = L.ppo_losses(z["logp"], z["old_logp"], z["adv"], z["mask"], z["values"], z["returns"], z["entropy"], z["det_actions"],
                     z["mirror_raw"], z["mirrored_acts"], float(z["clip"]), float(z["vf_coeff"]))
    for k in ("actor_loss", "entropy_penalty", "critic_loss", "approx_kl", "mirror_loss", "clip_fraction"):
        assert abs(o[k] - float(z[k])) < 2e-6 * max(1.0, abs(float(z[k]))), k
    assert 0.05 < o["clip_fraction"] < 0.5
    assert_close(o["dlogp"], z["dlogp"].ravel(), "d actor_loss / d logp", rtol=1e-5, atol=1e-9)
    assert_close(o["dvalues"], z["dvalues"].ravel(), "d critic_loss / d values", rtol=1e-5, atol=1e-9)
