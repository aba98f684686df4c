"""CPU: the plain-C oracle (oracle/c, engine-style formulation) against the NumPy oracle -- two independently
written restatements of the same third-party arithmetic must agree to float64 round-off."""
import numpy as np

from conftest import GOLDEN, a3_random_states, assert_close
from oracle import c_oracle
from oracle import h1 as OH
from oracle import kinematics as K
from oracle import learner as L


def test_forward_matches_numpy_oracle(h1_model, a3_model, h1_states):
    for model, (q, v) in ((h1_model, h1_states), (a3_model, a3_random_states(a3_model, 200, seed=4))):
        a = c_oracle.forward(c_oracle.CModel(model), q, v)
        b = K.forward(model, q, v)
        for k in ("xpos", "xquat", "site_xpos", "site_xmat", "cvel", "subtree_com"):
            assert_close(a[k], b[k], k, rtol=1e-12, atol=1e-12)


def test_h1_playback_matches_numpy_oracle(h1_model):
    tab = np.load(GOLDEN / "trajectory_ref.npz")["table"]
    n_env, n_steps, seed = 5, 130, 4242
    o = c_oracle.h1_play(c_oracle.CModel(h1_model), OH.perm(h1_model), tab, seed, 3, n_env, n_steps)
    for e in range(n_env):
        ref = OH.play_trajectory_from_velocity(h1_model, tab, 1, n_steps, seed=seed, env_id=3 + e)
        assert np.array_equal(o["traj_no"][e], ref["traj_no"]) and np.array_equal(o["step_no"][e], ref["step_no"])
        assert np.array_equal(o["fallen"][e].astype(bool), ref["fallen"])
        assert np.array_equal(o["obs"][e], ref["obs"])
        assert_close(o["reward"][e], ref["reward"], "reward", rtol=1e-13, atol=1e-13)
        for k, w in (("xpos", 63), ("xquat", 84), ("site_xpos", 3), ("cvel", 126)):
            assert_close(o[k][e], ref[k].reshape(n_steps, w), k, rtol=1e-12, atol=1e-12)


def test_gae_matches_numpy_oracle():
    rng = np.random.default_rng(2)
    n, T = 6, 80
    r, v, vn = rng.normal(0, 1, (3, n, T))
    last = rng.random((n, T)) < 0.1
    ab = last & (rng.random((n, T)) < 0.5)
    vt, adv = c_oracle.gae(r, v, vn, ab, last, 0.99, 0.97)
    for e in range(n):
        f_vt, f_adv = L.compute_gae(v[e], vn[e], r[e], ab[e], last[e], 0.99, 0.97)
        assert_close(adv[e], f_adv, "adv", rtol=1e-13, atol=1e-13)
        assert_close(vt[e], f_vt, "vt", rtol=1e-13, atol=1e-13)
