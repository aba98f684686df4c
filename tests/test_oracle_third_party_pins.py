"""CPU: pins for the two third-party pieces the reference calls but does not ship:

* ``transforms3d`` (default axes 'sxyz'; call sites StickFigureA3.py:160-161,227, walking_task.py:76,119,204-223,
  mujoco_robot_interface.py:344) -- ``oracle/tf3.py`` is checked against SciPy's ``Rotation`` (scipy IS installed here:
  an independent, widely used implementation of the same conventions: static-frame x-y-z Euler angles = scipy's
  lowercase "xyz" extrinsic sequence; scipy quaternions are scalar-LAST).
* ``mushroom_rl.utils.value_functions.compute_gae`` (call site imitation_lib/imitation/gail_TRPO.py:126-128) --
  ``oracle/learner.py`` restates its backward loop; here it is checked against the DEFINITION of GAE(lambda) written as
  explicit forward sums over each segment (Schulman et al. 2016, eq. 16: A_t = sum_l (gamma lambda)^l delta_{t+l}), which
  shares no recurrence with it.
"""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from conftest import assert_close
from oracle import learner as L
from oracle import tf3


def _wxyz(r):
    q = r.as_quat()                     # scalar-last
    return np.concatenate([q[..., 3:], q[..., :3]], axis=-1)


def _same_rotation(a, b, name):
    s = np.sign(np.sum(a * b, axis=-1, keepdims=True))
    assert_close(a, b * np.where(s == 0, 1.0, s), name, rtol=1e-12, atol=1e-12)


def test_tf3_matches_scipy_rotation():
    rng = np.random.default_rng(0)
    n = 5000
    ai, aj, ak = rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi / 2 + 1e-3, np.pi / 2 - 1e-3, n), rng.uniform(-np.pi, np.pi, n)
    ref = Rotation.from_euler("xyz", np.stack([ai, aj, ak], axis=1))          # extrinsic x, then y, then z
    # euler2quat
    _same_rotation(tf3.euler2quat(ai, aj, ak), _wxyz(ref), "euler2quat('sxyz') vs scipy")
    # quat2mat, on unnormalised quaternions too (transforms3d scales by 2/|q|^2)
    q = _wxyz(ref) * rng.uniform(0.3, 3.0, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1))
    assert_close(tf3.quat2mat(q), ref.as_matrix(), "quat2mat vs scipy", rtol=1e-12, atol=1e-12)
    # quat2euler / mat2euler: the same angles back (pitch inside (-pi/2, pi/2): unique)
    ex, ey, ez = tf3.quat2euler(q)
    got = Rotation.from_euler("xyz", np.stack([ex, ey, ez], axis=1))
    assert_close(got.as_matrix(), ref.as_matrix(), "quat2euler round trip through scipy", rtol=1e-11, atol=1e-11)
    assert_close(np.stack([ex, ey, ez], 1), ref.as_euler("xyz"), "quat2euler angles vs scipy", rtol=1e-9, atol=1e-9)
    # euler2mat(0, 0, yaw) and mat2quat
    assert_close(tf3.rotz(ak), Rotation.from_euler("z", ak[:, None]).as_matrix(), "euler2mat(0,0,yaw)", rtol=1e-12, atol=1e-12)
    for i in range(200):
        _same_rotation(tf3.mat2quat(ref[i].as_matrix())[None], _wxyz(ref[i])[None], "mat2quat vs scipy")
        assert tf3.mat2quat(ref[i].as_matrix())[0] >= 0
    # the composition the A3 observation uses (StickFigureA3.py:160-161): euler2quat(roll, pitch, 0) of quat2euler(q)
    roll, pitch, _ = tf3.quat2euler(q)
    want = _wxyz(Rotation.from_euler("xyz", np.stack([ref.as_euler("xyz")[:, 0], ref.as_euler("xyz")[:, 1], np.zeros(n)], 1)))
    _same_rotation(tf3.euler2quat(roll, pitch, np.zeros(n)), want, "root orientation without yaw")


def _gae_by_definition(v, v_next, r, absorbing, last, gamma, lam):
    """GAE(lambda) as explicit sums: segments end where last is set (and at the end of the data); inside a segment
    delta_t = r_t + gamma v'_t - v_t, at its end the bootstrap is dropped iff the transition is absorbing."""
    n = len(r)
    adv = np.zeros(n)
    ends = [k for k in range(n) if last[k] or k == n - 1]
    start = 0
    for e in ends:
        delta = r[start:e + 1] + gamma * v_next[start:e + 1] - v[start:e + 1]
        if absorbing[e]:
            delta[-1] = r[e] - v[e]
        for t in range(start, e + 1):
            l = np.arange(0, e - t + 1)
            adv[t] = np.sum((gamma * lam) ** l * delta[t - start:])
        start = e + 1
    return adv + v, adv


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_compute_gae_matches_the_definition(seed):
    rng = np.random.default_rng(seed)
    n = 400
    v, v_next, r = rng.normal(0, 1, n), rng.normal(0, 1, n), rng.normal(0, 1, n)
    last = rng.random(n) < 0.04
    absorbing = last & (rng.random(n) < 0.5)
    if seed == 2:
        absorbing[-1] = True                      # the data ends on an absorbing transition
    vt, adv = L.compute_gae(v, v_next, r, absorbing, last, 0.99, 0.97)
    vt2, adv2 = _gae_by_definition(v, v_next, r, absorbing, last, 0.99, 0.97)
    assert_close(adv, adv2, "advantage", rtol=1e-11, atol=1e-11)
    assert_close(vt, vt2, "value target", rtol=1e-11, atol=1e-11)
    # the time-major batched form used for rollout buffers: each column is an independent flat dataset
    T, N = 60, 7
    vb, vnb, rb = rng.normal(0, 1, (T, N)), rng.normal(0, 1, (T, N)), rng.normal(0, 1, (T, N))
    lb = rng.random((T, N)) < 0.05
    ab = lb & (rng.random((T, N)) < 0.5)
    vtb, advb = L.compute_gae_batched(vb, vnb, rb, ab, lb, 0.99, 0.97)
    for j in range(N):
        _, a = _gae_by_definition(vb[:, j], vnb[:, j], rb[:, j], ab[:, j], lb[:, j], 0.99, 0.97)
        assert_close(advb[:, j], a, f"batched column {j}", rtol=1e-11, atol=1e-11)
