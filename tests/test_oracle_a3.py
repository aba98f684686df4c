"""The A3 oracle (oracle/a3.py) against the reference's OWN WalkingTask / reward code.

``tests/golden/a3_task_ref.npz`` was produced by ``tools/gen_golden.py:gen_a3_task`` by running
``olympic_mujoco/tasks/walking_task.py`` + ``tasks/rewards.py`` of the reference (loaded by path, with a fake
``MujocoRobotInterface`` client backed by the oracle's FK and a ``transforms3d`` shim backed by oracle/tf3.py).
It pins the task's control flow, integer state and reward arithmetic; FK and tf3 stay "parity unpinned".
"""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import a3 as OA


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "a3_task_ref.npz")


def _ints(ts):
    return [ts.phase, ts.t1, ts.t2, ts.target_reached_frames, ts.mode, ts.seq_len, int(ts.target_reached)]


def test_reset_matches_reference_task(a3_model, gold):
    seed, it = int(gold["seed"]), float(gold["iteration_count"])
    for e in range(gold["reset_qpos"].shape[0]):
        np.testing.assert_array_equal(OA.reset_uniforms(seed, e, 0), gold["reset_u"][e])
        qpos, qvel, ts, obs = OA.reset(a3_model, seed, e, 0, iteration_count=it)
        np.testing.assert_array_equal(qpos, gold["reset_qpos"][e])
        assert _ints(ts) == list(gold["reset_ints"][e])
        np.testing.assert_allclose(ts.sequence, gold["reset_sequence"][e], rtol=0, atol=1e-13)
        assert obs.shape == (41,)
    assert set(gold["reset_ints"][:, 4]) == {OA.STANDING, OA.FORWARD}
    assert set(gold["reset_ints"][:, 0]) == {0, OA.PERIOD // 2}


def test_step_tail_matches_reference_task(a3_model, gold):
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    lut = phase_clock_lut()
    seed, it = int(gold["seed"]), float(gold["iteration_count"])
    n_env, T = gold["step_done"].shape
    assert gold["step_ints"][..., 1].max() >= 3 and gold["step_done"].any() and not gold["step_done"].all()
    for e in range(n_env):
        _, _, ts, _ = OA.reset(a3_model, seed, e, 0, iteration_count=it)
        for t in range(T):
            c = gold["step_contact"][e, t].astype(np.float64)
            con = OA.Contact(l_grf=c[0], r_grf=c[1], min_z=c[2], foot_contact=bool(c[3]), bad_collision=bool(c[4]))
            obs, total, done, terms = OA.step_tail(a3_model, gold["step_qpos"][e, t].astype(np.float64),
                                                   gold["step_qvel"][e, t].astype(np.float64), ts, con, lut)
            assert _ints(ts) == list(gold["step_ints"][e, t]), (e, t)
            assert done == bool(gold["step_done"][e, t]), (e, t)
            np.testing.assert_allclose(terms, gold["step_terms"][e, t], rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(obs[33:], gold["step_goal"][e, t], rtol=1e-12, atol=1e-13)
            assert abs(total - gold["step_terms"][e, t].sum()) < 1e-12
            # get_obs (StickFigureA3.py:144-178) cannot be imported (MuJoCo); its structure is checked here
            ph = ts.phase
            np.testing.assert_allclose(obs[31:33], [np.sin(2 * np.pi * ph / 88), np.cos(2 * np.pi * ph / 88)])
            assert abs(np.linalg.norm(obs[0:4]) - 1) < 1e-12
