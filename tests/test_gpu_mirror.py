"""GPU parity of the N3 mirror transforms against fixtures produced by the reference's own SymmetricEnv
(rl/envs/wrappers.py, imported by path in tools/gen_golden.py -> tests/golden/mirror_ref.npz)."""
import numpy as np
import pytest

from conftest import GOLDEN, assert_close

pytestmark = pytest.mark.gpu


def test_mirror_obs_action_clock_vs_reference_wrapper():
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from olympics_mujoco_b200.environments.stick_figure_a3 import StickFigureA3
    g = np.load(GOLDEN / "mirror_ref.npz")
    env = StickFigureA3(n_envs=2)
    rb = env.robot
    assert list(g["mirrored_obs"]) == list(rb.mirrored_obs) and list(g["mirrored_acts"]) == list(rb.mirrored_acts)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a.T, dtype=np.float32), device="cuda")
    so = Kn.make_mirror_spec(rb.mirrored_obs)
    sc = Kn.make_mirror_spec(rb.mirrored_obs, clock_inds=rb.clock_inds)
    sa = Kn.make_mirror_spec(rb.mirrored_acts)
    assert np.array_equal(Kn.mirror(so, t(g["obs"])).cpu().numpy().T, g["mirror_obs"])          # signed permutation: exact
    assert np.array_equal(Kn.mirror(sa, t(g["act"])).cpu().numpy().T, g["mirror_act"])
    assert_close(Kn.mirror(sc, t(g["obs"])).cpu().numpy().T, g["mirror_clock_obs"], "mirror_clock_observation")
    x = t(g["obs"])
    assert torch.equal(Kn.mirror(so, Kn.mirror(so, x)), x)                                          # an involution
    with pytest.raises(Exception):
        Kn.mirror(so, x, out=x)
