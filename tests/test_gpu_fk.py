"""GPU parity: K1 (om_fk) through the C ABI vs the float64 oracle (oracle/kinematics.py)."""
import numpy as np
import pytest

from conftest import a3_random_states, assert_close

pytestmark = pytest.mark.gpu


def _run(model, qpos, qvel, force_generic):
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    dm = Kn.DeviceModel(model)
    out = Kn.fk(dm, Kn.to_soa(qpos), Kn.to_soa(qvel), force_generic=force_generic)
    torch.cuda.synchronize()
    return dm, {k: v.cpu().numpy() for k, v in out.items()}


def _check(model, qpos, qvel, out):
    from oracle import kinematics as K
    ref = K.forward(model, qpos, qvel)
    n = qpos.shape[0]
    assert_close(out["xpos"].T.reshape(n, model.nbody, 3), ref["xpos"], "xpos")
    assert_close(out["xquat"].T.reshape(n, model.nbody, 4), ref["xquat"], "xquat")
    assert_close(out["site_xpos"].T.reshape(n, model.nsite, 3), ref["site_xpos"], "site_xpos")
    assert_close(out["site_xmat"].T.reshape(n, model.nsite, 3, 3), ref["site_xmat"], "site_xmat")
    assert_close(out["cvel"].T.reshape(n, model.nbody, 6), ref["cvel"], "cvel")
    assert_close(out["subtree_com"].T, ref["subtree_com"][:, 1], "subtree_com")


@pytest.mark.parametrize("force_generic", [False, True])
def test_fk_h1_dataset_states(h1_model, h1_states, force_generic):
    qpos, qvel = h1_states
    dm, out = _run(h1_model, qpos, qvel, force_generic)
    assert dm.specialised
    _check(h1_model, qpos, qvel, out)


@pytest.mark.parametrize("force_generic", [False, True])
def test_fk_a3_random_states(a3_model, force_generic):
    qpos, qvel = a3_random_states(a3_model, 777, seed=1)     # ragged size: not a multiple of the block
    dm, out = _run(a3_model, qpos, qvel, force_generic)
    assert dm.specialised
    _check(a3_model, qpos, qvel, out)


def test_fk_generic_model_with_arms():
    """A model without a generated kernel (H1 with its 8 arm joints) runs the table-driven kernel."""
    from olympics_mujoco_b200 import mjcf
    model = mjcf.load_builtin("unitree_h1_arms")
    rng = np.random.default_rng(5)
    qpos = (rng.normal(0, 0.4, (300, model.nq))).astype(np.float32).astype(np.float64)
    qvel = (rng.normal(0, 1.5, (300, model.nv))).astype(np.float32).astype(np.float64)
    dm, out = _run(model, qpos, qvel, False)
    assert not dm.specialised
    _check(model, qpos, qvel, out)


def test_fk_edge_sizes(h1_model, h1_states):
    """n = 1 and n = 0 (empty batch)."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    qpos, qvel = h1_states
    dm, out = _run(h1_model, qpos[:1], qvel[:1], False)
    _check(h1_model, qpos[:1], qvel[:1], out)
    empty = Kn.fk(dm, torch.empty((17, 0), device="cuda"), torch.empty((17, 0), device="cuda"))
    assert empty["xpos"].shape == (63, 0)


def test_fk_properties_large(h1_model):
    """Size-independent properties at bench scale (2^20 envs): unit quaternions, pure root translation
    moves every body by the same amount and gives cvel = [0; v] for every body."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    n = 1 << 20
    dm = Kn.DeviceModel(h1_model)
    g = torch.Generator(device="cuda").manual_seed(0)
    q = (torch.rand((17, n), device="cuda", generator=g) - 0.5) * 0.8
    qd = torch.zeros((17, n), device="cuda")
    qd[0], qd[1], qd[2] = 0.7, -0.2, 0.1            # slides tx, tz, ty
    a = Kn.fk(dm, q, qd)
    qn = a["xquat"].view(21, 4, n).norm(dim=1)
    assert float((qn - 1).abs().max()) < 1e-5
    shift = torch.tensor([0.5, -0.25, 0.125], device="cuda")
    q2 = q.clone()
    q2[0] += shift[0]; q2[1] += shift[1]; q2[2] += shift[2]
    b = Kn.fk(dm, q2, qd)
    # joint axes: tx=(1,0,0), tz=(0,1,0), ty=(0,0,1) -> world shift (0.5, -0.25, 0.125)
    d = (b["xpos"] - a["xpos"]).view(21, 3, n)[1:]
    assert float((d - shift.view(1, 3, 1)).abs().max()) < 2e-5
    cv = a["cvel"].view(21, 6, n)[1:]
    assert float(cv[:, :3].abs().max()) == 0.0
    ref = torch.tensor([0.7, -0.2, 0.1], device="cuda").view(1, 3, 1)
    assert float((cv[:, 3:] - ref).abs().max()) < 1e-6


@pytest.mark.parametrize("name,builtin,force_generic", [("h1", "unitree_h1", False), ("h1", "unitree_h1", True),
                                                        ("h1_arms", "unitree_h1_arms", False),
                                                        ("a3", "stick_figure_a3", False), ("a3", "stick_figure_a3", True)])
def test_fk_vs_independent_checker_fixture(name, builtin, force_generic):
    """The CUDA kernels against the INDEPENDENT checker's fixture (tools/fk_independent.py: own MJCF reader, homogeneous
    transforms, numerically differentiated velocities) -- not against the oracle the kernels were developed with."""
    from conftest import GOLDEN
    from olympics_mujoco_b200 import mjcf
    g = np.load(GOLDEN / "fk_independent_ref.npz")
    model = mjcf.load_builtin(builtin)
    q, v = g[name + "_qpos"], g[name + "_qvel"]
    n = q.shape[0]
    dm, out = _run(model, q, v, force_generic)
    assert_close(out["xpos"].T.reshape(n, model.nbody, 3), g[name + "_xpos"], "xpos")
    quat = out["xquat"].T.reshape(n, model.nbody, 4)
    ref_q = g[name + "_xquat"]
    s = np.sign(np.sum(quat * ref_q, axis=-1, keepdims=True))
    assert_close(quat, ref_q * np.where(s == 0, 1.0, s), "xquat (up to the sign of the quaternion)")
    assert_close(out["site_xpos"].T.reshape(n, model.nsite, 3), g[name + "_site_xpos"], "site_xpos")
    assert_close(out["site_xmat"].T.reshape(n, model.nsite, 3, 3), g[name + "_site_xmat"], "site_xmat")
    assert_close(out["cvel"].T.reshape(n, model.nbody, 6), g[name + "_cvel"], "cvel")
    assert_close(out["subtree_com"].T, g[name + "_subtree_com"][:, 1], "subtree_com")
