"""GPU parity of the N1 row (action de-normalisation, JVRC target + PD law) against oracle/action.py."""
import numpy as np
import pytest

from conftest import assert_close

pytestmark = pytest.mark.gpu


def test_pd_torque_and_affine_action(a3_model, h1_model):
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from olympics_mujoco_b200.environments.stick_figure_a3 import StickFigureA3
    from oracle import action as OA
    rng = np.random.default_rng(0)
    for n in (1, 257, 5000):
        env = StickFigureA3(n_envs=n, seed=1)
        rb = env.robot
        a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        qpos = rng.normal(0, 0.5, (n, 25)).astype(np.float32)
        qvel = rng.normal(0, 2.0, (n, 24)).astype(np.float32)
        target = np.stack([OA.jvrc_target(a[e].astype(np.float64), rb.actuators, rb.motor_offset) for e in range(min(n, 64))])
        t32 = torch.as_tensor((a.astype(np.float64) + rb.motor_offset).astype(np.float32).T.copy(), device="cuda")
        ctrl = rb.pd_ctrl(t32, torch.as_tensor(qpos.T.copy(), device="cuda"), torch.as_tensor(qvel.T.copy(), device="cuda"))
        adr = [int(a3_model.jnt_qposadr[a3_model.jnt_names.index(j)]) for j in a3_model.actuator_joint]
        dadr = [int(a3_model.jnt_dofadr[a3_model.jnt_names.index(j)]) for j in a3_model.actuator_joint]
        for e in range(min(n, 64)):
            ref = OA.pd_ctrl(t32[:, e].cpu().numpy().astype(np.float64), qpos[e].astype(np.float64), qvel[e].astype(np.float64),
                             adr, dadr, rb.kp, rb.kd, np.ones(12))
            assert_close(ctrl[:, e].cpu().numpy(), ref, "pd ctrl", rtol=1e-5, atol=1e-4)      # torques are O(100)
            assert_close(t32[:, e].cpu().numpy(), target[e], "jvrc target")
    # add_offset inside the kernel == offset added by the caller
    sp = Kn.make_pd_spec(adr, dadr, rb.kp, rb.kd, np.ones(12), rb.motor_offset)
    raw = torch.as_tensor(a.T.copy(), device="cuda")
    c2 = Kn.pd_torque(sp, raw, torch.as_tensor(qpos.T.copy(), device="cuda"), torch.as_tensor(qvel.T.copy(), device="cuda"))
    assert torch.allclose(c2, ctrl, rtol=1e-5, atol=1e-3)
    # H1 de-normalisation through the env hook
    from olympics_mujoco_b200 import LocoEnvBase
    henv = LocoEnvBase.make("UnitreeH1.walk.real", n_envs=33, seed=0)
    nu = henv.info.action_space.shape[0]
    act = rng.uniform(-1, 1, (33, nu)).astype(np.float32)
    got = henv._preprocess_action(act).cpu().numpy()
    idx = [h1_model.actuator_names.index(x) for x in henv._action_spec]
    low, high = h1_model.actuator_ctrlrange[idx, 0], h1_model.actuator_ctrlrange[idx, 1]
    assert_close(got, OA.preprocess_action(act, low, high), "preprocess_action", rtol=1e-5, atol=1e-4)
