"""CPU: analytic known-answer checks that pin the kinematics oracle (the reference has no FK fixtures --
its FK lives in the MuJoCo C engine, which cannot be installed here; see oracle/kinematics.py)."""
import numpy as np

from conftest import a3_random_states, assert_close
from oracle import kinematics as K
from oracle import tf3


def test_model_tables_match_survey(h1_model, a3_model):
    assert (h1_model.nbody, h1_model.njnt, h1_model.nq, h1_model.nv, h1_model.nsite) == (21, 17, 17, 17, 1)
    assert (a3_model.nbody, a3_model.njnt, a3_model.nq, a3_model.nv, a3_model.nsite) == (17, 19, 25, 24, 2)
    assert abs(h1_model.total_mass - 51.437) < 1e-9
    assert abs(a3_model.total_mass - 40.8214) < 1e-3
    assert h1_model.jnt_names[:6] == ["pelvis_tx", "pelvis_tz", "pelvis_ty", "pelvis_tilt", "pelvis_list", "pelvis_rotation"]
    # re-oriented arm quaternions are normalised by the compiler (UnitreeH1.py:281-288)
    q = h1_model.body_quat[h1_model.body_id("left_shoulder_pitch_link")]
    assert_close(q, np.array([1, 0.25, 0.1, 0]) / np.linalg.norm([1, 0.25, 0.1, 0]), "arm quat", rtol=1e-12, atol=1e-12)
    assert a3_model.site_names == ["rf_force", "lf_force"] and h1_model.site_names == ["imu"]


def test_zero_pose_positions_are_sums_of_offsets(h1_model):
    out = K.forward(h1_model, h1_model.qpos0[None], np.zeros((1, 17)))
    ankle = out["xpos"][0, h1_model.body_id("left_ankle_link")]
    assert_close(ankle, [0.039468, 0.0875 + 0.11536, 1.045 - 0.1742 - 0.4 - 0.4], "left ankle", rtol=1e-12, atol=1e-12)
    assert_close(out["xquat"][0, :13], np.tile([1.0, 0, 0, 0], (13, 1)), "identity quats", rtol=0, atol=1e-15)
    assert_close(out["site_xpos"][0, 0], np.array([0, 0, 1.045]) + [-0.04452, -0.01891, 0.27756], "imu", rtol=1e-12, atol=1e-12)


def test_single_joint_rotations(h1_model):
    q = h1_model.qpos0.copy()
    q[h1_model.jnt_qposadr[h1_model.joint_id("knee_angle_l")]] = 0.5          # hinge about local y
    out = K.forward(h1_model, q[None], np.zeros((1, 17)))
    knee = out["xpos"][0, h1_model.body_id("left_knee_link")]
    ankle = out["xpos"][0, h1_model.body_id("left_ankle_link")]
    # rotation of (0,0,-0.4) by +0.5 rad about y: (x, z) = (-0.4 sin, -0.4 cos)
    assert_close(ankle - knee, [-0.4 * np.sin(0.5), 0.0, -0.4 * np.cos(0.5)], "shank", rtol=1e-12, atol=1e-12)
    # pelvis_tilt has axis (0,-1,0): positive tilt pitches the other way
    q = h1_model.qpos0.copy()
    q[3] = 0.3
    out = K.forward(h1_model, q[None], np.zeros((1, 17)))
    assert_close(out["xquat"][0, 1], [np.cos(0.15), 0, -np.sin(0.15), 0], "tilt quat", rtol=1e-12, atol=1e-12)


def test_off_centre_joint_correction(a3_model):
    """A3 knee has pos (0,0,0.02): rotating it keeps the anchor fixed, not the body origin."""
    q = a3_model.qpos0.copy()
    q[a3_model.jnt_qposadr[a3_model.joint_id("right_knee")]] = 1.0
    a = K.forward(a3_model, a3_model.qpos0[None], np.zeros((1, 24)))
    b = K.forward(a3_model, q[None], np.zeros((1, 24)))
    j = a3_model.joint_id("right_knee")
    assert_close(b["xanchor"][0, j], a["xanchor"][0, j], "anchor fixed", rtol=1e-12, atol=1e-12)
    shin = a3_model.body_id("right_shin")
    d = b["xpos"][0, shin] - b["xanchor"][0, j]
    assert_close(np.linalg.norm(d), 0.02, "origin stays 0.02 from the anchor", rtol=1e-12, atol=1e-12)
    assert np.linalg.norm(b["xpos"][0, shin] - a["xpos"][0, shin]) > 0.015


def test_root_translation_velocity_and_com(h1_model, h1_states):
    qpos, _ = h1_states
    qvel = np.zeros_like(qpos)
    qvel[:, 0], qvel[:, 1], qvel[:, 2] = 0.7, -0.2, 0.1
    out = K.forward(h1_model, qpos, qvel)
    assert_close(out["cvel"][:, 1:, :3], 0 * out["cvel"][:, 1:, :3], "omega", rtol=0, atol=1e-15)
    assert_close(out["cvel"][:, 1:, 3:], np.broadcast_to([0.7, -0.2, 0.1], out["cvel"][:, 1:, 3:].shape), "v", rtol=1e-12, atol=1e-12)
    # subtree_com[1] is the mass-weighted mean of xipos
    com = (out["xipos"] * h1_model.body_mass[None, :, None]).sum(1) / h1_model.total_mass
    assert_close(out["subtree_com"][:, 1], com, "com", rtol=1e-12, atol=1e-12)
    assert_close(np.linalg.norm(out["xquat"], axis=-1), np.ones(out["xquat"].shape[:2]), "unit quats", rtol=0, atol=1e-12)


def _fd_check(model, q, v, q_minus, q_plus, eps):
    """Central differences: (x(q+) - x(q-)) / 2eps equals the velocity the engine reports at q."""
    a = K.forward(model, q, v)
    m, p = K.forward(model, q_minus, v), K.forward(model, q_plus, v)
    fd = (p["xpos"] - m["xpos"]) / (2 * eps)
    pv = np.stack([K.mj_objectVelocity_xbody(model, a["xpos"], a["subtree_com"], a["cvel"], i)[:, 3:]
                   for i in range(model.nbody)], axis=1)
    assert_close(fd, pv, "finite-difference point velocity", rtol=1e-6, atol=1e-6)
    # angular part: R(t+eps) R(t-eps)^T ~ I + 2 eps [w]x
    dR = np.einsum("nbij,nbkj->nbik", p["xmat"], m["xmat"])
    w_fd = np.stack([dR[..., 2, 1] - dR[..., 1, 2], dR[..., 0, 2] - dR[..., 2, 0], dR[..., 1, 0] - dR[..., 0, 1]], -1) / (4 * eps)
    assert_close(w_fd, a["cvel"][..., :3], "finite-difference angular velocity", rtol=1e-6, atol=1e-6)


def test_velocities_are_time_derivatives_h1(h1_model, h1_states):
    qpos, qvel = h1_states
    eps = 1e-5
    _fd_check(h1_model, qpos[:64], qvel[:64], qpos[:64] - eps * qvel[:64], qpos[:64] + eps * qvel[:64], eps)


def test_velocities_are_time_derivatives_a3(a3_model):
    q, v = a3_random_states(a3_model, 64, seed=2)
    q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
    eps = 1e-5

    def advance(h):
        q2 = q.copy()
        q2[:, :3] += h * v[:, :3]
        half = 0.5 * h * v[:, 3:6]                          # body-frame angular velocity: q <- q * exp(h w / 2)
        ang = np.linalg.norm(half, axis=1, keepdims=True)
        dq = np.concatenate([np.cos(ang), half * np.sinc(ang / np.pi)], axis=1)
        q2[:, 3:7] = K.quat_mul(q[:, 3:7], dq)
        q2[:, 7:] += h * v[:, 6:]
        return q2

    _fd_check(a3_model, q, v, advance(-eps), advance(eps), eps)


def test_free_joint_quaternion_is_normalised(a3_model):
    q, v = a3_random_states(a3_model, 16, seed=3)
    a = K.forward(a3_model, q, v)
    qn = q.copy()
    qn[:, 3:7] /= np.linalg.norm(qn[:, 3:7], axis=1, keepdims=True)
    b = K.forward(a3_model, qn, v)
    assert_close(a["xpos"], b["xpos"], "xpos", rtol=1e-12, atol=1e-12)
    assert_close(np.linalg.norm(a["xquat"], axis=-1), np.ones((16, 17)), "unit", rtol=0, atol=1e-12)


def test_tf3_round_trips():
    rng = np.random.default_rng(0)
    ang = rng.uniform(-1.4, 1.4, (200, 3))
    q = tf3.euler2quat(ang[:, 0], ang[:, 1], ang[:, 2])
    assert_close(np.linalg.norm(q, axis=1), np.ones(200), "unit", rtol=0, atol=1e-12)
    back = np.stack(tf3.quat2euler(q), axis=1)
    assert_close(back, ang, "euler round trip", rtol=1e-10, atol=1e-10)
    # quat2mat agrees with MuJoCo's quat2mat on unit quaternions and tolerates scaling
    assert_close(tf3.quat2mat(q), K.quat2mat(q), "quat2mat", rtol=1e-12, atol=1e-12)
    assert_close(tf3.quat2mat(3.0 * q), K.quat2mat(q), "scaled", rtol=1e-12, atol=1e-12)
    assert_close(tf3.rotz(ang[:, 2]), tf3.quat2mat(tf3.euler2quat(0 * ang[:, 0], 0 * ang[:, 0], ang[:, 2])), "rotz", rtol=1e-12, atol=1e-12)
    for i in range(10):
        m = tf3.quat2mat(q[i])
        q2 = tf3.mat2quat(m)
        assert_close(q2 * np.sign(q2[0]) , q[i] * np.sign(q[i][0]), "mat2quat", rtol=1e-9, atol=1e-9)
