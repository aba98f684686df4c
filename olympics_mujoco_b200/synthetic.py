"""Synthetic, dataset-shaped inputs (the reference's datasets are not available offline).

The reference loads ``datasets/humanoids/real/02-constspeed_UnitreeH1.npz`` (reference
``real_humanoid_robots/UnitreeH1.py:225``): one 1-D float64 array per observation key, in
observation-spec order, sampled at 500 Hz, plus optional ``split_points``
(format: ``examples/random_npz.py:37-46``, ``utils/trajectory.py:61-96``).  These generators produce
files of exactly that shape: bounded, low-pass-filtered random walks inside the joint ranges, a root
that advances at the walk target speed, velocities by finite differences -- and non-terminal for
``_has_fallen`` as the reference requires of a dataset (``loco_env_base.py:949-957``).
"""
import numpy as np

from . import mjcf

H1_SPEC_JOINTS = ["pelvis_tx", "pelvis_tz", "pelvis_ty", "pelvis_tilt", "pelvis_list", "pelvis_rotation",
                  "back_bkz", "l_arm_shy", "l_arm_shx", "l_arm_shz", "left_elbow", "r_arm_shy", "r_arm_shx",
                  "r_arm_shz", "right_elbow", "hip_flexion_r", "hip_adduction_r", "hip_rotation_r",
                  "knee_angle_r", "ankle_angle_r", "hip_flexion_l", "hip_adduction_l", "hip_rotation_l",
                  "knee_angle_l", "ankle_angle_l"]

# root channel ranges: inside the _has_fallen thresholds (UnitreeH1.py:176-179) with margin
_H1_ROOT = {"pelvis_tz": (-0.05, 0.05), "pelvis_ty": (-0.06, 0.01), "pelvis_tilt": (-0.25, 0.15),
            "pelvis_list": (-0.15, 0.2), "pelvis_rotation": (-0.25, 0.25)}


def _smooth_walk(rng, n, lo, hi, sigma=0.02, taps=25):
    x = np.empty(n)
    x[0] = rng.uniform(lo, hi)
    steps = rng.normal(0.0, sigma * (hi - lo), n)
    for t in range(1, n):
        v = x[t - 1] + steps[t]
        if v > hi:                      # reflect at the bounds
            v = 2 * hi - v
        if v < lo:
            v = 2 * lo - v
        x[t] = min(max(v, lo), hi)
    k = np.hanning(taps)
    k /= k.sum()
    pad = np.concatenate([np.full(taps // 2, x[0]), x, np.full(taps // 2, x[-1])])
    return np.convolve(pad, k, mode="valid")[:n]


def h1_walk_dataset(n_traj=4, t_raw=2500, freq=500.0, seed=0, model=None, speed=1.25):
    """Dict with the 34 H1 keys (arms removed) + ``split_points``; each array has n_traj*t_raw samples."""
    model = model or mjcf.load_builtin("unitree_h1")
    rng = np.random.default_rng(seed)
    joints = [j for j in H1_SPEC_JOINTS if j in model.jnt_names]
    q = {}
    for j in joints:
        segs = []
        for _ in range(n_traj):
            if j == "pelvis_tx":
                x0 = rng.uniform(-1.0, 1.0)
                seg = x0 + speed * np.arange(t_raw) / freq + _smooth_walk(rng, t_raw, -0.02, 0.02)
            elif j in _H1_ROOT:
                seg = _smooth_walk(rng, t_raw, *_H1_ROOT[j])
            else:
                lo, hi = model.jnt_range[model.jnt_names.index(j)]
                mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * 0.95
                seg = _smooth_walk(rng, t_raw, mid - half, mid + half)
            segs.append(seg)
        q[j] = np.concatenate(segs)
    data = {}
    for j in joints:
        data["q_" + j] = q[j]
    for j in joints:
        dq = np.empty_like(q[j])
        for s in range(n_traj):
            seg = q[j][s * t_raw:(s + 1) * t_raw]
            dq[s * t_raw:(s + 1) * t_raw] = np.gradient(seg) * freq
        data["dq_" + j] = dq
    data["split_points"] = np.arange(n_traj + 1) * t_raw
    return data


def a3_rollout_inputs(n_env, horizon, seed=0, model=None):
    """Config-3 shaped inputs: per-step A3 ``qpos[T,N,25]`` / ``qvel[T,N,24]`` random walks around the
    nominal pose within the joint ranges, contact summaries, critic values."""
    model = model or mjcf.load_builtin("stick_figure_a3")
    rng = np.random.default_rng(seed)
    nominal = np.array([0, 0, 1.34, 1, 0, 0, 0] + [d * np.pi / 180 for d in
                       [-30, 0, 0, 50, 0, -24, -30, 0, 0, 50, 0, -24, -3, -9.74, -30, -3, 9.74, -30]])
    lo = np.full(model.nq, -np.inf)
    hi = np.full(model.nq, np.inf)
    for j in range(model.njnt):
        if model.jnt_type[j] == mjcf.JNT_HINGE and model.jnt_limited[j]:
            lo[model.jnt_qposadr[j]], hi[model.jnt_qposadr[j]] = model.jnt_range[j]
    qpos = np.empty((horizon, n_env, model.nq))
    cur = nominal + rng.uniform(-0.02, 0.02, (n_env, model.nq))
    cur[:, 0:2] = rng.uniform(-1, 1, (n_env, 2))
    for t in range(horizon):
        cur = cur + rng.normal(0, 0.01, cur.shape)
        cur[:, 7:] = np.clip(cur[:, 7:], lo[7:], hi[7:])
        cur[:, 2] = np.clip(cur[:, 2], 1.0, 1.5)
        cur[:, 3:7] /= np.linalg.norm(cur[:, 3:7], axis=1, keepdims=True)
        qpos[t] = cur
    qvel = np.clip(rng.normal(0, 1, (horizon, n_env, model.nv)), -10, 10)
    fmax = model.total_mass * 9.8 * 0.5
    contact = dict(l_grf=rng.uniform(0, 2 * fmax, (horizon, n_env)),
                   r_grf=rng.uniform(0, 2 * fmax, (horizon, n_env)),
                   min_z=rng.uniform(-0.01, 0.01, (horizon, n_env)),
                   foot_contact=rng.random((horizon, n_env)) < 0.9,
                   bad_collision=rng.random((horizon, n_env)) < 0.01)
    values = rng.normal(0, 1, (horizon + 1, n_env))
    return dict(qpos=qpos, qvel=qvel, contact=contact, values=values)
