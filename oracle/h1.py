"""Oracle: UnitreeH1 imitation-learning step path in float64 NumPy (TEST INFRASTRUCTURE).

Follows, statement by statement:
* observation spec (arms removed)      ``real_humanoid_robots/UnitreeH1.py:293-356`` + ``:70-84``
* ``ObservationHelper._build_obs``      mushroom_rl>=1.10 (not in tree): concatenation of
  ``data.joint(name).qpos / qvel`` in spec order
* ``_create_observation``               ``environments/loco_env_base.py:737-767`` (drop root x, y)
* ``_has_fallen``                       ``real_humanoid_robots/UnitreeH1.py:162-203``
* ``TargetVelocityReward``              ``utils/reward.py:66-74``; built ``loco_env_base.py:799-807``,
  target 1.25 ``base_robot/base_humanoid_robot.py:149-151``; called with the PREVIOUS observation
  (mushroom ``MuJoCo.step``: ``reward(self._obs, action, cur_obs, absorbing)``)
* ``set_sim_state``                     ``loco_env_base.py:659-684``
* ``play_trajectory_from_velocity``     ``loco_env_base.py:444-560``
"""
import numpy as np

from . import kinematics as K
from .trajectory import TrajectoryState

ARM_JOINTS = ["l_arm_shy", "l_arm_shx", "l_arm_shz", "left_elbow",
              "r_arm_shy", "r_arm_shx", "r_arm_shz", "right_elbow"]
_SPEC_JOINTS = ["pelvis_tx", "pelvis_tz", "pelvis_ty", "pelvis_tilt", "pelvis_list", "pelvis_rotation",
                "back_bkz"] + ARM_JOINTS + \
               ["hip_flexion_r", "hip_adduction_r", "hip_rotation_r", "knee_angle_r", "ankle_angle_r",
                "hip_flexion_l", "hip_adduction_l", "hip_rotation_l", "knee_angle_l", "ankle_angle_l"]

TARGET_VELOCITY_WALK = 1.25
DT = 0.01          # timestep 0.001 * n_intermediate_steps 1 * n_substeps 10 (loco_env_base.py:46,52)


def spec_joints(model):
    """Joint names in observation-spec order, restricted to joints the model still has."""
    return [j for j in _SPEC_JOINTS if j in model.jnt_names]


def keys(model):
    js = spec_joints(model)
    return ["q_" + j for j in js] + ["dq_" + j for j in js]


def perm(model):
    """P[k] = qpos/qvel address (MJCF order) of spec entry k."""
    return np.array([model.jnt_qposadr[model.jnt_names.index(j)] for j in spec_joints(model)])


def build_obs(model, qpos, qvel):
    """ObservationHelper._build_obs for the H1 spec: [q in spec order, dq in spec order]."""
    P = perm(model)
    return np.concatenate([qpos[..., P], qvel[..., P]], axis=-1)


def create_observation(obs_full):
    """loco_env_base.py:761-765."""
    return obs_full[..., 2:].copy()


def has_fallen(obs):
    """UnitreeH1.py:162-203; strict inequalities, float64 thresholds."""
    y, tilt, lst, rot = obs[..., 0], obs[..., 1], obs[..., 2], obs[..., 3]
    c_y = (y < -0.3) | (y > 0.1)
    c_t = (tilt < (-np.pi / 4.5)) | (tilt > (np.pi / 12))
    c_l = (lst < -np.pi / 12) | (lst > np.pi / 8)
    c_r = (rot < (-np.pi / 8)) | (rot > (np.pi / 8))
    return c_y | c_t | c_l | c_r


def x_vel_idx(model):
    return keys(model).index("dq_pelvis_tx") - 2


def target_velocity_reward(state, idx, target=TARGET_VELOCITY_WALK):
    """utils/reward.py:72-74."""
    return np.exp(-np.square(state[..., idx] - target))


def set_sim_state(model, sample):
    """loco_env_base.py:659-684: named scatter of a spec-ordered sample into qpos/qvel."""
    P = perm(model)
    n = len(P)
    qpos = np.zeros(sample.shape[:-1] + (model.nq,))
    qvel = np.zeros(sample.shape[:-1] + (model.nv,))
    qpos[..., P] = sample[..., :n]
    qvel[..., P] = sample[..., n:2 * n]
    return qpos, qvel


def step(model, qpos, qvel, prev_obs):
    """The IL ``step`` tail after physics: FK, obs, absorbing, reward (mushroom MuJoCo.step)."""
    fk = K.forward(model, qpos, qvel)
    obs = create_observation(build_obs(model, qpos, qvel))
    absorbing = has_fallen(obs)
    reward = target_velocity_reward(prev_obs, x_vel_idx(model))
    return dict(fk, obs=obs, absorbing=absorbing, reward=reward)


def play_trajectory_from_velocity(model, table, n_episodes, n_steps_per_episode, seed=0, env_id=0,
                                  record_fk=True, traj_state=None):
    """One env of loco_env_base.py:444-560 (render/record off), recording what each step computes.

    Returns dict of arrays indexed [episode*n_steps + step]: qpos, qvel (sim state handed to
    mj_forward), xpos, xquat, site_xpos, cvel (its outputs), obs/fallen (from the NEXT trajectory
    sample, :539-541), reward (TargetVelocityReward on the previous obs, the step() convention),
    traj_no / step_no / reset_count (integer state after the step).  ``traj_state``: the ``final["traj_state"]`` of a
    previous call -- a SECOND call on the same env object (it begins with its own reset(), :481).
    """
    nj = len(perm(model))
    xv = x_vel_idx(model)
    tr = TrajectoryState(table, seed=seed, env_id=env_id) if traj_state is None else traj_state
    rec = {k: [] for k in ("qpos", "qvel", "xpos", "xquat", "site_xpos", "cvel", "obs", "fallen",
                           "reward", "traj_no", "step_no", "reset_count")}
    tr.reset_trajectory()                                   # :481 reset()
    sample = tr.get_current_sample()                        # :483
    prev_obs = create_observation(sample)                   # reset(): self._obs (:603)
    curr_qpos = sample[:nj].copy()                          # :508
    for _ep in range(n_episodes):
        for _j in range(n_steps_per_episode):
            qvel_s = sample[nj:2 * nj]                      # :515
            qpos_s = curr_qpos + DT * qvel_s                # :517
            sample = sample.copy()
            sample[:nj] = qpos_s                            # :519
            qpos, qvel = set_sim_state(model, sample)       # :521
            rec["qpos"].append(qpos)
            rec["qvel"].append(qvel)
            if record_fk:
                fk = K.forward(model, qpos[None], qvel[None])   # :525 mj_forward (hot-path subset)
                for k in ("xpos", "xquat", "site_xpos", "cvel"):
                    rec[k].append(fk[k][0])
            curr_qpos = qpos_s.copy()                       # :529 (_get_joint_pos reads back qpos)
            sample = tr.get_next_sample()                   # :532
            if sample is None:                              # :534-537
                sample = tr.reset_trajectory()
                curr_qpos = sample[:nj].copy()
            obs = create_observation(sample)                # :539
            rec["obs"].append(obs)
            rec["fallen"].append(bool(has_fallen(obs)))     # :541
            rec["reward"].append(float(target_velocity_reward(prev_obs, xv)))
            prev_obs = obs
            rec["traj_no"].append(tr.traj_no)
            rec["step_no"].append(tr.step_no)
            rec["reset_count"].append(tr.reset_count)
        s = tr.reset_trajectory()                           # :555 reset()
        prev_obs = create_observation(s)
        curr_qpos = s[:nj].copy()                           # :557 (sample stays the stale one)
    out = {k: np.asarray(v) for k, v in rec.items() if len(v)}
    out["final"] = dict(traj_no=tr.traj_no, step_no=tr.step_no, reset_count=tr.reset_count,
                        curr_qpos=curr_qpos, pending_sample=sample, traj_state=tr)
    return out


def play_trajectory(model, table, n_episodes, n_steps_per_episode, seed=0, env_id=0, record_fk=True):
    """One env of ``LocoEnvBase.play_trajectory`` (loco_env_base.py:338-442, render/record off): every step forces the
    model to the current trajectory sample (:408), runs mj_forward (:410), fetches the next sample / resets at the end
    of a trajectory (:414-418) and checks has_fallen on it (:420-422); the episode ends with reset() (:432) while
    ``sample`` stays the stale one.  Same record layout as play_trajectory_from_velocity."""
    nj = len(perm(model))
    xv = x_vel_idx(model)
    tr = TrajectoryState(table, seed=seed, env_id=env_id)
    rec = {k: [] for k in ("qpos", "qvel", "xpos", "xquat", "site_xpos", "cvel", "obs", "fallen", "reward", "traj_no",
                           "step_no", "reset_count")}
    tr.reset_trajectory()                                   # :377 reset()
    sample = tr.get_current_sample()                        # :379
    prev_obs = create_observation(sample)
    for _ep in range(n_episodes):
        for _j in range(n_steps_per_episode):
            qpos, qvel = set_sim_state(model, sample)       # :408
            rec["qpos"].append(qpos)
            rec["qvel"].append(qvel)
            if record_fk:
                fk = K.forward(model, qpos[None], qvel[None])   # :410
                for k in ("xpos", "xquat", "site_xpos", "cvel"):
                    rec[k].append(fk[k][0])
            sample = tr.get_next_sample()                   # :414
            if sample is None:                              # :416-418
                tr.reset_trajectory()
                sample = tr.get_current_sample()
            obs = create_observation(sample)                # :420
            rec["obs"].append(obs)
            rec["fallen"].append(bool(has_fallen(obs)))
            rec["reward"].append(float(target_velocity_reward(prev_obs, xv)))
            prev_obs = obs
            rec["traj_no"].append(tr.traj_no)
            rec["step_no"].append(tr.step_no)
            rec["reset_count"].append(tr.reset_count)
        s_ = tr.reset_trajectory()                          # :432 (sample stays the stale one)
        prev_obs = create_observation(s_)
    out = {k: np.asarray(v) for k, v in rec.items() if len(v)}
    out["final"] = dict(traj_no=tr.traj_no, step_no=tr.step_no, reset_count=tr.reset_count, pending_sample=sample)
    return out
