"""Oracle: the action side of ``step`` (N1, the row immediately before the hot path), float64 NumPy (TEST
INFRASTRUCTURE).

* ``preprocess_action``   ``LocoEnvBase._preprocess_action`` ``olympic_mujoco/environments/loco_env_base.py:1050-1069``
  with ``norm_act_mean / norm_act_delta`` of ``:165-175``
* ``jvrc_target``          ``JVRC.step`` ``olympic_mujoco/environments/robot.py:88-93`` (scatter to actuators + motor offset)
* ``step_pd``              ``MujocoRobotInterface.step_pd`` ``olympic_mujoco/interfaces/mujoco_robot_interface.py:425-443``
* ``pd_ctrl``              ``JVRC.do_simulation`` inner body ``robot.py:109-115`` (tau / gear)
"""
import numpy as np


def preprocess_action(action, low, high):
    mean, delta = (high + low) / 2.0, (high - low) / 2.0
    return np.asarray(action, np.float64) * delta + mean


def jvrc_target(action, actuators, motor_offset):
    filtered = np.zeros(len(motor_offset))
    for idx, act_id in enumerate(actuators):
        filtered[act_id] = action[idx]
    return filtered + np.asarray(motor_offset, np.float64)


def step_pd(p, v, curr_angles, curr_speeds, kp, kv):
    return kp * (p - curr_angles) + kv * (v - curr_speeds)


def pd_ctrl(target, qpos, qvel, qposadr, dofadr, kp, kd, gear):
    tau = step_pd(target, np.zeros(len(target)), qpos[qposadr], qvel[dofadr], kp, kd)
    return np.array([i / j for i, j in zip(tau, gear)])
