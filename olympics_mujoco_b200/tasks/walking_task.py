"""Batched ``WalkingTask`` (reference ``olympic_mujoco/tasks/walking_task.py``): the constants of the footstep
task and typed views onto its per-env device state.  All arithmetic (``step`` :246-293, ``calc_reward`` :74-110,
``done`` :298-319, ``reset`` :321-397) runs inside the fused CUDA kernels of ``csrc/om_a3.cu``; this class only
names things the way the reference does."""
from __future__ import annotations

import numpy as np

from .phase_clock import phase_clock_lut

REWARD_KEYS = ("foot_frc_score", "foot_vel_score", "orient_cost", "height_error", "step_reward", "upper_body_reward")
STANDING, FORWARD = 0, 1


class WalkingTask:
    def __init__(self, dt=0.025, root_body="torso", lfoot_body="left_foot", rfoot_body="right_foot", head_body="head",
                 goal_height_ref=0.80, total_duration=1.1, swing_duration=0.75, stance_duration=0.35):
        self._control_dt = dt
        self._root_body_name, self._lfoot_body_name = root_body, lfoot_body
        self._rfoot_body_name, self._head_body_name = rfoot_body, head_body
        self._goal_speed_ref = 0.0                                     # walking_task.py:30
        self._goal_height_ref = goal_height_ref                        # StickFigureA3.py:110-113
        self._total_duration, self._swing_duration, self._stance_duration = total_duration, swing_duration, stance_duration
        self.target_radius = 0.20                                      # :333
        self.delay_frames = int(np.floor(swing_duration / dt))         # :336
        self._period = int(np.floor(2 * total_duration * (1 / dt)))    # :351
        self.clock_lut = phase_clock_lut(swing_duration, stance_duration, 0.1, "grounded", 1 / dt, period=self._period)
        self._dev = None                                               # kernels.A3Task, bound by the env

    def bind(self, dev):
        self._dev = dev

    # ---- per-env state, [n] int32 / bool views of the device rows
    @property
    def _phase(self):
        return self._dev.ints[0]

    @property
    def t1(self):
        return self._dev.ints[1]

    @property
    def t2(self):
        return self._dev.ints[2]

    @property
    def target_reached_frames(self):
        return self._dev.ints[3]

    @property
    def mode(self):
        return self._dev.ints[4]

    @property
    def target_reached(self):
        return self._dev.ints[6] != 0

    @property
    def sequence(self):
        """[n, 20, 4] footstep plans (rows beyond len(sequence) are zero); ``sequence_len`` [n]."""
        return self._dev.sequence.view(20, 4, -1).permute(2, 0, 1)

    @property
    def sequence_len(self):
        return self._dev.ints[5]
