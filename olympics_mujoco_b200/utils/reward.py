"""IL-mode reward callables (the interface of reference ``olympic_mujoco/utils/reward.py:5-74``: called as
``reward(state, action, next_state, absorbing)``, optional ``reset_state()``), batched: ``state`` is a device tensor
[n, D] (or [D]) and the result a device tensor [n] (or a scalar).  The fused kernels (``om_h1_step``, ``om_h1_live_step``,
``om_h1_play_from_velocity``) evaluate ``TargetVelocityReward`` in-kernel; these objects serve ``env.reward(...)`` calls
made outside ``step`` and carry the parameters the kernels read (``_target_vel``, ``_x_vel_idx``)."""
import torch


class RewardInterface:
    """Base of every reward: stateless unless a subclass says otherwise."""

    def __call__(self, state, action, next_state, absorbing):
        raise NotImplementedError

    def reset_state(self):
        return None


class _Column(RewardInterface):
    """A reward that is a function of ONE observation column of ``state``."""

    def __init__(self, idx):
        self._idx = int(idx)

    def _column(self, state):
        return torch.as_tensor(state)[..., self._idx]


class NoReward(RewardInterface):
    def __call__(self, state, action, next_state, absorbing):
        return 0


class PosReward(_Column):
    """The root x position itself (reward_type "x_pos", loco_env_base.py:809-815)."""

    def __init__(self, pos_idx):
        super().__init__(pos_idx)
        self._pos_idx = self._idx

    def __call__(self, state, action, next_state, absorbing):
        return self._column(state)


class CustomReward(RewardInterface):
    """``reward_callback(state, action, next_state)``; 0 without one."""

    def __init__(self, reward_callback=None):
        self._reward_callback = reward_callback

    def __call__(self, state, action, next_state, absorbing):
        cb = self._reward_callback
        return 0 if cb is None else cb(state, action, next_state)


class TargetVelocityReward(_Column):
    """exp(-(v_x - v_target)^2) on the root x velocity (reward.py:66-74; target 1.25 walk / 2.5 run)."""

    def __init__(self, target_velocity, x_vel_idx):
        super().__init__(x_vel_idx)
        self._target_vel, self._x_vel_idx = target_velocity, self._idx

    def __call__(self, state, action, next_state, absorbing):
        err = self._column(state) - self._target_vel
        return torch.exp(-(err * err))
