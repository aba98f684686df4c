"""IL-mode reward interface (reference ``olympic_mujoco/utils/reward.py:5-74``), batched: ``state`` is a
device tensor [n, D] (or [D]) and the result a device tensor [n] (or scalar tensor).  The fused kernels
(``om_h1_step`` / ``om_h1_play_from_velocity``) compute ``TargetVelocityReward`` in-kernel; these classes serve
``env.reward(state, action, next_state, absorbing)`` calls made outside ``step``."""
import torch


class RewardInterface:
    def __call__(self, state, action, next_state, absorbing):
        raise NotImplementedError

    def reset_state(self):
        pass


class NoReward(RewardInterface):
    def __call__(self, state, action, next_state, absorbing):
        return 0


class PosReward(RewardInterface):
    def __init__(self, pos_idx):
        self._pos_idx = pos_idx

    def __call__(self, state, action, next_state, absorbing):
        return state[..., self._pos_idx]


class CustomReward(RewardInterface):
    def __init__(self, reward_callback=None):
        self._reward_callback = reward_callback

    def __call__(self, state, action, next_state, absorbing):
        if self._reward_callback is not None:
            return self._reward_callback(state, action, next_state)
        return 0


class TargetVelocityReward(RewardInterface):
    def __init__(self, target_velocity, x_vel_idx):
        self._target_vel = target_velocity
        self._x_vel_idx = x_vel_idx

    def __call__(self, state, action, next_state, absorbing):
        x_vel = torch.as_tensor(state)[..., self._x_vel_idx]
        return torch.exp(-torch.square(x_vel - self._target_vel))
