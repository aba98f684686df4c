from .trajectory import Trajectory, resample_table  # noqa: F401
from .reward import (CustomReward, NoReward, PosReward, RewardInterface, TargetVelocityReward)  # noqa: F401
from .checks import check_validity_task_mode_dataset  # noqa: F401
