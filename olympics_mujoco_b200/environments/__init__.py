from .loco_env_base import LocoEnvBase, ValidTaskConf  # noqa: F401
from .unitree_h1 import BaseHumanoidRobot, UnitreeH1  # noqa: F401
from .stick_figure_a3 import StickFigureA3  # noqa: F401

UnitreeH1.register()
