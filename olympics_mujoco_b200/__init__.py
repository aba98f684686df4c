"""B200-native batched env-step hot path for olympics-mujoco (UnitreeH1 / StickFigureA3).

Host API mirrors the reference (``LocoEnvBase.make("UnitreeH1.walk.real")``, ``reset`` / ``step`` /
``play_trajectory_from_velocity``, ``ObservationHelper``, ``Trajectory``); all per-step arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/om_b200.h``.  There is no CPU fallback.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # heavy (torch-importing) modules load lazily so that host-only tools (mjcf, codegen, build) stay light
    if name in ("LocoEnvBase", "UnitreeH1", "StickFigureA3", "ValidTaskConf"):
        from . import environments
        return getattr(environments, name)
    if name in ("ObservationHelper", "ObservationType"):
        from . import observation_helper
        return getattr(observation_helper, name)
    if name == "Trajectory":
        from .utils import Trajectory
        return Trajectory
    raise AttributeError(name)
