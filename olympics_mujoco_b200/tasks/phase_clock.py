"""Phase-clock look-up table for the WalkingTask reward (reference ``olympic_mujoco/tasks/rewards.py:270-366``
``create_phase_reward``): four PCHIP splines (right/left foot x force/velocity clock) over one gait cycle,
which the task only ever evaluates at INTEGER phases (``rewards.py:76-77,95-96``).  Built once on the host with
SciPy, exactly like the reference, then sampled into a ``[period, 4]`` table (columns r_frc, r_vel, l_frc,
l_vel) that lives in GPU constant memory."""
import numpy as np
from scipy.interpolate import PchipInterpolator


def phase_clock_lut(swing_duration=0.75, stance_duration=0.35, strict_relaxer=0.1, stance_mode="grounded", freq=40.0,
                    period=None):
    # the gait cycle is four windows: right swing, double stance, left swing, double stance
    edges = np.array([0.0, swing_duration, swing_duration + stance_duration, 2 * swing_duration + stance_duration,
                      2 * (swing_duration + stance_duration)]) * freq
    knots = np.empty(8)
    for w in range(4):
        off = (edges[w + 1] - edges[w]) * strict_relaxer
        knots[2 * w], knots[2 * w + 1] = edges[w] + off, edges[w + 1] - off
    last_off = (edges[4] - edges[3]) * strict_relaxer
    stance_frc = {"aerial": -1.0, "zero": 0.0}.get(stance_mode, 1.0)
    # per window: (right foot force, right foot velocity); the left foot swaps the two swing windows
    r_frc = np.repeat([-1.0, stance_frc, 1.0, stance_frc], 2)
    r_vel = np.repeat([1.0, -stance_frc, -1.0, -stance_frc], 2)
    l_frc = np.repeat([1.0, stance_frc, -1.0, stance_frc], 2)
    l_vel = np.repeat([-1.0, -stance_frc, 1.0, -stance_frc], 2)
    shift = knots[-1] + last_off
    x = np.concatenate([knots - shift, knots, knots + shift])        # one cycle before and after: periodic ends
    period = int(np.floor(edges[4])) if period is None else int(period)
    ph = np.arange(period)
    cols = [PchipInterpolator(x, np.tile(y, 3))(ph) for y in (r_frc, r_vel, l_frc, l_vel)]
    return np.stack(cols, axis=1)
