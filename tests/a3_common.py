"""Shared helpers of the A3 parity tests (host harness on CPU, CUDA kernels on the GPU)."""
import numpy as np

from conftest import GOLDEN
from oracle import a3 as OA


def golden():
    return np.load(GOLDEN / "a3_task_ref.npz")


def contact4(c5):
    """[..., 5] (l_grf, r_grf, min_z, foot_contact, bad_collision) -> the C ABI's [..., 4] record."""
    out = np.empty(c5.shape[:-1] + (4,), np.float32)
    out[..., :3] = c5[..., :3]
    out[..., 3] = c5[..., 3] + 2 * c5[..., 4]
    return out


def lut6(period=OA.PERIOD):
    """What om_a3_task_create builds: the four phase clocks + sin/cos of the phase, fp32."""
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    ph = np.arange(period)
    return np.concatenate([phase_clock_lut(period=period), np.sin(2 * np.pi * ph / period)[:, None],
                           np.cos(2 * np.pi * ph / period)[:, None]], axis=1).astype(np.float32)


def oracle_obs(model, gold, e, upto=None):
    """Float64 oracle rollout of env e of the fixture -> (obs [T,41], total [T]); also re-checks the ints."""
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    lut = phase_clock_lut()
    _, _, ts, obs0 = OA.reset(model, int(gold["seed"]), e, 0, iteration_count=float(gold["iteration_count"]))
    T = gold["step_done"].shape[1] if upto is None else upto
    obs, total = np.empty((T, 41)), np.empty(T)
    for t in range(T):
        c = gold["step_contact"][e, t].astype(np.float64)
        con = OA.Contact(l_grf=c[0], r_grf=c[1], min_z=c[2], foot_contact=bool(c[3]), bad_collision=bool(c[4]))
        obs[t], total[t], _, _ = OA.step_tail(model, gold["step_qpos"][e, t].astype(np.float64),
                                              gold["step_qvel"][e, t].astype(np.float64), ts, con, lut)
    return obs0, obs, total
