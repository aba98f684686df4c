"""Per-GPU shard of configs[4] at N = 8 (131072 envs), one 64-step rollout as `calls` playback calls: for ncu launch lists."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from olympics_mujoco_b200 import kernels as Kn  # noqa: E402
from olympics_mujoco_b200.environments import LocoEnvBase  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
t_call = 64 // calls
model, table = bench.build_table()
env = LocoEnvBase.make("UnitreeH1.walk.real", n_envs=n, traj_params=dict(table=table), seed=1234)
roll = env.make_rollout_buffers(t_call)
mom = torch.zeros(65, dtype=torch.float64, device="cuda")
for it in range(4):
    mom.zero_()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for c in range(calls):
        env.play_trajectory_from_velocity(1, t_call, render=False, out=roll, continue_episode=c > 0, obs_moments=mom)
    Kn.moment_stats(mom, "ppo_obs")
    t1.record()
    torch.cuda.synchronize()
    print(f"rollout {it}: {t0.elapsed_time(t1):.3f} ms ({n * 64 * 1517 / (t0.elapsed_time(t1) * 1e-3) / 1e9 / 6454.6:.3f} of HBM)")
