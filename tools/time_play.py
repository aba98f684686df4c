"""Per-iteration device time of the H1 playback call (diagnostic)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from olympics_mujoco_b200.environments import LocoEnvBase

model, table = bench.build_table()
n, T = 4096, 500
env = LocoEnvBase.make("UnitreeH1.walk.real", n_envs=n, traj_params=dict(table=table), seed=1234)
rolls = [env.make_rollout_buffers(T) for _ in range(2)]
for which in (0, 1, 0, 1):
    ts = []
    for i in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.play_trajectory_from_velocity(n_episodes=1, n_steps_per_episode=T, render=False, out=rolls[which])
        b.record()
        torch.cuda.synchronize()
        ts.append(round(a.elapsed_time(b), 3))
    print("buffer set", which, ts)
ts = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    env.play_trajectory_from_velocity(n_episodes=1, n_steps_per_episode=T, render=False, out=rolls[0])
    b.record()
    ts.append((a, b))
torch.cuda.synchronize()
print("back to back", [round(a.elapsed_time(b), 3) for a, b in ts])
