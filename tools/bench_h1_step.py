"""Timing of the single-step H1 API at BASELINE.json configs[4] scale: 1 048 576 envs (what one GPU of the 1M-env job
runs when N = 1; N = 8 gives 131 072 per GPU): trajectory next-sample (K3) + fused FK / observation / absorbing / reward
step (K1 + K2) per env step, launched per step like a live rollout.  One JSON line (measurement aid)."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
BYTES_PER_ENV_STEP = 1517


def measure(envs=1 << 20, steps=20, warmup=3):
    import bench
    from olympics_mujoco_b200 import kernels as Kn
    from olympics_mujoco_b200.build import H1_SPEC_JOINTS
    model, table = bench.build_table()
    n = envs
    dm = Kn.DeviceModel(model)
    perm_list = [int(model.jnt_qposadr[model.jnt_names.index(j)]) for j in H1_SPEC_JOINTS]   # UnitreeH1.py:303-355
    spec = Kn.make_h1_spec(perm_list, 15)                  # dq_pelvis_tx in the emitted 32-entry observation
    traj = Kn.DeviceTrajectory(table, n, seed=5)
    sample = traj.reset()
    qpos, qvel = torch.empty((17, n), device="cuda"), torch.empty((17, n), device="cuda")
    out = None

    def step():
        nonlocal out
        traj.next(sample=sample)                       # K3: index + gather (+ wrap reset)
        Kn.set_sim_state(dm, spec, sample, qpos, qvel)   # A7 (the physics would run here)
        out = Kn.h1_step(dm, spec, qpos, qvel, sample[17], want_fk=True, out=out)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    # A live rollout launches three small kernels per env step; below ~10^6 envs the Python / launch overhead of the
    # loop exceeds the kernels' run time, so the device-side cost is measured on CUDA-graph replays of `steps`
    # consecutive env steps (what a launch-bound inner loop should be captured into anyway), and the kernel alone on
    # a graph of `steps` h1_step launches.
    side = torch.cuda.Stream()
    g_step, g_kernel = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g_step, stream=side):
            for _ in range(steps):
                step()
        with torch.cuda.graph(g_kernel, stream=side):
            for _ in range(steps):
                out = Kn.h1_step(dm, spec, qpos, qvel, sample[17], want_fk=True, out=out)
    torch.cuda.synchronize()

    def replay(g, reps=5):
        g.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            g.replay()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / (reps * steps)

    ms, kms = replay(g_step), replay(g_kernel)
    # the FUSED live step (om_h1_live_step: next sample + scatter + FK + obs + flag + reward in one kernel, one C call with
    # pre-bound arguments): graph-replayed and launched eagerly from Python
    pxv = sample[17].clone()
    live = Kn.H1LiveStep(dm, spec, traj, pxv, out=dict(out, qpos=qpos, qvel=qvel))
    for _ in range(warmup):
        live()
    g_live = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        live.rebind(stream=side.cuda_stream)
        live()
        side.synchronize()
        with torch.cuda.graph(g_live, stream=side):
            for _ in range(steps):
                live()
    torch.cuda.synchronize()
    live.rebind()
    live_graph_ms = replay(g_live)
    live_eager = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(4 * steps):
            live()
        t1.record()
        torch.cuda.synchronize()
        live_eager.append(t0.elapsed_time(t1) / (4 * steps))
    live_eager_ms = min(live_eager)
    # the same loop launched eagerly from Python, for the record
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    eager_ms = t0.elapsed_time(t1) / steps
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    kernel_bytes = 136 + 1104 + 128 + 4 + 1 + 4                     # qpos/qvel in, FK + obs + reward + flag out, prev x-vel
    # live_step_frac: the fused live step on SURVEY 8(d)'s 1517 B/env-step (it additionally writes the qpos / qvel mirrors)
    ach = kernel_bytes * n / (kms * 1e-3) / 1e9
    return {"workload": f"UnitreeH1 single env step, {n} envs (configs[4] per-GPU shard at N=1)", "value": n / (ms * 1e-3),
            "unit": "env-steps/s", "ms_per_step": ms, "h1_step_kernel_ms": kms, "eager_python_loop_ms_per_step": eager_ms,
            "live_step_ms": live_graph_ms, "live_step_eager_ms": live_eager_ms,
            "live_step_frac": BYTES_PER_ENV_STEP * n / (live_graph_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "timing": "CUDA-graph replay of `steps` consecutive env steps",
            "roofline": {"bound": "hbm", "kernel": "h1_step_kernel<WRITE_FK> (h1_step_split_kernel up to 32768 envs)", "achieved": ach, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "bytes_per_env_step": kernel_bytes}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(measure(a.envs, a.steps)))
