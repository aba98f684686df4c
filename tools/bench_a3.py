"""Timing of BASELINE.json configs[2]: A3 PPO walk rollout (obs / reward / done + discounted returns),
16384 envs x 64-step horizon on one B200.  Prints one JSON line (a measurement aid; the driver-facing headline
bench is bench.py)."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

BYTES_PER_ENV_STEP = 490          # SURVEY.md 8(d) config 3 (FK fused, not materialised)


def measure(envs=16384, horizon=64, steps=20, warmup=3, per_step=False, fused_returns=True):
    args = argparse.Namespace(envs=envs, horizon=horizon, steps=steps, warmup=warmup, per_step=per_step)
    from olympics_mujoco_b200 import kernels as Kn
    from olympics_mujoco_b200 import mjcf
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    from olympics_mujoco_b200.environments.stick_figure_a3 import HALF_SITTING_DEG
    model = mjcf.load_builtin("stick_figure_a3")
    n, T = args.envs, args.horizon
    dm = Kn.DeviceModel(model)
    init = np.array([0, 0, 0.81, 1, 0, 0, 0] + [q * np.pi / 180 for q in HALF_SITTING_DEG])
    task = Kn.A3Task(dm, n, phase_clock_lut(), init, seed=0)
    qpos0, qvel0 = Kn.soa(25, n), Kn.soa(24, n)
    task.reset(qpos0, qvel0)
    g = torch.Generator(device="cuda").manual_seed(0)
    qpos = qpos0[None] + 0.02 * torch.randn((T, 25, n), device="cuda", generator=g).cumsum(0)
    qvel = torch.randn((T, 24, n), device="cuda", generator=g).clamp_(-10, 10)
    fmax = model.total_mass * 9.8 * 0.5
    con = torch.stack([torch.rand((T, n), device="cuda", generator=g) * 2 * fmax, torch.rand((T, n), device="cuda", generator=g) * 2 * fmax,
                       (torch.rand((T, n), device="cuda", generator=g) - 0.5) * 0.02,
                       (torch.rand((T, n), device="cuda", generator=g) < 0.7).float()
                       + 2 * (torch.rand((T, n), device="cuda", generator=g) < 0.01).float()], dim=1).contiguous()
    values = torch.randn((T + 1, n), device="cuda", generator=g)
    out = dict(obs=torch.empty((T, 41, n), device="cuda"), terms=torch.empty((T, 6, n), device="cuda"),
               reward=torch.empty((T, n), device="cuda"), done=torch.empty((T, n), dtype=torch.uint8, device="cuda"))
    if fused_returns and not per_step:
        out.update(ret=torch.empty((T, n), device="cuda"), adv=torch.empty((T, n), device="cuda"))
    ints0, seq0 = task.ints.clone(), task.sequence.clone()
    mom = torch.zeros(3, dtype=torch.float64, device="cuda")

    def step(ev=None):
        task.ints.copy_(ints0)
        if ev:
            ev[0].record()
        if args.per_step:
            for t in range(T):
                task.step(qpos[t], qvel[t], con[t], out={k: v[t] for k, v in out.items()})
        elif fused_returns:
            # obs / reward / done AND the discounted returns enqueued by one call (om_a3_task_rollout)
            res = task.step(qpos, qvel, con, out=out, returns=dict(values=values[:-1], v_next=values[1:], gamma=0.99))
            ret, adv = res["ret"], res["adv"]
        else:
            task.step(qpos, qvel, con, out=out)
        if args.per_step or not fused_returns:
            ret, adv = Kn.ppo_returns(out["reward"], values[:-1], 0.99, path_end=out["done"], v_next=values[1:])
        if ev:
            ev[1].record()            # the 490 B/env-step of SURVEY 8(d) include the returns pass, so it is timed with the task
        mom.zero_()
        Kn.moments_scalar(adv, out=mom)
        stats = Kn.adv_stats(mom, unbiased=True, eps=1e-5)
        Kn.normalize(adv, stats, out=adv)
        return ret, adv

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Kn.reset_launch_count()
    import time
    t0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        step(evs[i])
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps     # time the HOST needs to enqueue one step
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    kms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    launches = Kn.launch_count()
    eager_kms, eager_ms, timing = kms, ms, "eager launches, CUDA events around om_a3_task_rollout"
    if not args.per_step and host_ms > 0.8 * ms:
        # The host needs as long to enqueue the step's kernels as the GPU needs to run them (small configs): the
        # device-side cost is then measured by replaying the same launches out of a CUDA graph.
        def rollout_only():
            task.ints.copy_(ints0)
            task.step(qpos, qvel, con, out=out, **({"returns": dict(values=values[:-1], v_next=values[1:], gamma=0.99)} if fused_returns else {}))
            if not fused_returns:
                Kn.ppo_returns(out["reward"], values[:-1], 0.99, path_end=out["done"], v_next=values[1:])
        side = torch.cuda.Stream()
        graphs = []
        for fn in (rollout_only, step):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                fn()
                side.synchronize()
                with torch.cuda.graph(gr, stream=side):
                    fn()
            graphs.append(gr)
        torch.cuda.synchronize()
        times = []
        for gr in graphs:
            for _ in range(args.warmup):
                gr.replay()
            t0.record()
            for _ in range(args.steps):
                gr.replay()
            t1.record()
            torch.cuda.synchronize()
            times.append(t0.elapsed_time(t1) / args.steps)
        kms, ms = times
        kms -= 0.002                      # the restore of the task's integer state inside the replayed graph (a 16 B/env copy)
        timing = "CUDA-graph replay of the same launches (the eager loop is bound by the host's enqueue rate at this size)"
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    achieved = BYTES_PER_ENV_STEP * n * T / (kms * 1e-3) / 1e9
    return {"workload": f"A3 PPO walk rollout {n} envs x {T} steps (configs[2])", "per_step_launches": args.per_step,
            "fused_returns": bool(fused_returns and not args.per_step),
            "value": n * T / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms, "task_kernel_ms": kms,
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms, "timing": timing,
            "eager_ms_per_step": eager_ms, "eager_task_kernel_ms": eager_kms,
            "roofline": {"bound": "hbm", "kernel": "a3_feat_kernel + a3_walk_kernel + a3_post_kernel + affine_scan_kernel (om_a3_task_rollout: one call)", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "bytes_per_env_step": BYTES_PER_ENV_STEP}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--horizon", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--per-step", action="store_true", help="launch the step kernel once per env step (the live-rollout pattern)")
    ap.add_argument("--separate-returns", action="store_true", help="om_a3_task_step + om_ppo_returns instead of om_a3_task_rollout")
    a = ap.parse_args()
    print(json.dumps(measure(a.envs, a.horizon, a.steps, a.warmup, a.per_step, not a.separate_returns)))


if __name__ == "__main__":
    main()
