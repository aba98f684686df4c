"""CPU baselines for bench.py (TEST INFRASTRUCTURE), all on the configs[1] workload -- H1
``play_trajectory_from_velocity`` rollout + GAE:

(iii) ``run``            the oracle's C restatement, float64, OpenMP across envs on every host core (one env per thread at
                         a time: the reference's one-env-per-Ray-worker model).  kind = "port".  Output arrays are
                         allocated ONCE outside the timed region, as the GPU arm does with its rollout buffers.
(i)   ``python_loop``    the float64 NumPy oracle in the reference's own form: ONE env, a Python loop over the steps
                         (loco_env_base.py:511-557), one process, one core.
(ii)  ``python_multiproc`` the same, one process per host core (the reference's Ray ``num_procs`` model,
                         examples/reinforcement_learning_ppo/a3/train_a3_walk.py:138).

The reference's own CPU path (MuJoCo ``mj_forward`` driven from Python) cannot be installed in this image.  All three
do LESS work per step than ``mj_forward`` (no collision / constraint / RNE stages), so they are conservative -- i.e.
fast -- stand-ins for the reference."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

from . import c_oracle
from . import h1 as OH

ROOT = Path(__file__).resolve().parent.parent


def run(model, table, steps=1, warmup=0, horizon=500, budget_s=15.0, n_env=None):
    cm = c_oracle.CModel(model)
    perm = OH.perm(model)
    cores = c_oracle.use_all_cores()
    # size the bounded sample: time a small probe, then pick n_env so one "step" takes ~budget_s
    probe = max(cores, 8)
    t0 = time.perf_counter()
    c_oracle.h1_play(cm, perm, table, 1, 0, probe, horizon, record=False)
    per_env_step = (time.perf_counter() - t0) / (probe * horizon)
    if n_env is None:
        n_env = int(max(cores, min(4096, budget_s / max(steps + warmup, 1) / (per_env_step * horizon))))
        n_env = max(cores, (n_env // cores) * cores)
    rng = np.random.default_rng(0)
    values = rng.normal(0, 1, (n_env, horizon + 1))
    last = np.zeros((n_env, horizon), np.uint8)
    # preallocated outputs (touched once here, so the timed steps see neither page faults nor zero fills)
    out = c_oracle.h1_play_buffers(model, table.shape[0], n_env, horizon)
    gae_out = c_oracle.gae_buffers(n_env, horizon)
    v, v_next = np.ascontiguousarray(values[:, :-1]), np.ascontiguousarray(values[:, 1:])

    def one_step(seed):
        c_oracle.h1_play(cm, perm, table, seed, 0, n_env, horizon, record=True, out=out)
        vt, adv = c_oracle.gae(out["reward"], v, v_next, out["fallen"], last, 0.99, 0.97, out=gae_out)
        np.subtract(adv, adv.mean(), out=adv)
        np.divide(adv, adv.std() + 1e-8, out=adv)
        return float(out["checksum"].sum() + adv[0, 0])

    one_step(7)                                            # first touch of every buffer, outside the timer
    for i in range(warmup):
        one_step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        one_step(100 + i)
    dt = time.perf_counter() - t0
    return dict(value=n_env * horizon * steps / dt, ms_per_step=dt / steps * 1e3, cores=cores, kind="port",
                sample=f"{n_env} envs x {horizon} steps per step (of 4096 x 500), C float64 + OpenMP, "
                       f"{cores} threads of {os.cpu_count()} host CPUs, outputs preallocated")


def _python_rollout(model, table, horizon, seed, env_id):
    """One env, reference control flow: Python loop over the steps + compute_gae over the episode."""
    from . import learner as L
    t0 = time.perf_counter()
    r = OH.play_trajectory_from_velocity(model, table, 1, horizon, seed=seed, env_id=env_id)
    v = np.random.default_rng(env_id).normal(0, 1, horizon + 1)
    L.compute_gae(v[:-1], v[1:], r["reward"], r["fallen"], np.zeros(horizon, bool), 0.99, 0.97)
    return time.perf_counter() - t0


def python_loop(model, table, horizon=500, episodes=1):
    """(i): single process, single env, Python loop.  Returns env-steps/s."""
    dt = sum(_python_rollout(model, table, horizon, 3, e) for e in range(episodes))
    return dict(value=episodes * horizon / dt, cores=1, kind="port",
                sample=f"{episodes} env x {horizon} steps, float64 NumPy oracle, Python loop over steps, 1 process")


def python_multiproc(horizon=500, procs=None, timeout=300):
    """(ii): one worker process per host core, each the single-env Python loop (started as fresh interpreters: the
    caller may hold a CUDA context, which does not survive fork).  Aggregate env-steps / slowest worker's timed span."""
    procs = procs or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    ps = [subprocess.Popen([sys.executable, "-m", "oracle.cpu_baseline", "--py-worker", str(i), str(horizon)], cwd=str(ROOT),
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for i in range(procs)]
    spans = []
    for p in ps:
        out, _ = p.communicate(timeout=timeout)
        if p.returncode == 0:
            spans.append(json.loads(out.strip().splitlines()[-1])["s"])
    if not spans:
        raise RuntimeError("no python-loop worker finished")
    return dict(value=len(spans) * horizon / max(spans), cores=len(spans), kind="port",
                sample=f"{len(spans)} processes x 1 env x {horizon} steps, float64 NumPy oracle, Python loop (Ray-worker model)")


def _py_worker(idx, horizon):
    sys.path.insert(0, str(ROOT))
    from olympics_mujoco_b200 import mjcf, synthetic
    from olympics_mujoco_b200.utils.trajectory import resample_table
    model = mjcf.load_builtin("unitree_h1")
    table = resample_table(synthetic.h1_walk_dataset(n_traj=4, t_raw=2500, seed=0, model=model), model)
    print(json.dumps({"s": _python_rollout(model, table, horizon, 3, idx)}))


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "--py-worker":
        _py_worker(int(sys.argv[2]), int(sys.argv[3]))
