"""CPU baseline for bench.py (TEST INFRASTRUCTURE): the oracle's C restatement of the configs[1] workload --
H1 ``play_trajectory_from_velocity`` rollout + GAE -- timed on the host cores with OpenMP across envs (one env
per thread at a time, the reference's one-env-per-Ray-worker model).  kind = "port": the reference's own CPU
path (MuJoCo ``mj_forward`` driven from Python) cannot be installed in this image.  The port does LESS work
per step than ``mj_forward`` (no collision / constraint / RNE stages, no Python interpreter overhead), so it is
a conservative -- i.e. fast -- stand-in for the reference."""
import os
import time

import numpy as np

from . import c_oracle
from . import h1 as OH


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

    def one_step(seed):
        out = c_oracle.h1_play(cm, perm, table, seed, 0, n_env, horizon, record=True)
        vt, adv = c_oracle.gae(out["reward"], values[:, :-1], values[:, 1:], out["fallen"], last, 0.99, 0.97)
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        return float(out["checksum"].sum() + adv[0, 0])

    for i in range(warmup):
        one_step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        one_step(100 + i)
    dt = time.perf_counter() - t0
    return dict(value=n_env * horizon * steps / dt, ms_per_step=dt / steps * 1e3, cores=cores, kind="port",
                sample=f"{n_env} envs x {horizon} steps per step (of 4096 x 500), C float64 + OpenMP, "
                       f"{cores} threads of {os.cpu_count()} host CPUs")
