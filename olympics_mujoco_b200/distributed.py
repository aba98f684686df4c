"""Multi-GPU plumbing of the path: environments shard by contiguous index range (one process per GPU), constants
are replicated, and the ONLY exchange is a small ``all_reduce(sum)`` of float64 moment partial sums -- where the
reference reduces across its Ray workers (``rl/envs/normalize.py:35-48``, ``rl/algos/ppo.py:334-336``) and across a
``fit`` batch (``imitation_lib/imitation/gail_TRPO.py:128``, ``imitation_lib/utils/networks.py:76-81``).

``torch.distributed`` is the transport (NCCL over NVLink on the GPUs, gloo in the CPU tests); nothing here touches
the data path.  All functions degrade to no-ops in a single process."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend=None, device=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).  Returns
    (rank, world, local_rank)."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local) if device is None else device
        dist.init_process_group(backend, **kw)
    return rank, world, local


def world():
    return dist.get_world_size() if dist.is_initialized() else 1


def env_shard(n_total, rank, world_size):
    """Contiguous, balanced shard of ``n_total`` envs -> (env_id0, n_local).  ``env_id0`` keys the Philox contract, so
    a sharded run draws exactly the resets of the single-process run."""
    base, rem = divmod(int(n_total), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    env_id0 = rank * base + min(rank, rem)
    return env_id0, n_local


def all_reduce_moments(mom):
    """Sum a float64 moment buffer (``om_moments`` layout: sum[C], sumsq[C], count) over the ranks, in place."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mom, op=dist.ReduceOp.SUM)
    return mom


def mean_std_from_moments(mom, kind):
    """``kind``: "standardizer" (networks.py:76-81: std = sqrt(max(E[x^2]-mean^2, 1e-2))), "ppo_obs"
    (normalize.py:48: sqrt(var + 1e-8)), "adv_ppo" (ppo.py:336: unbiased std + 1e-5), "adv_gail"
    (gail_TRPO.py:128: population std + 1e-8).  Returns (mean, denominator) float64."""
    c = (mom.numel() - 1) // 2
    s, ss, n = mom[:c], mom[c:2 * c], mom[2 * c]
    mean = s / n
    var = ss / n - mean * mean
    if kind == "standardizer":
        return mean, torch.sqrt(torch.clamp(var, min=1e-2))
    if kind == "ppo_obs":
        return mean, torch.sqrt(torch.clamp(var, min=0.0) + 1e-8)
    if kind == "adv_ppo":
        return mean, torch.sqrt(torch.clamp(var, min=0.0) * n / (n - 1)) + 1e-5
    if kind == "adv_gail":
        return mean, torch.sqrt(torch.clamp(var, min=0.0)) + 1e-8
    raise ValueError(kind)


class Standardizer:
    """Running standardiser of the discriminator input (``imitation_lib/utils/networks.py:48-81``) for sharded
    envs: the running sums live on the device as one float64 buffer [2C+1] that starts at (0, 1e-2, 1e-2) like the
    reference's; every update adds THIS rank's partial sums, all-reduces the increment and folds it in."""

    def __init__(self, n_features, device="cuda"):
        self.c = int(n_features)
        self.running = torch.zeros(2 * self.c + 1, dtype=torch.float64, device=device)
        self.running[self.c:] = 1e-2
        self.mean = torch.zeros(self.c, dtype=torch.float64, device=device)
        self.std = torch.ones(self.c, dtype=torch.float64, device=device)

    def update_from_moments(self, local_increment):
        inc = all_reduce_moments(local_increment.clone())
        self.running += inc
        self.mean, self.std = mean_std_from_moments(self.running, "standardizer")
        return self.mean, self.std

    def update(self, x_soa):
        """x_soa [C, n] float32 CUDA: partial sums by the K6 kernel, then ``update_from_moments``."""
        from . import kernels as Kn
        return self.update_from_moments(Kn.moments(x_soa))

    def snapshot_f32(self):
        return self.mean.to(torch.float32), self.std.to(torch.float32)
