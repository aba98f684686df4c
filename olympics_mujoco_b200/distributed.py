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


class MailboxTimeout(RuntimeError):
    """A peer did not deliver its partial sums within the timeout: the round's output on this rank is NaN."""


class Mailbox:
    """``om_mailbox_*``: sum of <= 128 float64 values over the ranks through mailboxes in each rank's HBM, mapped into
    every peer process with CUDA IPC and written over NVLink by one small kernel per rank (csrc/om_mailbox.cu).  The IPC
    handles travel once, through ``torch.distributed.all_gather_object``.

    Construction is collective and never raises on one rank only: the status of the local create step is gathered with
    the handle, so every rank takes the same decision (use the mailbox / fall back) at the same point."""

    MAX_N = 128

    def __init__(self, timeout_ms=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        self._lib, self._C = lib, C
        world, rank = dist.get_world_size(), dist.get_rank()
        h = C.c_void_p()
        mine = (C.c_ubyte * 64)()
        rc = lib.om_mailbox_create(world, rank, C.byref(h), C.cast(mine, C.c_void_p))
        err = lib.om_last_error().decode() if rc != 0 else ""
        self.handle = h if rc == 0 else None
        gathered = [None] * world
        dist.all_gather_object(gathered, (rc, bytes(mine)))                     # unconditional: no rank is left waiting
        created = all(g[0] == 0 for g in gathered)
        if created:
            blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(g[1] for g in gathered))
            rc = lib.om_mailbox_connect(self.handle, C.cast(blob, C.c_void_p))
            if rc != 0:
                err = lib.om_last_error().decode()
        ok = torch.tensor([1 if (created and rc == 0) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)              # all ranks use the mailbox, or none does
        if int(ok) == 0:
            self.close()
            raise RuntimeError("mailbox: unavailable on some rank" + (f" (this rank: {err})" if err else ""))
        if timeout_ms is not None:
            self.set_timeout_ms(timeout_ms)

    def set_timeout_ms(self, ms):
        """0 waits for ever (NCCL's behaviour); the default is 30 s (or OM_MAILBOX_TIMEOUT_MS at creation)."""
        from . import _lib
        _lib.check(self._lib.om_mailbox_set_timeout_ms(self.handle, float(ms)))

    def all_reduce(self, x, out=None):
        """Enqueue one round on the current stream.  In place by default.  A round that times out leaves NaN in every
        output value on the rank that gave up (never a partial sum); ``check()`` turns that into an exception at the
        caller's next host synchronisation point."""
        from . import _lib
        out = x if out is None else out
        for t in (x, out):
            assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous() and t.numel() <= self.MAX_N
        assert out.numel() == x.numel()
        C = self._C
        _lib.check(self._lib.om_mailbox_allreduce(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), x.numel(),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    def timed_out(self):
        """Synchronises; True if a round gave up since the last call (the device word is cleared)."""
        from . import _lib
        f = self._C.c_int(0)
        _lib.check(self._lib.om_mailbox_timed_out(self.handle, self._C.byref(f)))
        return bool(f.value)

    def check(self):
        if self.timed_out():
            raise MailboxTimeout("NVLink mailbox all-reduce: a peer did not deliver within the timeout; the sums of that "
                                 "round are NaN on this rank (redo the round, e.g. over NCCL)")

    def close(self):
        if getattr(self, "handle", None):
            self._lib.om_mailbox_destroy(self.handle)
            self.handle = None


_mailbox = None
_mailbox_state = "unset"          # "unset" | "on" | "off"


def enable_mailbox(enable=True, timeout_ms=None):
    """Route ``all_reduce_moments`` through the NVLink mailbox kernel (collective call: every rank must make it).
    Falls back to NCCL on every rank if any rank cannot create or map a mailbox.  Returns whether the mailbox is in use."""
    global _mailbox, _mailbox_state
    if _mailbox is not None:
        _mailbox.close()
        _mailbox = None
    _mailbox_state = "off"
    if enable and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl":
        try:
            _mailbox = Mailbox(timeout_ms=timeout_ms)
            _mailbox_state = "on"
        except Exception as e:                                  # IPC not permitted, no peer access ...: NCCL it is
            import warnings
            warnings.warn(f"NVLink mailbox all-reduce unavailable ({e}); using NCCL")
    return _mailbox_state == "on"


def mailbox_check():
    """Host synchronisation point of the mailbox route: raises ``MailboxTimeout`` if a round since the last check gave up
    (its output is NaN on this rank).  No-op when the mailbox is not in use."""
    if _mailbox is not None:
        _mailbox.check()


def all_reduce_moments(mom):
    """Sum a float64 moment buffer (``om_moments`` layout: sum[C], sumsq[C], count) over the ranks, in place: the NVLink
    mailbox kernel when ``enable_mailbox()`` set it up (CUDA float64, <= 128 values), else ``dist.all_reduce`` (NCCL on
    the GPUs, gloo in the CPU tests)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if (_mailbox is not None and mom.is_cuda and mom.dtype == torch.float64 and mom.is_contiguous()
                and mom.numel() <= Mailbox.MAX_N):
            _mailbox.all_reduce(mom)
        else:
            dist.all_reduce(mom, op=dist.ReduceOp.SUM)
    return mom


def mean_std_from_moments(mom, kind):
    """``kind``: "standardizer" (networks.py:76-81: std = sqrt(max(E[x^2]-mean^2, 1e-2))), "ppo_obs"
    (normalize.py:48: sqrt(var + 1e-8)), "adv_ppo" (ppo.py:336: unbiased std + 1e-5), "adv_gail"
    (gail_TRPO.py:128: population std + 1e-8).  Returns (mean, denominator) float64.  A CUDA buffer goes through ONE
    kernel (``om_moment_stats``); the torch expressions below serve host tensors (the gloo tests)."""
    if mom.is_cuda:
        from . import kernels as Kn
        if kind not in Kn.MOMENT_KINDS:
            raise ValueError(kind)
        return Kn.moment_stats(mom, kind)
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
