"""Multi-GPU check of the NVLink mailbox all-reduce (csrc/om_mailbox.cu) against NCCL, and their latencies.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_mailbox.py

``run_checks`` is also what tests/test_gpu_mailbox_multi.py runs (in-process under torchrun, or through a 2-rank
torchrun child when the box has two GPUs): sums bit-identical on every rank and bit-equal to a rank-ordered float64 sum
of the gathered inputs, equal to NCCL to rounding, late ranks inside the timeout, and a rank that is later than the
timeout (the ranks that gave up see NaN + MailboxTimeout, never a partial sum; the next round is correct again).
"""
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from olympics_mujoco_b200 import distributed as D       # noqa: E402


def rank_ordered_sum(x, world):
    """What the mailbox kernel computes: 0.0 + x_0 + x_1 + ... in rank order, float64 (NCCL's tree may round differently)."""
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    s = torch.zeros_like(x)
    for p in parts:
        s = s + p
    return s


def run_checks(rounds=200, timeout_case=True):
    rank, world = dist.get_rank(), dist.get_world_size()
    assert D._mailbox is not None, "mailbox not enabled"
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    ok = True
    for it in range(rounds):                               # many rounds back to back: parity slots, fast / slow ranks
        n = [68, 3, 128, 1, 65][it % 5]
        x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        if it % 7 == rank % 7:
            torch.cuda._sleep(200000)                      # this rank is late to the round (inside the timeout)
        ref = x.clone()
        dist.all_reduce(ref)
        exact = rank_ordered_sum(x, world)
        got = D.all_reduce_moments(x.clone())
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        same_bits = all(torch.equal(gathered[0], t) for t in gathered)
        ok = ok and same_bits and torch.equal(got, exact) and torch.allclose(got, ref, rtol=1e-13, atol=1e-13)
    # a rank that is MUCH later than the others, still inside the (default 30 s) timeout: ~0.3 s of device sleep
    x = torch.full((68,), float(rank + 1), dtype=torch.float64, device="cuda")
    if rank == world - 1:
        torch.cuda._sleep(600_000_000)
    got = D.all_reduce_moments(x.clone())
    torch.cuda.synchronize()
    ok = ok and bool((got == world * (world + 1) / 2).all())
    D.mailbox_check()                                      # nothing timed out so far
    if timeout_case and world > 1:
        dist.barrier()
        D._mailbox.set_timeout_ms(50.0)
        x = torch.full((5,), float(rank + 1), dtype=torch.float64, device="cuda")
        if rank == 0:
            torch.cuda._sleep(1_500_000_000)               # ~0.7 s: far beyond the 50 ms timeout
        got = D.all_reduce_moments(x.clone())
        torch.cuda.synchronize()
        if rank == 0:                                      # the late rank finds everyone's words: correct sum, no flag
            ok = ok and bool((got == world * (world + 1) / 2).all()) and not D._mailbox.timed_out()
        else:                                              # the ranks that gave up: NaN everywhere, flag set, then cleared
            ok = ok and bool(torch.isnan(got).all())
            try:
                D.mailbox_check()
                ok = False
            except D.MailboxTimeout:
                pass
            ok = ok and not D._mailbox.timed_out()
        dist.barrier()                                     # every rank is past the failed round before the next one starts
        D._mailbox.set_timeout_ms(30000.0)
        # two rounds so that both parity slots are exercised after the failed round
        for _ in range(2):
            x = torch.randn(68, dtype=torch.float64, device="cuda", generator=g)
            exact = rank_ordered_sum(x, world)
            got = D.all_reduce_moments(x.clone())
            torch.cuda.synchronize()
            ok = ok and torch.equal(got, exact)
        D.mailbox_check()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(int(flag))


def main():
    rank, world, local = D.init()
    torch.cuda.set_device(local)
    assert D.enable_mailbox(True), "mailbox could not be set up"
    ok = run_checks()
    # latency: 68 float64 (the bench's moment buffer), stream-ordered calls
    x = torch.randn(68, dtype=torch.float64, device="cuda")
    res = {}
    for name, fn in (("nccl", lambda: dist.all_reduce(x)), ("mailbox", lambda: D._mailbox.all_reduce(x))):
        for _ in range(20):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(200):
            fn()
        t1.record()
        torch.cuda.synchronize()
        res[name] = t0.elapsed_time(t1) / 200 * 1e3
    if rank == 0:
        print(f"world {world}: results identical on every rank, bit-equal to the rank-ordered sum and equal to NCCL, late "
              f"and timed-out ranks handled: {ok}; us per all-reduce of 68 float64: NCCL {res['nccl']:.1f}, "
              f"mailbox {res['mailbox']:.1f}")
    assert ok
    dist.barrier()
    D.enable_mailbox(False)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
