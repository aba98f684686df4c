"""Multi-GPU check of the NVLink mailbox all-reduce (csrc/om_mailbox.cu) against NCCL, and their latencies.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_mailbox.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from olympics_mujoco_b200 import distributed as D       # noqa: E402


def main():
    rank, world, local = D.init()
    torch.cuda.set_device(local)
    assert D.enable_mailbox(True), "mailbox could not be set up"
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    ok = True
    for it in range(200):                                  # many rounds back to back: parity slots, fast / slow ranks
        n = [68, 3, 128, 1, 65][it % 5]
        x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        if it % 7 == rank % 7:
            torch.cuda._sleep(200000)                      # this rank is late to the round
        ref = x.clone()
        dist.all_reduce(ref)
        got = D.all_reduce_moments(x.clone())
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        same_bits = all(torch.equal(gathered[0], t) for t in gathered)
        close = torch.allclose(got, ref, rtol=1e-13, atol=1e-13)
        ok = ok and same_bits and close
    torch.cuda.synchronize()
    assert not D._mailbox.timed_out()
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
        print(f"world {world}: results identical on every rank and equal to NCCL: {ok}; "
              f"us per all-reduce of 68 float64: NCCL {res['nccl']:.1f}, mailbox {res['mailbox']:.1f}")
    assert ok
    dist.barrier()
    D.enable_mailbox(False)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
