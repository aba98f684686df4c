"""Host-path ceiling of the end-to-end leg: concurrent device->host copy bandwidth per rank at N = 1, 2, 4, 8 ranks of ONE
torchrun job (ranks >= N idle), with bench.py's buffer size (288.8 MB per step), pinned host memory and CPU affinity.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/probe_d2h.py

Variants: "torch_pinned" (tensor.pin_memory() after nvmlDeviceSetCpuAffinity -- what bench.py does), "no_affinity" (the
same without the affinity call), "hostalloc_wc" (cudaHostAlloc write-combined), "hostalloc_portable" (default flags through
the runtime API).  One JSON line per (variant, N) on rank 0: per-rank GB/s (min / mean) and the aggregate."""
import ctypes
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
BYTES = 288_768_000


def host_alloc(nbytes, flags):
    rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc failed: {rc}")
    ctypes.memset(p, 0, nbytes)                           # first touch on this thread's NUMA node
    return rt, p


def timed_copies(src, dst_ptr_or_tensor, reps, rt=None):
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    t0.record()
    for _ in range(reps):
        if rt is None:
            dst_ptr_or_tensor.copy_(src, non_blocking=True)
        else:
            rt.cudaMemcpyAsync(dst_ptr_or_tensor, ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(BYTES), 2, ctypes.c_void_p(st))
    t1.record()
    torch.cuda.synchronize()
    return BYTES * reps / (t0.elapsed_time(t1) * 1e-3) / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    src = torch.empty(BYTES, dtype=torch.uint8, device="cuda")
    full = os.sched_getaffinity(0)
    results = []
    numa = None
    for variant in ("no_affinity", "torch_pinned", "hostalloc_wc", "hostalloc_portable"):
        os.sched_setaffinity(0, full)
        if variant != "no_affinity":
            try:
                import pynvml
                pynvml.nvmlInit()
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
                numa = sorted(os.sched_getaffinity(0))[:2]
            except Exception as e:
                numa = f"affinity failed: {e!r}"
        rt = None
        if variant.startswith("hostalloc"):
            rt, dst = host_alloc(BYTES, 0x04 if variant.endswith("wc") else 0x01)
        else:
            dst = torch.empty(BYTES, dtype=torch.uint8).pin_memory()
        timed_copies(src, dst, 2, rt)                      # warm-up
        for n in [k for k in (1, 2, 4, 8) if k <= world]:
            if world > 1:
                dist.barrier()
            gbs = timed_copies(src, dst, 8, rt) if rank < n else 0.0
            t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
            allv = [torch.zeros_like(t) for _ in range(world)]
            if world > 1:
                dist.all_gather(allv, t)
            else:
                allv = [t]
            vals = [float(v) for v in allv[:n]]
            if rank == 0:
                results.append(dict(variant=variant, ranks=n, per_rank_min=round(min(vals), 2), per_rank_mean=round(sum(vals) / n, 2),
                                    aggregate=round(sum(vals), 1), unit="GB/s", bytes_per_copy=BYTES))
        if rt is not None:
            rt.cudaFreeHost(dst)
        del dst
    if rank == 0:
        print(json.dumps(dict(host_cpus=len(full), first_cpus_near_gpu0=numa, results=results)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
