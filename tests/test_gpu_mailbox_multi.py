"""Multi-rank correctness of the NVLink mailbox all-reduce (csrc/om_mailbox.cu; the reduction sites it serves:
/root/reference rl/envs/normalize.py:35-48, rl/algos/ppo.py:334-336).

Two ways to run it:
  * under torchrun, one pytest per rank (every rank runs the same test in-process):
        python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
            -m pytest tests/test_gpu_mailbox_multi.py -m gpu -x -q
  * plain ``pytest -m gpu`` on a box with >= 2 GPUs: the test launches that torchrun line itself (tools/check_mailbox.py).
On a one-GPU box without torchrun it skips (tests/test_gpu_mailbox.py covers the world of one).
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_mailbox_sums_bit_identical_on_every_rank_late_and_timed_out_ranks():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:                                           # in-process, this pytest is one rank of a torchrun job
        from olympics_mujoco_b200 import distributed as D
        sys.path.insert(0, str(ROOT / "tools"))
        import check_mailbox
        rank, world, local = D.init()
        torch.cuda.set_device(local)
        assert D.enable_mailbox(True), "the mailbox could not be set up on every rank"
        try:
            assert check_mailbox.run_checks(rounds=60)
        finally:
            D.enable_mailbox(False)
        return
    if torch.cuda.device_count() < 2:
        pytest.skip("needs WORLD_SIZE > 1 (torchrun) or a box with two GPUs")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(ROOT / "tools" / "check_mailbox.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "handled: True" in r.stdout
