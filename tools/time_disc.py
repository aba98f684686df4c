"""Discriminator kernel time vs. number of tiles (diagnostic: in-CTA latency vs. cross-SM bandwidth contention)."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from olympics_mujoco_b200 import kernels as Kn
g = np.load(ROOT / "tests/golden/discriminator_ref.npz")
for kind in ("vail", "gail"):
    pref = "v_" if kind == "vail" else "g_"
    names = ("w1", "b1", "w2", "b2", "wmu", "bmu", "wlv", "blv", "wd", "bd") if kind == "vail" else ("w1", "b1", "w2", "b2", "w3", "b3")
    disc = Kn.Discriminator(kind, {k: g[pref + k] for k in names})
    mean, std = torch.zeros(32, device="cuda"), torch.ones(32, device="cuda")
    for tiles in (1, 2, 8, 37, 74, 148, 296, 592):
        n = 128 * tiles
        s = torch.randn((32, n), device="cuda")
        eps = torch.randn((128, n), device="cuda") if kind == "vail" else None
        for _ in range(3):
            disc.reward(s, mean, std, eps=eps)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); disc.reward(s, mean, std, eps=eps); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(kind, "tiles", tiles, "us min %.1f med %.1f" % (min(ts), sorted(ts)[5]))
