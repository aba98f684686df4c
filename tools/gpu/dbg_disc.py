import sys, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo")
from olympics_mujoco_b200 import kernels as Kn, _lib
from pathlib import Path; GOLDEN = Path("/root/repo/tests/golden")
n = int(sys.argv[1])
lib = _lib.load()
g = np.load(GOLDEN / "discriminator_ref.npz")
p = {k: g["v_" + k] for k in ("w1", "b1", "w2", "b2", "wmu", "bmu", "wlv", "blv", "wd", "bd")}
disc = Kn.Discriminator("vail", p)
s = torch.randn(32, n, device="cuda"); eps = torch.randn(128, n, device="cuda")
mean = torch.zeros(32, device="cuda"); std = torch.ones(32, device="cuda")
for _ in range(3): disc.reward(s, mean, std, eps=eps)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
lib.om_debug_dump.argtypes = [ctypes.c_void_p]
lib.om_debug_dump(buf)
a = np.array(buf)
t0 = a[0]
def rel(x): return [int(v - t0) if v else None for v in x]
print("issuer: a_full(k) passed      ", rel(a[0:13]))
print("issuer: b_full(q) passed      ", rel(a[36:54]))
print("issuer: chunk q issued        ", rel(a[16:34]))
print("producer: st(k) issued        ", rel(a[56:69]))
print("producer: publish(k) arrived  ", rel(a[72:85]))
print("producer: ready R1A R1B R2 R3A R3B passed", rel(a[88:93]), "tile end", rel(a[95:96]))
