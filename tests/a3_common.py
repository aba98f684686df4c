"""Shared helpers of the A3 parity tests (host harness on CPU, CUDA kernels on the GPU)."""
import numpy as np

from conftest import GOLDEN
from oracle import a3 as OA


def golden():
    return np.load(GOLDEN / "a3_task_ref.npz")


def contact4(c5):
    """[..., 5] (l_grf, r_grf, min_z, foot_contact, bad_collision) -> the C ABI's [..., 4] record."""
    out = np.empty(c5.shape[:-1] + (4,), np.float32)
    out[..., :3] = c5[..., :3]
    out[..., 3] = c5[..., 3] + 2 * c5[..., 4]
    return out


def lut6(period=OA.PERIOD):
    """What om_a3_task_create builds: the four phase clocks + sin/cos of the phase, fp32."""
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    ph = np.arange(period)
    return np.concatenate([phase_clock_lut(period=period), np.sin(2 * np.pi * ph / period)[:, None],
                           np.cos(2 * np.pi * ph / period)[:, None]], axis=1).astype(np.float32)


def oracle_obs(model, gold, e, upto=None):
    """Float64 oracle rollout of env e of the fixture -> (obs [T,41], total [T]); also re-checks the ints."""
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    lut = phase_clock_lut()
    _, _, ts, obs0 = OA.reset(model, int(gold["seed"]), e, 0, iteration_count=float(gold["iteration_count"]))
    T = gold["step_done"].shape[1] if upto is None else upto
    obs, total = np.empty((T, 41)), np.empty(T)
    for t in range(T):
        c = gold["step_contact"][e, t].astype(np.float64)
        con = OA.Contact(l_grf=c[0], r_grf=c[1], min_z=c[2], foot_contact=bool(c[3]), bad_collision=bool(c[4]))
        obs[t], total[t], _, _ = OA.step_tail(model, gold["step_qpos"][e, t].astype(np.float64),
                                              gold["step_qvel"][e, t].astype(np.float64), ts, con, lut)
    return obs0, obs, total


def threshold_cases(model, n=160, seed=7):
    """Adversarial inputs for the two float64 threshold decisions of the A3 task (walking_task.py:266-283 target_reached,
    :298-319 done): states whose decision margin is a few fp32 ulps -- far inside the error of any fp32 forward pass --
    found by bisection IN FP32 INPUT SPACE on the float64 oracle.  Half of the cases put `root z - lowest foot-site z`
    astride 0.6 (bisection on the right knee angle), half put `|left foot site - target|` astride the 0.2 m radius
    (bisection on the target's x).  Returns fp32 inputs and the float64 oracle's decisions for them."""
    from oracle import kinematics as K
    rng = np.random.default_rng(seed)
    ls, rs = model.site_id("lf_force"), model.site_id("rf_force")
    knee = int(model.jnt_qposadr[model.joint_id("right_knee")])

    def sites(q32):
        fk = K.forward(model, q32.astype(np.float64), np.zeros((len(q32), model.nv)))
        return fk["xpos"][:, 1, 2], fk["site_xpos"][:, ls], fk["site_xpos"][:, rs]

    def bisect(f, lo, hi, iters=40):
        """fp32 bisection of a float64 function with f(lo) < 0 <= f(hi): returns adjacent floats (a, b), f(a) < 0 <= f(b)."""
        lo, hi = lo.astype(np.float32), hi.astype(np.float32)
        for _ in range(iters):
            mid = (lo.astype(np.float64) + hi.astype(np.float64)) / 2
            mid = mid.astype(np.float32)
            neg = f(mid) < 0
            lo, hi = np.where(neg, mid, lo), np.where(neg, hi, mid)
        return lo, hi

    half = n // 2
    q = np.tile(model.qpos0, (n, 1)) + rng.normal(0, 0.08, (n, model.nq))
    q[:, :2] = rng.uniform(-2, 2, (n, 2)); q[:, 2] = rng.uniform(1.0, 1.4, n)
    quat = np.array([1.0, 0, 0, 0]) + rng.normal(0, 0.08, (n, 4))
    q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    q = q.astype(np.float32)
    seq = np.zeros((n, 20, 4), np.float32)
    seq[:, :, :3] = 1000.0                                              # targets out of reach unless planted below
    # ---- done: the right leg folds at the knee until the right site is the low one and root z - site z crosses 0.6
    def h_of_knee(kv):
        qq = q[:half].copy(); qq[:, knee] = kv
        rz, l, r = sites(qq)
        return (rz - np.minimum(l[:, 2], r[:, 2])) - 0.6
    # both thighs raised past the horizontal (the feet come up under the root: h ~ 0.35); unfolding the right knee lowers the
    # right foot again until h crosses 0.6
    for j in ("left_hip_y", "right_hip_y"):
        q[:half, int(model.jnt_qposadr[model.joint_id(j)])] = np.float32(-1.8)
    q[:half, int(model.jnt_qposadr[model.joint_id("left_knee")])] = 0.0
    lo_k, hi_k = np.full(half, 0.0, np.float32), np.full(half, 1.8, np.float32)     # h(lo) < 0.6 <= h(hi)
    ok = (h_of_knee(lo_k) < 0) & (h_of_knee(hi_k) >= 0)
    a, b = bisect(h_of_knee, lo_k, hi_k)
    pick = rng.integers(0, 4, half)                                      # the two adjacent floats and their neighbours
    kv = np.select([pick == 0, pick == 1, pick == 2, pick == 3],
                   [a, b, np.nextafter(a, np.float32(10)), np.nextafter(b, np.float32(-10))]).astype(np.float32)
    q[:half, knee] = np.where(ok, kv, q[:half, knee])
    # ---- near: the target sits 0.2 m in front (+x) of the left foot site; bisection on its x
    rz, l, r = sites(q[half:])
    ty, tz = l[:, 1].astype(np.float32), l[:, 2].astype(np.float32)

    def d_of_x(tx):
        return np.sqrt((l[:, 0] - tx.astype(np.float64)) ** 2 + (l[:, 1] - ty.astype(np.float64)) ** 2
                       + (l[:, 2] - tz.astype(np.float64)) ** 2) - 0.2
    a, b = bisect(d_of_x, (l[:, 0] + 0.1).astype(np.float32), (l[:, 0] + 0.3).astype(np.float32))
    pick = rng.integers(0, 4, n - half)
    tx = np.select([pick == 0, pick == 1, pick == 2, pick == 3],
                   [a, b, np.nextafter(a, np.float32(-1e9)), np.nextafter(b, np.float32(1e9))]).astype(np.float32)
    seq[half:, 0, 0], seq[half:, 0, 1], seq[half:, 0, 2] = tx, ty, tz
    ints = np.tile(np.array([0, 0, 1, 0, 1, 20, 0], np.int32), (n, 1))    # phase, t1, t2, frames, FORWARD, len, reached
    # ---- the float64 oracle's decisions on exactly these fp32 inputs
    rz, l, r = sites(q)
    h = rz - np.minimum(l[:, 2], r[:, 2])
    p = seq[:, 0, :3].astype(np.float64)
    dl, dr = np.linalg.norm(l - p, axis=1), np.linalg.norm(r - p, axis=1)
    margin = np.minimum(np.abs(h - 0.6), np.minimum(np.abs(dl - 0.2), np.abs(dr - 0.2)))
    return dict(qpos=q, seq=seq.reshape(n, 80), ints=ints, want_done=h < 0.6, want_reached=(dl < 0.2) | (dr < 0.2),
                margin=margin, planted=np.concatenate([ok, np.ones(n - half, bool)]))
