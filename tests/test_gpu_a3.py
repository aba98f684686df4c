"""GPU parity: StickFigureA3 step tail (K1 fused + K2/A3), its multi-step replay and the randomised reset, through
the C ABI, against (a) the fixture produced by the reference's own WalkingTask code and (b) the float64 oracle."""
import numpy as np
import pytest

import a3_common as A
from conftest import a3_random_states, assert_close

pytestmark = pytest.mark.gpu


def _task(model, n, seed=0, env_id0=0):
    from olympics_mujoco_b200 import kernels as Kn
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    from oracle import a3 as OA
    dm = Kn.DeviceModel(model)
    return Kn.A3Task(dm, n, phase_clock_lut(), OA.init_qpos(), seed=seed, env_id0=env_id0)


def _soa_t(x):
    """[n, T, C] numpy -> [T, C, n] float32 CUDA."""
    import torch
    return torch.as_tensor(np.ascontiguousarray(np.transpose(x, (1, 2, 0)), dtype=np.float32), device="cuda")


@pytest.mark.parametrize("multi_step", [False, "fused", "time_parallel"])
def test_a3_step_vs_reference_task_fixture(a3_model, multi_step, om_knob):
    """One step per call; T steps in one call through the fused one-thread-per-env kernel; and through the
    time-parallel pair (FK pass over (env, t) + sequential task pass)."""
    import torch
    if multi_step:
        om_knob("a3_split", int("1" if multi_step == "time_parallel" else "0"))
    gold = A.golden()
    n, T = gold["step_done"].shape
    task = _task(a3_model, n)
    task.ints.copy_(torch.as_tensor(gold["reset_ints"].T.astype(np.int32)))
    task.sequence.copy_(torch.as_tensor(gold["reset_sequence"].reshape(n, 80).T.astype(np.float32)))
    qpos, qvel, con = _soa_t(gold["step_qpos"]), _soa_t(gold["step_qvel"]), _soa_t(A.contact4(gold["step_contact"]))
    if multi_step:
        out = task.step(qpos, qvel, con)
        o = {k: v.cpu().numpy() for k, v in out.items()}
        ints_last = task.ints.cpu().numpy().T
        np.testing.assert_array_equal(ints_last, gold["step_ints"][:, -1])
    else:
        o = dict(obs=np.zeros((T, 41, n), np.float32), terms=np.zeros((T, 6, n), np.float32),
                 reward=np.zeros((T, n), np.float32), done=np.zeros((T, n), np.uint8))
        for t in range(T):
            out = task.step(qpos[t], qvel[t], con[t])
            for k in o:
                o[k][t] = out[k].cpu().numpy()
            np.testing.assert_array_equal(task.ints.cpu().numpy().T, gold["step_ints"][:, t], err_msg=f"step {t}")
    np.testing.assert_array_equal(o["done"].T.astype(bool), gold["step_done"])
    assert_close(np.transpose(o["terms"], (2, 0, 1)), gold["step_terms"], "terms")
    assert_close(o["reward"].T, gold["step_terms"].sum(axis=2), "reward")
    assert_close(np.transpose(o["obs"], (2, 0, 1))[:, :, 33:], gold["step_goal"], "goal steps")
    for e in range(2):                                                 # full observation vs the float64 oracle
        _, ref_obs, _ = A.oracle_obs(a3_model, gold, e, upto=120)
        assert_close(o["obs"][:120, :, e], ref_obs, "obs")


def test_a3_reset_vs_reference_task_fixture(a3_model):
    import torch
    gold = A.golden()
    n = gold["reset_qpos"].shape[0]
    task = _task(a3_model, n, seed=int(gold["seed"]))
    qpos, qvel = torch.zeros((25, n), device="cuda"), torch.zeros((24, n), device="cuda")
    obs = task.reset(qpos, qvel, iteration_count=float(gold["iteration_count"]))
    np.testing.assert_array_equal(task.ints.cpu().numpy().T, gold["reset_ints"])
    assert_close(qpos.cpu().numpy().T, gold["reset_qpos"], "qpos")
    assert_close(qvel.cpu().numpy().T, gold["reset_qvel"], "qvel")
    assert_close(task.sequence.cpu().numpy().T.reshape(n, 20, 4), gold["reset_sequence"], "sequence")
    assert_close(obs.cpu().numpy().T, gold["reset_obs"], "obs")
    assert np.array_equal(task.reset_count.cpu().numpy(), np.ones(n, np.int32))
    # masked reset: only env 1 and 4 draw again (counter 1), the others keep their state
    before = task.ints.clone()
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    mask[[1, 4]] = 1
    task.reset(qpos, qvel, mask=mask, iteration_count=float(gold["iteration_count"]))
    assert np.array_equal(task.reset_count.cpu().numpy(), [1, 2, 1, 1, 2, 1])
    keep = [0, 2, 3, 5]
    assert torch.equal(task.ints[:, keep], before[:, keep])
    from oracle import a3 as OA
    for e in (1, 4):
        q, v, ts, ob = OA.reset(a3_model, int(gold["seed"]), e, 1, iteration_count=float(gold["iteration_count"]))
        assert_close(qpos[:, e].cpu().numpy(), q, "qpos (second reset)")
        assert list(task.ints[:, e].cpu().numpy()) == [ts.phase, ts.t1, ts.t2, ts.target_reached_frames, ts.mode,
                                                       ts.seq_len, int(ts.target_reached)]


def test_a3_random_states_large_with_fk_outputs(a3_model):
    """4096 envs: reset on the GPU, then one step on random post-physics states with the MjData fields written;
    a sample of envs is replayed through the float64 oracle.  Integer / bool outputs must be identical (no tolerance on a
    flag: the kernels re-take near-threshold decisions in float64)."""
    import torch
    from oracle import a3 as OA
    from oracle import kinematics as K
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    n, seed = 4096, 77
    task = _task(a3_model, n, seed=seed, env_id0=1000)
    qpos, qvel = torch.zeros((25, n), device="cuda"), torch.zeros((24, n), device="cuda")
    task.reset(qpos, qvel, iteration_count=5000.0)
    ints0 = task.ints.cpu().numpy().T.copy()
    rng = np.random.default_rng(5)
    q0 = qpos.cpu().numpy().T.astype(np.float64)
    q = q0 + rng.normal(0, 0.05, q0.shape)
    q[:, 2] = rng.uniform(0.5, 1.5, n)                                  # some envs below the done height
    v = rng.normal(0, 1.0, (n, 24))
    q, v = q.astype(np.float32), v.astype(np.float32)
    fmax = a3_model.total_mass * 9.8 * 0.5
    c5 = np.stack([rng.uniform(0, 2 * fmax, n), rng.uniform(0, 2 * fmax, n), rng.uniform(-0.01, 0.01, n),
                   rng.random(n) < 0.7, rng.random(n) < 0.05], axis=1).astype(np.float32)
    want = ("obs", "terms", "reward", "done", "xpos", "xquat", "site_xpos", "site_xmat", "cvel")
    out = task.step(torch.as_tensor(q.T.copy(), device="cuda"), torch.as_tensor(v.T.copy(), device="cuda"),
                    torch.as_tensor(A.contact4(c5).T.copy(), device="cuda"), want=want)
    o = {k: x.cpu().numpy() for k, x in out.items()}
    ints1 = task.ints.cpu().numpy().T
    assert set(ints0[:, 4]) == {0, 1} and 0.1 < (ints0[:, 4] == 0).mean() < 0.3          # 20 % STANDING
    assert o["done"].any() and not o["done"].all()
    ref = K.forward(a3_model, q.astype(np.float64), v.astype(np.float64))
    assert_close(o["xpos"].T.reshape(n, 17, 3), ref["xpos"], "xpos")
    assert_close(o["xquat"].T.reshape(n, 17, 4), ref["xquat"], "xquat")
    assert_close(o["site_xpos"].T.reshape(n, 2, 3), ref["site_xpos"], "site_xpos")
    assert_close(o["site_xmat"].T.reshape(n, 2, 3, 3), ref["site_xmat"], "site_xmat")
    assert_close(o["cvel"].T.reshape(n, 17, 6), ref["cvel"], "cvel")
    lut = phase_clock_lut()
    sample = rng.choice(n, 96, replace=False)
    seq_dev = task.sequence.cpu().numpy().T.reshape(n, 20, 4).astype(np.float64)
    for e in sample:
        _, _, ts, _ = OA.reset(a3_model, seed, 1000 + int(e), 0, iteration_count=5000.0)
        assert [ts.phase, ts.t1, ts.t2, ts.target_reached_frames, ts.mode, ts.seq_len, int(ts.target_reached)] == list(ints0[e])
        L = len(ts.sequence)
        assert_close(np.asarray(ts.sequence)[:ts.seq_len], seq_dev[e, :ts.seq_len], "footstep plan of the reset")
        # the flags are exact functions of the fp32 INPUTS (qpos, contact summary, footstep plan): the oracle steps on the
        # plan the device holds (its own float64 plan differs from it by fp32 rounding)
        ts.sequence = [seq_dev[e, k].copy() if k < ts.seq_len else np.asarray(ts.sequence[k]).copy() for k in range(L)]
        c = c5[e].astype(np.float64)
        con = OA.Contact(l_grf=c[0], r_grf=c[1], min_z=c[2], foot_contact=bool(c[3]), bad_collision=bool(c[4]))
        obs, total, done, terms = OA.step_tail(a3_model, q[e].astype(np.float64), v[e].astype(np.float64), ts, con, lut)
        assert [ts.phase, ts.t1, ts.t2, ts.target_reached_frames, ts.mode, ts.seq_len, int(ts.target_reached)] == list(ints1[e])
        assert bool(o["done"][e]) == done
        assert_close(o["obs"][:, e], obs, "obs")
        assert_close(o["terms"][:, e], terms, "terms")
        assert_close(o["reward"][e], total, "reward")


@pytest.mark.parametrize("path", ["fused", "time_parallel"])
def test_a3_threshold_flags_equal_float64_oracle_on_adversarial_inputs(a3_model, path, om_knob):
    """A10 / A12 bit-exact WITHOUT excusals: inputs whose decision margin is a few fp32 ulps (1e-8 m, found by bisection
    on the float64 oracle in fp32 input space; the fp32 chain's own error is ~1e-6 m).  The kernels re-decide such cases
    from a float64 forward pass (om_a3_task.cuh: a3_done_height / a3_near_exact), so done and target_reached equal the
    float64 oracle's on every one of them."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    cases = A.threshold_cases(a3_model, n=4000, seed=11)
    n = len(cases["qpos"])
    assert cases["planted"].mean() > 0.95 and np.median(cases["margin"]) < 1e-6
    task = _task(a3_model, n, seed=1)
    t32 = lambda a: torch.as_tensor(np.ascontiguousarray(a.T), device="cuda")
    con = torch.tensor([100.0, 100.0, 0.0, 1.0], device="cuda")[:, None].repeat(1, n).contiguous()
    qpos, qvel = t32(cases["qpos"]), torch.zeros((24, n), device="cuda")
    if path == "fused":
        om_knob("a3_split", 0)
        task.ints.copy_(t32(cases["ints"])); task.sequence.copy_(t32(cases["seq"]))
        out = task.step(qpos, qvel, con)
        done, reached = out["done"].bool().cpu().numpy(), task.ints[6].bool().cpu().numpy()
    else:                              # two identical steps through the (env, t)-parallel kernels; the first one decides
        om_knob("a3_split", 1)
        task.ints.copy_(t32(cases["ints"])); task.sequence.copy_(t32(cases["seq"]))
        out = task.step(qpos[None].repeat(2, 1, 1).contiguous(), qvel[None].repeat(2, 1, 1).contiguous(),
                        con[None].repeat(2, 1, 1).contiguous())
        done = out["done"][0].bool().cpu().numpy()
        reached = (task.ints[3] > 0).cpu().numpy()       # frames counts the consecutive in-target steps (2 < delay)
        assert np.array_equal(out["done"][1].bool().cpu().numpy(), done)
    assert np.array_equal(done, cases["want_done"]), np.flatnonzero(done != cases["want_done"])[:10]
    assert np.array_equal(reached, cases["want_reached"]), np.flatnonzero(reached != cases["want_reached"])[:10]


def test_a3_edge_sizes(a3_model):
    import torch
    for n in (0, 1, 63, 65):
        task = _task(a3_model, n)
        qpos, qvel = torch.zeros((25, n), device="cuda"), torch.zeros((24, n), device="cuda")
        obs = task.reset(qpos, qvel)
        out = task.step(qpos, qvel, torch.zeros((4, n), device="cuda"))
        torch.cuda.synchronize()
        assert obs.shape == (41, n) and out["obs"].shape == (41, n) and out["done"].shape == (n,)
        if n:
            assert torch.isfinite(out["obs"]).all() and torch.isfinite(out["reward"]).all()
    task = _task(a3_model, 8)
    out = task.step(torch.zeros((0, 25, 8), device="cuda"), torch.zeros((0, 24, 8), device="cuda"),
                    torch.zeros((0, 4, 8), device="cuda"))
    assert out["obs"].shape == (0, 41, 8)


def test_stick_figure_a3_env_api(a3_model):
    """reset / step through the reference-facing class with an attached (scripted) physics callable."""
    import torch
    from olympics_mujoco_b200 import StickFigureA3
    from oracle import a3 as OA
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    n = 5
    env = StickFigureA3(n_envs=n, seed=3, algorithm_type="AlgorithmType.REINFORCEMENT_LEARNING")
    assert env.observation_space.shape == (41,) and env.action_space.shape == (12,)
    assert len(env.robot.mirrored_obs) == 41 and env.task._period == 88 and env.task.delay_frames == 30
    with pytest.raises(RuntimeError):
        env.step(np.zeros((n, 12)))
    obs0 = env.reset()
    assert obs0.shape == (n, 41)
    rng = np.random.default_rng(1)
    script = []

    def physics(e, target):
        assert target.shape == (n, 12)
        q = e.qpos.clone() + torch.as_tensor(rng.normal(0, 0.01, (25, n)), dtype=torch.float32, device="cuda")
        v = e.qvel.clone() + torch.as_tensor(rng.normal(0, 0.1, (24, n)), dtype=torch.float32, device="cuda")
        c = torch.as_tensor(np.stack([rng.uniform(0, 300, n), rng.uniform(0, 300, n), rng.uniform(-0.01, 0.01, n),
                                      rng.integers(0, 2, n).astype(np.float64)]), dtype=torch.float32, device="cuda")
        script.append((q.cpu().numpy().T.astype(np.float64), v.cpu().numpy().T.astype(np.float64), c.cpu().numpy().T))
        return q, v, c
    env.attach_dynamics(physics)
    lut = phase_clock_lut()
    states = [OA.reset(a3_model, 3, e, 0, iteration_count=np.inf) for e in range(n)]
    for e in range(n):
        assert_close(obs0[e].cpu().numpy(), states[e][3], "reset obs")
    for _ in range(3):
        obs, rew, done, rewards = env.step(np.zeros((n, 12), np.float32))
        assert obs.shape == (n, 41) and rew.shape == (n,) and done.dtype == torch.bool and len(rewards) == 6
        q, v, c = script[-1]
        for e in range(n):
            con = OA.Contact(l_grf=c[e, 0], r_grf=c[e, 1], min_z=c[e, 2], foot_contact=bool(int(c[e, 3]) & 1),
                             bad_collision=bool(int(c[e, 3]) & 2))
            ro, rt, rd, rterms = OA.step_tail(a3_model, q[e], v[e], states[e][2], con, lut)
            assert_close(obs[e].cpu().numpy(), ro, "obs"); assert_close(rew[e].cpu().numpy(), rt, "reward")
            assert bool(done[e]) == rd
            assert_close(rewards["step_reward"][e].cpu().numpy(), rterms[4], "step_reward")
    single = StickFigureA3(seed=3)
    assert single.reset().shape == (41,)


@pytest.mark.parametrize("delay,radius", [(0, 5.0), (1, 5.0), (2, 0.9), (5, 0.6), (30, 5.0), (7, 0.35)])
def test_a3_replay_state_machine_stress(a3_model, delay, radius, om_knob):
    """The candidate-bit state machine of the time-parallel replay (feat -> walk -> post kernels) against the fused
    one-thread-per-env kernel under short delays and radii that make "target near" frequent or permanent: the target
    advances every few steps, the candidate chain clamps at the last target, calls split into sub-calls.  Integer state
    and flags identical, floats to 1e-6."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from olympics_mujoco_b200.tasks.phase_clock import phase_clock_lut
    from oracle import a3 as OA
    gold = A.golden()
    n, T = gold["step_done"].shape
    dm = Kn.DeviceModel(a3_model)
    qpos, qvel, con = _soa_t(gold["step_qpos"]), _soa_t(gold["step_qvel"]), _soa_t(A.contact4(gold["step_contact"]))
    res = {}
    for split in ("0", "1"):
        om_knob("a3_split", int(split))
        task = Kn.A3Task(dm, n, phase_clock_lut(), OA.init_qpos(), delay_frames=delay, target_radius=radius)
        ints = gold["reset_ints"].T.astype(np.int32).copy()
        ints[:, 3:] = np.array([[3, 4, 5, 1, 1, 20, 1]], np.int32).T         # some envs start mid-plan with frames pending
        ints[:, 4] = [7, 2, 3, max(delay - 1, 0), 1, 20, 1]                   # one step short of an advance
        ints[:, 5] = [7, 18, 19, delay + 3, 1, 20, 1]                         # overdue, at the end of the plan
        task.ints.copy_(torch.as_tensor(ints))
        task.sequence.copy_(torch.as_tensor(gold["reset_sequence"].reshape(n, 80).T.astype(np.float32)))
        out = task.step(qpos, qvel, con)
        torch.cuda.synchronize()
        res[split] = ({k: v.cpu().numpy() for k, v in out.items()}, task.ints.cpu().numpy())
    (a, ia), (b, ib) = res["0"], res["1"]
    np.testing.assert_array_equal(ia, ib)
    np.testing.assert_array_equal(a["done"], b["done"])
    assert (ia[1] != gold["reset_ints"].T[1]).any()                            # targets did advance
    assert_close(b["obs"], a["obs"], "obs", rtol=1e-6, atol=1e-6)
    assert_close(b["terms"], a["terms"], "terms", rtol=1e-6, atol=1e-6)
    assert_close(b["reward"], a["reward"], "reward", rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("T,n", [(64, 1000), (37, 130), (230, 96)])
def test_a3_rollout_with_fused_returns_equals_step_then_ppo_returns(a3_model, T, n):
    """om_a3_task_rollout (configs[2]: obs / reward / done + PPOBuffer.finish_path returns, one call) ==
    om_a3_task_step followed by om_ppo_returns, bit for bit on every output; T = 230 takes the several-sub-call route;
    the returns are also checked against the reference's PPOBuffer restatement."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from oracle import learner as L
    g = torch.Generator(device="cuda").manual_seed(5)
    res = {}
    for fused in (False, True):
        task = _task(a3_model, n, seed=3)
        q0, v0 = Kn.soa(25, n), Kn.soa(24, n)
        task.reset(q0, v0, iteration_count=5000.0)
        if not fused:
            qpos = q0[None] + 0.01 * torch.randn((T, 25, n), device="cuda", generator=g).cumsum(0)
            fold = torch.linspace(0, 1, T, device="cuda")[:, None] * (torch.rand(n, device="cuda", generator=g) < 0.3)
            for hip, knee in ((7, 10), (13, 16)):
                qpos[:, hip] -= 1.5 * fold
                qpos[:, knee] -= 1.6 * fold
            qvel = torch.randn((T, 24, n), device="cuda", generator=g)
            con = torch.stack([torch.rand((T, n), device="cuda", generator=g) * 400, torch.rand((T, n), device="cuda", generator=g) * 400,
                               (torch.rand((T, n), device="cuda", generator=g) - 0.5) * 0.02,
                               (torch.rand((T, n), device="cuda", generator=g) < 0.7).float()], dim=1).contiguous()
            values = torch.randn((T + 1, n), device="cuda", generator=g)
            v_last = torch.randn(n, device="cuda", generator=g)
            out = task.step(qpos, qvel, con)
            ret, adv = Kn.ppo_returns(out["reward"], values[:-1].contiguous(), 0.99, path_end=out["done"], v_next=values[1:].contiguous(),
                                      v_last=v_last)
            res[fused] = dict(out, ret=ret, adv=adv, ints=task.ints.clone())
        else:
            out = task.step(qpos, qvel, con, returns=dict(values=values[:-1].contiguous(), v_next=values[1:].contiguous(),
                                                          v_last=v_last, gamma=0.99))
            res[fused] = dict(out, ints=task.ints.clone())
    torch.cuda.synchronize()
    for k in ("obs", "terms", "reward", "done", "ret", "adv", "ints"):
        assert torch.equal(res[False][k], res[True][k]), k
    assert bool(res[True]["done"].any())
    r, d, v = (res[True][k].cpu().numpy() for k in ("reward", "done", "ret"))
    vals, vl = values.cpu().numpy(), v_last.cpu().numpy()
    for e in range(0, n, max(1, n // 7)):
        want = L.ppo_returns_path_end(r[:, e], d[:, e], 0.99, v_last=vl[e]) if hasattr(L, "ppo_returns_path_end") else None
        if want is None:                                   # plain restatement of finish_path over the done-delimited paths
            want = np.zeros(T)
            R = float(vl[e])
            for t in range(T - 1, -1, -1):
                if d[t, e] == 1:
                    R = 0.0
                R = float(r[t, e]) + 0.99 * R
                want[t] = R
        assert_close(v[:, e], want, "returns vs finish_path", rtol=2e-5, atol=2e-5)


def test_a3_replay_is_deterministic_under_repetition(a3_model):
    """The post pass of the replay starts while the walk pass runs and each of its threads waits for bit 7 of its own
    per-step byte (om_a3.cu: a3_walk_kernel / a3_post_kernel).  A hole in that hand-over -- a post thread reading a
    candidate byte as a state code, or a stale byte of the previous call -- shows as run-to-run differences: 30
    back-to-back rollouts of 4096 envs x 64 steps from the same task state, every output bit-identical to the first,
    with a smaller and a longer call in between so that stale bytes of other shapes lie in the scratch."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    n, T = 4096, 64
    g = torch.Generator(device="cuda").manual_seed(9)
    task = _task(a3_model, n, seed=11)
    q0, v0 = Kn.soa(25, n), Kn.soa(24, n)
    task.reset(q0, v0, iteration_count=5000.0)
    qpos = q0[None] + 0.01 * torch.randn((T + 40, 25, n), device="cuda", generator=g).cumsum(0)
    qvel = torch.randn((T + 40, 24, n), device="cuda", generator=g)
    con = torch.stack([torch.rand((T + 40, n), device="cuda", generator=g) * 400, torch.rand((T + 40, n), device="cuda", generator=g) * 400,
                       (torch.rand((T + 40, n), device="cuda", generator=g) - 0.5) * 0.02,
                       (torch.rand((T + 40, n), device="cuda", generator=g) < 0.7).float()], dim=1).contiguous()
    ints0 = task.ints.clone()
    first = None
    for it in range(30):
        task.ints.copy_(ints0)
        steps = T if it % 5 else (T + 40 if it % 2 else 7)           # every fifth call has another length
        out = task.step(qpos[:steps], qvel[:steps], con[:steps])
        if steps != T:
            continue
        cur = {k: v.clone() for k, v in out.items()}
        cur["ints"] = task.ints.clone()
        if first is None:
            first = cur
        else:
            for k, v in cur.items():
                assert torch.equal(v, first[k]), f"{k} differs on rollout {it}"
    torch.cuda.synchronize()
    assert first is not None
