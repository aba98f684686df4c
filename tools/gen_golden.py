"""Generate tests/golden/* by running the reference's OWN Python modules (authoring container only).

Each reference file is loaded by path (never ``import olympic_mujoco``: its package ``__init__`` pulls
MuJoCo), with tiny stubs for absent third-party modules.  Outputs are small ``.npz`` fixtures that
travel to the GPU box; /root/reference itself does not.

    python tools/gen_golden.py [/root/reference]
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, str(REF / rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def flat(sample):
    return np.concatenate([np.atleast_1d(x) for x in sample])


def stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


# ---------------------------------------------------------------- 1. Trajectory (utils/trajectory.py)
def gen_trajectory():
    from olympics_mujoco_b200 import mjcf, synthetic
    ref_traj = load("ref_traj", "olympic_mujoco/utils/trajectory.py")
    model = mjcf.load_builtin("unitree_h1")
    data = synthetic.h1_walk_dataset(n_traj=3, t_raw=250, seed=7, model=model)
    # push two channels out of range so that clipping is exercised
    data["q_knee_angle_r"] = data["q_knee_angle_r"] + 0.8
    data["q_hip_rotation_l"] = data["q_hip_rotation_l"] * 1.5
    keys = [k for k in data if k != "split_points"]
    joints = [k[2:] for k in keys[:17]]
    low = np.full(34, -np.inf)
    high = np.full(34, np.inf)
    for i, j in enumerate(joints):
        jid = model.jnt_names.index(j)
        if model.jnt_limited[jid]:
            low[i], high[i] = model.jnt_range[jid]
    tr = ref_traj.Trajectory(keys=list(keys), low=low[2:], high=high[2:], joint_pos_idx=np.arange(17),
                             interpolate_map=lambda t: np.array(t), interpolate_remap=lambda t: [o for o in t],
                             traj_files={k: v.copy() for k, v in data.items()}, traj_dt=1 / 500.0,
                             control_dt=1 / 100.0, clip_trajectory_to_joint_ranges=True, warn=False)
    table = np.array(tr.trajectories)                       # [K, n_traj, T]
    script, samples, state = [], [], []
    rng = np.random.default_rng(3)
    for _ in range(6):
        traj_no, sub = int(rng.integers(0, table.shape[1])), int(rng.integers(0, table.shape[2]))
        s = tr.reset_trajectory(substep_no=sub, traj_no=traj_no)
        script.append((0, traj_no, sub)); samples.append(flat(s)); state.append((tr.traj_no, tr.subtraj_step_no))
        s = tr.get_current_sample()
        script.append((1, -1, -1)); samples.append(flat(s)); state.append((tr.traj_no, tr.subtraj_step_no))
        for _ in range(int(rng.integers(3, 60))):
            s = tr.get_next_sample()
            script.append((2, -1, -1))
            samples.append(np.full(34, np.nan) if s is None else flat(s))
            state.append((tr.traj_no, tr.subtraj_step_no))
            if s is None:
                break
    ds = tr.create_dataset(ignore_keys=["q_pelvis_tx", "q_pelvis_tz"])
    np.savez_compressed(OUT / "trajectory_ref.npz", keys=np.array(keys), low=low, high=high,
                        table=table, split_points=tr.split_points, script=np.array(script),
                        samples=np.array(samples), state=np.array(state),
                        ds_states=ds["states"], ds_next_states=ds["next_states"],
                        ds_absorbing=ds["absorbing"], ds_last=ds["last"],
                        **{"in_" + k: v for k, v in data.items()})
    print("trajectory_ref: table", table.shape, "script", len(script))


# ---------------------------------------------------------------- 2. phase clocks (tasks/rewards.py)
def gen_phase_clock():
    rw = load("ref_rewards", "olympic_mujoco/tasks/rewards.py")
    right, left = rw.create_phase_reward(0.75, 0.35, 0.1, "grounded", 1 / 0.025)
    ph = np.arange(88)
    lut = np.stack([right[0](ph), right[1](ph), left[0](ph), left[1](ph)], axis=1)
    np.savez(OUT / "phase_clock_ref.npz", lut=lut)
    print("phase_clock_ref:", lut.shape, lut[:4, 0])


# ---------------------------------------------------------------- 3. PPOBuffer (rl/algos/ppo.py)
def gen_ppo():
    stub("ray", remote=lambda f: f)
    stub("matplotlib"); stub("matplotlib.pyplot")
    stub("rl"); stub("rl.envs", WrapEnv=object)
    ppo = load("ref_ppo", "rl/algos/ppo.py")
    rng = np.random.default_rng(11)
    T, N, gamma = 64, 8, 0.99
    rewards = rng.normal(0, 1, (T, N))
    values = rng.normal(0, 1, (T, N))
    done_last = rng.random(N) < 0.5
    v_boot = rng.normal(0, 1, N)
    ret = np.empty((T, N))
    for e in range(N):
        buf = ppo.PPOBuffer(gamma, 0.95)
        for t in range(T):
            buf.store(np.zeros((1, 1)), np.zeros((1, 1)), np.array([rewards[t, e]]), np.array([values[t, e]]))
        buf.finish_path(last_val=(not done_last[e]) * np.array([v_boot[e]]))
        ret[:, e] = np.array(buf.returns).reshape(-1)
    import torch
    r_t, v_t = torch.Tensor(ret.T.reshape(-1)), torch.Tensor(values.T.reshape(-1))
    adv = r_t - v_t
    adv_n = (adv - adv.mean()) / (adv.std() + 1e-5)                      # ppo.py:335-336
    kat = ppo.PPOBuffer(0.99, 0.95)
    for r in (1.0, 2.0, 3.0):
        kat.store(np.zeros((1, 1)), np.zeros((1, 1)), np.array([r]), np.array([0.0]))
    kat.finish_path(last_val=np.array([10.0]))
    np.savez(OUT / "ppo_returns_ref.npz", rewards=rewards, values=values, done_last=done_last, v_boot=v_boot,
             gamma=gamma, returns=ret, adv_norm=adv_n.numpy().reshape(N, T).T,
             kat_returns=np.array(kat.returns).reshape(-1))
    print("ppo_returns_ref: kat", np.array(kat.returns).reshape(-1))


# ---------------------------------------------------------------- 4. discriminator nets (networks.py)
def gen_networks():
    import torch
    stub("mushroom_rl"); stub("mushroom_rl.utils")
    stub("mushroom_rl.utils.preprocessors", RunningStandardization=object)
    nets = load("ref_networks", "imitation_lib/utils/networks.py")
    torch.manual_seed(0)
    std = nets.Standardizer()
    enc = nets.FullyConnectedNetwork(input_shape=(32,), output_shape=(128,), n_features=[256],
                                     activations=["relu", "relu"], standardizer=None, squeeze_out=False)
    dec = nets.FullyConnectedNetwork(input_shape=(128,), output_shape=(1,), n_features=[],
                                     activations=["identity"], standardizer=None,
                                     initializers=[nets.NormcInitializer(std=0.1)], squeeze_out=False)
    vail = nets.VariationalNet(input_shape=(32,), output_shape=(1,), z_size=128, encoder_net=enc,
                               decoder_net=dec, use_next_states=False, use_actions=False, standardizer=std)
    gstd = nets.Standardizer()
    gail = nets.DiscriminatorNetwork(input_shape=(32,), output_shape=(1,), n_features=[512, 256],
                                     activations=["tanh", "tanh", "identity"], squeeze_out=False,
                                     standardizer=gstd, use_actions=False, use_next_states=False)
    n_par_vail = sum(p.numel() for p in vail.parameters())
    n_par_gail = sum(p.numel() for p in gail.parameters())
    rng = np.random.default_rng(5)
    B = 256
    s = (rng.normal(0, 1, (B, 32)) * rng.uniform(0.2, 3, 32) + rng.normal(0, 1, 32)).astype(np.float32)
    st = torch.from_numpy(s)
    torch.manual_seed(1)
    eps = torch.randn(B, 128)
    torch.manual_seed(1)                                                  # reparameterize draws the same eps
    with torch.no_grad():
        d, mu, logvar = vail(st)
        dg = gail(st)
    r_vail = np.squeeze(-np.log(1 - 1 / (1 + np.exp(-d.numpy())) + 1e-8)).astype(np.float32)   # gail_TRPO.py:326-327
    r_gail = np.squeeze(-np.log(1 - 1 / (1 + np.exp(-dg.numpy())) + 1e-8)).astype(np.float32)
    g = lambda t: t.detach().numpy()
    np.savez_compressed(
        OUT / "discriminator_ref.npz", s=s, eps=eps.numpy(),
        vail_mean=np.asarray(std.mean, np.float64), vail_std=np.asarray(std.std, np.float64),
        gail_mean=np.asarray(gstd.mean, np.float64), gail_std=np.asarray(gstd.std, np.float64),
        v_w1=g(enc._linears[0].weight), v_b1=g(enc._linears[0].bias),
        v_w2=g(enc._linears[1].weight), v_b2=g(enc._linears[1].bias),
        v_wmu=g(vail.mu_out.weight), v_bmu=g(vail.mu_out.bias),
        v_wlv=g(vail.logvar_out.weight), v_blv=g(vail.logvar_out.bias),
        v_wd=g(dec._linears[0].weight), v_bd=g(dec._linears[0].bias),
        g_w1=g(gail._linears[0].weight), g_b1=g(gail._linears[0].bias),
        g_w2=g(gail._linears[1].weight), g_b2=g(gail._linears[1].bias),
        g_w3=g(gail._linears[2].weight), g_b3=g(gail._linears[2].bias),
        vail_d=d.numpy(), vail_mu=mu.numpy(), vail_logvar=logvar.numpy(), vail_reward=r_vail,
        gail_d=dg.numpy(), gail_reward=r_gail, n_par_vail=n_par_vail, n_par_gail=n_par_gail)
    # Standardizer running sums over three batches
    st2 = nets.Standardizer()
    xs = [rng.normal(1, 2, (n, 5)) for n in (7, 11, 3)]
    for x in xs:
        st2.update_mean_std(x)
    np.savez(OUT / "standardizer_ref.npz", x0=xs[0], x1=xs[1], x2=xs[2], mean=st2.mean, std=st2.std)
    print("discriminator_ref: params", n_par_vail, n_par_gail, "reward", r_vail[:3], r_gail[:3])


# ---------------------------------------------------------------- 5. IL reward + recorded rollouts
def gen_saved_rollouts():
    stub("olympic_mujoco"); stub("olympic_mujoco.utils")
    stub("olympic_mujoco.utils.math", mat2angle_xy=lambda m: 0.0)
    rew = load("ref_reward", "olympic_mujoco/utils/reward.py")
    out = {}
    for name in ("vail_unprocessed_0", "gail_unprocessed_0", "vail_processed_0", "gail_processed_0"):
        d = np.load(REF / "saved_npz" / f"{name}.npz")
        keys = list(d.keys())
        arr = np.stack([d[k] for k in keys], axis=1)                       # [500, 34] spec order
        out[name] = arr[::5].copy()                                        # each sample is repeated 5x
        f = rew.TargetVelocityReward(target_velocity=1.25, x_vel_idx=15)
        out[name + "_reward"] = np.array([f(row[2:], None, None, False) for row in out[name]])
    out["keys"] = np.array(keys)
    np.savez_compressed(OUT / "saved_rollouts_ref.npz", **out)
    r = out["vail_unprocessed_0_reward"]
    print("saved_rollouts_ref: vail reward min/max/mean", r.min(), r.max(), r.mean())


# ---------------------------------------------------------------- 6. RunningMeanStd (rl/envs/normalize.py)
def gen_running_mean_std():
    stub("ray", remote=lambda f: f, wait=None, get=None)
    pkg = stub("rl_envs_pkg"); pkg.__path__ = []
    stub("rl_envs_pkg.wrappers", WrapEnv=object)
    spec = importlib.util.spec_from_file_location("rl_envs_pkg.normalize", str(REF / "rl/envs/normalize.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rl_envs_pkg.normalize"] = mod
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(9)
    xs = [rng.normal(0.5, 1.5, (n, 4)) for n in (3, 4, 5)]
    rms = mod.RunningMeanStd(epsilon=1e-4, shape=(4,))
    for x in xs:
        rms.update(x)
    np.savez(OUT / "running_mean_std_ref.npz", x0=xs[0], x1=xs[1], x2=xs[2], mean=rms.mean, var=rms.var,
             count=rms.count)
    print("running_mean_std_ref: mean", rms.mean)

# ---------------------------------------------------------------- 7. WalkingTask (tasks/walking_task.py)
class _FakeRandom:
    """np.random stand-in for WalkingTask.reset / generate_step_sequence: returns the Philox-contract draws
    u[53..57] (oracle/a3.py reset) in the order the reference consumes them."""

    def __init__(self, u):
        self.u = u

    def choice(self, a, p=None):
        a = list(a)
        if p is not None:                      # mode: STANDING 0.2 / FORWARD 0.8 (walking_task.py:362-364)
            return a[0] if self.u[54] < 0.2 else a[3]
        if len(a) == 2 and a[0] == 0:          # initial phase in {0, period/2} (:354)
            return a[0] if self.u[53] < 0.5 else a[1]
        return a[0] if self.u[55] < 0.5 else a[1]          # step-height sign (:379)

    def uniform(self, lo, hi):
        return lo + (hi - lo) * self.u[56]     # first-step y (:154,158)

    def randint(self, lo, hi):
        return lo if self.u[57] < 0.5 else lo + 1          # c in {2,3} (:164)


class _NpProxy:
    """`np` as seen by the reference module, with `.random` replaced."""

    def __init__(self, rnd):
        self.random = rnd

    def __getattr__(self, k):
        return getattr(np, k)


def gen_a3_task():
    import os
    from olympics_mujoco_b200 import mjcf
    from oracle import a3 as OA
    from oracle import kinematics as K
    from oracle import tf3 as T3
    model = mjcf.load_builtin("stick_figure_a3")
    ids = OA.A3Ids(model)
    # transforms3d stand-in built on the oracle restatement (oracle/tf3.py): pins the task CONTROL FLOW and
    # reward arithmetic to the reference's own code; the tf3 formulas themselves stay "parity unpinned".
    tf = stub("transforms3d")
    tf.euler = stub("transforms3d.euler", euler2quat=lambda a, b, c: T3.euler2quat(a, b, c),
                    quat2euler=lambda q: tuple(float(x) for x in T3.quat2euler(np.asarray(q))),
                    euler2mat=lambda a, b, c: T3.rotz(c),
                    mat2euler=lambda m: tuple(float(x) for x in T3.mat2euler(np.asarray(m))))
    tf.quaternions = stub("transforms3d.quaternions", quat2mat=lambda q: T3.quat2mat(np.asarray(q)),
                          mat2quat=lambda m: T3.mat2quat(m))

    def compose(t, r, z):
        a = np.eye(4)
        a[:3, :3] = np.asarray(r) * np.asarray(z)[None, :]
        a[:3, 3] = t
        return a
    tf.affines = stub("transforms3d.affines", compose=compose)
    rw = load("ref_rewards_a3", "olympic_mujoco/tasks/rewards.py")
    pkg = stub("olympic_mujoco"); pkg.__path__ = []
    tasks = stub("olympic_mujoco.tasks", rewards=rw); tasks.__path__ = []
    sys.modules["olympic_mujoco.tasks.rewards"] = rw
    cwd = os.getcwd()
    os.chdir(REF)                                   # WalkingTask.__init__ opens ./footstep_plans.txt
    try:
        wt = load("ref_walking_task", "olympic_mujoco/tasks/walking_task.py")
    finally:
        os.chdir(cwd)

    class Con:
        def __init__(self, z):
            self.pos = np.array([0.0, 0.0, z])

    class Client:
        """MujocoRobotInterface getters (interfaces/mujoco_robot_interface.py:239-399) over the oracle's FK; the
        contact-solver outputs are injected per step."""

        def __init__(self):
            self.model = types.SimpleNamespace(geom=lambda name: types.SimpleNamespace(pos=np.zeros(3)))

        def set(self, qpos, qvel, contact):
            self.fk = K.forward(model, qpos[None], qvel[None])
            self.c = contact

        def get_robot_mass(self):
            return model.total_mass

        def get_object_xpos_by_name(self, name, typ):
            if typ == "OBJ_BODY":
                return self.fk["xpos"][0, model.body_id(name)]
            return self.fk["site_xpos"][0, model.site_id(name)]

        def get_object_xquat_by_name(self, name, typ):
            if typ == "OBJ_BODY":
                return self.fk["xquat"][0, model.body_id(name)]
            return T3.mat2quat(self.fk["site_xmat"][0, model.site_id(name)])

        def _vel(self, b):
            v = K.mj_objectVelocity_xbody(model, self.fk["xpos"], self.fk["subtree_com"], self.fk["cvel"], b)[0]
            return [v[3:6], v[0:3]]

        def get_lfoot_body_vel(self):
            return self._vel(ids.lfoot)

        def get_rfoot_body_vel(self):
            return self._vel(ids.rfoot)

        def get_lfoot_body_pos(self):
            return self.fk["xpos"][0, ids.lfoot]

        def get_rfoot_body_pos(self):
            return self.fk["xpos"][0, ids.rfoot]

        def get_lfoot_grf(self):
            return self.c.l_grf

        def get_rfoot_grf(self):
            return self.c.r_grf

        def check_bad_collisions(self):
            return self.c.bad_collision

        def check_rfoot_floor_collision(self):
            return self.c.foot_contact

        def check_lfoot_floor_collision(self):
            return False

        def get_rfoot_floor_contacts(self):
            return [(0, Con(self.c.min_z)), (1, Con(self.c.min_z + 0.003))] if self.c.foot_contact else []

        def get_lfoot_floor_contacts(self):
            return []

    seed, n_env, T = 25, 6, 330
    rng = np.random.default_rng(5)
    rec = {k: [] for k in ("qpos", "qvel", "contact", "terms", "done", "ints", "goal", "obs")}
    reset_rec = {k: [] for k in ("qpos", "qvel", "ints", "sequence", "obs", "u")}
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    for e in range(n_env):
        u = OA.reset_uniforms(seed, e, 0)
        qpos0, qvel0, ts0, obs0 = OA.reset(model, seed, e, 0, iteration_count=7000)
        client = Client()
        client.set(qpos0, qvel0, OA.Contact())
        os.chdir(REF)
        try:
            task = wt.WalkingTask(client=client, dt=0.025, neutral_foot_orient=np.array([1, 0, 0, 0]),
                                  root_body="torso", lfoot_body="left_foot", rfoot_body="right_foot", head_body="head")
        finally:
            os.chdir(cwd)
        task._goal_height_ref, task._total_duration = 0.80, 1.1          # StickFigureA3.py:110-113
        task._swing_duration, task._stance_duration = 0.75, 0.35
        wt.np = _NpProxy(_FakeRandom(u))
        task.reset(iter_count=7000)
        wt.np = np
        seq = np.zeros((OA.MAX_STEPS, 4))
        seq[:len(task.sequence)] = np.array(task.sequence)
        mode = OA.STANDING if task.mode == wt.WalkModes.STANDING else OA.FORWARD
        reset_rec["qpos"].append(qpos0); reset_rec["qvel"].append(qvel0); reset_rec["u"].append(u)
        reset_rec["ints"].append([task._phase, task.t1, task.t2, task.target_reached_frames, mode,
                                  len(task.sequence), int(task.target_reached)])
        reset_rec["sequence"].append(seq); reset_rec["obs"].append(obs0)
        # a dataset-shaped walk: the root advances along its heading so that feet pass over the targets
        yaw = T3.quat2euler(qpos0[3:7])[2]
        speed = rng.uniform(0.1, 0.2) * 0.025
        q, v = qpos0.copy(), qvel0.copy()
        for k in ("qpos", "qvel", "contact", "terms", "done", "ints", "goal", "obs"):
            rec[k].append([])
        for t in range(T):
            q = q.copy()
            q[0] += speed * np.cos(yaw); q[1] += speed * np.sin(yaw)
            q[2] = 1.34 + 0.02 * np.sin(0.2 * t) - (0.75 if (e == 3 and t > 300) else 0.0)     # env 3 falls
            q[7:] = np.clip(q[7:] + rng.normal(0, 0.01, 18), OA.init_qpos()[7:] - 0.5, OA.init_qpos()[7:] + 0.5)
            dq = rng.normal(0, 0.002, 4); dq[3] += 0.002 * np.sin(0.05 * t)
            q[3:7] = q[3:7] + dq                                           # un-normalised on purpose
            v = np.clip(v + rng.normal(0, 0.2, 24), -10, 10)
            if t % 37 == 5:
                v[6:18] *= 0.01                                            # near-still feet: velocity clock saturates
            q, v = f32(q), f32(v)
            fmax = model.total_mass * 9.8 * 0.5
            con = OA.Contact(l_grf=float(f32(rng.uniform(0, 2 * fmax))), r_grf=float(f32(rng.uniform(0, 2 * fmax))),
                             min_z=float(f32(rng.uniform(-0.01, 0.01))), foot_contact=bool(rng.random() < 0.7),
                             bad_collision=bool(rng.random() < 0.02))
            client.set(q, v, con)
            task.step()
            terms = task.calc_reward(None, None, None)
            done = task.done()
            rec["qpos"][-1].append(q); rec["qvel"][-1].append(v)
            rec["contact"][-1].append([con.l_grf, con.r_grf, con.min_z, float(con.foot_contact), float(con.bad_collision)])
            rec["terms"][-1].append([float(x) for x in terms.values()])
            rec["done"][-1].append(bool(done))
            rec["ints"][-1].append([task._phase, task.t1, task.t2, task.target_reached_frames, mode, len(task.sequence),
                                    int(task.target_reached)])
            rec["goal"][-1].append(np.concatenate([task._goal_steps_x, task._goal_steps_y, task._goal_steps_z,
                                                   task._goal_steps_theta]))
    out = {"step_" + k: np.array(v) for k, v in rec.items() if k != "obs"}
    for k in ("step_qpos", "step_qvel", "step_contact"):          # fp32-representable by construction
        assert np.array_equal(out[k], out[k].astype(np.float32).astype(np.float64))
        out[k] = out[k].astype(np.float32)
    out.update({"reset_" + k: np.array(v) for k, v in reset_rec.items()})
    np.savez_compressed(OUT / "a3_task_ref.npz", seed=seed, iteration_count=7000, **out)
    ints = out["step_ints"]
    print("a3_task_ref: steps", out["step_terms"].shape, "max t1", ints[..., 1].max(), "done frac",
          out["step_done"].mean(), "modes", out["reset_ints"][:, 4])

# ---------------------------------------------------------------- 8. mirror symmetry (rl/envs/wrappers.py)
def gen_mirror():
    import torch
    wr = load("ref_wrappers", "rl/envs/wrappers.py")
    base_mir_obs = [0.1, -1, 2, -3, -4, 5, -6, 13, -14, -15, 16, -17, 18, 7, -8, -9, 10, -11, 12,
                    25, -26, -27, 28, -29, 30, 19, -20, -21, 22, -23, 24]                 # StickFigureA3.py:118-125
    mirrored_obs = base_mir_obs + [len(base_mir_obs) + i for i in range(10)]
    mirrored_acts = [6, -7, -8, 9, -10, 11, 0.1, -1, -2, 3, -4, 5]
    clock_inds = [31, 32]
    env = wr.SymmetricEnv(lambda: types.SimpleNamespace(base_obs_len=41), mirrored_obs=mirrored_obs,
                          mirrored_act=mirrored_acts, clock_inds=clock_inds)
    rng = np.random.default_rng(9)
    obs = rng.normal(0, 1, (37, 41)).astype(np.float32)
    ph = rng.integers(0, 88, 37)
    obs[:, 31], obs[:, 32] = np.sin(2 * np.pi * ph / 88), np.cos(2 * np.pi * ph / 88)
    act = rng.normal(0, 1, (37, 12)).astype(np.float32)
    to, ta = torch.from_numpy(obs), torch.from_numpy(act)
    np.savez(OUT / "mirror_ref.npz", mirrored_obs=np.array(mirrored_obs), mirrored_acts=np.array(mirrored_acts),
             clock_inds=np.array(clock_inds), obs=obs, act=act, mirror_obs=env.mirror_observation(to).numpy(),
             mirror_act=env.mirror_action(ta).numpy(), mirror_clock_obs=env.mirror_clock_observation(to).numpy())
    print("mirror_ref: obs", obs.shape)


# ---------------------------------------------------------------- 9. discriminator fit losses (imitation_lib/utils/math.py)
def gen_disc_loss():
    """GailDiscriminatorLoss / VDBLoss exactly as GAIL._fit_discriminator and _discriminator_logging use them
    (gail_TRPO.py:167-258, vail_TRPO.py:23-33): [policy; expert] logits, 0/1 and noisy targets, beta update."""
    import torch
    stub("mushroom_rl"); stub("mushroom_rl.utils")
    stub("mushroom_rl.utils.angles", euler_to_quat=None)
    stub("mushroom_rl.utils.torch", to_float_tensor=lambda x, *a: torch.as_tensor(x, dtype=torch.float32))
    m = load("ref_il_math", "imitation_lib/utils/math.py")
    rng = np.random.default_rng(11)
    B = 96
    logits = np.concatenate([rng.normal(-1.0, 2.5, B), rng.normal(1.5, 2.5, B)]).astype(np.float32)[:, None]
    logits[3, 0], logits[B + 5, 0] = -30.0, 40.0                       # saturated samples
    t01 = np.concatenate([np.zeros((B, 1)), np.ones((B, 1))]).astype(np.float32)
    tnoisy = np.concatenate([rng.uniform(0.01, 0.10, (B, 1)), rng.uniform(0.80, 0.99, (B, 1))]).astype(np.float32)
    mu = rng.normal(0, 0.7, (2 * B, 128)).astype(np.float32)
    logvar = rng.normal(-0.5, 0.8, (2 * B, 128)).astype(np.float32)
    gl = m.GailDiscriminatorLoss(entcoeff=1e-3)
    tl = torch.from_numpy(logits).requires_grad_(True)
    loss01 = gl(tl, torch.from_numpy(t01))
    loss01.backward()
    grad01 = tl.grad.numpy().copy()
    loss_noisy = gl(torch.from_numpy(logits), torch.from_numpy(tnoisy)).item()
    ent = gl.logit_bernoulli_entropy(torch.from_numpy(logits)).numpy()
    vl = m.VDBLoss(info_constraint=0.5, lr_beta=1e-5)
    kl = vl.kl_divergence(torch.from_numpy(mu), torch.from_numpy(logvar)).numpy()
    betas, vlosses = [vl._beta], []
    for _ in range(3):                                                  # three fits: beta moves
        vlosses.append(vl((torch.from_numpy(logits), torch.from_numpy(mu), torch.from_numpy(logvar)), torch.from_numpy(t01)).item())
        betas.append(float(vl._beta))
    np.savez(OUT / "disc_loss_ref.npz", logits=logits, t01=t01, tnoisy=tnoisy, mu=mu, logvar=logvar, gail_loss01=loss01.item(),
             gail_grad01=grad01, gail_loss_noisy=loss_noisy, ent=ent, kl=kl, vdb_losses=np.array(vlosses), betas=np.array(betas),
             info_constraint=0.5, lr_beta=1e-5, entcoeff=1e-3)
    print("disc_loss_ref: gail", loss01.item(), loss_noisy, "vdb", vlosses, "beta", betas)


# ---------------------------------------------------------------- 10. PPO minibatch losses (rl/algos/ppo.py:231-282)
def gen_ppo_loss():
    """PPO.update_policy of the reference, executed with small stand-in actor / critic modules (a linear-Gaussian policy
    with a state-independent std, a linear critic) and the reference's own SymmetricEnv mirror functions: the inputs the
    loss kernel consumes (log-probs, advantages, mask, values, returns, entropies, deterministic and mirrored actions) and
    the six numbers update_policy returns, plus autograd gradients with respect to log_probs and values."""
    import torch
    stub("ray", remote=lambda f: f)
    stub("matplotlib"); stub("matplotlib.pyplot")
    stub("rl"); stub("rl.envs", WrapEnv=object)
    ppo = load("ref_ppo2", "rl/algos/ppo.py")
    wr = load("ref_wrappers2", "rl/envs/wrappers.py")
    base_mir_obs = [0.1, -1, 2, -3, -4, 5, -6, 13, -14, -15, 16, -17, 18, 7, -8, -9, 10, -11, 12,
                    25, -26, -27, 28, -29, 30, 19, -20, -21, 22, -23, 24]
    mirrored_obs = base_mir_obs + [len(base_mir_obs) + i for i in range(10)]
    mirrored_acts = [6, -7, -8, 9, -10, 11, 0.1, -1, -2, 3, -4, 5]
    env = wr.SymmetricEnv(lambda: types.SimpleNamespace(base_obs_len=41), mirrored_obs=mirrored_obs,
                          mirrored_act=mirrored_acts, clock_inds=[31, 32])
    torch.manual_seed(3)
    B, nobs, nu = 200, 41, 12

    class Actor(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(nobs, nu)
            self.logstd = torch.nn.Parameter(torch.full((nu,), -0.7))
        def forward(self, x):
            return torch.tanh(self.lin(x))
        def distribution(self, x):
            return torch.distributions.Normal(self.forward(x), self.logstd.exp())

    actor, old, critic = Actor(), Actor(), torch.nn.Linear(nobs, 1)
    with torch.no_grad():
        for p_new, p_old in zip(actor.parameters(), old.parameters()):
            p_old.copy_(p_new + 0.004 * torch.randn_like(p_new))
    algo = ppo.PPO.__new__(ppo.PPO)
    algo.policy, algo.critic, algo.old_policy, algo.clip, algo.vf_coeff = actor, critic, old, 0.2, 0.5
    obs = torch.randn(B, nobs)
    ph = torch.randint(0, 88, (B,)).double()
    obs[:, 31], obs[:, 32] = torch.sin(2 * np.pi * ph / 88).float(), torch.cos(2 * np.pi * ph / 88).float()
    act = old.distribution(obs).sample()
    ret, adv = torch.randn(B, 1), torch.randn(B, 1)
    mask = (torch.rand(B, 1) > 0.1).float()
    # what update_policy computes internally, recorded as the kernel's inputs (leaf copies for the gradients)
    pdf = actor.distribution(obs)
    logp = pdf.log_prob(act).sum(-1, keepdim=True).detach().requires_grad_(True)
    old_logp = old.distribution(obs).log_prob(act).sum(-1, keepdim=True).detach()
    values = critic(obs).detach().requires_grad_(True)
    out = algo.update_policy(obs, act, ret, adv, mask, mirror_observation=env.mirror_clock_observation,
                             mirror_action=env.mirror_action)
    actor_loss, entropy_penalty, critic_loss, approx_kl, mirror_loss, clip_fraction = out
    # the same expressions on the leaf copies -> gradients with respect to log_probs and values (ppo.py:244-256)
    ratio = (logp - old_logp).exp()
    a2 = -torch.min(ratio * adv * mask, ratio.clamp(1 - 0.2, 1 + 0.2) * adv * mask).mean()
    c2 = 0.5 * torch.nn.functional.mse_loss(ret, values)
    (a2 + c2).backward()
    assert abs(a2.item() - actor_loss.item()) < 1e-7 and abs(c2.item() - critic_loss.item()) < 1e-7
    with torch.no_grad():
        det = actor(obs)
        mir_raw = actor(env.mirror_clock_observation(obs))                 # policy(mirror(obs)) BEFORE mirror_action
        ent = pdf.entropy()
    np.savez(OUT / "ppo_loss_ref.npz", logp=logp.detach().numpy(), old_logp=old_logp.numpy(), adv=adv.numpy(), mask=mask.numpy(),
             values=values.detach().numpy(), returns=ret.numpy(), entropy=ent.numpy(), det_actions=det.numpy(),
             mirror_raw=mir_raw.numpy(), mirrored_acts=np.array(mirrored_acts), clip=0.2, vf_coeff=0.5,
             actor_loss=actor_loss.item(), entropy_penalty=entropy_penalty.item(), critic_loss=critic_loss.item(),
             approx_kl=approx_kl.item(), mirror_loss=mirror_loss.item(), clip_fraction=clip_fraction,
             dlogp=logp.grad.numpy(), dvalues=values.grad.numpy())
    print("ppo_loss_ref:", [float(x) if not hasattr(x, "item") else x.item() for x in out])


if __name__ == "__main__":
    gen_trajectory()
    gen_phase_clock()
    gen_ppo()
    gen_networks()
    gen_saved_rollouts()
    gen_running_mean_std()
    gen_a3_task()
    gen_mirror()
    gen_disc_loss()
    gen_ppo_loss()
