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


if __name__ == "__main__":
    gen_trajectory()
    gen_phase_clock()
    gen_ppo()
    gen_networks()
    gen_saved_rollouts()
    gen_running_mean_std()
