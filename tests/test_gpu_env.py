"""GPU: the reference-facing Python API (LocoEnvBase.make / reset / step / play_trajectory_from_velocity,
ObservationHelper, Trajectory) on top of the kernels, checked against the oracle."""
import numpy as np
import pytest

from conftest import GOLDEN, assert_close

pytestmark = pytest.mark.gpu


def _table():
    tab = np.load(GOLDEN / "trajectory_ref.npz")["table"].copy()
    tab[2:] = tab[2:].astype(np.float32).astype(np.float64)
    return tab


def _make(n_envs, **kw):
    import olympics_mujoco_b200 as om
    return om.LocoEnvBase.make("UnitreeH1.walk.real", n_envs=n_envs, traj_params=dict(table=_table()), seed=77, **kw)


def test_make_reset_shapes_and_spaces():
    from oracle import h1 as OH
    env = _make(16)
    assert env.info.observation_space.shape == (32,) and env.info.action_space.shape == (11,)
    assert env.info.gamma == 0.99 and env.info.horizon == 1000 and abs(env.dt - 0.01) < 1e-12
    assert env.get_all_observation_keys() == OH.keys(env._model)
    assert env.get_obs_idx("dq_pelvis_tx") == [15] and env._len_qpos_qvel() == (17, 17)
    assert list(env.get_kinematic_obs_mask()) == list(range(32))
    obs = env.reset()
    assert tuple(obs.shape) == (16, 32)
    # reset puts every env on its (Philox) trajectory sample, x and y re-centred to 0
    tr = env.trajectories
    tab = _table()
    tn, sn = tr.traj_no.cpu().numpy(), tr.subtraj_step_no.cpu().numpy()
    ref = np.stack([tab[2:, a, b] for a, b in zip(tn, sn)])
    assert np.array_equal(obs.cpu().numpy(), ref.astype(np.float32))
    assert float(env.data.qpos[:2].abs().max()) == 0.0
    single = _make(None)
    assert tuple(single.reset().shape) == (32,)


def test_step_with_attached_dynamics_matches_oracle():
    import torch
    from oracle import h1 as OH
    env = _make(32)
    obs0 = env.reset().clone()
    g = torch.Generator(device="cuda").manual_seed(0)

    def dynamics(e, ctrl):                       # stand-in physics: a deterministic perturbation of the state
        q = e.data.qpos.t() + 0.01 * torch.randn((e.n_envs, 17), device="cuda", generator=g)
        v = e.data.qvel.t() + 0.1 * torch.randn((e.n_envs, 17), device="cuda", generator=g)
        return q, v

    with pytest.raises(RuntimeError, match="dynamics"):
        env.step(torch.zeros((32, 11), device="cuda"))
    env.attach_dynamics(dynamics)
    prev = obs0
    for _ in range(3):
        obs, reward, absorbing, info = env.step(torch.zeros((32, 11), device="cuda"))
        q = env.data.qpos.t().double().cpu().numpy()
        v = env.data.qvel.t().double().cpu().numpy()
        ref = OH.step(env._model, q, v, prev.double().cpu().numpy())
        assert np.array_equal(obs.cpu().numpy(), ref["obs"].astype(np.float32))
        assert np.array_equal(absorbing.cpu().numpy(), ref["absorbing"])
        assert_close(reward.cpu().numpy(), ref["reward"], "reward")
        assert_close(env.data.xpos.t().cpu().numpy().reshape(32, 21, 3), ref["xpos"], "xpos")
        assert_close(env.data.cvel.t().cpu().numpy().reshape(32, 21, 6), ref["cvel"], "cvel")
        # reward(state, ...) called by hand agrees with the fused kernel
        assert_close(env.reward(prev, None, obs, absorbing).cpu().numpy(), ref["reward"], "reward api")
        assert np.array_equal(env.is_absorbing(obs).cpu().numpy(), ref["absorbing"])
        prev = obs.clone()


def test_play_trajectory_from_velocity_api_matches_oracle():
    from olympics_mujoco_b200 import kernels as Kn
    from oracle import h1 as OH
    n, T = 8, 90
    env = _make(n)
    out = env.play_trajectory_from_velocity(n_episodes=2, n_steps_per_episode=T, render=False)
    xpos = Kn.env_major(out["xpos"], 21, 3).cpu().numpy()           # [T, n, 21, 3] view of the SoA buffer
    for e in range(n):
        ref = OH.play_trajectory_from_velocity(env._model, _table(), 2, T, seed=77, env_id=e)
        assert np.array_equal(out["step_no_t"][:, e].cpu().numpy(), ref["step_no"][T:])
        assert_close(xpos[:, e], ref["xpos"][T:], "xpos (second episode)")
        assert np.array_equal(out["obs"][:, :, e].cpu().numpy(), ref["obs"][T:].astype(np.float32))
    with pytest.raises(NotImplementedError):
        env.play_trajectory_from_velocity(1, 1, render=True)


def test_play_trajectory_and_dataset_and_trajectory_api():
    import torch
    from oracle import h1 as OH
    env = _make(4)
    res = env.play_trajectory(n_episodes=1, n_steps_per_episode=60, render=False)           # fused kernel
    assert tuple(res["obs"].shape) == (60, 32, 4) and not bool(res["fallen"].any())
    from olympics_mujoco_b200.environments.loco_env_base import LocoEnvBase as Base
    loop = Base.play_trajectory(_make(4), n_episodes=1, n_steps_per_episode=60, render=False)  # per-step kernels, same seed
    assert torch.equal(res["obs"][-1].t(), loop["obs"]) and not bool(loop["has_fallen"].any())
    ds = env.create_dataset()
    z = np.load(GOLDEN / "trajectory_ref.npz")
    assert ds["states"].shape == (149, 32) and np.array_equal(ds["last"], z["ds_last"])
    # a dataset with a terminal state is rejected (loco_env_base.py:949-957)
    env._dataset = None
    bad = env.trajectories.trajectories[2]
    bad[0, 10] = 5.0                                               # q_pelvis_ty far above the 0.1 threshold
    with pytest.raises(ValueError, match="terminal states"):
        env.create_dataset()
    # Trajectory object: reset(substep, traj) / current / next / None at the end (single env)
    single = _make(None)
    tr = single.trajectories
    s = tr.reset_trajectory(substep_no=47, traj_no=1)
    tab = _table()
    assert_close(s[0, 2:].cpu().numpy(), tab[2:, 1, 47], "forced reset sample", rtol=1e-7, atol=1e-7)
    assert tr.get_next_sample() is not None and tr.get_next_sample() is not None       # 48, 49
    assert tr.get_next_sample() is None and int(tr.subtraj_step_no[0]) == 50           # trajectory_length


def test_observation_helper_generic_spec():
    """Upstream ObservationHelper semantics on body/site entries (BODY_POS, BODY_ROT, BODY_VEL, SITE_POS)."""
    import torch
    from olympics_mujoco_b200.observation_helper import BatchedData, ObservationHelper, ObservationType
    from olympics_mujoco_b200 import kernels as Kn, mjcf
    from oracle import kinematics as K
    model = mjcf.load_builtin("unitree_h1")
    n = 6
    data = BatchedData(model, n)
    spec = [("q_knee", "knee_angle_l", ObservationType.JOINT_POS), ("torso_pos", "torso_link", ObservationType.BODY_POS),
            ("torso_rot", "torso_link", ObservationType.BODY_ROT), ("ankle_vel", "left_ankle_link", ObservationType.BODY_VEL),
            ("imu", "imu", ObservationType.SITE_POS), ("dq_knee", "knee_angle_l", ObservationType.JOINT_VEL)]
    oh = ObservationHelper(spec, model, data)
    assert oh.obs_idx_map["torso_rot"] == [4, 5, 6, 7] and oh.joint_pos_idx == [0] and oh.joint_vel_idx == [17]
    assert oh.obs_low[0] == -0.26 and oh.obs_high[0] == 2.05 and np.isinf(oh.obs_low[1])
    rng = np.random.default_rng(0)
    q = rng.normal(0, 0.3, (n, 17)).astype(np.float32); v = rng.normal(0, 1, (n, 17)).astype(np.float32)
    data.qpos.copy_(Kn.to_soa(q)); data.qvel.copy_(Kn.to_soa(v))
    dm = Kn.DeviceModel(model)
    Kn.fk(dm, data.qpos, data.qvel, out=dict(xpos=data.xpos, xquat=data.xquat, cvel=data.cvel, site_xpos=data.site_xpos,
                                             site_xmat=data.site_xmat, subtree_com=data.subtree_com))
    obs = oh._build_obs(data).cpu().numpy()
    ref = K.forward(model, q.astype(np.float64), v.astype(np.float64))
    t, a = model.body_id("torso_link"), model.body_id("left_ankle_link")
    exp = np.concatenate([q[:, [9]], ref["xpos"][:, t], ref["xquat"][:, t], ref["cvel"][:, a], ref["site_xpos"][:, 0],
                          v[:, [9]]], axis=1)
    assert_close(obs, exp, "generic obs")
    assert_close(oh.get_from_obs(torch.as_tensor(obs), "imu").numpy(), ref["site_xpos"][:, 0], "get_from_obs")
    # _modify_data is the inverse gather: a fresh data object written from the observation rebuilds the same observation
    data2 = BatchedData(model, n)
    oh._modify_data(data2, torch.as_tensor(obs))
    assert np.array_equal(oh._build_obs(data2).cpu().numpy(), obs)
    assert np.array_equal(data2.qpos[9].cpu().numpy(), q[:, 9]) and np.array_equal(data2.qvel[9].cpu().numpy(), v[:, 9])


def test_set_sim_state_kernel_matches_named_scatter():
    """UnitreeH1.set_sim_state (one kernel) == the reference's per-key scatter (LocoEnvBase.set_sim_state)."""
    import torch
    from olympics_mujoco_b200 import LocoEnvBase
    from olympics_mujoco_b200.environments.loco_env_base import LocoEnvBase as Base
    env = LocoEnvBase.make("UnitreeH1.walk.real", n_envs=19, seed=2)
    rng = np.random.default_rng(0)
    sample = torch.as_tensor(rng.normal(0, 1, (19, 34)).astype(np.float32), device="cuda")
    env.set_sim_state(sample)
    q1, v1 = env.data.qpos.clone(), env.data.qvel.clone()
    env.data.qpos.zero_(); env.data.qvel.zero_()
    Base.set_sim_state(env, sample)
    assert torch.equal(q1, env.data.qpos) and torch.equal(v1, env.data.qvel)
    from oracle import h1 as OH
    qo, vo = OH.set_sim_state(env._model, sample.cpu().numpy().astype(np.float64))
    assert np.array_equal(q1.cpu().numpy().T, qo.astype(np.float32)) and np.array_equal(v1.cpu().numpy().T, vo.astype(np.float32))


def test_playback_resets_at_the_start_of_every_call_like_the_reference():
    """loco_env_base.py:481: EVERY call of play_trajectory_from_velocity begins with reset() (a fresh draw), and every
    episode ends with one.  Two calls on one env object == the oracle called twice on one trajectory state; the second
    call's reset runs inside the playback kernel.  continue_episode=True (extension) carries the state on instead: two
    one-episode calls == one two-episode call."""
    import torch
    from oracle import h1 as OH
    n, T = 6, 70
    for chunk_note, steps in (("sequential kernel", T), ("time-parallel kernel", 400)):
        env = _make(n)
        a = {k: v.clone() for k, v in env.play_trajectory_from_velocity(1, steps, render=False).items()}
        b = env.play_trajectory_from_velocity(1, steps, render=False)
        for e in range(n):
            r1 = OH.play_trajectory_from_velocity(env._model, _table(), 1, steps, seed=77, env_id=e, record_fk=(e == 0))
            r2 = OH.play_trajectory_from_velocity(env._model, _table(), 1, steps, seed=77, env_id=e, record_fk=(e == 0),
                                                  traj_state=r1["final"]["traj_state"])
            for out, ref in ((a, r1), (b, r2)):
                assert np.array_equal(out["traj_no_t"][:, e].cpu().numpy(), ref["traj_no"]), chunk_note
                assert np.array_equal(out["step_no_t"][:, e].cpu().numpy(), ref["step_no"]), chunk_note
                assert np.array_equal(out["obs"][:, :, e].cpu().numpy(), ref["obs"].astype(np.float32))
                assert_close(out["reward"][:, e].cpu().numpy(), ref["reward"], "reward")
                if e == 0:
                    assert_close(out["xpos"][:, :, e].cpu().numpy(), ref["xpos"].reshape(steps, 63), "xpos")
            assert int(env.trajectories.device_state.reset_count[e]) == r2["final"]["reset_count"]
        # the env is left in the state of the last reset(): its observation is the pending sample
        tr = env.trajectories.device_state
        tab = _table()
        ref = np.stack([tab[2:, int(x), int(y)] for x, y in zip(tr.traj_no.cpu(), tr.step_no.cpu())]).astype(np.float32)
        assert np.array_equal(env._obs.cpu().numpy(), ref)
    env1, env2 = _make(n), _make(n)
    two = env1.play_trajectory_from_velocity(2, T, render=False)
    env2.play_trajectory_from_velocity(1, T, render=False)
    cont = env2.play_trajectory_from_velocity(1, T, render=False, continue_episode=True)
    for k in ("traj_no_t", "step_no_t", "obs", "xpos", "reward"):
        assert torch.equal(two[k], cont[k]), k


def test_playback_fused_observation_moments():
    """S1 fused into the playback kernels: obs_moments == om_moments over the emitted observation buffer == float64 sums."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    for n, T in ((40, 33), (300, 260)):                       # sequential-in-time kernel, time-parallel kernel
        env = _make(n)
        mom = torch.zeros(65, dtype=torch.float64, device="cuda")
        out = env.play_trajectory_from_velocity(1, T, render=False, obs_moments=mom)
        ref = Kn.moments(out["obs"])
        torch.cuda.synchronize()
        o = out["obs"].double()
        exact = torch.cat([o.sum(dim=(0, 2)), (o * o).sum(dim=(0, 2)), torch.tensor([float(n * T)], dtype=torch.float64, device="cuda")])
        assert float(mom[64]) == n * T
        assert_close(mom.cpu().numpy(), exact.cpu().numpy(), "fused moments vs float64 sums", rtol=1e-12, atol=1e-9)
        assert_close(mom.cpu().numpy(), ref.cpu().numpy(), "fused moments vs om_moments", rtol=1e-12, atol=1e-9)
        out2 = env.play_trajectory_from_velocity(1, T, render=False, obs_moments=mom)      # accumulates
        assert float(mom[64]) == 2 * n * T


def test_fused_live_step_matches_three_kernel_path_and_oracle():
    """om_h1_live_step (one kernel: next sample / wrap reset + set_sim_state + FK + obs + has_fallen + reward) against
    (a) om_traj_next -> om_set_sim_state -> om_h1_step bit for bit and (b) the oracle; step_graph replays == eager steps."""
    import torch
    from olympics_mujoco_b200 import kernels as Kn
    from oracle import h1 as OH
    from oracle.trajectory import TrajectoryState
    n, steps = 48, 130                                        # table T = 50: every env wraps (and resets) at least twice
    env, env_b = _make(n), _make(n)
    obs0 = env.reset().clone()
    env_b.reset()
    trb = env_b.trajectories.device_state
    sample = Kn.soa(34, n)
    pxv = env_b._prev_x_vel.clone()
    oracle_states = []
    for e in range(3):
        ts = TrajectoryState(_table(), seed=77, env_id=e)
        ts.reset_trajectory()
        oracle_states.append((ts, OH.create_observation(ts.get_current_sample())))
    for s in range(steps):
        obs, reward, absorbing, _ = env.step_trajectory()
        trb.next(sample=sample)
        Kn.set_sim_state(env_b._dm, env_b._spec, sample, env_b.data.qpos, env_b.data.qvel)
        ref = Kn.h1_step(env_b._dm, env_b._spec, env_b.data.qpos, env_b.data.qvel, pxv)
        assert torch.equal(obs, ref["obs"].t()) and torch.equal(reward, ref["reward"]), s
        assert torch.equal(absorbing, ref["absorbing"].bool())
        for k in ("qpos", "qvel"):
            assert torch.equal(getattr(env.data, k), getattr(env_b.data, k)), (k, s)
        for k in ("xpos", "xquat", "site_xpos", "cvel"):     # same generated FK inlined into two kernels (the three-kernel
            # path may take the three-threads-per-env variant): equal to fp32 rounding, not necessarily bit for bit
            assert_close(getattr(env.data, k).cpu().numpy(), ref[k].cpu().numpy(), f"{k} step {s}", rtol=2e-6, atol=2e-6)
        pxv = ref["obs"][15].clone()
        for e, (ts, prev) in enumerate(oracle_states):
            smp = ts.get_next_sample()
            if smp is None:
                smp = ts.reset_trajectory()
            q, v = OH.set_sim_state(env._model, smp.astype(np.float32).astype(np.float64))
            o = OH.step(env._model, q[None], v[None], prev[None])
            assert np.array_equal(obs[e].cpu().numpy(), o["obs"][0].astype(np.float32))
            assert_close(float(reward[e]), o["reward"][0], "reward")
            assert_close(env.data.xpos[:, e].cpu().numpy().reshape(21, 3), o["xpos"][0], "xpos")
            assert int(env.trajectories.device_state.step_no[e]) == ts.step_no
            oracle_states[e] = (ts, o["obs"][0])
    # graph: 2 replays x 16 captured steps == 32 eager steps from the same state
    env_g, env_e = _make(n), _make(n)
    env_g.reset(); env_e.reset()
    replay, bufs = env_g.step_graph(16, want=("obs", "reward", "absorbing", "xpos", "wrapped"))
    for r in range(2):
        replay()
        torch.cuda.synchronize()
        for t in range(16):
            obs, reward, absorbing, _ = env_e.step_trajectory()
            assert torch.equal(bufs["obs"][t].t(), obs) and torch.equal(bufs["reward"][t], reward)
            assert torch.equal(bufs["xpos"][t], env_e.data.xpos)
    assert torch.equal(env_g.trajectories.device_state.step_no, env_e.trajectories.device_state.step_no)


def test_step_with_soa_dynamics_needs_no_transposes():
    import torch
    env_a, env_s = _make(20), _make(20)
    env_a.reset(); env_s.reset()
    ga, gs = (torch.Generator(device="cuda").manual_seed(5) for _ in range(2))

    def dyn_aos(e, ctrl):
        dq = 0.01 * torch.randn((17, e.n_envs), device="cuda", generator=ga)
        return (e.data.qpos + dq).t(), (e.data.qvel + 10 * dq).t()

    def dyn_soa(e, ctrl):
        assert tuple(ctrl.shape) == (11, e.n_envs)
        dq = 0.01 * torch.randn((17, e.n_envs), device="cuda", generator=gs)
        return e.data.qpos + dq, e.data.qvel + 10 * dq

    env_a.attach_dynamics(dyn_aos)
    env_s.attach_dynamics(dyn_soa, soa=True)
    act = torch.rand((20, 11), device="cuda") * 2 - 1
    for _ in range(3):
        oa, ra, aa, _ = env_a.step(act)
        os_, rs, as_, _ = env_s.step(act.t().contiguous())
        assert torch.equal(oa, os_) and torch.equal(ra, rs) and torch.equal(aa, as_)
        assert torch.equal(env_a.data.xpos, env_s.data.xpos)


def _variant_table(model):
    """The golden 34-key table restricted to the keys this model's spec keeps (e.g. without the back joint)."""
    from oracle import h1 as OH
    full = ["q_" + j for j in OH._SPEC_JOINTS] + ["dq_" + j for j in OH._SPEC_JOINTS]
    full = [k for k in full if not any(a in k for a in OH.ARM_JOINTS)]
    keep = [full.index(k) for k in OH.keys(model)]
    return _table()[keep]


@pytest.mark.parametrize("kw", [dict(disable_back_joint=True), dict(hold_weight=True, weight_mass=5.0), dict(use_foot_forces=True)])
def test_unitree_h1_variants_step_and_generic_playback(kw):
    """UnitreeH1.py:38-111 variants: the back joint removed, a carried weight, foot forces in the observation.  step() on
    attached dynamics against the oracle on the variant's model; play_trajectory_from_velocity through the per-step
    kernels of the base class against the oracle's playback."""
    import torch
    import olympics_mujoco_b200 as om
    from oracle import h1 as OH
    from oracle import kinematics as K
    n = 12
    from olympics_mujoco_b200 import mjcf
    table = _variant_table(mjcf.unitree_h1_variant(**{k: v for k, v in kw.items() if k != "use_foot_forces"}))
    env = om.LocoEnvBase.make("UnitreeH1.walk.real", n_envs=n, traj_params=dict(table=table), seed=31, **kw)
    model = env._model
    nq = model.nq
    nobs = 2 * nq - 2 + (6 if kw.get("use_foot_forces") else 0)
    assert env.info.observation_space.shape == (nobs,) and not env._dm.specialised or kw.get("use_foot_forces")
    if kw.get("hold_weight"):
        assert model.body_names[-1] == "weight" and abs(model.total_mass - 61.437) < 1e-9
    obs0 = env.reset().clone()
    assert tuple(obs0.shape) == (n, nobs)
    g = torch.Generator(device="cuda").manual_seed(1)
    grf = torch.rand((n, 6), device="cuda", generator=g) * 400

    def dynamics(e, ctrl):
        q = e.data.qpos.t() + 0.02 * torch.randn((n, nq), device="cuda", generator=g)
        v = e.data.qvel.t() + 0.2 * torch.randn((n, nq), device="cuda", generator=g)
        return (q, v, grf) if kw.get("use_foot_forces") else (q, v)

    env.attach_dynamics(dynamics)
    prev = obs0
    for _ in range(2):
        obs, reward, absorbing, _ = env.step(torch.zeros((n, len(env._action_spec)), device="cuda"))
        q = env.data.qpos.t().double().cpu().numpy()
        v = env.data.qvel.t().double().cpu().numpy()
        ref = OH.step(model, q, v, prev[:, :2 * nq - 2].double().cpu().numpy())
        assert np.array_equal(obs[:, :2 * nq - 2].cpu().numpy(), ref["obs"].astype(np.float32))
        if kw.get("use_foot_forces"):
            assert torch.equal(obs[:, -6:], grf / 1000.0)                       # _create_observation :737-767
        assert np.array_equal(absorbing.cpu().numpy(), ref["absorbing"])
        assert_close(reward.cpu().numpy(), ref["reward"], "reward")
        assert_close(env.data.xpos.t().cpu().numpy().reshape(n, model.nbody, 3), ref["xpos"], "xpos")
        assert_close(env.data.cvel.t().cpu().numpy().reshape(n, model.nbody, 6), ref["cvel"], "cvel (COM includes the weight)")
        prev = obs.clone()
    if kw.get("use_foot_forces"):
        return
    # generic playback (per-step kernels) against the oracle's playback on the same variant
    T = 70
    env2 = om.LocoEnvBase.make("UnitreeH1.walk.real", n_envs=n, traj_params=dict(table=table), seed=31, **kw)
    res = env2.play_trajectory_from_velocity(n_episodes=1, n_steps_per_episode=T, render=False)
    for e in range(3):
        ref = OH.play_trajectory_from_velocity(model, table, 1, T, seed=31, env_id=e, record_fk=False)
        assert np.array_equal(res["obs"][e].cpu().numpy(), ref["obs"][-1].astype(np.float32))
        assert bool(res["has_fallen"][e]) == bool(ref["fallen"].any())
        assert int(env2.trajectories.device_state.reset_count[e]) == ref["final"]["reset_count"]


def test_multi_model_reset_switches_the_carried_weight():
    """MultiMuJoCo (loco_env_base.py:586-599): hold_weight with weight_mass=None builds one model per valid weight; every
    reset() moves the env object to the next one (random_env_reset=False) or to a Philox-drawn one."""
    import torch
    import olympics_mujoco_b200 as om
    from olympics_mujoco_b200.utils.philox import STREAM_MODEL_RESET, philox_randint
    env = om.LocoEnvBase.make("UnitreeH1.carry.real", n_envs=4, traj_params=dict(table=_table()), seed=9, random_env_reset=False)
    assert len(env._models) == 4 and [m.body_mass[-1] for m in env._models] == [0.2, 2.0, 10.0, 20.0]
    seen = []
    for _ in range(6):
        env.reset()
        seen.append(env._current_model_idx)
        env.forward()
        com_z = float(env.data.subtree_com[2, 0])
        seen[-1] = (seen[-1], round(com_z, 6))
    assert [s[0] for s in seen] == [1, 2, 3, 0, 1, 2]
    assert len({s[1] for s in seen[:4]}) == 4                              # a heavier weight moves the centre of mass
    rnd = om.LocoEnvBase.make("UnitreeH1.carry.real", n_envs=4, traj_params=dict(table=_table()), seed=9)
    got = []
    for k in range(8):
        rnd.reset()
        got.append(rnd._current_model_idx)
    assert got == [philox_randint(9, 0, k, STREAM_MODEL_RESET, 4) for k in range(8)] and len(set(got)) > 1
