"""Reference-trajectory store behind the reference's ``Trajectory`` interface, batched over n envs.

Mirrors ``olympic_mujoco/utils/trajectory.py`` (constructor :16-127, ``reset_trajectory`` :289-323,
``get_current_sample`` :381-387, ``get_next_sample`` :389-401, ``create_dataset`` :129-193,
``check_if_trajectory_is_in_range`` :325-366, ``_interpolate_trajectories`` :230-287).

Split of work: everything that happens ONCE at load time (npz parsing, range clipping, splitting, cubic
re-sampling to the control rate, dataset flattening) is host NumPy/SciPy float64 exactly as in the
reference; everything that happens PER STEP (index state, reset draws, sample gather, x/y re-centring)
runs on the GPU through ``om_traj_*`` (K3).  Reset draws follow the Philox contract instead of NumPy's
global MT19937 stream (hard part 4).
"""
from __future__ import annotations

import warnings
from copy import deepcopy

import numpy as np
from scipy import interpolate


class Trajectory:
    def __init__(self, keys, low, high, joint_pos_idx, interpolate_map=None, interpolate_remap=None, traj_path=None,
                 traj_files=None, interpolate_map_params=None, interpolate_remap_params=None, traj_dt=0.002,
                 control_dt=0.01, ignore_keys=None, clip_trajectory_to_joint_ranges=False, traj_info=None, warn=True,
                 table=None, n_envs=1, seed=0, env_id0=0, device="cuda"):
        given = sum(x is not None for x in (traj_path, traj_files, table))
        assert given == 1, "Please specify either traj_path or traj_files, but not both."
        self.keys = list(keys)
        self._traj_info = traj_info
        self.traj_dt, self.control_dt = traj_dt, control_dt
        if table is not None:
            # an already resampled [K, n_traj, T] table (the layout of np.array(Trajectory.trajectories))
            tab = np.asarray(table, dtype=np.float64)
            assert tab.ndim == 3 and tab.shape[0] == len(self.keys)
            self.trajectories = [tab[k] for k in range(tab.shape[0])]
            self.split_points = np.arange(tab.shape[1] + 1) * tab.shape[2]
        else:
            files = np.load(traj_path, allow_pickle=True) if traj_path is not None else traj_files
            self._trajectory_files = {k: np.asarray(d) for k, d in files.items()}
            self._clip_to_range(low, high, joint_pos_idx, warn, clip_trajectory_to_joint_ranges)
            self.keys += [k for k in self._trajectory_files if k.startswith("goal") and k not in self.keys]
            if ignore_keys is not None:
                for ik in ignore_keys:
                    self.keys.remove(ik)
            if "split_points" in self._trajectory_files:
                self.split_points = np.asarray(self._trajectory_files["split_points"])
            else:
                self.split_points = np.array([0, len(next(iter(self._trajectory_files.values())))])
            self.trajectories = self._extract_trajectory_from_files()
            if traj_info is not None:
                assert len(traj_info) == self.number_of_trajectories, \
                    "The number of trajectory infos/labels need to be equal to the number of trajectories."
            if self.traj_dt != control_dt:
                self._interpolate_trajectories(interpolate_map, interpolate_remap, interpolate_map_params,
                                               interpolate_remap_params)
        for obs in self.trajectories:
            if obs.ndim != 2:
                raise ValueError("only scalar observations per key are supported on the device path "
                                 f"(got shape {obs.shape})")
        self._n_envs, self._seed, self._env_id0, self._device = int(n_envs), int(seed), int(env_id0), device
        self._dev = None

    # ------------------------------------------------------------------ load-time host logic
    def _clip_to_range(self, low, high, j_idx, warn, clip):
        """trajectory.py:325-366 (relies, like the reference, on the file's key order = spec order)."""
        if not (warn or clip):
            return
        j_idx = list(np.asarray(j_idx)[2:])
        for i, (k, d) in enumerate(list(self._trajectory_files.items())):
            if i in j_idx:
                high_i, low_i = high[i - 2], low[i - 2]
                if warn:
                    msg = "Clipping the trajectory into range!" if clip else ""
                    if np.max(d) > high_i:
                        warnings.warn("Trajectory violates joint range in %s. Maximum in trajectory is %f "
                                      "and maximum range is %f. %s" % (self.keys[i], np.max(d), high_i, msg), RuntimeWarning)
                    elif np.min(d) < low_i:
                        warnings.warn("Trajectory violates joint range in %s. Minimum in trajectory is %f "
                                      "and minimum range is %f. %s" % (self.keys[i], np.min(d), low_i, msg), RuntimeWarning)
                if clip:
                    self._trajectory_files[k] = np.clip(d, low_i, high_i)

    def _extract_trajectory_from_files(self):
        """trajectory.py:195-228."""
        trajectories = [self._trajectory_files[key] for key in self.keys]
        lens = np.array([len(o) for o in trajectories])
        assert np.all(lens == lens[0]), "Some observations have different lengths than others. Trajectory is corrupted. "
        out = []
        for obs in trajectories:
            parts = np.split(obs, self.split_points[1:-1])
            pl = np.array([len(p) for p in parts])
            assert np.all(pl == pl[0]), "Only trajectories of equal length are currently supported."
            out.append(np.array(parts))
        return out

    def _interpolate_trajectories(self, map_funct, re_map_funct, map_params, re_map_params):
        """trajectory.py:230-287: cubic re-sampling of every trajectory to the control rate."""
        assert (map_funct is None) == (re_map_funct is None)
        L = self.trajectory_length
        x = np.arange(L)
        x_new = np.linspace(0, L - 1, round(L * (self.traj_dt / self.control_dt)), endpoint=True)
        new_trajs = []
        for i in range(self.number_of_trajectories):
            traj = [obs[i] for obs in self.trajectories]
            if map_funct is not None:
                traj = map_funct(traj) if map_params is None else map_funct(traj, **map_params)
            else:
                traj = np.array(traj)
            new = interpolate.interp1d(x, traj, kind="cubic", axis=1)(x_new)
            if re_map_funct is not None:
                new = re_map_funct(new) if re_map_params is None else re_map_funct(new, **re_map_params)
            new_trajs.append(new)
        self.trajectories = [np.array([t[k] for t in new_trajs]) for k in range(self.number_obs_trajectory)]
        self.split_points = np.arange(self.number_of_trajectories + 1) * self.trajectories[0].shape[1]

    def create_dataset(self, ignore_keys=None, state_callback=None, state_callback_params=None):
        """trajectory.py:129-193."""
        flat = self.flattened_trajectories()
        all_data = dict(zip(self.keys, deepcopy(list(flat))))
        if ignore_keys is not None:
            for ikey in ignore_keys:
                del all_data[ikey]
        states = np.concatenate(list(all_data.values()), axis=1)
        if state_callback is not None:
            states = np.array([state_callback(s, **state_callback_params) for s in states])
        absorbing = np.zeros(len(states) - 1)
        last = np.zeros(len(states))
        last[self.split_points[1:] - 1] = 1.0
        out = dict(states=states[:-1], next_states=states[1:], absorbing=absorbing, last=last)
        if self._traj_info is not None:
            out["info"] = np.array([[lab] * self.trajectory_length for lab in self._traj_info]).reshape(-1)
        return out

    def flattened_trajectories(self):
        return [obs.reshape((-1, 1)) for obs in self.trajectories]

    def table(self):
        """[K, n_traj, T] float64."""
        return np.array(self.trajectories)

    # ------------------------------------------------------------------ per-step device logic (K3)
    def bind(self, n_envs=None, seed=None, env_id0=None, device=None):
        """(Re)create the device table and the per-env state for n_envs environments."""
        from .. import kernels as Kn
        if n_envs is not None:
            self._n_envs = int(n_envs)
        if seed is not None:
            self._seed = int(seed)
        if env_id0 is not None:
            self._env_id0 = int(env_id0)
        if device is not None:
            self._device = device
        self._dev = Kn.DeviceTrajectory(self.table(), self._n_envs, seed=self._seed, env_id0=self._env_id0,
                                        device=self._device)
        self._sample = Kn.soa(len(self.keys), self._n_envs, device=self._device)
        return self._dev

    @property
    def device_state(self):
        if self._dev is None:
            self.bind()
        return self._dev

    def _as_out(self, sample):
        """SoA [K, n] -> what the caller of the reference would see: [n, K] (a list-like row per env)."""
        return sample.t()

    def reset_trajectory(self, substep_no=None, traj_no=None, mask=None):
        """trajectory.py:289-323 for every env (or the masked subset); ints may be scalars or [n] tensors."""
        import torch
        d = self.device_state
        def per_env(v, hi):
            if v is None:
                return None
            if not torch.is_tensor(v):
                assert 0 <= int(v) <= hi
                v = torch.full((d.n,), int(v), dtype=torch.int32, device=d.traj_no.device)
            return v.to(torch.int32)
        d.reset(mask=mask, traj_no=per_env(traj_no, self.number_of_trajectories),
                substep_no=per_env(substep_no, self.trajectory_length), sample=self._sample)
        return self._as_out(self._sample)

    def get_current_sample(self):
        self.device_state.current(sample=self._sample)
        return self._as_out(self._sample)

    def get_next_sample(self):
        """trajectory.py:389-401.  Returns None when a single env ran off its trajectory; with n > 1 envs the
        finished envs keep their previous sample row and ``self.ended`` ([n] uint8) marks them."""
        import torch
        d = self.device_state
        if not hasattr(self, "ended") or self.ended.numel() != d.n:
            self.ended = torch.zeros(d.n, dtype=torch.uint8, device=d.traj_no.device)
        self.ended.zero_()
        d.next(sample=self._sample, wrapped=self.ended, auto_reset=False)
        if d.n == 1 and bool(self.ended[0]):
            return None
        return self._as_out(self._sample)

    def get_from_sample(self, sample, key):
        assert sample.shape[-1] == len(self.keys)
        return sample[..., self.get_idx(key)]

    def get_idx(self, key):
        return self.keys.index(key)

    @property
    def traj_no(self):
        return self.device_state.traj_no

    @property
    def subtraj_step_no(self):
        return self.device_state.step_no

    @property
    def number_obs_trajectory(self):
        return len(self.trajectories)

    @property
    def trajectory_length(self):
        return self.trajectories[0].shape[1]

    @property
    def number_of_trajectories(self):
        return self.trajectories[0].shape[0]


def resample_table(data, model, traj_dt=1 / 500.0, control_dt=1 / 100.0, clip=True):
    """Dataset dict (34 H1 keys in spec order [+ split_points]) -> resampled [K, n_traj, T] float64 table,
    through exactly the load-time path ``LocoEnvBase.load_trajectory`` uses."""
    keys = [k for k in data if k != "split_points"]
    nq = len(keys) // 2
    low = np.full(len(keys), -np.inf)
    high = np.full(len(keys), np.inf)
    for i, k in enumerate(keys[:nq]):
        j = model.jnt_names.index(k[2:])
        if model.jnt_limited[j]:
            low[i], high[i] = model.jnt_range[j]
    tr = Trajectory(keys=keys, low=low[2:], high=high[2:], joint_pos_idx=np.arange(nq), traj_files=dict(data),
                    traj_dt=traj_dt, control_dt=control_dt, clip_trajectory_to_joint_ranges=clip, warn=False)
    return tr.table()
