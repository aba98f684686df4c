"""``ObservationHelper`` / ``ObservationType`` of mushroom_rl (>=1.10, not in the reference tree; used at
``loco_env_base.py:329,603,667,699,886,997,1136,1182-1192,1200,1230``), over batched device data.

Same constructor arguments, attributes (``observation_spec, obs_idx_map, joint_pos_idx, joint_vel_idx,
obs_low, obs_high, build_omit_idx``) and methods as upstream.  ``data`` is a
:class:`BatchedData` (the device-resident stand-in for ``MjData``); ``_build_obs`` returns ``[n, D]``.
"""
from __future__ import annotations

from enum import Enum

import numpy as np
import torch

from .mjcf import JNT_FREE, JNT_BALL


class ObservationType(Enum):
    __order__ = "BODY_POS BODY_ROT BODY_VEL JOINT_POS JOINT_VEL SITE_POS SITE_ROT"
    BODY_POS = 0
    BODY_ROT = 1
    BODY_VEL = 2
    JOINT_POS = 3
    JOINT_VEL = 4
    SITE_POS = 5
    SITE_ROT = 6


class BatchedData:
    """Device-resident ``MjData`` subset for n envs, structure-of-arrays ([C, n], env index contiguous).
    ``xpos``/``xquat``/``cvel``/``site_xpos``/``site_xmat`` hold the last K1 result."""

    FIELDS = ("qpos", "qvel", "xpos", "xquat", "cvel", "site_xpos", "site_xmat", "subtree_com")

    def __init__(self, model, n, device="cuda"):
        self.model, self.n = model, n
        z = lambda c: torch.zeros((c, n), dtype=torch.float32, device=device)
        self.qpos, self.qvel = z(model.nq), z(model.nv)
        self.qpos[:] = torch.as_tensor(model.qpos0, dtype=torch.float32, device=device)[:, None]
        self.xpos, self.xquat, self.cvel = z(model.nbody * 3), z(model.nbody * 4), z(model.nbody * 6)
        self.site_xpos, self.site_xmat, self.subtree_com = z(model.nsite * 3), z(model.nsite * 9), z(3)

    def joint_rows(self, name):
        m = self.model
        j = m.jnt_names.index(name)
        nq = {JNT_FREE: 7, JNT_BALL: 4}.get(int(m.jnt_type[j]), 1)
        nv = {JNT_FREE: 6, JNT_BALL: 3}.get(int(m.jnt_type[j]), 1)
        return slice(int(m.jnt_qposadr[j]), int(m.jnt_qposadr[j]) + nq), slice(int(m.jnt_dofadr[j]), int(m.jnt_dofadr[j]) + nv)

    def rows(self, name, otype):
        """SoA rows [c, n] of one observation entry."""
        m = self.model
        if otype == ObservationType.JOINT_POS:
            return self.qpos[self.joint_rows(name)[0]]
        if otype == ObservationType.JOINT_VEL:
            return self.qvel[self.joint_rows(name)[1]]
        if otype in (ObservationType.BODY_POS, ObservationType.BODY_ROT, ObservationType.BODY_VEL):
            b = m.body_names.index(name)
            arr, w = {ObservationType.BODY_POS: (self.xpos, 3), ObservationType.BODY_ROT: (self.xquat, 4),
                      ObservationType.BODY_VEL: (self.cvel, 6)}[otype]
            return arr[b * w:(b + 1) * w]
        s = m.site_names.index(name)
        if otype == ObservationType.SITE_POS:
            return self.site_xpos[3 * s:3 * s + 3]
        return self.site_xmat[9 * s:9 * s + 9]


class ObservationHelper:
    def __init__(self, observation_spec, model, data, max_joint_velocity=None):
        if len(observation_spec) == 0:
            raise AttributeError("No Environment observations were specified. Add at least one observation to "
                                 "the observation_spec.")
        self.obs_low, self.obs_high = [], []
        self.joint_pos_idx, self.joint_vel_idx = [], []
        self.joint_mujoco_idx = []
        self.obs_idx_map, self.build_omit_idx = {}, {}
        self.observation_spec = observation_spec
        self._model = model
        cur = 0
        for key, name, ot in observation_spec:
            assert key not in self.obs_idx_map.keys(), 'Found duplicate key in observation specification: "%s"' % key
            count = int(data.rows(name, ot).shape[0])
            self.obs_idx_map[key] = list(range(cur, cur + count))
            self.build_omit_idx[key] = []
            if count == 1 and ot == ObservationType.JOINT_POS:
                self.joint_pos_idx.append(cur)
                j = model.jnt_names.index(name)
                self.joint_mujoco_idx.append(j)
                if model.jnt_limited[j]:
                    self.obs_low.append(model.jnt_range[j][0]); self.obs_high.append(model.jnt_range[j][1])
                else:
                    self.obs_low.append(-np.inf); self.obs_high.append(np.inf)
            elif count == 1 and ot == ObservationType.JOINT_VEL:
                self.joint_vel_idx.append(cur)
                if max_joint_velocity is None:
                    self.obs_low.append(-np.inf); self.obs_high.append(np.inf)
                else:
                    self.obs_low.append(-max_joint_velocity[len(self.joint_vel_idx) - 1])
                    self.obs_high.append(max_joint_velocity[len(self.joint_vel_idx) - 1])
            else:
                self.obs_low.extend([-np.inf] * count); self.obs_high.extend([np.inf] * count)
            cur += count
        self.obs_low, self.obs_high = np.array(self.obs_low), np.array(self.obs_high)

    def remove_obs(self, key, index):
        self.build_omit_idx[key].append(index)

    def add_obs(self, key, length, min_value=-np.inf, max_value=np.inf):
        self.obs_idx_map[key] = list(range(len(self.obs_low), len(self.obs_low) + length))
        if hasattr(min_value, "__len__"):
            self.obs_low = np.append(self.obs_low, min_value)
        else:
            self.obs_low = np.append(self.obs_low, [min_value] * length)
        if hasattr(max_value, "__len__"):
            self.obs_high = np.append(self.obs_high, max_value)
        else:
            self.obs_high = np.append(self.obs_high, [max_value] * length)

    def get_from_obs(self, obs, key):
        return obs[..., self.obs_idx_map[key]]

    def get_joint_pos_from_obs(self, obs):
        return obs[..., self.joint_pos_idx]

    def get_joint_vel_from_obs(self, obs):
        return obs[..., self.joint_vel_idx]

    def get_obs_limits(self):
        return self.obs_low, self.obs_high

    def get_state(self, data, name, o_type):
        """[n, c] view of one entry."""
        return data.rows(name, o_type).t()

    def get_all_observation_keys(self):
        return list(self.obs_idx_map.keys())

    def _modify_data(self, data, obs):
        """mushroom_rl ``ObservationHelper._modify_data``: the inverse of ``_build_obs`` -- write a full observation
        ([n, D] or [D], every spec entry present) back into the data arrays it was gathered from."""
        obs = torch.as_tensor(obs, dtype=torch.float32, device=data.qpos.device)
        if obs.dim() == 1:
            obs = obs[None].expand(data.n, -1)
        cur = 0
        for key, name, ot in self.observation_spec:
            rows = data.rows(name, ot)                       # a view into data.qpos / qvel / xpos ...
            c = rows.shape[0]
            rows.copy_(obs[:, cur:cur + c].t())
            cur += c
        return data

    def _build_obs_soa(self, data):
        parts = []
        for key, name, ot in self.observation_spec:
            rows = data.rows(name, ot)
            omit = self.build_omit_idx[key]
            if omit:
                keep = [i for i in range(rows.shape[0]) if i not in omit]
                rows = rows[keep]
            parts.append(rows)
        return torch.cat(parts, dim=0)

    def _build_obs(self, data):
        """Spec-driven gather on the device -> [n, D]."""
        return self._build_obs_soa(data).t()
