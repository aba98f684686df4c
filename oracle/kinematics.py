"""Oracle: MuJoCo forward kinematics / COM / COM velocity in float64 NumPy (TEST INFRASTRUCTURE).

Restates ``mj_kinematics``, ``mj_comPos``, ``mj_comVel`` and ``mj_objectVelocity`` of the MuJoCo C
engine pinned by the reference (``mujoco==2.3.6``, reference ``requirements.txt:4``; source files
``engine_core_smooth.c`` / ``engine_support.c`` are NOT in /root/reference).  Reference call sites:
``mujoco.mj_forward`` at ``olympic_mujoco/environments/loco_env_base.py:410,525,1160`` and
``mj_objectVelocity`` at ``olympic_mujoco/interfaces/mujoco_robot_interface.py:299-327``.

Conventions (MuJoCo): quaternions ``[w,x,y,z]``; ``xmat`` row-major; spatial vectors
``[rot(3); lin(3)]``; ``cvel[i]`` is expressed at ``subtree_com[body_rootid[i]]``, world aligned.

All functions are batched over a leading env axis ``N`` and loop in Python only over bodies/joints,
in the engine's order, so a single env (``N=1``) follows the engine statement by statement.
PARITY UNPINNED by the reference (it has no FK fixtures, and MuJoCo cannot be installed here).  Pinned instead by an
independent second checker that shares no code with this file (``tools/fk_independent.py``: own MJCF reader, 4x4
homogeneous transforms, numerically differentiated velocities; ``tests/test_fk_independent.py``) and by the analytic
known answers of ``tests/test_oracle_fk.py``.
"""
from __future__ import annotations

import numpy as np

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
MJ_MINVAL = 1e-15


# ------------------------------------------------------------------ quaternion helpers (mju_*)
def quat_mul(a, b):
    """mju_mulQuat."""
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], axis=-1)


def quat_rot(q, v):
    """mju_rotVecQuat: rotate v by unit quaternion q."""
    w, u = q[..., :1], q[..., 1:]
    t = w * v + np.cross(u, v)
    return v + 2.0 * np.cross(u, t)


def quat_normalize(q):
    """mju_normalize4 (a zero quaternion becomes the identity)."""
    n = np.linalg.norm(q, axis=-1, keepdims=True)
    out = np.where(n < MJ_MINVAL, np.array([1.0, 0, 0, 0]), q / np.maximum(n, MJ_MINVAL))
    return out


def quat2mat(q):
    """mju_quat2Mat -> [...,3,3]."""
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0] = w * w + x * x - y * y - z * z
    m[..., 1, 1] = w * w - x * x + y * y - z * z
    m[..., 2, 2] = w * w - x * x - y * y + z * z
    m[..., 0, 1] = 2 * (x * y - w * z)
    m[..., 0, 2] = 2 * (x * z + w * y)
    m[..., 1, 0] = 2 * (x * y + w * z)
    m[..., 1, 2] = 2 * (y * z - w * x)
    m[..., 2, 0] = 2 * (x * z - w * y)
    m[..., 2, 1] = 2 * (y * z + w * x)
    return m


def axis_angle_quat(axis, angle):
    """mju_axisAngle2Quat with a constant axis [3] and batched angle [N]."""
    half = 0.5 * angle
    s = np.sin(half)
    return np.stack([np.cos(half), axis[0] * s, axis[1] * s, axis[2] * s], axis=-1)


# ------------------------------------------------------------------ mj_kinematics
def mj_kinematics(model, qpos):
    """Restates mj_kinematics.  qpos [N,nq] -> dict of xpos, xquat, xmat, xipos, xanchor, xaxis,
    site_xpos, site_xmat (all float64, env axis first)."""
    qpos = np.asarray(qpos, dtype=np.float64)
    N, nb, nj = qpos.shape[0], model.nbody, model.njnt
    xpos = np.zeros((N, nb, 3))
    xquat = np.zeros((N, nb, 4))
    xquat[:, 0, 0] = 1.0
    xmat = np.zeros((N, nb, 3, 3))
    xmat[:, 0] = np.eye(3)
    xanchor = np.zeros((N, nj, 3))
    xaxis = np.zeros((N, nj, 3))
    for i in range(1, nb):
        jadr, jnum = int(model.body_jntadr[i]), int(model.body_jntnum[i])
        if jnum == 1 and model.jnt_type[jadr] == JNT_FREE:
            qa = int(model.jnt_qposadr[jadr])
            pos = qpos[:, qa:qa + 3].copy()
            quat = quat_normalize(qpos[:, qa + 3:qa + 7])
            xanchor[:, jadr] = pos
            xaxis[:, jadr] = model.jnt_axis[jadr]
        else:
            pid = int(model.body_parentid[i])
            pos = xpos[:, pid] + np.einsum("nij,j->ni", xmat[:, pid], model.body_pos[i])
            quat = quat_mul(xquat[:, pid], np.broadcast_to(model.body_quat[i], (N, 4)))
            for j in range(jadr, jadr + jnum):
                qa = int(model.jnt_qposadr[j])
                jt = int(model.jnt_type[j])
                jaxis = np.broadcast_to(model.jnt_axis[j], (N, 3))
                jpos = np.broadcast_to(model.jnt_pos[j], (N, 3))
                xaxis[:, j] = quat_rot(quat, jaxis)
                xanchor[:, j] = quat_rot(quat, jpos) + pos
                if jt == JNT_SLIDE:
                    pos = pos + xaxis[:, j] * (qpos[:, qa] - model.qpos0[qa])[:, None]
                elif jt in (JNT_HINGE, JNT_BALL):
                    if jt == JNT_BALL:
                        qloc = quat_normalize(qpos[:, qa:qa + 4])
                    else:
                        qloc = axis_angle_quat(model.jnt_axis[j], qpos[:, qa] - model.qpos0[qa])
                    quat = quat_mul(quat, qloc)
                    pos = xanchor[:, j] - quat_rot(quat, jpos)
                else:
                    raise ValueError("free joint must be the only joint of its body")
        quat = quat_normalize(quat)
        xquat[:, i] = quat
        xpos[:, i] = pos
        xmat[:, i] = quat2mat(quat)
    # inertial frames and sites (mj_local2Global)
    xipos = xpos + np.einsum("nbij,bj->nbi", xmat, model.body_ipos)
    sb = model.site_bodyid
    site_xpos = xpos[:, sb] + np.einsum("nsij,sj->nsi", xmat[:, sb], model.site_pos)
    site_xquat = quat_mul(xquat[:, sb], np.broadcast_to(model.site_quat, (N, model.nsite, 4)))
    site_xmat = quat2mat(site_xquat)
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, xanchor=xanchor, xaxis=xaxis,
                site_xpos=site_xpos, site_xmat=site_xmat)


def body_subtreemass(model):
    m = model.body_mass.astype(np.float64).copy()
    for i in range(model.nbody - 1, 0, -1):
        m[model.body_parentid[i]] += m[i]
    return m


# ------------------------------------------------------------------ mj_comPos
def mj_comPos(model, kin):
    """Restates mj_comPos: subtree_com [N,nbody,3] and cdof [N,nv,6]."""
    xipos, xmat = kin["xipos"], kin["xmat"]
    N, nb = xipos.shape[0], model.nbody
    sub = np.zeros((N, nb, 3))
    stm = body_subtreemass(model)
    for i in range(nb - 1, -1, -1):
        sub[:, i] += xipos[:, i] * model.body_mass[i]
        if i:
            sub[:, model.body_parentid[i]] += sub[:, i]
        if stm[i] < MJ_MINVAL:
            sub[:, i] = xipos[:, i]
        else:
            sub[:, i] = sub[:, i] * (1.0 / max(MJ_MINVAL, stm[i]))
    cdof = np.zeros((N, model.nv, 6))
    for j in range(model.njnt):
        da, bi, jt = int(model.jnt_dofadr[j]), int(model.jnt_bodyid[j]), int(model.jnt_type[j])
        offset = sub[:, model.body_rootid[bi]] - kin["xanchor"][:, j]
        if jt == JNT_FREE:
            for k in range(3):
                cdof[:, da + k, 3 + k] = 1.0
            da += 3
        if jt in (JNT_FREE, JNT_BALL):
            for k in range(3):
                axis = xmat[:, bi, :, k]          # column k of xmat = body axis k in the world
                cdof[:, da + k, :3] = axis
                cdof[:, da + k, 3:] = np.cross(axis, offset)
        elif jt == JNT_SLIDE:
            cdof[:, da, 3:] = kin["xaxis"][:, j]
        else:
            axis = kin["xaxis"][:, j]
            cdof[:, da, :3] = axis
            cdof[:, da, 3:] = np.cross(axis, offset)
    return sub, cdof


# ------------------------------------------------------------------ mj_comVel
def mj_comVel(model, cdof, qvel):
    """Restates mj_comVel: cvel [N,nbody,6] = [omega; v at subtree_com[root]]."""
    qvel = np.asarray(qvel, dtype=np.float64)
    N = qvel.shape[0]
    cvel = np.zeros((N, model.nbody, 6))
    for i in range(1, model.nbody):
        v = cvel[:, model.body_parentid[i]].copy()
        jadr, jnum = int(model.body_jntadr[i]), int(model.body_jntnum[i])
        for j in range(jadr, jadr + jnum):
            da = int(model.jnt_dofadr[j])
            ndof = {JNT_FREE: 6, JNT_BALL: 3}.get(int(model.jnt_type[j]), 1)
            for d in range(da, da + ndof):
                v = v + cdof[:, d] * qvel[:, d:d + 1]
        cvel[:, i] = v
    return cvel


def mj_objectVelocity_xbody(model, xpos, subtree_com, cvel, body_id):
    """mj_objectVelocity(mjOBJ_XBODY, flg_local=0): [omega; v at xpos[body]] world aligned."""
    w = cvel[:, body_id, :3]
    v = cvel[:, body_id, 3:]
    tran = xpos[:, body_id] - subtree_com[:, model.body_rootid[body_id]]
    return np.concatenate([w, v - np.cross(tran, w)], axis=-1)


def forward(model, qpos, qvel):
    """Hot-path subset of mj_forward: kinematics + comPos + comVel.  Returns the engine fields the
    reference consumes (``xpos, xquat, site_xpos, site_xmat, cvel``) plus intermediates."""
    kin = mj_kinematics(model, qpos)
    sub, cdof = mj_comPos(model, kin)
    cvel = mj_comVel(model, cdof, qvel)
    out = dict(kin)
    out.update(subtree_com=sub, cdof=cdof, cvel=cvel)
    return out
