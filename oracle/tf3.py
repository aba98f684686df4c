"""Oracle: the ``transforms3d`` functions the reference calls (TEST INFRASTRUCTURE).

``transforms3d`` is imported by the reference (``loco_env_base.py:8``, ``walking_task.py:3``,
``StickFigureA3.py:3``, ``mujoco_robot_interface.py:2``) but is neither vendored nor pinned, and is
not installed here.  These are restatements of its published algorithms for the default axes
``'sxyz'`` (static frame, x-y-z order): ``quaternions.quat2mat``, ``euler.mat2euler``,
``euler.quat2euler``, ``euler.euler2quat``, ``euler.euler2mat(0,0,yaw)``, ``quaternions.mat2quat``.
Unpinned by the reference; pinned against SciPy's ``Rotation`` (an independent implementation of the same
conventions) in ``tests/test_oracle_third_party_pins.py``, plus round-trip properties.
"""
import numpy as np

_FLOAT_EPS = np.finfo(np.float64).eps
_EPS4 = _FLOAT_EPS * 4.0


def quat2mat(q):
    """Batched [...,4] -> [...,3,3]; handles unnormalised q (scales by 2/|q|^2)."""
    q = np.asarray(q, np.float64)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    Nq = w * w + x * x + y * y + z * z
    small = Nq < _FLOAT_EPS
    s = 2.0 / np.where(small, 1.0, Nq)
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0] = 1.0 - (yY + zZ); m[..., 0, 1] = xY - wZ; m[..., 0, 2] = xZ + wY
    m[..., 1, 0] = xY + wZ; m[..., 1, 1] = 1.0 - (xX + zZ); m[..., 1, 2] = yZ - wX
    m[..., 2, 0] = xZ - wY; m[..., 2, 1] = yZ + wX; m[..., 2, 2] = 1.0 - (xX + yY)
    m[small] = np.eye(3)
    return m


def mat2euler(M):
    """axes='sxyz' -> (ax, ay, az), batched."""
    M = np.asarray(M, np.float64)
    cy = np.sqrt(M[..., 0, 0] ** 2 + M[..., 1, 0] ** 2)
    reg = cy > _EPS4
    ax = np.where(reg, np.arctan2(M[..., 2, 1], M[..., 2, 2]), np.arctan2(-M[..., 1, 2], M[..., 1, 1]))
    ay = np.arctan2(-M[..., 2, 0], cy)
    az = np.where(reg, np.arctan2(M[..., 1, 0], M[..., 0, 0]), 0.0)
    return ax, ay, az


def quat2euler(q):
    return mat2euler(quat2mat(q))


def euler2quat(ai, aj, ak):
    """axes='sxyz' -> [...,4] (w first)."""
    ai, aj, ak = np.asarray(ai, np.float64) / 2.0, np.asarray(aj, np.float64) / 2.0, np.asarray(ak, np.float64) / 2.0
    ci, si, cj, sj, ck, sk = np.cos(ai), np.sin(ai), np.cos(aj), np.sin(aj), np.cos(ak), np.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.stack([cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc], axis=-1)


def rotz(theta):
    """euler2mat(0, 0, theta)."""
    theta = np.asarray(theta, np.float64)
    c, s = np.cos(theta), np.sin(theta)
    m = np.zeros(theta.shape + (3, 3))
    m[..., 0, 0] = c; m[..., 0, 1] = -s; m[..., 1, 0] = s; m[..., 1, 1] = c; m[..., 2, 2] = 1.0
    return m


def mat2quat(M):
    """Bar-Itzhack largest-eigenvector method, sign fixed so that w >= 0 (single matrix)."""
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = np.asarray(M, np.float64).reshape(9)
    Kmat = np.array([[Qxx - Qyy - Qzz, 0, 0, 0],
                     [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0],
                     [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
                     [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(Kmat)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q = -q
    return q
