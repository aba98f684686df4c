"""An INDEPENDENT second checker for forward kinematics / COM / COM velocity (TEST INFRASTRUCTURE).

`oracle/kinematics.py` restates MuJoCo's recursions (quaternion chains, cdof, cvel recursion).  This file shares NO code,
helper or table with it (nor with `olympics_mujoco_b200/mjcf.py`): it
  * reads the MJCF files itself (ElementTree; only what kinematics and mass properties need),
  * poses every body with 4x4 HOMOGENEOUS TRANSFORMS composed down the tree (rotation matrices from Rodrigues' formula,
    never a quaternion product),
  * gets velocities NUMERICALLY: q(t) = q (+) t * qdot (free-joint orientation integrated with the angular velocity in the
    BODY frame, positions in the world frame -- MuJoCo's qvel convention), central differences of each body's pose, and only
    then transports the linear velocity to the subtree centre of mass of the tree's root, the point MuJoCo's `cvel` refers to.
So a convention error shared by the two restatements of the recursions (the cvel reference point, the free-joint angular
velocity frame, the hinge anchor correction, geom-derived masses) cannot pass both.

The reference's call sites: `mujoco.mj_forward` at /root/reference/olympic_mujoco/environments/loco_env_base.py:410,525,1160;
`mj_objectVelocity` at olympic_mujoco/interfaces/mujoco_robot_interface.py:299-327.  Model files:
olympic_mujoco/environments/data/unitree_h1/h1.xml, data/stickFigure_A3/a3.xml (edited as UnitreeH1.py:71-109,134-160,
245-291 does: arm joints removed and arm bodies re-oriented, back joint removed, the carried weight added).

    python tools/fk_independent.py            # regenerates tests/golden/fk_independent_ref.npz (needs /root/reference)
"""
import math
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

REF_DATA = Path("/root/reference/olympic_mujoco/environments/data")
H1_XML = REF_DATA / "unitree_h1" / "h1.xml"
A3_XML = REF_DATA / "stickFigure_A3" / "a3.xml"
H1_ARM_JOINTS = ("l_arm_shy", "l_arm_shx", "l_arm_shz", "left_elbow", "r_arm_shy", "r_arm_shx", "r_arm_shz", "right_elbow")
# UnitreeH1._reorient_arms (UnitreeH1.py:275-291); MuJoCo's compiler normalises body quaternions
H1_REORIENT = {"left_shoulder_pitch_link": (1.0, 0.25, 0.1, 0.0), "right_elbow_link": (1.0, 0.0, 0.25, 0.0),
               "right_shoulder_pitch_link": (1.0, -0.25, 0.1, 0.0), "left_elbow_link": (1.0, 0.0, 0.25, 0.0)}


def _floats(text, n=None):
    v = [float(x) for x in text.split()]
    assert n is None or len(v) == n, (text, n)
    return np.array(v)


def rot_from_quat_attr(q):
    """Rotation matrix of an MJCF `quat` attribute (w x y z), normalised first, by Rodrigues' formula on its axis-angle."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q)
    s = float(np.linalg.norm(q[1:]))              # |sin(angle / 2)| (not sqrt(1 - w^2): that cancels near the identity)
    if s < 1e-300:
        return np.eye(3)
    return rodrigues(q[1:] / s, 2.0 * math.atan2(s, q[0]))


def rodrigues(axis, angle):
    k = np.asarray(axis, dtype=np.float64)
    k = k / np.linalg.norm(k)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1.0 - math.cos(angle)) * (K @ K)


def homog(R=None, p=None):
    T = np.eye(4)
    if R is not None:
        T[:3, :3] = R
    if p is not None:
        T[:3, 3] = p
    return T


class Body:
    def __init__(self, name, parent):
        self.name, self.parent = name, parent
        self.T0 = np.eye(4)            # pose in the parent frame at the reference configuration
        self.joints = []               # dicts(name, type, axis, pos, qadr, vadr)
        self.mass = 0.0
        self.ipos = np.zeros(3)        # centre of mass in the body frame
        self.sites = []                # (name, pos, R)


def _geom_mass_and_centre(g, default_type):
    """Mass (density 1000, or the explicit mass attribute) and centre of a primitive geom: what MuJoCo's compiler uses
    to build a body's inertial frame when the body has no <inertial>."""
    gtype = g.get("type", default_type)
    size = _floats(g.get("size", "0"))
    if g.get("fromto") is not None:
        ft = _floats(g.get("fromto"), 6)
        a, b = ft[:3], ft[3:]
        centre, half = 0.5 * (a + b), 0.5 * np.linalg.norm(b - a)
        r = size[0]
        if gtype == "capsule":
            vol = math.pi * r * r * (2.0 * half) + 4.0 / 3.0 * math.pi * r ** 3
        elif gtype == "cylinder":
            vol = math.pi * r * r * (2.0 * half)
        else:
            raise ValueError(f"fromto on a {gtype}")
    else:
        centre = _floats(g.get("pos", "0 0 0"), 3)
        if gtype == "sphere":
            vol = 4.0 / 3.0 * math.pi * size[0] ** 3
        elif gtype == "box":
            vol = 8.0 * size[0] * size[1] * size[2]
        elif gtype == "capsule":
            vol = math.pi * size[0] ** 2 * (2.0 * size[1]) + 4.0 / 3.0 * math.pi * size[0] ** 3
        elif gtype == "cylinder":
            vol = math.pi * size[0] ** 2 * (2.0 * size[1])
        else:
            return 0.0, centre             # plane / mesh: carries no mass here
    mass = float(g.get("mass")) if g.get("mass") is not None else 1000.0 * vol
    return mass, centre


class Tree:
    """Kinematic tree + mass properties read straight from an MJCF file."""

    def __init__(self, xml_path, remove_joints=(), body_quat=None, extra_bodies=(), default_geom_type="sphere"):
        root = ET.parse(str(xml_path)).getroot()
        # the one default this reader needs: a3.xml's class "body" makes capsule the default geom type
        for d in root.iter("default"):
            if d.get("class") == "body" and d.find("geom") is not None and d.find("geom").get("type"):
                default_geom_type = d.find("geom").get("type")
        self.bodies = [Body("world", -1)]
        self.nq = self.nv = 0
        body_quat = dict(body_quat or {})

        def add(elem, parent):
            b = Body(elem.get("name"), parent)
            R = rot_from_quat_attr(body_quat.get(b.name, _floats(elem.get("quat", "1 0 0 0"), 4)))
            b.T0 = homog(R, _floats(elem.get("pos", "0 0 0"), 3))
            idx = len(self.bodies)
            self.bodies.append(b)
            for j in list(elem.findall("joint")) + list(elem.findall("freejoint")):
                if j.get("name") in remove_joints:
                    continue
                jt = "free" if j.tag == "freejoint" else j.get("type", "hinge")
                jd = dict(name=j.get("name"), type=jt, axis=_floats(j.get("axis", "0 0 1"), 3),
                          pos=_floats(j.get("pos", "0 0 0"), 3), qadr=self.nq, vadr=self.nv)
                self.nq += 7 if jt == "free" else 1
                self.nv += 6 if jt == "free" else 1
                b.joints.append(jd)
            inertial = elem.find("inertial")
            if inertial is not None:
                b.mass = float(inertial.get("mass"))
                b.ipos = _floats(inertial.get("pos", "0 0 0"), 3)
            else:
                ms, cs = zip(*[_geom_mass_and_centre(g, default_geom_type) for g in elem.findall("geom")]) \
                    if elem.findall("geom") else ((), ())
                b.mass = float(sum(ms))
                b.ipos = sum(m * c for m, c in zip(ms, cs)) / b.mass if b.mass > 0 else np.zeros(3)
            for s in elem.findall("site"):
                b.sites.append((s.get("name"), _floats(s.get("pos", "0 0 0"), 3),
                                rot_from_quat_attr(_floats(s.get("quat", "1 0 0 0"), 4))))
            for child in elem.findall("body"):
                add(child, idx)
            for (pname, name, geoms) in extra_bodies:                 # UnitreeH1._add_weight: a jointless child body
                if pname == b.name:
                    wb = Body(name, idx)
                    ms = [float(g["mass"]) for g in geoms]
                    cs = [np.asarray(g["pos"], dtype=np.float64) for g in geoms]
                    wb.mass = float(sum(ms))
                    wb.ipos = sum(m * c for m, c in zip(ms, cs)) / wb.mass
                    self.bodies.append(wb)

        for top in root.find("worldbody").findall("body"):
            add(top, 0)
        self.names = [b.name for b in self.bodies]
        self.site_names = [s[0] for b in self.bodies for s in b.sites]

    # ------------------------------------------------------------------ pose
    def poses(self, q):
        """World 4x4 pose of every body frame for configuration q (MuJoCo qpos layout)."""
        T = [np.eye(4) for _ in self.bodies]
        for i, b in enumerate(self.bodies[1:], start=1):
            if len(b.joints) == 1 and b.joints[0]["type"] == "free":
                a = b.joints[0]["qadr"]
                T[i] = homog(rot_from_quat_attr(q[a + 3:a + 7]), q[a:a + 3])     # a free body ignores its parent and pos
                continue
            X = T[b.parent] @ b.T0
            for j in b.joints:
                x = q[j["qadr"]]
                if j["type"] == "slide":
                    M = homog(None, j["axis"] / np.linalg.norm(j["axis"]) * x)
                else:                                   # hinge about the axis through the anchor `pos` (body frame)
                    M = homog(None, j["pos"]) @ homog(rodrigues(j["axis"], x)) @ homog(None, -j["pos"])
                X = X @ M
            T[i] = X
        return T

    def integrate(self, q, v, t):
        """q (+) t v: slides / hinges add; a free joint moves its position along the WORLD-frame linear velocity and turns
        its orientation by the BODY-frame angular velocity (R <- R exp([w t]))."""
        out = np.array(q, dtype=np.float64)
        for b in self.bodies[1:]:
            for j in b.joints:
                qa, va = j["qadr"], j["vadr"]
                if j["type"] == "free":
                    out[qa:qa + 3] = q[qa:qa + 3] + t * v[va:va + 3]
                    w = v[va + 3:va + 6]
                    R = rot_from_quat_attr(q[qa + 3:qa + 7])
                    ang = np.linalg.norm(w) * t
                    Rn = R @ (rodrigues(w, ang) if abs(ang) > 0 else np.eye(3))
                    out[qa + 3:qa + 7] = quat_from_rot(Rn)
                else:
                    out[qa] = q[qa] + t * v[va]
        return out

    def root_of(self, i):
        while self.bodies[i].parent > 0:
            i = self.bodies[i].parent
        return i

    def forward(self, q, v, h=1e-6):
        """-> dict(xpos, xmat, xquat, xipos, site_xpos, site_xmat, subtree_com, cvel) for ONE configuration."""
        q, v = np.asarray(q, np.float64), np.asarray(v, np.float64)
        nb = len(self.bodies)
        T = self.poses(q)
        Tp, Tm = self.poses(self.integrate(q, v, h)), self.poses(self.integrate(q, v, -h))
        xpos = np.array([t[:3, 3] for t in T])
        xmat = np.array([t[:3, :3] for t in T])
        xipos = np.array([t[:3, :3] @ b.ipos + t[:3, 3] for t, b in zip(T, self.bodies)])
        mass = np.array([b.mass for b in self.bodies])
        # subtree centre of mass of every body (children have larger indices)
        msum, mcom = mass.copy(), mass[:, None] * xipos
        for i in range(nb - 1, 0, -1):
            p = self.bodies[i].parent
            msum[p] += msum[i]
            mcom[p] += mcom[i]
        sub = np.where(msum[:, None] > 1e-15, mcom / np.maximum(msum, 1e-300)[:, None], xipos)
        cvel = np.zeros((nb, 6))
        for i in range(1, nb):
            dR = (Tp[i][:3, :3] - Tm[i][:3, :3]) / (2 * h)
            W = dR @ xmat[i].T                          # [w]x, up to O(h^2)
            w = 0.5 * np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]])
            vo = (Tp[i][:3, 3] - Tm[i][:3, 3]) / (2 * h)           # velocity of the body-frame origin
            P = sub[self.root_of(i)]
            cvel[i, :3] = w
            cvel[i, 3:] = vo + np.cross(w, P - xpos[i])            # ... transported to the tree's centre of mass
        sites_p, sites_R = [], []
        for t, b in zip(T, self.bodies):
            for (_, sp, sR) in b.sites:
                sites_p.append(t[:3, :3] @ sp + t[:3, 3])
                sites_R.append(t[:3, :3] @ sR)
        return dict(xpos=xpos, xmat=xmat, xquat=np.array([quat_from_rot(R) for R in xmat]), xipos=xipos,
                    site_xpos=np.array(sites_p).reshape(-1, 3), site_xmat=np.array(sites_R).reshape(-1, 3, 3),
                    subtree_com=sub, cvel=cvel, total_mass=float(mass[1:].sum()), body_mass=mass)

    def object_velocity(self, out, i):
        """mj_objectVelocity(mjOBJ_XBODY, flg_local=0), reordered [lin, ang] as mujoco_robot_interface.py:327 does: the
        velocity of the body-frame ORIGIN, numerically (independent of the cvel transport above)."""
        w = out["cvel"][i, :3]
        P = out["subtree_com"][self.root_of(i)]
        return np.concatenate([out["cvel"][i, 3:] - np.cross(w, P - out["xpos"][i]), w])


def quat_from_rot(R):
    """Unit quaternion (w >= 0 branch by largest component) of a rotation matrix -- Shepperd's method."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    cand = [tr, R[0, 0], R[1, 1], R[2, 2]]
    k = int(np.argmax(cand))
    if k == 0:
        s = math.sqrt(1.0 + tr) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif k == 1:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif k == 2:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    return q / np.linalg.norm(q)


# ---------------------------------------------------------------------------------------------- model variants
def h1_tree(disable_arms=True, disable_back_joint=False, hold_weight=False, weight_mass=None):
    remove = tuple(H1_ARM_JOINTS if disable_arms else ()) + (("back_bkz",) if disable_back_joint else ())
    quat = H1_REORIENT if (disable_arms and not hold_weight) else None
    extra = ()
    if hold_weight:                                       # UnitreeH1._add_weight (UnitreeH1.py:245-261): two boxes
        extra = (("torso_link", "weight", (dict(mass=weight_mass, pos=(0.35, 0, 0.1)), dict(mass=weight_mass, pos=(0.9, 0, 0.1)))),)
    return Tree(H1_XML, remove_joints=remove, body_quat=quat, extra_bodies=extra)


def a3_tree():
    return Tree(A3_XML)


def _states(tree, n, rng, free=False):
    q = rng.normal(0, 0.4, (n, tree.nq))
    v = rng.normal(0, 1.5, (n, tree.nv))
    if free:
        q[:, :3] = rng.uniform(-2, 2, (n, 3)) + [0, 0, 1.3]
        quat = rng.normal(0, 1, (n, 4))
        q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    else:
        q[:, :3] = rng.uniform(-1, 1, (n, 3))            # H1 root slides
    q[0] = 0.0                                            # the reference configuration first
    if free:
        q[0, 3] = 1.0
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    q, v = f32(q), f32(v)
    if free:                                              # keep the free-joint quaternion a unit quaternion in float64
        q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
    return q, v


def generate(path):
    rng = np.random.default_rng(20260101)
    out = {}
    variants = dict(h1=h1_tree(), h1_arms=h1_tree(disable_arms=False), h1_noback=h1_tree(disable_back_joint=True),
                    h1_carry=h1_tree(hold_weight=True, weight_mass=5.0), a3=a3_tree())
    for name, tree in variants.items():
        n = 24
        q, v = _states(tree, n, rng, free=(name == "a3"))
        res = [tree.forward(q[i], v[i]) for i in range(n)]
        out[name + "_qpos"], out[name + "_qvel"] = q, v
        for k in ("xpos", "xquat", "xmat", "xipos", "site_xpos", "site_xmat", "subtree_com", "cvel"):
            out[f"{name}_{k}"] = np.stack([r[k] for r in res])
        out[name + "_body_mass"] = res[0]["body_mass"]
        out[name + "_body_names"] = np.array(tree.names)
        if name == "a3":
            feet = [tree.names.index("left_foot"), tree.names.index("right_foot")]
            out["a3_foot_objvel"] = np.stack([[tree.object_velocity(r, f) for f in feet] for r in res])
    np.savez_compressed(path, **out)
    return out


if __name__ == "__main__":
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "fk_independent_ref.npz"
    generate(dst)
    print("wrote", dst, dst.stat().st_size, "bytes")
