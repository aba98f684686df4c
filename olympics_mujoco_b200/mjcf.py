"""MJCF -> flat kinematic model tables (host side, no MuJoCo needed).

The reference hands its MJCF files to MuJoCo's C compiler through ``dm_control.mjcf``
(reference ``olympic_mujoco/environments/real_humanoid_robots/UnitreeH1.py:70-111``,
``loco_env_base.py:836-868``).  Neither package exists in this image, so this module
restates the part of the MJCF compiler the kinematic hot path needs: body tree in
depth-first document order, default classes, joint tables (type/axis/pos/range), explicit
``<inertial>`` or geom-derived mass + centre of mass (density 1000), sites, actuators.

The result, :class:`KinematicModel`, is what both the CPU oracle and the CUDA path consume
(``om_model_create`` in ``include/om_b200.h``).  It can be stored as JSON so that the GPU box,
which has no copy of the reference's XML files, loads the tables instead of the XML.

Only the MJCF subset used by ``h1.xml`` / ``a3.xml`` is supported; anything else raises.
"""
from __future__ import annotations

import json
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

# mjtJoint enum order of MuJoCo (engine value, kept so tables read like mjModel)
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
_JNT_NAMES = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}
_NQ = {JNT_FREE: 7, JNT_BALL: 4, JNT_SLIDE: 1, JNT_HINGE: 1}
_NV = {JNT_FREE: 6, JNT_BALL: 3, JNT_SLIDE: 1, JNT_HINGE: 1}


def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.asarray(default, dtype=np.float64)
    v = np.asarray([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


def _normalize(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    if n < 1e-15:
        raise ValueError("zero quaternion in MJCF")
    return q / n


@dataclass
class KinematicModel:
    """Flat tables named after the mjModel fields they stand in for."""

    name: str
    nq: int
    nv: int
    body_names: list
    body_parentid: np.ndarray      # [nbody] int32
    body_rootid: np.ndarray        # [nbody] int32 (top-level ancestor, world -> 0)
    body_pos: np.ndarray           # [nbody,3]
    body_quat: np.ndarray          # [nbody,4] normalised, w first
    body_ipos: np.ndarray          # [nbody,3] centre of mass in the body frame
    body_mass: np.ndarray          # [nbody]
    body_jntadr: np.ndarray        # [nbody] int32 (-1 when the body has no joint)
    body_jntnum: np.ndarray        # [nbody] int32
    jnt_names: list
    jnt_type: np.ndarray           # [njnt] int32 (mjtJoint)
    jnt_bodyid: np.ndarray         # [njnt] int32
    jnt_axis: np.ndarray           # [njnt,3] unit, local
    jnt_pos: np.ndarray            # [njnt,3] local anchor
    jnt_qposadr: np.ndarray        # [njnt] int32
    jnt_dofadr: np.ndarray         # [njnt] int32
    jnt_range: np.ndarray          # [njnt,2]
    jnt_limited: np.ndarray        # [njnt] bool
    qpos0: np.ndarray              # [nq]
    site_names: list
    site_bodyid: np.ndarray        # [nsite] int32
    site_pos: np.ndarray           # [nsite,3]
    site_quat: np.ndarray          # [nsite,4]
    actuator_names: list = field(default_factory=list)
    actuator_joint: list = field(default_factory=list)
    actuator_gear: np.ndarray = field(default_factory=lambda: np.zeros(0))
    actuator_ctrlrange: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    actuator_ctrllimited: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=bool))
    timestep: float = 0.002

    # ------------------------------------------------------------------ sizes
    @property
    def nbody(self):
        return len(self.body_names)

    @property
    def njnt(self):
        return len(self.jnt_names)

    @property
    def nsite(self):
        return len(self.site_names)

    @property
    def nu(self):
        return len(self.actuator_names)

    @property
    def total_mass(self):
        """mj_getTotalmass: sum of body masses excluding the world body."""
        return float(self.body_mass[1:].sum())

    def body_id(self, name):
        return self.body_names.index(name)

    def joint_id(self, name):
        return self.jnt_names.index(name)

    def site_id(self, name):
        return self.site_names.index(name)

    # ------------------------------------------------------------- table-level model edits
    # The reference edits the MJCF through dm_control before MuJoCo compiles it (loco_env_base.py:836-868
    # _delete_from_xml_handle; UnitreeH1.py:245-261 _add_weight).  The same edits on the COMPILED tables give the same
    # model (checked against a fresh compile of the edited MJCF in tests/test_host.py), so every variant of a robot can be
    # built from the tables shipped with the package, without the reference's XML at hand.
    def without_joints(self, names, name=None):
        """The model with the named (single-dof) joints and the actuators that drive them removed."""
        import copy
        drop = [self.jnt_names.index(n) for n in names]
        for j in drop:
            if int(self.jnt_type[j]) not in (JNT_SLIDE, JNT_HINGE):
                raise ValueError("without_joints handles slide / hinge joints")
        keep = [j for j in range(self.njnt) if j not in drop]
        m = copy.deepcopy(self)
        m.name = name or self.name
        qmap = {int(self.jnt_qposadr[j]): None for j in drop}
        keep_q = [k for k in range(self.nq) if k not in qmap]
        m.qpos0 = self.qpos0[keep_q].copy()
        m.nq, m.nv = self.nq - len(drop), self.nv - len(drop)
        for f in ("jnt_type", "jnt_bodyid", "jnt_axis", "jnt_pos", "jnt_range", "jnt_limited"):
            setattr(m, f, np.asarray(getattr(self, f))[keep].copy())
        m.jnt_names = [self.jnt_names[j] for j in keep]
        qa, da = 0, 0
        m.jnt_qposadr, m.jnt_dofadr = np.zeros(len(keep), np.int32), np.zeros(len(keep), np.int32)
        for i, j in enumerate(keep):
            m.jnt_qposadr[i], m.jnt_dofadr[i] = qa, da
            t = int(self.jnt_type[j])
            qa += 7 if t == JNT_FREE else 4 if t == JNT_BALL else 1
            da += _NV[t]
        m.body_jntadr, m.body_jntnum = np.full(self.nbody, -1, np.int32), np.zeros(self.nbody, np.int32)
        for i in range(len(keep)):
            b = int(m.jnt_bodyid[i])
            if m.body_jntnum[b] == 0:
                m.body_jntadr[b] = i
            m.body_jntnum[b] += 1
        ka = [a for a in range(self.nu) if self.actuator_joint[a] not in names]
        m.actuator_names = [self.actuator_names[a] for a in ka]
        m.actuator_joint = [self.actuator_joint[a] for a in ka]
        m.actuator_gear = np.asarray(self.actuator_gear)[ka].copy()
        m.actuator_ctrlrange = np.asarray(self.actuator_ctrlrange)[ka].copy()
        m.actuator_ctrllimited = np.asarray(self.actuator_ctrllimited)[ka].copy()
        return m

    def with_child_body(self, parent, body_name, mass, ipos, pos=(0.0, 0.0, 0.0), name=None):
        """The model with one more jointless body hanging off ``parent``.  Supported where the new body is the LAST one in
        MuJoCo's depth-first numbering (the parent's subtree ends the tree), which is where dm_control's ``add`` puts it
        for the H1 torso."""
        import copy
        pid = self.body_id(parent)
        b = self.nbody - 1
        while b != pid:                                  # the last body must descend from the parent
            if b == 0:
                raise ValueError(f"a child of {parent!r} would not be the last body of the tree")
            b = int(self.body_parentid[b])
        m = copy.deepcopy(self)
        m.name = name or self.name
        m.body_names = self.body_names + [body_name]
        app = lambda a, v: np.concatenate([np.asarray(a), np.asarray(v, dtype=np.asarray(a).dtype).reshape((1,) + np.asarray(a).shape[1:])])
        m.body_parentid = app(self.body_parentid, pid)
        m.body_rootid = app(self.body_rootid, self.body_rootid[pid])
        m.body_pos = app(self.body_pos, pos)
        m.body_quat = app(self.body_quat, [1.0, 0, 0, 0])
        m.body_ipos = app(self.body_ipos, ipos)
        m.body_mass = app(self.body_mass, mass)
        m.body_jntadr = app(self.body_jntadr, -1)
        m.body_jntnum = app(self.body_jntnum, 0)
        return m

    # ------------------------------------------------------------- (de)serialise
    _ARRAYS = ("body_parentid body_rootid body_pos body_quat body_ipos body_mass body_jntadr "
               "body_jntnum jnt_type jnt_bodyid jnt_axis jnt_pos jnt_qposadr jnt_dofadr jnt_range "
               "jnt_limited qpos0 site_bodyid site_pos site_quat actuator_gear actuator_ctrlrange "
               "actuator_ctrllimited").split()
    _INT = ("body_parentid body_rootid body_jntadr body_jntnum jnt_type jnt_bodyid jnt_qposadr "
            "jnt_dofadr site_bodyid").split()
    _BOOL = ("jnt_limited", "actuator_ctrllimited")

    def to_dict(self):
        d = {"name": self.name, "nq": self.nq, "nv": self.nv, "timestep": self.timestep,
             "body_names": self.body_names, "jnt_names": self.jnt_names,
             "site_names": self.site_names, "actuator_names": self.actuator_names,
             "actuator_joint": self.actuator_joint}
        for k in self._ARRAYS:
            d[k] = np.asarray(getattr(self, k)).tolist()
        return d

    @classmethod
    def from_dict(cls, d):
        kw = dict(d)
        for k in cls._ARRAYS:
            a = np.asarray(d[k])
            if k in cls._INT:
                a = a.astype(np.int32)
            elif k in cls._BOOL:
                a = a.astype(bool)
            else:
                a = a.astype(np.float64)
            kw[k] = a
        for k, w in (("body_pos", 3), ("body_quat", 4), ("body_ipos", 3), ("jnt_axis", 3),
                     ("jnt_pos", 3), ("jnt_range", 2), ("site_pos", 3), ("site_quat", 4),
                     ("actuator_ctrlrange", 2)):
            kw[k] = kw[k].reshape(-1, w)
        return cls(**kw)

    def save_json(self, path):
        Path(path).write_text(json.dumps(self.to_dict(), indent=1))

    @classmethod
    def load_json(cls, path):
        return cls.from_dict(json.loads(Path(path).read_text()))


# --------------------------------------------------------------------------- defaults
class _Defaults:
    """Resolved <default> tree: class name -> {element tag -> attribute dict}."""

    def __init__(self, root):
        self.classes = {"main": {}}
        for top in root.findall("default"):
            self._walk(top, "main", inherit={})

    def _walk(self, node, cls_name, inherit):
        cur = {tag: dict(attrs) for tag, attrs in inherit.items()}
        for child in node:
            if child.tag == "default":
                continue
            cur.setdefault(child.tag, {}).update(child.attrib)
        self.classes[cls_name] = cur
        for child in node.findall("default"):
            self._walk(child, child.get("class"), cur)

    def resolve(self, tag, elem, childclass):
        cls_name = elem.get("class") or childclass or "main"
        if cls_name not in self.classes:
            raise ValueError(f"unknown default class {cls_name!r}")
        attrs = dict(self.classes[cls_name].get(tag, {}))
        attrs.update(elem.attrib)
        return attrs


# --------------------------------------------------------------------------- geoms
def _geom_mass_com(a):
    """Mass and centre (body frame) of a primitive geom, MuJoCo conventions."""
    gtype = a.get("type", "sphere")
    if gtype in ("plane", "hfield"):
        return 0.0, np.zeros(3)
    if gtype == "mesh":
        # meshes need the STL volume; both in-scope models give explicit <inertial> for
        # mesh bodies, so a mesh geom contributes only if someone asks for geom inertia.
        raise ValueError("mesh geoms need an explicit <inertial>")
    size = _vec(a.get("size"), default=[0.0])
    pos = _vec(a.get("pos"), 3, default=[0, 0, 0])
    fromto = _vec(a.get("fromto"), 6)
    if fromto is not None:
        p0, p1 = fromto[:3], fromto[3:]
        pos = 0.5 * (p0 + p1)
        half = 0.5 * np.linalg.norm(p1 - p0)
    else:
        half = size[1] if size.size > 1 else 0.0
    if gtype == "sphere":
        vol = 4.0 / 3.0 * math.pi * size[0] ** 3
    elif gtype == "capsule":
        r = size[0]
        vol = math.pi * r * r * (2 * half) + 4.0 / 3.0 * math.pi * r ** 3
    elif gtype == "cylinder":
        vol = math.pi * size[0] ** 2 * (2 * half)
    elif gtype == "box":
        vol = 8.0 * size[0] * size[1] * size[2]
    elif gtype == "ellipsoid":
        vol = 4.0 / 3.0 * math.pi * size[0] * size[1] * size[2]
    else:
        raise ValueError(f"unsupported geom type {gtype!r}")
    if "mass" in a:
        return float(a["mass"]), pos
    density = float(a.get("density", 1000.0))
    return density * vol, pos


# --------------------------------------------------------------------------- compiler
def compile_mjcf(xml_path, name=None, remove_joints=(), remove_actuators=(), body_quat_overrides=None):
    """Compile an MJCF file into a :class:`KinematicModel`.

    ``remove_joints`` / ``remove_actuators`` / ``body_quat_overrides`` restate the XML edits the
    reference performs through dm_control before compiling (``loco_env_base.py:836-868``
    ``_delete_from_xml_handle``; ``UnitreeH1.py:268-290`` ``_reorient_arms``).
    """
    root = ET.parse(str(xml_path)).getroot()
    if root.tag != "mujoco":
        raise ValueError("not an MJCF file")
    comp = root.find("compiler")
    if comp is not None and comp.get("angle", "degree") != "radian":
        raise ValueError("only angle='radian' models are supported")
    autolimits = comp is not None and comp.get("autolimits", "false") == "true"
    opt = root.find("option")
    timestep = float(opt.get("timestep", 0.002)) if opt is not None else 0.002
    defaults = _Defaults(root)
    body_quat_overrides = dict(body_quat_overrides or {})
    remove_joints = set(remove_joints)
    remove_actuators = set(remove_actuators)

    B = dict(names=["world"], parent=[0], root=[0], pos=[np.zeros(3)], quat=[np.array([1.0, 0, 0, 0])],
             ipos=[np.zeros(3)], mass=[0.0], jntadr=[-1], jntnum=[0])
    J = dict(names=[], type=[], body=[], axis=[], pos=[], qposadr=[], dofadr=[], range=[], limited=[])
    S = dict(names=[], body=[], pos=[], quat=[])
    qpos0 = []
    nv = 0

    def add_body(elem, parent_id, childclass):
        nonlocal nv
        bid = len(B["names"])
        bname = elem.get("name", f"body{bid}")
        childclass = elem.get("childclass", childclass)
        for bad in ("euler", "axisangle", "xyaxes", "zaxis"):
            if elem.get(bad) is not None:
                raise ValueError(f"body orientation via {bad!r} is not supported")
        B["names"].append(bname)
        B["parent"].append(parent_id)
        B["root"].append(bid if parent_id == 0 else B["root"][parent_id])
        B["pos"].append(_vec(elem.get("pos"), 3, default=[0, 0, 0]))
        quat = body_quat_overrides.pop(bname, None)
        if quat is None:
            quat = _vec(elem.get("quat"), 4, default=[1, 0, 0, 0])
        B["quat"].append(_normalize(quat))
        B["jntadr"].append(-1)
        B["jntnum"].append(0)
        # joints, document order
        for j in elem:
            if j.tag not in ("joint", "freejoint"):
                continue
            if j.get("name") in remove_joints:
                remove_joints.discard(j.get("name"))
                continue
            if j.tag == "freejoint":
                a = dict(j.attrib)
                jt = JNT_FREE
            else:
                a = defaults.resolve("joint", j, childclass)
                jt = _JNT_NAMES[a.get("type", "hinge")]
            jid = len(J["names"])
            if B["jntnum"][bid] == 0:
                B["jntadr"][bid] = jid
            B["jntnum"][bid] += 1
            J["names"].append(a.get("name", f"joint{jid}"))
            J["type"].append(jt)
            J["body"].append(bid)
            axis = _vec(a.get("axis"), 3, default=[0, 0, 1])
            J["axis"].append(axis / np.linalg.norm(axis) if jt in (JNT_SLIDE, JNT_HINGE) else np.array([0.0, 0, 1]))
            J["pos"].append(np.zeros(3) if jt == JNT_FREE else _vec(a.get("pos"), 3, default=[0, 0, 0]))
            J["qposadr"].append(len(qpos0))
            J["dofadr"].append(nv)
            rng = _vec(a.get("range"), 2)
            lim = a.get("limited")
            if lim is None or lim == "auto":
                limited = autolimits and rng is not None
            else:
                limited = lim == "true"
            J["range"].append(rng if rng is not None else np.zeros(2))
            J["limited"].append(bool(limited))
            if jt == JNT_FREE:
                qpos0.extend(list(B["pos"][bid]) + list(B["quat"][bid]))
            elif jt == JNT_BALL:
                qpos0.extend([1.0, 0, 0, 0])
            else:
                qpos0.append(float(a.get("ref", 0.0)))
            nv += _NV[jt]
        # inertia: explicit <inertial> wins, else sum the geoms (inertiafromgeom="auto")
        inertial = elem.find("inertial")
        if inertial is not None:
            B["mass"].append(float(inertial.get("mass")))
            B["ipos"].append(_vec(inertial.get("pos"), 3, default=[0, 0, 0]))
        else:
            m_tot, mc = 0.0, np.zeros(3)
            for g in elem.findall("geom"):
                m, c = _geom_mass_com(defaults.resolve("geom", g, childclass))
                m_tot += m
                mc += m * c
            B["mass"].append(m_tot)
            B["ipos"].append(mc / m_tot if m_tot > 0 else np.zeros(3))
        for s in elem.findall("site"):
            a = defaults.resolve("site", s, childclass)
            S["names"].append(a.get("name", f"site{len(S['names'])}"))
            S["body"].append(bid)
            S["pos"].append(_vec(a.get("pos"), 3, default=[0, 0, 0]))
            S["quat"].append(_normalize(_vec(a.get("quat"), 4, default=[1, 0, 0, 0])))
        for child in elem.findall("body"):
            add_body(child, bid, childclass)

    world = root.find("worldbody")
    for s in world.findall("site"):
        a = defaults.resolve("site", s, None)
        S["names"].append(a.get("name", f"site{len(S['names'])}"))
        S["body"].append(0)
        S["pos"].append(_vec(a.get("pos"), 3, default=[0, 0, 0]))
        S["quat"].append(_normalize(_vec(a.get("quat"), 4, default=[1, 0, 0, 0])))
    for b in world.findall("body"):
        add_body(b, 0, None)
    if remove_joints:
        raise ValueError(f"joints to remove not found: {sorted(remove_joints)}")
    if body_quat_overrides:
        raise ValueError(f"bodies to re-orient not found: {sorted(body_quat_overrides)}")

    A = dict(names=[], joint=[], gear=[], ctrlrange=[], ctrllimited=[])
    act = root.find("actuator")
    if act is not None:
        for m in act:
            if m.get("name") in remove_actuators:
                remove_actuators.discard(m.get("name"))
                continue
            a = defaults.resolve(m.tag, m, None)
            if a.get("joint") not in J["names"]:
                raise ValueError(f"actuator {a.get('name')!r} drives a removed joint")
            A["names"].append(a.get("name"))
            A["joint"].append(a.get("joint"))
            A["gear"].append(float(a.get("gear", "1").split()[0]))
            cr = _vec(a.get("ctrlrange"), 2)
            cl = a.get("ctrllimited")
            A["ctrlrange"].append(cr if cr is not None else np.zeros(2))
            A["ctrllimited"].append((cl == "true") if cl not in (None, "auto") else (autolimits and cr is not None))
    if remove_actuators:
        raise ValueError(f"actuators to remove not found: {sorted(remove_actuators)}")

    i32 = lambda x: np.asarray(x, dtype=np.int32)
    f64 = lambda x, w: np.asarray(x, dtype=np.float64).reshape(-1, w)
    return KinematicModel(
        name=name or root.get("model", "model"), nq=len(qpos0), nv=nv,
        body_names=B["names"], body_parentid=i32(B["parent"]), body_rootid=i32(B["root"]),
        body_pos=f64(B["pos"], 3), body_quat=f64(B["quat"], 4), body_ipos=f64(B["ipos"], 3),
        body_mass=np.asarray(B["mass"], dtype=np.float64), body_jntadr=i32(B["jntadr"]),
        body_jntnum=i32(B["jntnum"]),
        jnt_names=J["names"], jnt_type=i32(J["type"]), jnt_bodyid=i32(J["body"]),
        jnt_axis=f64(J["axis"], 3), jnt_pos=f64(J["pos"], 3), jnt_qposadr=i32(J["qposadr"]),
        jnt_dofadr=i32(J["dofadr"]), jnt_range=f64(J["range"], 2),
        jnt_limited=np.asarray(J["limited"], dtype=bool), qpos0=np.asarray(qpos0, dtype=np.float64),
        site_names=S["names"], site_bodyid=i32(S["body"]), site_pos=f64(S["pos"], 3),
        site_quat=f64(S["quat"], 4),
        actuator_names=A["names"], actuator_joint=A["joint"],
        actuator_gear=np.asarray(A["gear"], dtype=np.float64),
        actuator_ctrlrange=f64(A["ctrlrange"], 2),
        actuator_ctrllimited=np.asarray(A["ctrllimited"], dtype=bool), timestep=timestep)


# ------------------------------------------------------------------ the two in-scope robots
H1_ARM_JOINTS = ["l_arm_shy", "l_arm_shx", "l_arm_shz", "left_elbow",
                 "r_arm_shy", "r_arm_shx", "r_arm_shz", "right_elbow"]      # UnitreeH1.py:149-155
H1_ARM_QUATS = {"left_shoulder_pitch_link": [1.0, 0.25, 0.1, 0.0],          # UnitreeH1.py:280-288
                "right_elbow_link": [1.0, 0.0, 0.25, 0.0],
                "right_shoulder_pitch_link": [1.0, -0.25, 0.1, 0.0],
                "left_elbow_link": [1.0, 0.0, 0.25, 0.0]}


H1_VALID_WEIGHTS = [0.1, 1.0, 5.0, 10.0]                                    # UnitreeH1.py:62
# UnitreeH1._add_weight (UnitreeH1.py:245-261): body "weight" under torso_link with TWO box geoms of `mass` each
H1_WEIGHT_GEOM_POS = ((0.35, 0.0, 0.1), (0.9, 0.0, 0.1))


def unitree_h1_variant(disable_arms=True, disable_back_joint=False, hold_weight=False, weight_mass=None, base=None):
    """Every UnitreeH1 variant the constructor switches select (UnitreeH1.py:38-111), from the shipped tables:
    arms disabled -> arm joints removed and, unless a weight is carried, the arms re-oriented (the default table);
    back joint removed; ``hold_weight`` -> a jointless "weight" body of 2 x ``weight_mass`` under torso_link, arms kept in
    their MJCF orientation.  ``base``: a compiled full model (arms + back) to derive from instead of the shipped one."""
    if hold_weight and not disable_arms:
        raise AssertionError("If you want Unitree H1 to carry a weight, please disable the arms. They will be kept fixed.")
    if disable_arms and not hold_weight and base is None:
        m = load_builtin("unitree_h1")
    else:
        m = base if base is not None else load_builtin("unitree_h1_arms")
        if disable_arms:
            m = m.without_joints(H1_ARM_JOINTS, name="UnitreeH1")
            if not hold_weight:
                for b, q in H1_ARM_QUATS.items():
                    m.body_quat[m.body_id(b)] = _normalize(np.asarray(q, dtype=np.float64))
    if disable_back_joint:
        m = m.without_joints(["back_bkz"], name=m.name)
    if hold_weight:
        w = float(weight_mass)
        com = np.mean(np.asarray(H1_WEIGHT_GEOM_POS), axis=0)               # two equal masses
        m = m.with_child_body("torso_link", "weight", 2.0 * w, com, name=m.name)
    return m


def compile_unitree_h1(xml_path, disable_arms=True, disable_back_joint=False, hold_weight=False, weight_mass=None):
    """UnitreeH1 model with the constructor's XML edits applied (UnitreeH1.py:70-88)."""
    if hold_weight:
        full = compile_mjcf(xml_path, name="UnitreeH1")
        return unitree_h1_variant(disable_arms, disable_back_joint, True, weight_mass, base=full)
    rm_j, rm_a, quats = [], [], None
    if disable_arms:
        rm_j += H1_ARM_JOINTS
        rm_a += [j + "_actuator" for j in H1_ARM_JOINTS]
        quats = H1_ARM_QUATS
    if disable_back_joint:
        rm_j.append("back_bkz")
        rm_a.append("back_bkz_actuator")
    return compile_mjcf(xml_path, name="UnitreeH1", remove_joints=rm_j, remove_actuators=rm_a,
                        body_quat_overrides=quats)


def compile_stick_figure_a3(xml_path):
    """StickFigureA3 model; the reference applies no XML edits (StickFigureA3.py:50-62)."""
    return compile_mjcf(xml_path, name="StickFigureA3")


_MODEL_DIR = Path(__file__).resolve().parent / "models"


def load_builtin(name):
    """Load one of the pre-compiled tables shipped in ``olympics_mujoco_b200/models``."""
    return KinematicModel.load_json(_MODEL_DIR / f"{name}.json")
