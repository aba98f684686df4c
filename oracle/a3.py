"""Oracle: StickFigureA3 RL step tail (obs / WalkingTask / reward / done / reset), float64 NumPy,
one env at a time in the reference's control flow (TEST INFRASTRUCTURE).

Follows:
* ``StickFigureA3.get_obs``            ``real_humanoid_robots/StickFigureA3.py:144-181``
* ``StickFigureA3.step`` tail           ``:187-202`` (everything after ``robot.step`` = mj_step, out of scope)
* ``StickFigureA3.reset_model``        ``:205-235``; nominal pose ``environments/robot.py:61-80``
* ``WalkingTask.step``                 ``tasks/walking_task.py:246-293``; ``update_target_steps`` :228-244;
  ``update_goal_steps`` :184-225; ``calc_reward`` :74-110; ``step_reward`` :56-72; ``done`` :298-319;
  ``reset`` :321-397; ``generate_step_sequence`` :137-182; ``transform_sequence`` :113-135
* reward terms ``tasks/rewards.py:27-40`` (height), ``:65-83`` (foot force clock), ``:85-102`` (foot velocity
  clock), ``:121-126`` (orientation); clocks ``:270-366`` evaluated at integer phase (LUT [period,4] with
  columns r_frc, r_vel, l_frc, l_vel; pinned by ``tests/golden/phase_clock_ref.npz``)
* body/site reads ``interfaces/mujoco_robot_interface.py:239-346``; contact-derived inputs
  (``get_lfoot_grf`` :287-297, ``get_rfoot_grf`` :275-285, foot-floor contact z ``rewards.py:29-31``,
  ``check_bad_collisions`` :393-399) come from the contact solver, which is out of the hot path; they
  enter as the per-env ``contact`` record.
"""
from dataclasses import dataclass, field

import numpy as np

from . import kinematics as K
from . import philox, tf3

STANDING, FORWARD = 0, 1
CONTROL_DT = 0.025
SWING, STANCE, TOTAL = 0.75, 0.35, 1.1            # StickFigureA3.py:110-113
GOAL_HEIGHT_REF = 0.80
GOAL_SPEED_REF = 0.0                              # walking_task.py:30
TARGET_RADIUS = 0.20
DELAY_FRAMES = int(np.floor(SWING / CONTROL_DT))  # 30
PERIOD = int(np.floor(2 * TOTAL * (1 / CONTROL_DT)))  # 88
MAX_STEPS = 20

NOMINAL_DEG = [-30, 0, 0, 50, 0, -24, -30, 0, 0, 50, 0, -24, -3, -9.74, -30, -3, 9.74, -30]


def init_qpos():
    """robot.py:61-80 (base z 0.81 is overwritten to 1.34 by reset_model)."""
    return np.array([0, 0, 0.81, 1, 0, 0, 0] + [q * np.pi / 180.0 for q in NOMINAL_DEG], dtype=np.float64)


@dataclass
class Contact:
    """What the task reads from the contact solver each step."""
    l_grf: float = 0.0
    r_grf: float = 0.0
    min_z: float = 0.0            # min contact-point z over foot-floor contacts
    foot_contact: bool = False    # any foot-floor contact
    bad_collision: bool = False   # (n_foot_floor_contacts != ncon)


@dataclass
class TaskState:
    phase: int = 0
    t1: int = 0
    t2: int = 0
    target_reached: bool = False
    target_reached_frames: int = 0
    mode: int = FORWARD
    seq_len: int = 0
    sequence: np.ndarray = field(default_factory=lambda: np.zeros((MAX_STEPS, 4)))
    goal_x: np.ndarray = field(default_factory=lambda: np.zeros(2))
    goal_y: np.ndarray = field(default_factory=lambda: np.zeros(2))
    goal_z: np.ndarray = field(default_factory=lambda: np.zeros(2))
    goal_theta: np.ndarray = field(default_factory=lambda: np.zeros(2))


class A3Ids:
    def __init__(self, model):
        self.root = model.body_id("torso")
        self.head = model.body_id("head")
        self.lfoot = model.body_id("left_foot")
        self.rfoot = model.body_id("right_foot")
        self.lsite = model.site_id("lf_force")
        self.rsite = model.site_id("rf_force")


def get_obs(qpos, qvel, ts):
    """StickFigureA3.py:144-178 -> [41]."""
    clock = [np.sin(2 * np.pi * ts.phase / PERIOD), np.cos(2 * np.pi * ts.phase / PERIOD)]
    ext = np.concatenate([clock, ts.goal_x, ts.goal_y, ts.goal_z, ts.goal_theta])
    r, p, _ = tf3.quat2euler(qpos[3:7])
    root_orient = tf3.euler2quat(r, p, 0.0)
    state = np.concatenate([root_orient, qvel[3:6], qpos[7:19], qvel[6:18], ext])
    assert state.shape == (41,)
    return state


def _foot_state(model, ids, fk):
    """walking_task.py:254-263 (site positions, XBODY linear velocities of the foot bodies)."""
    l_pos = fk["site_xpos"][0, ids.lsite]
    r_pos = fk["site_xpos"][0, ids.rsite]
    l_vel = K.mj_objectVelocity_xbody(model, fk["xpos"], fk["subtree_com"], fk["cvel"], ids.lfoot)[0, 3:]
    r_vel = K.mj_objectVelocity_xbody(model, fk["xpos"], fk["subtree_com"], fk["cvel"], ids.rfoot)[0, 3:]
    return l_pos, r_pos, l_vel, r_vel


def update_goal_steps(ts, root_pos, root_quat):
    """walking_task.py:184-225."""
    ts.goal_x[:] = 0; ts.goal_y[:] = 0; ts.goal_z[:] = 0; ts.goal_theta[:] = 0
    ref = np.eye(4)
    ref[:3, :3] = tf3.quat2mat(root_quat)
    ref[:3, 3] = root_pos
    for idx, t in enumerate([ts.t1, ts.t2]):
        tgt = np.eye(4)
        tgt[:3, :3] = tf3.rotz(ts.sequence[t][3])
        tgt[:3, 3] = ts.sequence[t][0:3]
        rel = np.linalg.inv(ref).dot(tgt)
        if ts.mode != STANDING:
            ts.goal_x[idx], ts.goal_y[idx], ts.goal_z[idx] = rel[0, 3], rel[1, 3], rel[2, 3]
            ts.goal_theta[idx] = tf3.mat2euler(rel[:3, :3])[2]


def task_step(model, ids, fk, ts):
    """walking_task.py:246-293.  Returns the foot quantities the reward needs."""
    ts.phase += 1
    if ts.phase >= PERIOD:
        ts.phase = 0
    l_pos, r_pos, l_vel, r_vel = _foot_state(model, ids, fk)
    target = ts.sequence[ts.t1][0:3]
    l_in = np.linalg.norm(l_pos - target) < TARGET_RADIUS
    r_in = np.linalg.norm(r_pos - target) < TARGET_RADIUS
    if l_in or r_in:
        ts.target_reached = True
        ts.target_reached_frames += 1
    else:
        ts.target_reached = False
        ts.target_reached_frames = 0
    if ts.target_reached and ts.target_reached_frames >= DELAY_FRAMES:
        ts.t1 = ts.t2                                   # update_target_steps :228-244
        ts.t2 += 1
        if ts.t2 == ts.seq_len:
            ts.t2 = ts.seq_len - 1
        ts.target_reached = False
        ts.target_reached_frames = 0
    update_goal_steps(ts, fk["xpos"][0, ids.root], fk["xquat"][0, ids.root])
    return l_pos, r_pos, l_vel, r_vel


def calc_reward(model, ids, fk, ts, contact, clock_lut, feet):
    """walking_task.py:74-110 -> array of the 6 weighted terms in dict order."""
    l_pos, r_pos, l_vel, r_vel = feet
    mass = model.total_mass
    orient = tf3.euler2quat(0.0, 0.0, ts.sequence[ts.t1][3])
    if ts.mode == STANDING:
        r_frc_c, r_vel_c, l_frc_c, l_vel_c = 1.0, -1.0, 1.0, -1.0
    else:
        r_frc_c, r_vel_c, l_frc_c, l_vel_c = clock_lut[ts.phase]
    root_pos = fk["xpos"][0, ids.root]
    head_pos = fk["xpos"][0, ids.head]
    # rewards.py:65-83
    fmax = mass * 9.8 * 0.5
    nl = min(contact.l_grf, fmax) / fmax * 2 - 1
    nr = min(contact.r_grf, fmax) / fmax * 2 - 1
    frc = (np.tan(np.pi / 4 * l_frc_c * nl) + np.tan(np.pi / 4 * r_frc_c * nr)) / 2
    # rewards.py:85-102
    vmax = 0.2
    vl = min(np.linalg.norm(l_vel), vmax) / vmax * 2 - 1
    vr = min(np.linalg.norm(r_vel), vmax) / vmax * 2 - 1
    vel = (np.tan(np.pi / 4 * l_vel_c * vl) + np.tan(np.pi / 4 * r_vel_c * vr)) / 2
    # rewards.py:121-126
    orient_r = np.exp(-10 * (1 - np.inner(orient, fk["xquat"][0, ids.root]) ** 2))
    # rewards.py:27-40
    contact_point = contact.min_z if contact.foot_contact else 0.0
    err = np.abs(root_pos[2] - contact_point - GOAL_HEIGHT_REF)
    if err < 0.01 + 0.05 * GOAL_SPEED_REF:
        err = 0.0
    height = np.exp(-40 * np.square(err))
    # walking_task.py:56-72
    target = ts.sequence[ts.t1][0:3]
    fd = min(np.linalg.norm(l_pos - target), np.linalg.norm(r_pos - target))
    hit = np.exp(-fd / 0.25) if ts.target_reached else 0.0
    mp = (ts.sequence[ts.t1][0:2] + ts.sequence[ts.t2][0:2]) / 2
    progress = np.exp(-np.linalg.norm(root_pos[0:2] - mp) / 2)
    step_r = 0.8 * hit + 0.2 * progress
    upper = np.exp(-10 * np.square(np.linalg.norm(head_pos[0:2] - root_pos[0:2])))
    return np.array([0.150 * frc, 0.150 * vel, 0.050 * orient_r, 0.050 * height, 0.450 * step_r, 0.050 * upper])


def done(ids, fk, contact, feet):
    """walking_task.py:298-319."""
    l_pos, r_pos = feet[0], feet[1]
    root_rel_height = fk["xpos"][0, ids.root][2] - min(l_pos[2], r_pos[2])
    return bool(root_rel_height < 0.6) or bool(contact.bad_collision)


def step_tail(model, qpos, qvel, ts, contact, clock_lut):
    """StickFigureA3.step after robot.step: returns (obs[41], total, done, terms[6]); mutates ts."""
    ids = A3Ids(model)
    fk = K.forward(model, qpos[None], qvel[None])
    feet = task_step(model, ids, fk, ts)
    terms = calc_reward(model, ids, fk, ts, contact, clock_lut, feet)
    total = float(sum(float(t) for t in terms))
    d = done(ids, fk, contact, feet)
    qn = qpos.copy()
    qn[3:7] = qn[3:7] / np.linalg.norm(qn[3:7])      # mj_kinematics normalises qpos quaternions in place
    return get_obs(qn, qvel, ts), total, d, terms


# ------------------------------------------------------------------ reset (A13)
N_UNIFORM = 60


def reset_uniforms(seed, env_id, reset_count):
    """The contract's uniform draws u[0..59] in [0,1) for one A3 reset (15 Philox blocks)."""
    u = []
    for s in range(N_UNIFORM // 4):
        w = philox.draw(seed, np.uint32(env_id), np.uint32(reset_count), philox.STREAM_A3_RESET + s)
        u.extend(float(philox.to_unit(x)) for x in w)
    return np.array(u)


def reset(model, seed, env_id, reset_count, iteration_count=np.inf):
    """reset_model (StickFigureA3.py:205-235) + WalkingTask.reset (walking_task.py:321-397).
    Draw order: u[0:25] qpos noise, u[25:49] qvel noise, u[49],u[50] root x,y, u[51] pitch, u[52] yaw,
    u[53] phase, u[54] mode, u[55] step-height sign, u[56] first-step y, u[57] c."""
    u = reset_uniforms(seed, env_id, reset_count)
    c = 0.02
    qpos = init_qpos() + (u[0:25] * 2 - 1) * c
    qvel = (u[25:49] * 2 - 1) * c
    qpos[0] = u[49] * 2 - 1
    qpos[1] = u[50] * 2 - 1
    qpos[2] = 1.34
    qpos[3:7] = tf3.euler2quat(0.0, (u[51] * 10 - 5) * np.pi / 180, (u[52] * 2 - 1) * np.pi)
    ids = A3Ids(model)
    fk = K.forward(model, qpos[None], qvel[None])            # set_state -> mj_forward
    ts = TaskState()
    ts.phase = 0 if u[53] < 0.5 else PERIOD // 2
    ts.mode = STANDING if u[54] < 0.2 else FORWARD
    step_size, step_gap, step_height, num_steps = 0.3, 0.15, 0.0, MAX_STEPS
    if ts.mode == STANDING:
        num_steps = 1
    else:
        h = float(np.clip((iteration_count - 3000) / 8000, 0, 1) * 0.1)
        step_height = -h if u[55] < 0.5 else h
    # generate_step_sequence :137-182 (straight path)
    first_y = 0.095 + 0.01 * u[56]
    seq = []
    if ts.phase == PERIOD // 2:
        seq.append(np.array([0, -first_y, 0, 0]))
        y = -step_gap
    else:
        seq.append(np.array([0, first_y, 0, 0]))
        y = step_gap
    x = z = 0.0
    cc = 2 if u[57] < 0.5 else 3
    for i in range(1, num_steps):
        x += step_size
        y *= -1
        if i > cc:
            z += step_height
        seq.append(np.array([x, y, z, 0]))
    # transform_sequence :113-135
    lf, rf = fk["xpos"][0, ids.lfoot], fk["xpos"][0, ids.rfoot]
    yaw = tf3.quat2euler(fk["xquat"][0, ids.root])[2]
    mid = (lf + rf) / 2
    ts.seq_len = len(seq)
    for k, (sx, sy, sz, th) in enumerate(seq):
        ts.sequence[k] = [mid[0] + sx * np.cos(yaw) - sy * np.sin(yaw),
                          mid[1] + sx * np.sin(yaw) + sy * np.cos(yaw), sz, yaw + th]
    ts.t1 = ts.t2                                            # update_target_steps
    ts.t2 += 1
    if ts.t2 == ts.seq_len:
        ts.t2 = ts.seq_len - 1
    qn = qpos.copy()
    qn[3:7] = qn[3:7] / np.linalg.norm(qn[3:7])
    return qpos, qvel, ts, get_obs(qn, qvel, ts)
