"""WBC problem definitions of the reference's robots, as data.

Each ``*_problem`` restates what the reference's ``<Robot>TCIContainer`` +
``<Robot>Controller.__init__`` declare (tasks, gains, contacts, internal
constraints, actuated set, torque limits, regularisation):

* Atlas    pnc/atlas_pnc/atlas_tci_container.py:10-99, atlas_controller.py:10-49,
           config/atlas_config.py:32-68
* Laikago  pnc/laikago_pnc/laikago_tci_container.py:10-81, config/laikago_config.py:23-48
* Draco3   pnc/draco3_pnc/draco3_tci_container.py:11-96, draco3_controller.py:17-44,
           config/draco3_config.py:30-72
* Draco3-LB pnc/draco3_lb_pnc/draco3_lb_tci_container.py:18-81, config/draco3_lb_config.py:28-56
* Valkyrie  no valkyrie_pnc exists in the reference; the task set is synthesised on
           the Atlas pattern with config/valkyrie_config.py:19-47 weights.

Also: nominal poses and the synthetic-state generator of SURVEY.md section 8(d).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from .urdf_model import RobotModel

TASK_JOINT, TASK_SELECTED_JOINT, TASK_LINK_XYZ, TASK_LINK_ORI, TASK_COM = range(5)
CONTACT_POINT, CONTACT_SURFACE = 0, 1
TASK_TYPE_NAMES = {"JOINT": TASK_JOINT, "SELECTED_JOINT": TASK_SELECTED_JOINT, "LINK_XYZ": TASK_LINK_XYZ,
                   "LINK_ORI": TASK_LINK_ORI, "COM": TASK_COM}

_MODELS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'models')


def load_model(name):
    """Load a compiled model table shipped in pypnc_b200/models (see compile_models.py)."""
    return RobotModel.load(os.path.join(_MODELS_DIR, name + '.json'))


@dataclass
class ProblemSpec:
    """Flat description of one IHWBC problem (mirrors pnc_problem_desc)."""
    model: RobotModel
    tasks: list = field(default_factory=list)
    contacts: list = field(default_factory=list)
    joint_sel: list = field(default_factory=list)
    kp: list = field(default_factory=list)
    kd: list = field(default_factory=list)
    ic_jac: np.ndarray = None
    n_ic: int = 0
    # IHWBC2 transmission rows (ihwbc2.py b_transmission_constraint): v indices of the (distal, passive) joints
    trans_distal: list = field(default_factory=list)
    trans_passive: list = field(default_factory=list)
    act_idx: list = field(default_factory=list)
    b_trq_limit: bool = False
    trq_limit: np.ndarray = None
    lambda_q_ddot: float = 0.
    lambda_rf: float = 0.
    ndes: int = 0
    w_default: list = field(default_factory=list)
    task_names: list = field(default_factory=list)

    # ---- builders --------------------------------------------------------
    def add_task(self, task_type, dim, target=None, kp=None, kd=None, w=1.0):
        m = self.model
        t = TASK_TYPE_NAMES[task_type] if isinstance(task_type, str) else task_type
        frame, joint_off = -1, 0
        if t in (TASK_LINK_XYZ, TASK_LINK_ORI):
            frame = m.frame_id(target)
        elif t == TASK_SELECTED_JOINT:
            joint_off = len(self.joint_sel)
            self.joint_sel += [m.get_q_dot_idx(j) for j in target]
            assert len(target) == dim
        kp = np.broadcast_to(np.asarray(kp, dtype=np.float64), (dim,))
        kd = np.broadcast_to(np.asarray(kd, dtype=np.float64), (dim,))
        npos = 4 if t == TASK_LINK_ORI else dim
        self.tasks.append(dict(type=t, dim=dim, frame=frame, joint_off=joint_off, des_off=self.ndes,
                               gain_off=len(self.kp), target=target))
        self.kp += list(kp)
        self.kd += list(kd)
        self.ndes += npos + 2 * dim
        self.w_default.append(float(w))
        self.task_names.append("%s:%s" % (task_type, target if isinstance(target, str) else 'joints'))
        return len(self.tasks) - 1

    def add_point_contact(self, link, mu):
        self.contacts.append(dict(type=CONTACT_POINT, frame=self.model.frame_id(link), mu=mu, x=0., y=0.,
                                  link=link))

    def add_surface_contact(self, link, x, y, mu):
        self.contacts.append(dict(type=CONTACT_SURFACE, frame=self.model.frame_id(link), mu=mu, x=x, y=y,
                                  link=link))

    def add_rolling_joint_constraint(self, pairs):
        """rows qdot_jp - qdot_jd = 0 (draco3_rolling_joint_constraint.py:9-16)"""
        nv = self.model.nv
        rows = np.zeros((len(pairs), nv))
        for r, (jp, jd) in enumerate(pairs):
            rows[r, self.model.get_q_dot_idx(jp)] = 1.
            rows[r, self.model.get_q_dot_idx(jd)] = -1.
        self.ic_jac = rows if self.ic_jac is None else np.concatenate([self.ic_jac, rows], 0)
        self.n_ic = self.ic_jac.shape[0]

    def set_actuation(self, passive_joints=(), b_trq_limit=False):
        """Sa / trq_limit as <Robot>Controller.__init__ builds them (atlas_controller.py:15-37)."""
        m = self.model
        passive = set(m.get_q_dot_idx(j) for j in passive_joints)
        self.act_idx = [i for i in range(6, m.nv) if i not in passive]
        self.b_trq_limit = bool(b_trq_limit)
        self.trq_limit = m.joint_trq_limit[[i - 6 for i in self.act_idx], :].copy()

    def finalize(self):
        if self.ic_jac is None:
            self.ic_jac = np.zeros((0, self.model.nv))
        if self.trq_limit is None:
            self.trq_limit = np.zeros((len(self.act_idx), 2))
        return self

    # ---- derived sizes ---------------------------------------------------
    @property
    def nc(self):
        return sum(6 if c['type'] == CONTACT_SURFACE else 3 for c in self.contacts)

    @property
    def ncone(self):
        return sum(18 if c['type'] == CONTACT_SURFACE else 6 for c in self.contacts)

    @property
    def nact(self):
        return len(self.act_idx)

    @property
    def qp_dims(self):
        n = self.model.nv + self.nc
        p = 6 + self.n_ic + len(self.trans_distal)
        m = self.ncone + (2 * self.nact if self.b_trq_limit else 0)
        return n, p, m

    @property
    def frames(self):
        """distinct model frames the problem needs materialised (task + contact links)"""
        out = []
        for t in self.tasks:
            if t['frame'] >= 0 and t['frame'] not in out:
                out.append(t['frame'])
        for c in self.contacts:
            if c['frame'] not in out:
                out.append(c['frame'])
        return out

    def algorithmic_flops_assembly(self):
        """F_asm of SURVEY.md 8(d): sum_tasks (2 d nv^2 + 2 d nv) + nv^2 (+ IC terms)."""
        nv = self.model.nv
        f = sum(2 * t['dim'] * nv * nv + 2 * t['dim'] * nv for t in self.tasks) + nv * nv
        if self.n_ic:
            f += 2 * self.n_ic * nv * nv + 2 * nv * nv * self.n_ic + 2 * nv ** 3
        return float(f)


# ---------------------------------------------------------------------------
def atlas_problem(model=None):
    m = model or load_model('atlas')
    s = ProblemSpec(m)
    s.add_task("COM", 3, 'com', [100.] * 3, [10.] * 3, 10.0)
    s.add_task("LINK_ORI", 3, 'pelvis_com', [100.] * 3, [10.] * 3, 20.0)
    sel = ["back_bkx", "back_bky", "back_bkz", "l_arm_elx", "l_arm_ely", "l_arm_shx", "l_arm_shz",
           "l_arm_wrx", "l_arm_wry", "l_arm_wry2", "neck_ry", "r_arm_elx", "r_arm_ely", "r_arm_shx",
           "r_arm_shz", "r_arm_wrx", "r_arm_wry", "r_arm_wry2"]
    s.add_task("SELECTED_JOINT", len(sel), sel, 100., 10., 0.1)
    s.add_task("LINK_XYZ", 3, 'r_sole', [400.] * 3, [40.] * 3, 60.0)
    s.add_task("LINK_XYZ", 3, 'l_sole', [400.] * 3, [40.] * 3, 60.0)
    s.add_task("LINK_ORI", 3, 'r_sole', [400.] * 3, [40.] * 3, 60.0)
    s.add_task("LINK_ORI", 3, 'l_sole', [400.] * 3, [40.] * 3, 60.0)
    s.add_task("LINK_XYZ", 3, 'r_hand', [70.] * 3, [5.] * 3, 0.0)
    s.add_task("LINK_XYZ", 3, 'l_hand', [70.] * 3, [5.] * 3, 0.0)
    s.add_surface_contact('r_sole', 0.11, 0.065, 0.3)
    s.add_surface_contact('l_sole', 0.11, 0.065, 0.3)
    s.set_actuation(b_trq_limit=True)
    s.lambda_q_ddot, s.lambda_rf = 1e-8, 1e-7
    return s.finalize()


def laikago_problem(model=None):
    m = model or load_model('laikago')
    s = ProblemSpec(m)
    s.add_task("COM", 3, 'com', [20.] * 3, [2.] * 3, 10.0)
    s.add_task("LINK_ORI", 3, 'chassis', [20.] * 3, [2.] * 3, 20.0)
    for toe in ('toeFL', 'toeFR', 'toeRR', 'toeRL'):
        s.add_task("LINK_XYZ", 3, toe, [40.] * 3, [4.] * 3, 60.0)
    for toe in ('toeFL', 'toeFR', 'toeRR', 'toeRL'):
        s.add_point_contact(toe, 0.4)
    s.set_actuation(b_trq_limit=False)
    s.lambda_q_ddot, s.lambda_rf = 1e-8, 1e-7
    return s.finalize()


def draco3_problem(model=None, b_trq_limit=False):
    m = model or load_model('draco3')
    s = ProblemSpec(m)
    s.add_task("COM", 3, 'com', [400.] * 3, [20.] * 3, 80.0)
    s.add_task("LINK_ORI", 3, 'torso_com_link', [100.] * 3, [10.] * 3, 80.0)
    upper = ['neck_pitch', 'l_shoulder_fe', 'l_shoulder_aa', 'l_shoulder_ie', 'l_elbow_fe', 'l_wrist_ps',
             'l_wrist_pitch', 'r_shoulder_fe', 'r_shoulder_aa', 'r_shoulder_ie', 'r_elbow_fe',
             'r_wrist_ps', 'r_wrist_pitch']
    s.add_task("SELECTED_JOINT", len(upper), upper,
               [40., 100., 100., 100., 50., 40., 50., 100., 100., 100., 50., 40., 50.],
               [2., 8., 8., 8., 3., 2., 3., 8., 8., 8., 3., 2., 3.], 20.0)
    s.add_task("LINK_XYZ", 3, 'r_foot_contact', [300.] * 3, [30.] * 3, 60.0)
    s.add_task("LINK_XYZ", 3, 'l_foot_contact', [300.] * 3, [30.] * 3, 60.0)
    s.add_task("LINK_ORI", 3, 'r_foot_contact', [300.] * 3, [30.] * 3, 60.0)
    s.add_task("LINK_ORI", 3, 'l_foot_contact', [300.] * 3, [30.] * 3, 60.0)
    s.add_surface_contact('r_foot_contact', 0.115, 0.065, 0.3)
    s.add_surface_contact('l_foot_contact', 0.115, 0.065, 0.3)
    s.add_rolling_joint_constraint([('l_knee_fe_jp', 'l_knee_fe_jd'), ('r_knee_fe_jp', 'r_knee_fe_jd')])
    s.set_actuation(passive_joints=('l_knee_fe_jd', 'r_knee_fe_jd'), b_trq_limit=b_trq_limit)
    s.lambda_q_ddot, s.lambda_rf = 1e-8, 1e-7
    return s.finalize()


def draco3_transmission_problem(model=None):
    """Draco3 through IHWBC2 with b_transmission_constraint (pnc/wbc/ihwbc/ihwbc2.py:131-139,249-256,386-403)
    as DracoManipulationController would set it up (draco_manipulation_controller.py:17-51, the commented
    IHWBC2 line): the PROXIMAL knee joints are passive (Sv), Sd selects the distal ones, the equality rows
    (Sd - Sv)(M qddot + c + g - Jc^T rf) = 0 replace the rolling-joint rows and nothing is projected."""
    m = model or load_model('draco3')
    s = draco3_problem(m)
    s.ic_jac, s.n_ic = np.zeros((0, m.nv)), 0
    s.set_actuation(passive_joints=('l_knee_fe_jp', 'r_knee_fe_jp'), b_trq_limit=False)
    s.trans_distal = [m.get_q_dot_idx('l_knee_fe_jd'), m.get_q_dot_idx('r_knee_fe_jd')]
    s.trans_passive = [m.get_q_dot_idx('l_knee_fe_jp'), m.get_q_dot_idx('r_knee_fe_jp')]
    return s.finalize()


def draco3_lb_problem(model=None):
    m = model or load_model('draco3_lb')
    s = ProblemSpec(m)
    s.add_task("COM", 3, 'com', [400.] * 3, [40.] * 3, 20.0)
    s.add_task("LINK_ORI", 3, 'torso_com_link', [100.] * 3, [10.] * 3, 20.0)
    s.add_task("LINK_XYZ", 3, 'r_foot_contact', [300.] * 3, [30.] * 3, 60.0)
    s.add_task("LINK_XYZ", 3, 'l_foot_contact', [300.] * 3, [30.] * 3, 60.0)
    s.add_task("LINK_ORI", 3, 'r_foot_contact', [300.] * 3, [30.] * 3, 60.0)
    s.add_task("LINK_ORI", 3, 'l_foot_contact', [300.] * 3, [30.] * 3, 60.0)
    s.add_surface_contact('r_foot_contact', 0.08, 0.02, 0.5)
    s.add_surface_contact('l_foot_contact', 0.08, 0.02, 0.5)
    s.add_rolling_joint_constraint([('l_knee_fe_jp', 'l_knee_fe_jd'), ('r_knee_fe_jp', 'r_knee_fe_jd')])
    s.set_actuation(passive_joints=('l_knee_fe_jd', 'r_knee_fe_jd'), b_trq_limit=True)
    s.lambda_q_ddot, s.lambda_rf = 1e-8, 1e-7
    return s.finalize()


def valkyrie_problem(model=None):
    m = model or load_model('valkyrie')
    s = ProblemSpec(m)
    s.add_task("COM", 3, 'com', [100.] * 3, [10.] * 3, 10.0)
    s.add_task("LINK_ORI", 3, 'pelvis', [100.] * 3, [10.] * 3, 20.0)
    upper = ['torsoYaw', 'torsoPitch', 'torsoRoll', 'lowerNeckPitch', 'neckYaw', 'upperNeckPitch',
             'leftShoulderPitch', 'leftShoulderRoll', 'leftShoulderYaw', 'leftElbowPitch',
             'rightShoulderPitch', 'rightShoulderRoll', 'rightShoulderYaw', 'rightElbowPitch']
    s.add_task("SELECTED_JOINT", len(upper), upper, 100., 10., 20.0)
    s.add_task("LINK_XYZ", 3, 'rightCOP_Frame', [400.] * 3, [40.] * 3, 40.0)
    s.add_task("LINK_XYZ", 3, 'leftCOP_Frame', [400.] * 3, [40.] * 3, 40.0)
    s.add_task("LINK_ORI", 3, 'rightCOP_Frame', [400.] * 3, [40.] * 3, 40.0)
    s.add_task("LINK_ORI", 3, 'leftCOP_Frame', [400.] * 3, [40.] * 3, 40.0)
    s.add_task("LINK_XYZ", 3, 'rightElbowPitchLink', [70.] * 3, [5.] * 3, 0.0)
    s.add_task("LINK_XYZ", 3, 'leftElbowPitchLink', [70.] * 3, [5.] * 3, 0.0)
    s.add_surface_contact('rightCOP_Frame', 0.11, 0.065, 0.3)
    s.add_surface_contact('leftCOP_Frame', 0.11, 0.065, 0.3)
    s.set_actuation(b_trq_limit=True)
    s.lambda_q_ddot, s.lambda_rf = 1e-8, 1e-8
    return s.finalize()


PROBLEMS = {'atlas': atlas_problem, 'laikago': laikago_problem, 'draco3': draco3_problem,
            'draco3_lb': draco3_lb_problem, 'valkyrie': valkyrie_problem,
            'draco3_trans': draco3_transmission_problem}


# ---------------------------------------------------------------------------
# nominal configurations and synthetic states (section 8d)
def nominal_pose(name, model):
    name = 'draco3' if name == 'draco3_trans' else name  # same robot, other WBC formulation
    jp = np.zeros(model.na)
    base_z, base_quat = 0.9, np.array([0., 0., 0., 1.])

    def setj(j, val):
        jp[model.joint_id[j]] = val

    if name == 'atlas':  # simulator/pybullet/atlas_dynamics_main.py:25-43
        setj('l_arm_shx', -np.pi / 4); setj('r_arm_shx', np.pi / 4)
        setj('l_arm_ely', -np.pi / 2); setj('r_arm_ely', np.pi / 2)
        setj('l_arm_elx', -np.pi / 2); setj('r_arm_elx', -np.pi / 2)
        for s in 'lr':
            setj(s + '_leg_hpy', -np.pi / 4); setj(s + '_leg_kny', np.pi / 2); setj(s + '_leg_aky', -np.pi / 4)
        base_z = 1.5 - 0.761
    elif name == 'laikago':  # simulator/pybullet/laikago_main.py:31-38
        for leg in ('FL', 'FR', 'RL', 'RR'):
            setj(leg + '_lower_leg_2_upper_leg_joint', -0.56)
        base_z, base_quat = 0.5, np.array([0.5, 0.5, 0.5, 0.5])
    elif name in ('draco3', 'draco3_lb'):  # simulator/pybullet/draco3_dynamics_main.py:32-57
        for s in 'lr':
            setj(s + '_hip_fe', -np.pi / 4); setj(s + '_knee_fe_jp', np.pi / 4); setj(s + '_knee_fe_jd', np.pi / 4)
            setj(s + '_ankle_fe', -np.pi / 4)
        if name == 'draco3':
            setj('l_shoulder_aa', np.pi / 6); setj('r_shoulder_aa', -np.pi / 6)
            setj('l_elbow_fe', -np.pi / 2); setj('r_elbow_fe', -np.pi / 2)
        base_z = 1.5 - 0.757
    elif name == 'valkyrie':  # simulator/pybullet/valkyrie_kinematics_main.py:58-70
        for s in ('left', 'right'):
            setj(s + 'HipPitch', -0.6); setj(s + 'KneePitch', 1.2); setj(s + 'AnklePitch', -0.6)
            setj(s + 'ShoulderPitch', 0.2)
        setj('leftShoulderRoll', -1.1); setj('rightShoulderRoll', 1.1)
        setj('leftElbowPitch', -1.57); setj('rightElbowPitch', 1.57)
        base_z = 1.1
    return jp, base_z, base_quat


def _quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


def _rotvec_to_quat(rv):
    th = np.linalg.norm(rv, axis=-1, keepdims=True)
    th_safe = np.where(th < 1e-12, 1.0, th)
    return np.concatenate([np.sin(th / 2) * rv / th_safe, np.cos(th / 2)], -1)


def synth_states(name, spec, B, seed=0):
    """Random-but-reasonable robot states of SURVEY.md 8(d).  Returns a dict of
    numpy fp64 arrays: base_pos [B,3], base_quat [B,4], base_lin_vel, base_ang_vel,
    joint_pos [B,na], joint_vel [B,na]."""
    model = spec.model
    rng = np.random.default_rng(seed)
    jp0, z0, q0 = nominal_pose(name, model)
    jp = jp0[None, :] + rng.uniform(-0.2, 0.2, (B, model.na))
    lo, hi = model.joint_pos_limit[:, 0], model.joint_pos_limit[:, 1]
    ok = hi > lo
    jp = np.where(ok[None, :], np.clip(jp, lo[None, :], hi[None, :]), jp)
    jv = rng.normal(0, 0.5, (B, model.na))
    if name.startswith('draco3'):  # keep the rolling contact consistent: jp == jd, qd_jp == qd_jd
        for s in 'lr':
            a, b = model.joint_id[s + '_knee_fe_jp'], model.joint_id[s + '_knee_fe_jd']
            jp[:, b] = jp[:, a]
            jv[:, b] = jv[:, a]
    base_pos = np.array([0., 0., z0])[None, :] + rng.normal(0, 0.05, (B, 3))
    base_quat = _quat_mul(_rotvec_to_quat(rng.normal(0, 0.15, (B, 3))), np.broadcast_to(q0, (B, 4)))
    return dict(base_pos=base_pos, base_quat=base_quat, base_lin_vel=rng.normal(0, 0.2, (B, 3)),
                base_ang_vel=rng.normal(0, 0.3, (B, 3)), joint_pos=jp, joint_vel=jv)


def synth_desireds(name, spec, states, actuals, seed=0, use_current_share=0.5):
    """Task desireds, w_hierarchy and rf_z_max per SURVEY.md 8(d).

    actuals: dict with per-task actual pos / vel ([B,dim] or quaternion [B,4] for ORI),
    i.e. task_pos[i], task_vel[i], computed by the caller from the robot system
    (oracle on CPU, or the CUDA workspace)."""
    rng = np.random.default_rng(seed + 7919)
    B = states['joint_pos'].shape[0]
    model = spec.model
    des = np.zeros((B, spec.ndes))
    w = np.tile(np.asarray(spec.w_default)[None, :], (B, 1))
    jp0, _, _ = nominal_pose(name, model)
    contact_frames = set(c['frame'] for c in spec.contacts)
    cur = rng.uniform(size=B) < use_current_share
    for i, t in enumerate(spec.tasks):
        d, o = t['dim'], t['des_off']
        pos_act, vel_act = actuals['task_pos'][i], actuals['task_vel'][i]
        if t['type'] == TASK_LINK_ORI:
            pos = _quat_mul(_rotvec_to_quat(rng.normal(0, 0.05, (B, 3))), pos_act)
            npos = 4
        elif t['type'] in (TASK_SELECTED_JOINT, TASK_JOINT):
            idx = [j - model.n_floating for j in spec.joint_sel[t['joint_off']:t['joint_off'] + d]] \
                if t['type'] == TASK_SELECTED_JOINT else list(range(d))
            pos = np.tile(jp0[idx][None, :], (B, 1))
            npos = d
        else:
            pos = pos_act + rng.normal(0, 0.02, (B, d))
            npos = d
        vel = vel_act + rng.normal(0, 0.05, (B, d))
        acc = rng.normal(0, 0.5, (B, d))
        if t['type'] in (TASK_SELECTED_JOINT, TASK_JOINT):
            vel = np.zeros((B, d))
            acc = np.zeros((B, d))
        if t['frame'] in contact_frames:  # feet: use_current() for a share of the states
            pos = np.where(cur[:, None], pos_act, pos)
            vel = np.where(cur[:, None], vel_act, vel)
            acc = np.where(cur[:, None], 0.0, acc)
            w[:, i] = rng.uniform(40., 60., B) if spec.w_default[i] == 60.0 else spec.w_default[i]
        des[:, o:o + npos] = pos
        des[:, o + npos:o + npos + d] = vel
        des[:, o + npos + d:o + npos + 2 * d] = acc
    # rf_z_max ~ U(200, 2000) N per foot on Atlas (182 kg, 2 feet); scaled by weight per contact elsewhere
    scale = model.total_mass / 182.4198 * 2.0 / max(1, len(spec.contacts))
    rfz = rng.uniform(200. * scale, 2000. * scale, (B, len(spec.contacts)))
    return des, w, rfz
