"""Batched tasks, contacts, internal constraints and ``IHWBC`` — the reference's WBC API.

Mirrors (reference paths): ``Task`` / ``BasicTask`` pnc/wbc/task.py:7-124, pnc/wbc/basic_task.py:10-137;
``Contact`` / ``PointContact`` / ``SurfaceContact`` pnc/wbc/contact.py:7-66, pnc/wbc/basic_contact.py:13-173;
``InternalConstraint`` pnc/wbc/internal_constraint.py:6-33 and the Draco3 rolling joint
pnc/draco3_pnc/draco3_rolling_joint_constraint.py:6-20; ``IHWBC`` pnc/wbc/ihwbc/ihwbc.py:12-379.

Names, argument order and error behaviour are the reference's; every per-state array has a
leading batch dimension and is a CUDA ``torch.float64`` tensor.  The objects only hold data;
``IHWBC.solve`` compiles them into one problem descriptor (cached) and makes ONE call into the
C ABI (``pnc_ihwbc_solve``), where task evaluation, assembly and the QP run fused.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import engine
from .robots import ProblemSpec, TASK_TYPE_NAMES, TASK_LINK_ORI, CONTACT_POINT, CONTACT_SURFACE

_vp = C.c_void_p


def _as_batch(val, width, B, device):
    t = torch.as_tensor(val, dtype=torch.float64, device=device)
    if t.dim() == 1:
        t = t[None, :].expand(B, -1)
    assert t.shape[-1] == width
    return t


class BasicTask:
    """basic_task.py:10-137.  task_type in JOINT / SELECTED_JOINT / LINK_XYZ / LINK_ORI / COM."""

    def __init__(self, robot, task_type, dim, target_id=None, data_save=False):
        if task_type not in TASK_TYPE_NAMES:
            raise ValueError("Wrong task type")  # basic_task.py:105
        self._robot, self._task_type, self._dim, self._target_id = robot, task_type, dim, target_id
        self._w_hierarchy = 1.0
        self._kp, self._kd = np.zeros(dim), np.zeros(dim)
        self._pos_des = self._vel_des = self._acc_des = None
        self._eval = None
        self._owner = None  # (IHWBC, index) once compiled into a problem

    target_id = property(lambda s: s._target_id)
    dim = property(lambda s: s._dim)
    pos_des = property(lambda s: s._pos_des)

    @property
    def kp(self):
        return self._kp

    @kp.setter
    def kp(self, value):
        assert value.shape[0] == self._dim
        self._kp = np.asarray(value, dtype=np.float64)

    @property
    def kd(self):
        return self._kd

    @kd.setter
    def kd(self, value):
        assert value.shape[0] == self._dim
        self._kd = np.asarray(value, dtype=np.float64)

    @property
    def w_hierarchy(self):
        return self._w_hierarchy

    @w_hierarchy.setter
    def w_hierarchy(self, value):
        """float, or a [B] tensor of per-state weights"""
        self._w_hierarchy = value

    def update_desired(self, pos_des, vel_des, acc_des):  # task.py:85-105
        assert vel_des.shape[-1] == self._dim
        assert acc_des.shape[-1] == self._dim
        self._pos_des, self._vel_des, self._acc_des = pos_des, vel_des, acc_des
        self._eval = None

    # op_cmd / pos_err / jacobian / jacobian_dot_q_dot are evaluated on demand by pnc_task_eval
    def _evaluate(self):
        if self._owner is None:
            raise RuntimeError("the task is evaluated by IHWBC.solve; call it (or IHWBC.compile) first")
        self._eval = self._owner[0]._task_eval(self._owner[1])

    def update_cmd(self):
        self._evaluate()

    def update_jacobian(self):
        if self._eval is None:
            self._evaluate()

    def _field(self, k):
        if self._eval is None:
            self._evaluate()
        return self._eval[k]

    op_cmd = property(lambda s: s._field('op_cmd'))
    pos_err = property(lambda s: s._field('pos_err'))
    jacobian = property(lambda s: s._field('jacobian'))
    jacobian_dot_q_dot = property(lambda s: s._field('jacobian_dot_q_dot'))


class _Contact:
    def __init__(self, robot, dim):
        self._robot, self._dim_contact = robot, dim
        self._rf_z_max = 0.0
        self._owner = None

    dim_contact = property(lambda s: s._dim_contact)

    @property
    def rf_z_max(self):
        return self._rf_z_max

    @rf_z_max.setter
    def rf_z_max(self, value):
        """float or [B] tensor; values <= 0 become 1e-3 (contact.py:37-41, applied in the kernel too)"""
        if not torch.is_tensor(value):
            value = 1e-3 if value <= 0. else value
        self._rf_z_max = value

    def update_contact(self):
        pass  # evaluated inside pnc_ihwbc_solve

    def _field(self, k):
        if self._owner is None:
            raise RuntimeError("the contact is evaluated by IHWBC.solve; call it (or IHWBC.compile) first")
        return self._owner[0]._contact_eval(self._owner[1])[k]

    jacobian = property(lambda s: s._field('jacobian'))
    cone_constraint_mat = property(lambda s: s._field('cone_constraint_mat'))
    cone_constraint_vec = property(lambda s: s._field('cone_constraint_vec'))


class PointContact(_Contact):  # basic_contact.py:13-51
    def __init__(self, robot, link_id, mu, data_save=False):
        super().__init__(robot, 3)
        self._link_id, self._mu, self._x, self._y, self._type = link_id, mu, 0., 0., CONTACT_POINT


class SurfaceContact(_Contact):  # basic_contact.py:54-173
    def __init__(self, robot, link_id, x, y, mu, data_save=False):
        super().__init__(robot, 6)
        self._link_id, self._mu, self._x, self._y, self._type = link_id, mu, x, y, CONTACT_SURFACE


class RollingJointConstraint:
    """draco3_rolling_joint_constraint.py:6-20: qdot[jp] - qdot[jd] = 0 for each (proximal, distal) pair."""

    def __init__(self, robot, pairs=(('l_knee_fe_jp', 'l_knee_fe_jd'), ('r_knee_fe_jp', 'r_knee_fe_jd'))):
        self._robot, self._pairs, self._dim = robot, list(pairs), len(pairs)
        nv = robot.n_q_dot
        self._jacobian = np.zeros((self._dim, nv))
        for r, (jp, jd) in enumerate(self._pairs):
            self._jacobian[r, robot.get_q_dot_idx(jp)] = 1.
            self._jacobian[r, robot.get_q_dot_idx(jd)] = -1.
        self._jacobian_dot_q_dot = np.zeros(self._dim)

    jacobian = property(lambda s: s._jacobian)
    jacobian_dot_q_dot = property(lambda s: s._jacobian_dot_q_dot)

    def update_internal_constraint(self):
        pass


class IHWBC:
    """ihwbc.py:12-379.  Usage as in the reference: update_setting -> solve."""

    def __init__(self, sf, sa, sv, data_save=False):
        sa = np.asarray(sa)
        self._n_q_dot, self._n_active, self._n_passive = sa.shape[1], sa.shape[0], np.asarray(sv).shape[0]
        self._sf, self._sa, self._sv = sf, sa, sv
        assert np.all((sa == 0) | (sa == 1)) and np.all(sa.sum(1) == 1), "sa must select joints (one-hot rows)"
        self._act_idx = [int(np.argmax(r)) for r in sa]
        self._trq_limit = None
        self._lambda_q_ddot = self._lambda_rf = self._w_rf = 0.
        self._w_hierarchy = None  # per-task weights set on the IHWBC object (ihwbc.py:76-78); None -> the tasks' own
        self._key, self._problem, self._spec = None, None, None
        self._robot = None
        self.status = self.iters = self.active = None

    trq_limit = property(lambda s: s._trq_limit)
    lambda_q_ddot = property(lambda s: s._lambda_q_ddot)
    lambda_rf = property(lambda s: s._lambda_rf)
    w_hierarchy = property(lambda s: s._w_hierarchy)
    w_rf = property(lambda s: s._w_rf)

    @trq_limit.setter
    def trq_limit(self, val):
        assert val is None or val.shape[0] == self._n_active
        self._trq_limit = None if val is None else np.array(val, dtype=np.float64)
        self._key = None

    @lambda_q_ddot.setter
    def lambda_q_ddot(self, val):
        self._lambda_q_ddot, self._key = val, None

    @lambda_rf.setter
    def lambda_rf(self, val):
        self._lambda_rf, self._key = val, None

    @w_hierarchy.setter
    def w_hierarchy(self, val):
        """Sequence of per-task weights in task_list order (floats or [B] tensors), as the controllers
        set it (atlas_controller.py:64-69).  These are what the cost uses (ihwbc.py:171-173), not the
        tasks' own w_hierarchy; while it is unset the tasks' weights are used."""
        self._w_hierarchy = val

    @w_rf.setter
    def w_rf(self, val):
        if val != 0.:
            raise NotImplementedError("w_rf != 0 is never used by the reference's controllers (ihwbc.py:80-82)")
        self._w_rf = val

    def update_setting(self, mass_matrix=None, mass_matrix_inv=None, coriolis=None, gravity=None):
        """Kept for API compatibility (ihwbc.py:84-88): the kernels read M, c, g from the robot's
        workspace written by update_system; an explicit inverse is never needed."""

    # ---- descriptor compilation ----------------------------------------------------------
    def compile(self, task_list, contact_list, internal_constraint_list):
        contact_list = [] if contact_list is None else contact_list  # ihwbc.py:177,200-203: the no-contact QP
        internal_constraint_list = [] if internal_constraint_list is None else internal_constraint_list
        if not contact_list and internal_constraint_list:
            # the reference's no-contact torque recovery uses Ni instead of Ni^T (ihwbc.py:349-354), a quirk
            # no shipped controller reaches; refuse instead of silently picking one of the two
            raise NotImplementedError("internal constraints without contacts (ihwbc.py:349-354)")
        for ic in internal_constraint_list:
            # ihwbc.py:134-135,287-296 add S Ji^T Lambda (Jidot qdot) to the torque-limit rows; the kernels
            # carry no such term (it is zero for every constraint the reference ships)
            if np.any(np.asarray(ic.jacobian_dot_q_dot) != 0.):
                raise NotImplementedError("internal constraint with non-zero jacobian_dot_q_dot")
        robot = (task_list[0] if task_list else contact_list[0])._robot
        key = (tuple((t._task_type, t._dim, str(t._target_id), tuple(t._kp), tuple(t._kd)) for t in task_list),
               tuple((c._type, c._link_id, c._mu, c._x, c._y) for c in contact_list),
               tuple(ic.jacobian.tobytes() for ic in internal_constraint_list), id(robot), self._key_extra())
        if key == self._key:
            return
        spec = ProblemSpec(robot.model)
        for t in task_list:
            spec.add_task(t._task_type, t._dim, t._target_id if t._target_id is not None else 'com', t._kp, t._kd,
                          1.0)
        for c in contact_list:
            if c._type == CONTACT_POINT:
                spec.add_point_contact(c._link_id, c._mu)
            else:
                spec.add_surface_contact(c._link_id, c._x, c._y, c._mu)
        for ic in internal_constraint_list:
            spec.ic_jac = ic.jacobian if spec.ic_jac is None else np.concatenate([spec.ic_jac, ic.jacobian], 0)
            spec.n_ic = spec.ic_jac.shape[0]
        spec.act_idx = list(self._act_idx)
        spec.b_trq_limit = self._trq_limit is not None
        spec.trq_limit = self._trq_limit
        spec.lambda_q_ddot, spec.lambda_rf = float(self._lambda_q_ddot), float(self._lambda_rf)
        self._spec_hook(spec)
        spec.finalize()
        self._spec, self._robot = spec, robot
        self._problem = engine.Problem(robot.dev_model, spec)
        self._key = key
        self._tasks, self._contacts = list(task_list), list(contact_list)
        for i, t in enumerate(task_list):
            t._owner = (self, i)
        for i, c in enumerate(contact_list):
            c._owner = (self, i)

    def _key_extra(self):  # subclasses: whatever else the compiled problem depends on
        return None

    def _spec_hook(self, spec):  # subclasses: last changes to the problem description
        pass

    def _pack(self):
        robot, spec = self._robot, self._spec
        B = robot.batch
        dev = 'cuda:%d' % robot._device
        des = torch.zeros((B, spec.ndes), dtype=torch.float64, device=dev)
        w = torch.zeros((B, len(self._tasks)), dtype=torch.float64, device=dev)
        for i, (t, d) in enumerate(zip(self._tasks, spec.tasks)):
            npos = 4 if d['type'] == TASK_LINK_ORI else d['dim']
            o = d['des_off']
            if t._pos_des is None:
                raise RuntimeError("task %d has no desired values: call update_desired" % i)
            des[:, o:o + npos] = _as_batch(t._pos_des, npos, B, dev)
            des[:, o + npos:o + npos + d['dim']] = _as_batch(t._vel_des, d['dim'], B, dev)
            des[:, o + npos + d['dim']:o + npos + 2 * d['dim']] = _as_batch(t._acc_des, d['dim'], B, dev)
            wi = t._w_hierarchy if self._w_hierarchy is None else self._w_hierarchy[i]
            w[:, i] = torch.as_tensor(wi, dtype=torch.float64, device=dev)
        rfz = torch.zeros((B, max(1, len(self._contacts))), dtype=torch.float64, device=dev)
        for i, c in enumerate(self._contacts):
            rfz[:, i] = torch.as_tensor(c._rf_z_max, dtype=torch.float64, device=dev)
        return des, w, rfz

    def solve(self, task_list, contact_list, internal_constraint_list, rf_des=None, verbose=False):
        """-> (joint_trq_cmd [B, n_active], joint_acc_cmd [B, n_active], sol_rf [B, dim_contacts]).
        Per-state solver outcomes are in self.status / self.iters / self.active (bit masks)."""
        if rf_des is not None and self._w_rf != 0.:
            raise NotImplementedError("rf_des is only used with w_rf != 0")
        if self._w_hierarchy is not None and len(self._w_hierarchy) != len(task_list):
            raise IndexError("w_hierarchy has %d entries for %d tasks" % (len(self._w_hierarchy), len(task_list)))
        self.compile(task_list, contact_list, internal_constraint_list)
        ws = self._robot.workspace(self._spec.frames)
        des, w, rfz = self._pack()
        out = engine.ihwbc_solve(self._problem, ws, des, w, rfz)
        self.status, self.iters, self.active = out['status'], out['iters'], out['active']
        self._last = (des, rfz)
        if verbose:
            st = self.status.cpu().numpy()
            print("IHWBC: %d states, status counts %s, mean iterations %.2f" %
                  (st.size, np.bincount(st, minlength=5).tolist(), float(self.iters.double().mean())))
        return out['trq'], out['acc'], out['rf']

    # ---- accessors used by Task / Contact properties ------------------------------------
    def _task_eval(self, it):
        ws = self._robot.workspace(self._spec.frames)
        des, _, _ = self._pack()
        B, d, nv = des.shape[0], self._spec.tasks[it]['dim'], self._robot.n_q_dot
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=des.device)
        jac, jd, op, pe = z(B, d, nv), z(B, d), z(B, d), z(B, d)
        P = lambda t: _vp(t.data_ptr())
        engine._check(self._problem.lib.pnc_task_eval(self._problem.h, ws.h, B, it, P(des), P(jac), P(jd), P(op),
                                                      P(pe), engine._stream()), 'pnc_task_eval')
        return dict(jacobian=jac, jacobian_dot_q_dot=jd, op_cmd=op, pos_err=pe)

    def _contact_eval(self, ic):
        ws = self._robot.workspace(self._spec.frames)
        _, _, rfz = self._pack()
        c = self._spec.contacts[ic]
        d, rows = (6, 18) if c['type'] == CONTACT_SURFACE else (3, 6)
        B, nv = rfz.shape[0], self._robot.n_q_dot
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=rfz.device)
        jac, cm, cv = z(B, d, nv), z(B, rows, d), z(B, rows)
        P = lambda t: _vp(t.data_ptr())
        engine._check(self._problem.lib.pnc_contact_eval(self._problem.h, ws.h, B, ic, P(rfz), P(jac), P(cm), P(cv),
                                                         engine._stream()), 'pnc_contact_eval')
        return dict(jacobian=jac, cone_constraint_mat=cm, cone_constraint_vec=cv)


class IHWBC2(IHWBC):
    """pnc/wbc/ihwbc/ihwbc2.py:12-430: IHWBC plus the transmission-constraint variant.  ``sd`` selects the distal
    joint of every transmission; with ``solve(..., b_transmission_constraint=True)`` the equality rows
    (Sd - Sv)(M qddot + c + g - Jc^T rf) = 0 replace the internal-constraint rows, nothing is projected
    (Ni = I) and the internal force Sv(...) comes back through Ji^T in the torque recovery (:386-403).
    Without the flag it behaves exactly as IHWBC."""

    def __init__(self, sf, sa, sv, sd, data_save=False):
        super().__init__(sf, sa, sv, data_save)
        self._sd = np.asarray(sd)
        self._b_trans = False

    def _transmission_pairs(self, internal_constraint_list):
        if self._trq_limit is not None:
            # ihwbc2.py:296-333 refers to sa_ni_trc_bar_tr, which the transmission branch never defines
            raise NotImplementedError("the reference's transmission branch has no torque-limit rows (NameError in ihwbc2.py)")
        sv, sd = np.asarray(self._sv), self._sd
        assert sv.shape == sd.shape and np.all(sv.sum(1) == 1) and np.all(sd.sum(1) == 1)
        ji = np.concatenate([ic.jacobian for ic in internal_constraint_list], 0)
        pas, dis = [int(np.argmax(r)) for r in sv], [int(np.argmax(r)) for r in sd]
        # the torque recovery B Snf (g - Ji^T Sv g) reduces to "distal joint += internal force" when row k of Ji
        # carries +1 at passive joint k and -1 at distal joint k (the rolling joints, the reference's only case)
        for k in range(len(pas)):
            row = np.zeros(ji.shape[1])
            row[pas[k]], row[dis[k]] = 1., -1.
            if not np.array_equal(ji[k], row):
                raise NotImplementedError("transmission constraint with a general internal-constraint Jacobian")
        return pas, dis

    def _key_extra(self):
        return ('trans', tuple(self._pairs[0]), tuple(self._pairs[1])) if self._b_trans else None

    def _spec_hook(self, spec):
        if self._b_trans:
            spec.ic_jac, spec.n_ic = None, 0
            spec.trans_passive, spec.trans_distal = list(self._pairs[0]), list(self._pairs[1])

    def compile(self, task_list, contact_list, internal_constraint_list):
        if self._b_trans:
            self._pairs = self._transmission_pairs(internal_constraint_list)
        super().compile(task_list, contact_list, internal_constraint_list)

    def solve(self, task_list, contact_list, internal_constraint_list, rf_des=None, verbose=False,
              b_transmission_constraint=False):
        self._b_trans = bool(b_transmission_constraint)
        return super().solve(task_list, contact_list, internal_constraint_list, rf_des, verbose)


class JointIntegrator:
    """joint_integrator.py:6-82 on [B, n_joint] device tensors."""

    def __init__(self, n_joint, dt, device=None):
        from . import _cabi
        self._lib = _cabi.load_library()
        self._n_joint, self._dt = n_joint, dt
        self._device = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
        self._alpha_pos = self._alpha_vel = 0.
        self._joint_pos_limit = self._joint_vel_limit = None
        self._pos = self._vel = None

    @staticmethod
    def _alpha(hz, dt):  # util/filters.py:4-8
        omega = 2 * np.pi * hz
        return float(np.clip((omega * dt) / (1. + (omega * dt)), 0., 1.))

    def _set_pos_cutoff(self, val):
        self._pos_cutoff_freq, self._alpha_pos = val, self._alpha(val, self._dt)

    def _set_vel_cutoff(self, val):
        self._vel_cutoff_freq, self._alpha_vel = val, self._alpha(val, self._dt)

    pos_cutoff_freq = property(lambda s: s._pos_cutoff_freq, _set_pos_cutoff)
    vel_cutoff_freq = property(lambda s: s._vel_cutoff_freq, _set_vel_cutoff)

    def _set_pos_limit(self, val):
        assert val.shape[0] == self._n_joint
        self._joint_pos_limit = torch.as_tensor(np.ascontiguousarray(val), dtype=torch.float64).to(self._device)

    def _set_vel_limit(self, val):
        assert val.shape[0] == self._n_joint
        self._joint_vel_limit = torch.as_tensor(np.ascontiguousarray(val), dtype=torch.float64).to(self._device)

    joint_pos_limit = property(lambda s: s._joint_pos_limit, _set_pos_limit)
    joint_vel_limit = property(lambda s: s._joint_vel_limit, _set_vel_limit)

    def initialize_states(self, init_vel, init_pos):
        assert init_vel.shape[-1] == self._n_joint and init_pos.shape[-1] == self._n_joint
        self._vel, self._pos = init_vel.clone().contiguous(), init_pos.clone().contiguous()

    def integrate(self, acc, vel, pos):
        B = acc.shape[0]
        P = lambda t: _vp(t.data_ptr())
        acc_c, pos_c = acc.contiguous(), pos.contiguous()  # kept alive until the launch is enqueued
        assert acc_c.device == self._vel.device == self._joint_pos_limit.device, "integrator state on another device"
        with torch.cuda.device(acc_c.device):
            engine._check(self._lib.pnc_joint_integrate(B, self._n_joint, P(acc_c), P(pos_c), P(self._vel),
                                                        P(self._pos), P(self._joint_pos_limit),
                                                        P(self._joint_vel_limit), self._alpha_vel, self._alpha_pos,
                                                        self._dt, engine._stream()), 'pnc_joint_integrate')
        del acc_c, pos_c
        return self._vel, self._pos
