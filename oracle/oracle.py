"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  See oracle/pnc_oracle.h for the parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pypnc_b200 import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

vp, i32, dbl = C.c_void_p, C.c_int, C.c_double


def build(force=False):
    """(Re)build the oracle libraries with oracle/Makefile."""
    if force or not os.path.exists(os.path.join(_HERE, 'libpnc_oracle.so')):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'CC=/usr/bin/gcc', 'CXX=/usr/bin/g++'])


def _load(fast=False):
    key = 'fast' if fast else 'strict'
    if key in _LIBS:
        return _LIBS[key]
    path = os.path.join(_HERE, 'libpnc_oracle_fast.so' if fast else 'libpnc_oracle.so')
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.orc_model_create.argtypes = [C.POINTER(_cabi.ModelDesc)]
    lib.orc_model_create.restype = vp
    lib.orc_model_destroy.argtypes = [vp]
    lib.orc_problem_create.argtypes = [vp, C.POINTER(_cabi.ProblemDesc)]
    lib.orc_problem_create.restype = vp
    lib.orc_problem_destroy.argtypes = [vp]
    lib.orc_solve_quadprog.restype = dbl
    lib.orc_solve_quadprog.argtypes = [i32, i32, i32] + [vp] * 13
    lib.orc_wbc_tick_batch.restype = dbl
    lib.orc_wbc_tick_batch.argtypes = [vp, i32, i32] + [vp] * 17
    lib.orc_osc_step_batch.restype = dbl
    lib.orc_osc_step_batch.argtypes = [vp, i32, i32] + [vp] * 8
    lib.orc_wbc_tick_batch_margin.restype = dbl
    lib.orc_wbc_tick_batch_margin.argtypes = [vp, i32, i32] + [vp] * 20
    _LIBS[key] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def load_ref_quadprog():
    """The REFERENCE's QuadProg++.cc compiled into oracle/_ref (None when absent)."""
    path = os.path.join(_HERE, '_ref', 'libref_quadprog.so')
    if not os.path.exists(path):
        try:
            build(force=True)
        except Exception:
            return None
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_solve_quadprog.restype = dbl
    lib.ref_solve_quadprog.argtypes = [i32, i32, i32] + [vp] * 8
    return lib


def ref_solve_quadprog(G, g0, CE, ce0, CI, ci0):
    """QuadProg++ convention: min .5 x'Gx + g0'x  s.t. CE'x + ce0 = 0, CI'x + ci0 >= 0."""
    lib = load_ref_quadprog()
    n = G.shape[0]
    p = CE.shape[1] if CE is not None and CE.size else 0
    m = CI.shape[1] if CI is not None and CI.size else 0
    G, g0 = _f64(G).copy(), _f64(g0)
    CE = _f64(CE) if p else np.zeros((n, 0))
    ce0 = _f64(ce0) if p else np.zeros(0)
    CI = _f64(CI) if m else np.zeros((n, 0))
    ci0 = _f64(ci0) if m else np.zeros(0)
    x = np.zeros(n)
    st = C.c_int(0)
    f = lib.ref_solve_quadprog(n, p, m, _p(G), _p(g0), _p(CE), _p(ce0), _p(CI), _p(ci0), _p(x), C.byref(st))
    return x, f, st.value


def solve_quadprog(G, g0, CE, ce0, CI, ci0):
    """oracle port of QuadProg++ (orc_solve_quadprog); returns dict(x,f,active,u,iters,status,flops)."""
    lib = _load()
    n = G.shape[0]
    p = CE.shape[1] if CE is not None and CE.size else 0
    m = CI.shape[1] if CI is not None and CI.size else 0
    G, g0 = _f64(G).copy(), _f64(g0)
    CE = _f64(CE) if p else np.zeros((n, 1))
    ce0 = _f64(ce0) if p else np.zeros(1)
    CI = _f64(CI) if m else np.zeros((n, 1))
    ci0 = _f64(ci0) if m else np.zeros(1)
    x = np.zeros(n)
    A = np.zeros(m + p + 1, dtype=np.int32)
    u = np.zeros(m + p + 1)
    iq, iters, status = C.c_int(0), C.c_int(0), C.c_int(0)
    flops = C.c_double(0)
    f = lib.orc_solve_quadprog(n, p, m, _p(G), _p(g0), _p(CE), _p(ce0), _p(CI), _p(ci0), _p(x), _p(A), _p(u),
                               C.byref(iq), C.byref(iters), C.byref(status), C.byref(flops))
    return dict(x=x, f=f, active=sorted(int(a) for a in A[p:iq.value]), A=A[:iq.value].copy(),
                u=u[:iq.value].copy(), iters=iters.value, status=status.value, flops=flops.value)


def solve_qp_qpsolvers(P, q, G=None, h=None, A=None, b=None, use_ref=False):
    """qpsolvers.solve_qp(P,q,G,h,A,b,solver='quadprog') signature on top of Goldfarb-Idnani
    (mapping of the qpsolvers quadprog shim: CE = A^T, ce0 = -b, CI = -G^T, ci0 = h)."""
    CE = np.ascontiguousarray(A.T) if A is not None else None
    ce0 = -np.asarray(b) if b is not None else None
    CI = np.ascontiguousarray(-G.T) if G is not None else None
    ci0 = np.asarray(h) if h is not None else None
    if use_ref:
        x, f, st = ref_solve_quadprog(P, q, CE, ce0, CI, ci0)
        return None if st != 0 else x
    r = solve_quadprog(P, q, CE, ce0, CI, ci0)
    return None if r['status'] != 0 else r['x']


class OracleModel:
    def __init__(self, model, fast=False):
        self.model = model
        self.lib = _load(fast)
        self._desc, self._keep = _cabi.make_model_desc(model)
        self.h = self.lib.orc_model_create(C.byref(self._desc))

    def __del__(self):
        try:
            self.lib.orc_model_destroy(self.h)
        except Exception:
            pass

    def pack_state(self, base_pos, base_quat, base_lin_vel, base_ang_vel, joint_pos, joint_vel):
        m = self.model
        q, v = np.zeros(m.nq), np.zeros(m.nv)
        args = [None if a is None else _f64(a) for a in (base_pos, base_quat, base_lin_vel, base_ang_vel)]
        jp, jv = _f64(joint_pos), _f64(joint_vel)
        self.lib.orc_pack_state(vp(self.h), *[_p(a) for a in args], _p(jp), _p(jv), _p(q), _p(v))
        return q, v

    def _call(self, name, q, v, shape, *extra):
        out = np.zeros(shape)
        args = [vp(self.h), _p(_f64(q))]
        if v is not None:
            args.append(_p(_f64(v)))
        args += list(extra) + [_p(out)]
        getattr(self.lib, name)(*args)
        return out

    def mass_matrix(self, q):
        return self._call('orc_mass_matrix', q, None, (self.model.nv, self.model.nv))

    def gravity(self, q):
        return self._call('orc_gravity', q, None, self.model.nv)

    def coriolis(self, q, v):
        return self._call('orc_coriolis', q, v, self.model.nv)

    def com(self, q, v):
        com, vcom = np.zeros(3), np.zeros(3)
        self.lib.orc_com(vp(self.h), _p(_f64(q)), _p(_f64(v)), _p(com), _p(vcom))
        return com, vcom

    def com_jacobian(self, q):
        return self._call('orc_com_jacobian', q, None, (3, self.model.nv))

    def com_jacobian_dot(self, q, v):
        return self._call('orc_com_jacobian_dot', q, v, (3, self.model.nv))

    def link_iso(self, q, frame):
        return self._call('orc_link_iso', q, None, (4, 4), i32(frame))

    def link_vel(self, q, v, frame):
        return self._call('orc_link_vel', q, v, 6, i32(frame))

    def link_jacobian(self, q, frame):
        return self._call('orc_link_jacobian', q, None, (6, self.model.nv), i32(frame))

    def link_jdqd(self, q, v, frame):
        return self._call('orc_link_jdqd', q, v, 6, i32(frame))

    def rnea(self, q, v, a, with_gravity=True):
        tau = np.zeros(self.model.nv)
        self.lib.orc_rnea(vp(self.h), _p(_f64(q)), _p(_f64(v)), _p(_f64(a)), i32(int(with_gravity)), _p(tau))
        return tau

    def centroidal(self, q, v):
        nv = self.model.nv
        Ag, hg, Ig = np.zeros((6, nv)), np.zeros(6), np.zeros((6, 6))
        self.lib.orc_centroidal(vp(self.h), _p(_f64(q)), _p(_f64(v)), _p(Ag), _p(hg), _p(Ig))
        return Ag, hg, Ig

    def osc_step(self, ee_frame, q, v, pos_des, vel_des, acc_des, kp, kd):
        nv = self.model.nv
        trq, qdd = np.zeros(nv), np.zeros(nv)
        self.lib.orc_osc_step(vp(self.h), i32(ee_frame), _p(_f64(q)), _p(_f64(v)), _p(_f64(pos_des)),
                              _p(_f64(vel_des)), _p(_f64(acc_des)), _p(_f64(kp)), _p(_f64(kd)), _p(trq), _p(qdd))
        return trq, qdd


    def osc_step_batch(self, ee_frame, q, v, pos_des, vel_des, kp, kd):
        """B OSC steps in a C loop -> (trq [B,nv], qddot [B,nv], seconds)."""
        B, nv = q.shape[0], self.model.nv
        trq, qdd = np.zeros((B, nv)), np.zeros((B, nv))
        sec = self.lib.orc_osc_step_batch(vp(self.h), i32(ee_frame), i32(B), _p(_f64(q)), _p(_f64(v)), _p(_f64(pos_des)),
                                          _p(_f64(vel_des)), _p(_f64(kp)), _p(_f64(kd)), _p(trq), _p(qdd))
        return trq, qdd, sec


class OracleProblem:
    def __init__(self, spec, fast=False):
        self.spec = spec
        self.om = OracleModel(spec.model, fast)
        self.lib = self.om.lib
        self._desc, self._keep = _cabi.make_problem_desc(spec)
        self.h = self.lib.orc_problem_create(vp(self.om.h), C.byref(self._desc))

    def __del__(self):
        try:
            self.lib.orc_problem_destroy(self.h)
        except Exception:
            pass

    def task_eval(self, q, v, des, it):
        d, nv = self.spec.tasks[it]['dim'], self.spec.model.nv
        jac, jdqd, op, perr = np.zeros((d, nv)), np.zeros(d), np.zeros(d), np.zeros(d)
        self.lib.orc_task_eval(vp(self.h), _p(_f64(q)), _p(_f64(v)), _p(_f64(des)), i32(it), _p(jac), _p(jdqd),
                               _p(op), _p(perr))
        return dict(jacobian=jac, jacobian_dot_q_dot=jdqd, op_cmd=op, pos_err=perr)

    def contact_eval(self, q, v, rf_z_max, ic):
        c = self.spec.contacts[ic]
        d, rows = (6, 18) if c['type'] == 1 else (3, 6)
        nv = self.spec.model.nv
        jac, jdqd, cm, cv = np.zeros((d, nv)), np.zeros(d), np.zeros((rows, d)), np.zeros(rows)
        self.lib.orc_contact_eval(vp(self.h), _p(_f64(q)), _p(_f64(v)), dbl(rf_z_max), i32(ic), _p(jac), _p(jdqd),
                                  _p(cm), _p(cv))
        return dict(jacobian=jac, jacobian_dot_q_dot=jdqd, cone_constraint_mat=cm, cone_constraint_vec=cv)

    def assemble(self, q, v, des, w, rf_z_max):
        n, p, m = self.spec.qp_dims
        P, qv, Gm, h = np.zeros((n, n)), np.zeros(n), np.zeros((max(m, 1), n)), np.zeros(max(m, 1))
        Am, b = np.zeros((p, n)), np.zeros(p)
        self.lib.orc_ihwbc_assemble(vp(self.h), _p(_f64(q)), _p(_f64(v)), _p(_f64(des)), _p(_f64(w)),
                                    _p(_f64(rf_z_max)), _p(P), _p(qv), _p(Gm), _p(h), _p(Am), _p(b))
        return dict(P=P, q=qv, G=Gm[:m], h=h[:m], A=Am, b=b)

    def tick_batch(self, states, des, w, rf_z_max, nthreads=1):
        """B controller ticks; returns dict of outputs + 'seconds'.  'margin' / 'margin_kind' are the
        decision margins of the QP solve (pnc_oracle.c: orc_last_qp_margin)."""
        spec, m = self.spec, self.spec.model
        B = states['joint_pos'].shape[0]
        n, p, mi = spec.qp_dims
        out = dict(trq=np.zeros((B, spec.nact)), acc=np.zeros((B, spec.nact)), rf=np.zeros((B, spec.nc)),
                   qddot=np.zeros((B, m.nv)), active=np.zeros((B, max(mi, 1)), dtype=np.uint8),
                   status=np.zeros(B, dtype=np.int32), iters=np.zeros(B, dtype=np.int32), flops=np.zeros(B),
                   margin=np.ones(B), margin_kind=np.zeros(B, dtype=np.int32), psi_rel=np.zeros(B))
        fl = m.n_floating != 0
        a = [_f64(states[k]) if fl else None for k in ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel')]
        jp, jv = _f64(states['joint_pos']), _f64(states['joint_vel'])
        des, w, rfz = _f64(des), _f64(w), _f64(rf_z_max)
        sec = self.lib.orc_wbc_tick_batch_margin(vp(self.h), B, nthreads, *[_p(x) for x in a], _p(jp), _p(jv),
                                                 _p(des), _p(w), _p(rfz), _p(out['trq']), _p(out['acc']),
                                                 _p(out['rf']), _p(out['qddot']), _p(out['active']),
                                                 _p(out['status']), _p(out['iters']), _p(out['flops']),
                                                 _p(out['margin']), _p(out['margin_kind']), _p(out['psi_rel']))
        out['active'] = out['active'][:, :mi]
        out['seconds'] = sec
        return out

    def task_actuals(self, states):
        """per-task actual pos/vel for synth_desireds (pos: [B,dim] or quat [B,4] for ORI)."""
        from scipy.spatial.transform import Rotation
        spec, m = self.spec, self.spec.model
        B = states['joint_pos'].shape[0]
        pos = [np.zeros((B, 4 if t['type'] == 3 else t['dim'])) for t in spec.tasks]
        vel = [np.zeros((B, t['dim'])) for t in spec.tasks]
        for b in range(B):
            q, v = self.om.pack_state(*[states[k][b] for k in ('base_pos', 'base_quat', 'base_lin_vel',
                                                               'base_ang_vel', 'joint_pos', 'joint_vel')])
            for i, t in enumerate(spec.tasks):
                if t['type'] == 2:
                    pos[i][b] = self.om.link_iso(q, t['frame'])[:3, 3]
                    vel[i][b] = self.om.link_vel(q, v, t['frame'])[3:]
                elif t['type'] == 3:
                    pos[i][b] = Rotation.from_matrix(self.om.link_iso(q, t['frame'])[:3, :3]).as_quat()
                    vel[i][b] = self.om.link_vel(q, v, t['frame'])[:3]
                elif t['type'] == 4:
                    pos[i][b], vel[i][b] = self.om.com(q, v)
                elif t['type'] == 1:
                    idx = [j - m.n_floating for j in spec.joint_sel[t['joint_off']:t['joint_off'] + t['dim']]]
                    pos[i][b], vel[i][b] = states['joint_pos'][b, idx], states['joint_vel'][b, idx]
                else:
                    pos[i][b], vel[i][b] = states['joint_pos'][b], states['joint_vel'][b]
        return dict(task_pos=pos, task_vel=vel)
