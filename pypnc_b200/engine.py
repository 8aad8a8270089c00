"""Thin Python handles over the C ABI (include/pnc_b200.h).

torch is used for device memory and streams only; every computation is a kernel of
``pypnc_b200/csrc/libpnc_b200.so``.  Nothing here has a CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi

vp = C.c_void_p


def _check(rc, what):
    if rc != 0:
        names = {-1: 'PNC_E_ARG', -2: 'PNC_E_CUDA', -3: 'PNC_E_SIZE', -4: 'PNC_E_NOGPU'}
        raise RuntimeError("%s failed: %s" % (what, names.get(rc, rc)))


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return vp(t.data_ptr())


def _stream():
    return vp(torch.cuda.current_stream().cuda_stream)


class Model:
    """Device copy of the compiled joint tables (PinocchioRobotSystem._config_robot)."""

    def __init__(self, model, device=0):
        self.lib = _cabi.load_library()
        self.model = model
        self.device = int(device)
        self._desc, self._keep = _cabi.make_model_desc(model)
        h = vp()
        _check(self.lib.pnc_model_create(C.byref(self._desc), self.device, C.byref(h)), 'pnc_model_create')
        self.h = h

    def __del__(self):
        try:
            self.lib.pnc_model_destroy(self.h)
        except Exception:
            pass


class Problem:
    """Tasks / contacts / internal constraints / actuation of one controller."""

    def __init__(self, dev_model, spec):
        self.lib = dev_model.lib
        self.dev_model = dev_model
        self.spec = spec
        self._desc, self._keep = _cabi.make_problem_desc(spec)
        h = vp()
        _check(self.lib.pnc_problem_create(dev_model.h, C.byref(self._desc), C.byref(h)), 'pnc_problem_create')
        self.h = h
        d = [C.c_int32() for _ in range(5)]
        _check(self.lib.pnc_problem_dims(h, *[C.byref(x) for x in d]), 'pnc_problem_dims')
        self.nc, self.ncone, self.n, self.p, self.m = [x.value for x in d]
        self.mw = max(1, (self.m + 63) // 64)

    def __del__(self):
        try:
            self.lib.pnc_problem_destroy(self.h)
        except Exception:
            pass


class Workspace:
    """Per-state rigid-body quantities in HBM for up to `capacity` states."""

    _FIELDS = dict(q='q', qdot='qdot', M='mass_matrix', grav='gravity', cori='coriolis', com='com_pos',
                   vcom='com_vel', jcom='com_jac', jcomdot='com_jacdot', jcom_jdqd='com_jdqd',
                   frame_iso='frame_iso', frame_vel='frame_vel', frame_jac='frame_jac', frame_jdqd='frame_jdqd')

    def __init__(self, dev_model, capacity, frames):
        self.lib = dev_model.lib
        self.dev_model = dev_model
        self.capacity = int(capacity)
        self.frames = [int(f) for f in frames]
        arr = np.ascontiguousarray(self.frames, dtype=np.int32)
        h = vp()
        _check(self.lib.pnc_workspace_create(dev_model.h, self.capacity, arr.ctypes.data_as(_cabi.c_int32_p),
                                             len(self.frames), C.byref(h)), 'pnc_workspace_create')
        self.h = h
        self.B = 0

    def __del__(self):
        try:
            self.lib.pnc_workspace_destroy(self.h)
        except Exception:
            pass

    @property
    def nbytes(self):
        return int(self.lib.pnc_workspace_bytes(self.h))

    def update_system(self, base_pos, base_quat, base_lin_vel, base_ang_vel, joint_pos, joint_vel, b_cent=False):
        B = joint_pos.shape[0]
        _check(self.lib.pnc_update_system_cent(self.h, B, _ptr(base_pos), _ptr(base_quat), _ptr(base_lin_vel),
                                               _ptr(base_ang_vel), _ptr(joint_pos), _ptr(joint_vel),
                                               1 if b_cent else 0, _stream()), 'pnc_update_system_cent')
        self.B = B

    def centroidal(self, B=None):
        """(Ag [B,6,nv], hg [B,6], Ig [B,6,6]) views; valid after update_system(..., b_cent=True)."""
        B = self.B if B is None else B
        nv, dev = self.dev_model.model.nv, self.dev_model.device
        return (_from_ptr(self.lib.pnc_ws_centroidal_ag(self.h), (B, 6, nv), dev),
                _from_ptr(self.lib.pnc_ws_centroidal_hg(self.h), (B, 6), dev),
                _from_ptr(self.lib.pnc_ws_centroidal_ig(self.h), (B, 6, 6), dev))

    def view(self, name, B=None):
        """A torch view (no copy) of one workspace array for the first B states."""
        m = self.dev_model.model
        B = self.B if B is None else B
        nfr, nv, nq = len(self.frames), m.nv, m.nq
        shapes = dict(q=(B, nq), qdot=(B, nv), M=(B, nv, nv), grav=(B, nv), cori=(B, nv), com=(B, 3), vcom=(B, 3),
                      jcom=(B, 3, nv), jcomdot=(B, 3, nv), jcom_jdqd=(B, 3), frame_iso=(B, nfr, 12),
                      frame_vel=(B, nfr, 6), frame_jac=(B, nfr, 6, nv), frame_jdqd=(B, nfr, 6))
        ptr = getattr(self.lib, 'pnc_ws_' + self._FIELDS[name])(self.h)
        return _from_ptr(ptr, shapes[name], self.dev_model.device)


def _from_ptr(ptr, shape, device):
    n = int(np.prod(shape))
    if n == 0:
        return torch.empty(shape, dtype=torch.float64, device='cuda:%d' % device)
    iface = {'shape': (n,), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}

    class _H:
        __cuda_array_interface__ = iface

    with torch.cuda.device(device):
        t = torch.as_tensor(_H(), device='cuda:%d' % device)
    return t.view(*shape)


TERM_REFINED, TERM_REFERENCE = 0, 1


def set_termination(mode):
    """Process-wide termination of the active-set loop: TERM_REFINED (default) or TERM_REFERENCE
    (stop exactly where QuadProg++.cc:306-312 stops)."""
    _check(_cabi.load_library().pnc_set_termination(int(mode)), 'pnc_set_termination')


PREC_FP64, PREC_FP32_DYNAMICS = 0, 1


def set_precision(mode):
    """Process-wide arithmetic of update_system: PREC_FP64 (default) or PREC_FP32_DYNAMICS, the optional fp32
    mode (rigid-body quantities computed in fp32, factorisations and active-set iterations in fp64)."""
    _check(_cabi.load_library().pnc_set_precision(int(mode)), 'pnc_set_precision')


def get_precision():
    return int(_cabi.load_library().pnc_get_precision())


def ihwbc_solve(problem, ws, des, w, rf_z_max, want_qddot=True):
    """pnc_ihwbc_solve_ref on device tensors.  Returns a dict of device tensors; 'active_ref' is the
    active set at the reference solver's own termination point (see include/pnc_b200.h)."""
    spec = problem.spec
    B = des.shape[0]
    dev = des.device
    f64 = dict(dtype=torch.float64, device=dev)
    out = dict(trq=torch.empty((B, spec.nact), **f64), acc=torch.empty((B, spec.nact), **f64),
               rf=torch.empty((B, problem.nc), **f64),
               qddot=torch.empty((B, spec.model.nv), **f64) if want_qddot else None,
               status=torch.empty(B, dtype=torch.int32, device=dev),
               iters=torch.empty(B, dtype=torch.int32, device=dev),
               active=torch.empty((B, problem.mw), dtype=torch.int64, device=dev),
               active_ref=torch.empty((B, problem.mw), dtype=torch.int64, device=dev))
    _check(problem.lib.pnc_ihwbc_solve_ref(problem.h, ws.h, B, _ptr(des), _ptr(w), _ptr(rf_z_max), _ptr(out['trq']),
                                           _ptr(out['acc']), _ptr(out['rf']), _ptr(out['qddot']),
                                           _ptr(out['status']), _ptr(out['iters']), _ptr(out['active']),
                                           _ptr(out['active_ref']), _stream()), 'pnc_ihwbc_solve_ref')
    return out


def unpack_active(active_words, m):
    """[B, mw] int64 bit masks -> [B, m] uint8 numpy."""
    a = active_words.detach().cpu().numpy().astype(np.uint64)
    B = a.shape[0]
    out = np.zeros((B, max(m, 1)), dtype=np.uint8)
    for i in range(m):
        out[:, i] = (a[:, i >> 6] >> np.uint64(i & 63)) & np.uint64(1)
    return out[:, :m]


def task_actuals_from_workspace(spec, ws):
    """Per-task actual position / velocity ([B,dim], quaternion [B,4] for LINK_ORI) read back from
    the workspace after update_system -- the inputs robots.synth_desireds needs."""
    from scipy.spatial.transform import Rotation
    m = spec.model
    iso = ws.view('frame_iso').cpu().numpy()
    vel = ws.view('frame_vel').cpu().numpy()
    com, vcom = ws.view('com').cpu().numpy(), ws.view('vcom').cpu().numpy()
    q, v = ws.view('q').cpu().numpy(), ws.view('qdot').cpu().numpy()
    B = q.shape[0]
    pos, vl = [], []
    for t in spec.tasks:
        if t['type'] in (2, 3):
            f = ws.frames.index(t['frame'])
            if t['type'] == 2:
                pos.append(iso[:, f, 9:12].copy()); vl.append(vel[:, f, 3:6].copy())
            else:
                pos.append(Rotation.from_matrix(iso[:, f, :9].reshape(B, 3, 3)).as_quat())
                vl.append(vel[:, f, 0:3].copy())
        elif t['type'] == 4:
            pos.append(com.copy()); vl.append(vcom.copy())
        else:
            idx = spec.joint_sel[t['joint_off']:t['joint_off'] + t['dim']] if t['type'] == 1 else \
                list(range(6, 6 + t['dim']))
            pos.append(q[:, [i + 1 for i in idx]].copy()); vl.append(v[:, idx].copy())
    return dict(task_pos=pos, task_vel=vl)


def quadprog_batch(G, g0, CE=None, ce0=None, CI=None, ci0=None, want_active=False):
    """Batched ``solve_quadprog(G, g0, CE, ce0, CI, ci0, x)`` (reference external_source/Goldfarb/
    QuadProg++.hh:70): min 1/2 x'Gx + g0'x  s.t. CE'x + ce0 = 0, CI'x + ci0 >= 0 for B problems.
    CUDA fp64 tensors: G [B,n,n], g0 [B,n], CE [B,n,p], ce0 [B,p], CI [B,n,m], ci0 [B,m] (one COLUMN
    per constraint, as QuadProg++'s matrices).  Returns (x [B,n], f [B], status [B] int32, iters [B]
    int32[, active [B, ceil(m/64)] int64 bit mask])."""
    lib = _cabi.load_library()
    assert G.is_cuda and G.dtype == torch.float64
    B, n = G.shape[0], G.shape[1]
    dev = G.device
    p = 0 if CE is None else CE.shape[2]
    m = 0 if CI is None else CI.shape[2]
    c = lambda t: None if t is None else t.contiguous()
    G, g0, CE, ce0, CI, ci0 = c(G), c(g0), c(CE), c(ce0), c(CI), c(ci0)
    x = torch.empty((B, n), dtype=torch.float64, device=dev)
    f = torch.empty((B,), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    iters = torch.empty((B,), dtype=torch.int32, device=dev)
    mw = max(1, (m + 63) // 64)
    active = torch.zeros((B, mw), dtype=torch.int64, device=dev) if want_active else None
    P = lambda t: None if t is None else _ptr(t)
    stream = vp(torch.cuda.current_stream(dev).cuda_stream)
    _check(lib.pnc_quadprog_batch(dev.index or 0, B, n, p, m, P(G), P(g0), P(CE), P(ce0), P(CI), P(ci0), P(x), P(f),
                                  P(status), P(iters), P(active), stream), 'pnc_quadprog_batch')
    return (x, f, status, iters, active) if want_active else (x, f, status, iters)


def tick_host_multi(problems, workspaces, host_in, host_out):
    """pnc_wbc_tick_host_multi: one host batch sharded over several GPUs from this process.
    problems / workspaces: one engine.Problem / engine.Workspace per device (same problem spec);
    host_in: the nine host arrays (numpy or pinned torch CPU tensors) base_pos, base_quat, base_lin_vel,
    base_ang_vel, joint_pos, joint_vel, des, w, rf_z_max with the GLOBAL batch first;
    host_out: trq, acc, rf, status (int32), iters (int32), active (uint64) host arrays (any may be None)."""
    lib = problems[0].lib
    n = len(problems)
    P = (vp * n)(*[p.h for p in problems])
    W = (vp * n)(*[w.h for w in workspaces])
    B = int(host_in[4].shape[0])

    def hp(a):
        if a is None:
            return None
        return vp(a.data_ptr()) if torch.is_tensor(a) else a.ctypes.data_as(vp)

    _check(lib.pnc_wbc_tick_host_multi(n, P, W, B, *[hp(a) for a in host_in], *[hp(a) for a in host_out]),
           'pnc_wbc_tick_host_multi')
