"""ctypes mirror of include/pnc_b200.h (descriptor structs + library loader).

The CUDA library is mandatory for the product path: :func:`load_library` raises
if ``libpnc_b200.so`` has not been built -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PNC_B200_LIB points at another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get('PNC_B200_LIB') or os.path.join(_HERE, 'csrc', 'libpnc_b200.so')

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class ModelDesc(C.Structure):
    _fields_ = [('nb', C.c_int32), ('nq', C.c_int32), ('nv', C.c_int32), ('na', C.c_int32),
                ('n_floating', C.c_int32), ('nf', C.c_int32), ('total_mass', C.c_double),
                ('mass_dyn', C.c_double), ('parent', c_int32_p), ('jtype', c_int32_p),
                ('idx_v', c_int32_p), ('subtree', c_int32_p), ('place_R', c_double_p),
                ('place_p', c_double_p), ('axis', c_double_p), ('mass', c_double_p), ('com', c_double_p),
                ('inertia', c_double_p), ('frame_parent', c_int32_p), ('frame_R', c_double_p),
                ('frame_p', c_double_p)]


class TaskDesc(C.Structure):
    _fields_ = [('type', C.c_int32), ('dim', C.c_int32), ('frame', C.c_int32), ('joint_off', C.c_int32),
                ('des_off', C.c_int32), ('gain_off', C.c_int32)]


class ContactDesc(C.Structure):
    _fields_ = [('type', C.c_int32), ('frame', C.c_int32), ('mu', C.c_double), ('x', C.c_double),
                ('y', C.c_double)]


class ProblemDesc(C.Structure):
    _fields_ = [('ntask', C.c_int32), ('tasks', C.POINTER(TaskDesc)), ('ncontact', C.c_int32),
                ('contacts', C.POINTER(ContactDesc)), ('n_joint_sel', C.c_int32), ('joint_sel', c_int32_p),
                ('n_gain', C.c_int32), ('kp', c_double_p), ('kd', c_double_p), ('n_ic', C.c_int32),
                ('ic_jac', c_double_p), ('nact', C.c_int32), ('act_idx', c_int32_p),
                ('b_trq_limit', C.c_int32), ('trq_limit', c_double_p), ('lambda_q_ddot', C.c_double),
                ('lambda_rf', C.c_double), ('ndes', C.c_int32), ('n_trans', C.c_int32),
                ('trans_distal', c_int32_p), ('trans_passive', c_int32_p)]


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


def make_model_desc(model):
    """RobotModel -> (ModelDesc, keepalive list)."""
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    keep = dict(parent=i32(model.parent), jtype=i32(model.jtype), idx_v=i32(model.idx_v),
                subtree=i32(model.subtree), place_R=f64(model.place_R), place_p=f64(model.place_p),
                axis=f64(model.axis), mass=f64(model.mass), com=f64(model.com), inertia=f64(model.inertia),
                frame_parent=i32(model.frame_parent), frame_R=f64(model.frame_R), frame_p=f64(model.frame_p))
    d = ModelDesc(nb=model.nb, nq=model.nq, nv=model.nv, na=model.na, n_floating=model.n_floating,
                  nf=len(model.frame_names), total_mass=model.total_mass, mass_dyn=model.mass_dyn,
                  parent=_ip(keep['parent']), jtype=_ip(keep['jtype']), idx_v=_ip(keep['idx_v']),
                  subtree=_ip(keep['subtree']), place_R=_dp(keep['place_R']), place_p=_dp(keep['place_p']),
                  axis=_dp(keep['axis']), mass=_dp(keep['mass']), com=_dp(keep['com']),
                  inertia=_dp(keep['inertia']), frame_parent=_ip(keep['frame_parent']),
                  frame_R=_dp(keep['frame_R']), frame_p=_dp(keep['frame_p']))
    return d, keep


def make_problem_desc(spec):
    """ProblemSpec -> (ProblemDesc, keepalive)."""
    tasks = (TaskDesc * max(1, len(spec.tasks)))()
    for i, t in enumerate(spec.tasks):
        tasks[i] = TaskDesc(t['type'], t['dim'], t['frame'], t['joint_off'], t['des_off'], t['gain_off'])
    contacts = (ContactDesc * max(1, len(spec.contacts)))()
    for i, c in enumerate(spec.contacts):
        contacts[i] = ContactDesc(c['type'], c['frame'], c['mu'], c['x'], c['y'])
    keep = dict(tasks=tasks, contacts=contacts,
                joint_sel=np.ascontiguousarray(spec.joint_sel, dtype=np.int32),
                kp=np.ascontiguousarray(spec.kp, dtype=np.float64),
                kd=np.ascontiguousarray(spec.kd, dtype=np.float64),
                ic_jac=np.ascontiguousarray(spec.ic_jac, dtype=np.float64),
                act_idx=np.ascontiguousarray(spec.act_idx, dtype=np.int32),
                trq_limit=np.ascontiguousarray(spec.trq_limit, dtype=np.float64),
                trans_d=np.ascontiguousarray(list(getattr(spec, 'trans_distal', [])) or [0], dtype=np.int32),
                trans_p=np.ascontiguousarray(list(getattr(spec, 'trans_passive', [])) or [0], dtype=np.int32))
    d = ProblemDesc(ntask=len(spec.tasks), tasks=tasks, ncontact=len(spec.contacts), contacts=contacts,
                    n_joint_sel=len(keep['joint_sel']), joint_sel=_ip(keep['joint_sel']),
                    n_gain=len(keep['kp']), kp=_dp(keep['kp']), kd=_dp(keep['kd']),
                    n_ic=spec.n_ic, ic_jac=_dp(keep['ic_jac']), nact=len(keep['act_idx']),
                    act_idx=_ip(keep['act_idx']), b_trq_limit=int(spec.b_trq_limit),
                    trq_limit=_dp(keep['trq_limit']), lambda_q_ddot=spec.lambda_q_ddot,
                    lambda_rf=spec.lambda_rf, ndes=spec.ndes, n_trans=len(getattr(spec, 'trans_distal', [])),
                    trans_distal=_ip(keep['trans_d']), trans_passive=_ip(keep['trans_p']))
    return d, keep


class Integrator(C.Structure):
    """pnc_integrator (include/pnc_b200.h)"""
    _fields_ = [('d_vel_state', C.c_void_p), ('d_pos_state', C.c_void_p), ('d_pos_limit', C.c_void_p),
                ('d_vel_limit', C.c_void_p), ('alpha_vel', C.c_double), ('alpha_pos', C.c_double), ('dt', C.c_double)]


TICK_DES_ON_DEVICE = 1

_lib = None


def load_library():
    """Load the CUDA C-ABI library; fail loudly when it is missing (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "pypnc_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the product path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.pnc_model_create.argtypes = [C.POINTER(ModelDesc), C.c_int, C.POINTER(vp)]
    lib.pnc_model_destroy.argtypes = [vp]
    lib.pnc_model_destroy.restype = None
    lib.pnc_problem_create.argtypes = [vp, C.POINTER(ProblemDesc), C.POINTER(vp)]
    lib.pnc_problem_destroy.argtypes = [vp]
    lib.pnc_problem_destroy.restype = None
    lib.pnc_problem_dims.argtypes = [vp] + [C.POINTER(i32)] * 5
    lib.pnc_workspace_create.argtypes = [vp, i32, c_int32_p, i32, C.POINTER(vp)]
    lib.pnc_workspace_destroy.argtypes = [vp]
    lib.pnc_workspace_destroy.restype = None
    lib.pnc_workspace_bytes.argtypes = [vp]
    lib.pnc_workspace_bytes.restype = i64
    lib.pnc_update_system.argtypes = [vp, i32] + [vp] * 6 + [vp]
    lib.pnc_update_system_cent.argtypes = [vp, i32] + [vp] * 6 + [i32, vp]
    for name in ('ag', 'hg', 'ig'):
        f = getattr(lib, 'pnc_ws_centroidal_' + name)
        f.argtypes = [vp]
        f.restype = vp
    for name in ('q', 'qdot', 'mass_matrix', 'gravity', 'coriolis', 'com_pos', 'com_vel', 'com_jac',
                 'com_jacdot', 'com_jdqd', 'frame_iso', 'frame_vel', 'frame_jac', 'frame_jdqd'):
        f = getattr(lib, 'pnc_ws_' + name)
        f.argtypes = [vp]
        f.restype = vp
    lib.pnc_ihwbc_solve.argtypes = [vp, vp, i32] + [vp] * 10 + [vp]
    lib.pnc_ihwbc_solve_ref.argtypes = [vp, vp, i32] + [vp] * 11 + [vp]
    lib.pnc_set_termination.argtypes = [i32]
    lib.pnc_get_termination.argtypes = []
    lib.pnc_set_precision.argtypes = [i32]
    lib.pnc_get_precision.argtypes = []
    lib.pnc_task_eval.argtypes = [vp, vp, i32, i32] + [vp] * 5 + [vp]
    lib.pnc_contact_eval.argtypes = [vp, vp, i32, i32] + [vp] * 4 + [vp]
    lib.pnc_osc_step.argtypes = [vp, i32, i32] + [vp] * 4 + [dbl, vp, vp, vp]
    lib.pnc_joint_integrate.argtypes = [i32, i32] + [vp] * 6 + [dbl, dbl, dbl, vp]
    lib.pnc_quadprog_batch.argtypes = [i32] * 5 + [vp] * 11 + [vp]
    lib.pnc_ramp_update.argtypes = [i32, vp, dbl, vp, vp, dbl, vp, vp]
    lib.pnc_foot_use_current.argtypes = [vp, i32, i32] + [vp] * 6 + [vp]
    lib.pnc_swing_init.argtypes = [vp, i32, i32, vp, vp, vp, dbl, vp, vp]
    lib.pnc_swing_eval.argtypes = [i32, vp, vp, vp, dbl] + [vp] * 6 + [vp]
    lib.pnc_hand_init.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.pnc_hand_eval.argtypes = [i32, vp, vp, vp, dbl, i32] + [vp] * 6 + [vp]
    lib.pnc_floating_base_init.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    lib.pnc_floating_base_eval.argtypes = [i32, vp, vp, vp, dbl] + [vp] * 6 + [vp]
    lib.pnc_sinusoid_eval.argtypes = [i32, vp, vp, c_double_p, c_double_p, dbl, vp, vp, vp, vp]
    lib.pnc_dcm_update.argtypes = [vp, i32, dbl, dbl, vp, vp, vp, vp]
    lib.pnc_wbc_tick_host.argtypes = [vp, vp, i32] + [vp] * 15
    lib.pnc_wbc_tick_host_ex.argtypes = [vp, vp, i32] + [vp] * 9 + [i32, vp, vp] + [vp] * 9
    lib.pnc_wbc_tick_host_multi.argtypes = [i32, vp, vp, i32] + [vp] * 15
    lib.pnc_launch_count.argtypes = []
    lib.pnc_launch_count.restype = i64
    lib.pnc_measure_fp64_peak.argtypes = [C.c_int, c_double_p]
    lib.pnc_debug_set_dump.argtypes = [vp, C.c_int]
    lib.pnc_debug_setup_record.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]
    lib.pnc_version.argtypes = []
    lib.pnc_version.restype = C.c_char_p
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    'pnc_model_create', 'pnc_model_destroy', 'pnc_problem_create', 'pnc_problem_destroy',
    'pnc_problem_dims', 'pnc_workspace_create', 'pnc_workspace_destroy', 'pnc_workspace_bytes',
    'pnc_update_system', 'pnc_update_system_cent', 'pnc_ws_centroidal_ag', 'pnc_ws_centroidal_hg',
    'pnc_ws_centroidal_ig', 'pnc_ws_q', 'pnc_ws_qdot', 'pnc_ws_mass_matrix', 'pnc_ws_gravity',
    'pnc_ws_coriolis', 'pnc_ws_com_pos', 'pnc_ws_com_vel', 'pnc_ws_com_jac', 'pnc_ws_com_jacdot',
    'pnc_ws_com_jdqd', 'pnc_ws_frame_iso', 'pnc_ws_frame_vel', 'pnc_ws_frame_jac', 'pnc_ws_frame_jdqd',
    'pnc_ihwbc_solve', 'pnc_ihwbc_solve_ref', 'pnc_set_termination', 'pnc_get_termination', 'pnc_set_precision', 'pnc_get_precision', 'pnc_task_eval', 'pnc_contact_eval', 'pnc_joint_integrate', 'pnc_osc_step', 'pnc_wbc_tick_host', 'pnc_wbc_tick_host_ex', 'pnc_wbc_tick_host_multi',
    'pnc_quadprog_batch', 'pnc_ramp_update', 'pnc_foot_use_current', 'pnc_swing_init', 'pnc_swing_eval', 'pnc_hand_init', 'pnc_hand_eval', 'pnc_floating_base_init',
    'pnc_floating_base_eval', 'pnc_sinusoid_eval', 'pnc_dcm_update',
    'pnc_launch_count', 'pnc_measure_fp64_peak', 'pnc_version', 'pnc_debug_set_dump', 'pnc_debug_setup_record',
]
