"""Batched ramp / trajectory managers and the DCM update — the generators that feed the tasks.

Mirrors (reference paths): ``TaskHierarchyManager`` pnc/wbc/manager/task_hierarchy_manager.py:4-35,
``ReactionForceManager`` reaction_force_manager.py:4-35, ``FootTrajectoryManager``
foot_trajectory_manager.py:11-117, ``FloatingBaseTrajectoryManager``
floating_base_trajectory_manager.py:10-115 and ``AtlasStateEstimator._update_dcm``
pnc/atlas_pnc/atlas_state_estimator.py:35-44.  Same class and method names and the same call
order as the reference's state machines use; times may be floats (all robots in step) or ``[B]``
tensors (every robot at its own phase).  Every number comes from the kernels of
``csrc/managers.cu`` through the C ABI; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import engine
from ._cabi import load_library

_vp = C.c_void_p
SWING_REC, FB_REC, HAND_REC = 26, 14, 24  # PNC_SWING_REC / PNC_FB_REC / PNC_HAND_REC of include/pnc_b200.h (checked by tests/test_cabi_and_host.py)


def _P(t):
    return _vp(t.data_ptr())


def _vecB(val, B, device):
    """float or [B] tensor -> contiguous [B] CUDA fp64 tensor."""
    if torch.is_tensor(val):
        t = val.to(device=device, dtype=torch.float64).reshape(-1)
        return t.expand(B).contiguous() if t.numel() == 1 else t.contiguous()
    return torch.full((B,), float(val), dtype=torch.float64, device=device)


class _Ramp:
    """out = (target - start) / duration * (clip(t, t0, t0 + duration) - t0) + start   (per state)"""

    def __init__(self):
        self._lib = load_library()
        self._start_time, self._duration, self._starting = 0., 0., None

    def _init(self, start_time, duration, current_value, B, device):
        self._start_time = _vecB(start_time, B, device)
        self._duration = _vecB(duration, B, device)
        self._starting = _vecB(current_value, B, device).clone()

    def _update(self, target, current_time):
        B = self._starting.shape[0]
        out = torch.empty_like(self._starting)
        engine._check(self._lib.pnc_ramp_update(B, _P(self._starting), float(target), _P(self._start_time),
                                                _P(self._duration), float(current_time), _P(out), None), 'pnc_ramp_update')
        return out


class TaskHierarchyManager(_Ramp):
    def __init__(self, task, w_max, w_min, batch=None, device='cuda:0'):
        super().__init__()
        self._task, self._w_max, self._w_min = task, w_max, w_min
        self._B, self._device = batch, device

    def _batch(self):
        w = self._task.w_hierarchy
        return self._B or (w.shape[0] if torch.is_tensor(w) else 1)

    def initialize_ramp_to_min(self, start_time, duration):  # :13-16
        self._init(start_time, duration, self._task.w_hierarchy, self._batch(), self._device)

    initialize_ramp_to_max = initialize_ramp_to_min  # :18-21, identical body in the reference

    def update_ramp_to_min(self, current_time):  # :23-28
        self._task.w_hierarchy = self._update(self._w_min, current_time)

    def update_ramp_to_max(self, current_time):  # :30-35
        self._task.w_hierarchy = self._update(self._w_max, current_time)


class ReactionForceManager(_Ramp):
    def __init__(self, contact, maximum_rf_z_max, batch=None, device='cuda:0'):
        super().__init__()
        self._contact, self._maximum_rf_z_max, self._minimum_rf_z_max = contact, maximum_rf_z_max, 0.001
        self._B, self._device = batch, device

    def _batch(self):
        r = self._contact.rf_z_max
        return self._B or (r.shape[0] if torch.is_tensor(r) else 1)

    def initialize_ramp_to_min(self, start_time, duration):  # :13-16
        self._init(start_time, duration, self._contact.rf_z_max, self._batch(), self._device)

    initialize_ramp_to_max = initialize_ramp_to_min

    def update_ramp_to_min(self, current_time):  # :23-28
        self._contact.rf_z_max = self._update(self._minimum_rf_z_max, current_time)

    def update_ramp_to_max(self, current_time):  # :30-35
        self._contact.rf_z_max = self._update(self._maximum_rf_z_max, current_time)


def _outputs(B, device):
    mk = lambda w: torch.empty((B, w), dtype=torch.float64, device=device)
    return mk(3), mk(3), mk(3), mk(4), mk(3), mk(3)


class FootTrajectoryManager:
    """use_current, or initialize_swing_foot_trajectory -> update_swing_foot_desired."""

    def __init__(self, pos_task, ori_task, robot):
        assert pos_task.target_id == ori_task.target_id
        self._pos_task, self._ori_task, self._robot = pos_task, ori_task, robot
        self._target_id = pos_task.target_id
        self._lib = load_library()
        self._swing_start_time = self._swing_duration = self._rec = None
        self._swing_height = 0.05

    swing_height = property(lambda s: s._swing_height, lambda s, v: setattr(s, '_swing_height', v))

    def _ws(self):
        ws, _ = self._robot._slot(self._target_id)
        return ws, self._robot.model.frame_id(self._target_id)

    def use_current(self):  # :41-52
        ws, frame = self._ws()
        B = ws.B
        pos, vel, acc, quat, w, wd = _outputs(B, ws.view('com').device)
        engine._check(self._lib.pnc_foot_use_current(ws.h, B, frame, _P(pos), _P(vel), _P(acc), _P(quat), _P(w), _P(wd), None), 'pnc_foot_use_current')
        self._pos_task.update_desired(pos, vel, acc)
        self._ori_task.update_desired(quat, w, wd)

    def initialize_swing_foot_trajectory(self, start_time, swing_duration, landing_pos, landing_quat):  # :54-84
        """landing_pos [B,3], landing_quat [B,4] (the reference passes a Footstep; its pos / quat)."""
        ws, frame = self._ws()
        B, dev = ws.B, ws.view('com').device
        self._swing_start_time, self._swing_duration = _vecB(start_time, B, dev), _vecB(swing_duration, B, dev)
        lp = torch.as_tensor(landing_pos, dtype=torch.float64, device=dev).expand(B, 3).contiguous()
        lq = torch.as_tensor(landing_quat, dtype=torch.float64, device=dev).expand(B, 4).contiguous()
        self._rec = torch.empty((B, SWING_REC), dtype=torch.float64, device=dev)
        engine._check(self._lib.pnc_swing_init(ws.h, B, frame, _P(lp), _P(lq), _P(self._swing_duration),
                                               float(self._swing_height), _P(self._rec), None), 'pnc_swing_init')

    def update_swing_foot_desired(self, curr_time):  # :86-110
        B, dev = self._rec.shape[0], self._rec.device
        pos, vel, acc, quat, w, wd = _outputs(B, dev)
        engine._check(self._lib.pnc_swing_eval(B, _P(self._rec), _P(self._swing_start_time), _P(self._swing_duration),
                                               float(curr_time), _P(pos), _P(vel), _P(acc), _P(quat), _P(w), _P(wd), None), 'pnc_swing_eval')
        self._pos_task.update_desired(pos, vel, acc)
        self._ori_task.update_desired(quat, w, wd)


class FloatingBaseTrajectoryManager:
    def __init__(self, com_task, base_ori_task, robot):
        self._com_task, self._base_ori_task, self._robot = com_task, base_ori_task, robot
        self._lib = load_library()
        self._start_time = self._duration = self._rec = None
        self._amp, self._freq = np.zeros(3), np.zeros(3)
        self._b_swaying = False

    b_swaying = property(lambda s: s._b_swaying, lambda s, v: setattr(s, '_b_swaying', v))

    def _ws(self):
        ws, _ = self._robot._slot(self._base_ori_task.target_id)
        return ws, self._robot.model.frame_id(self._base_ori_task.target_id)

    def initialize_floating_base_interpolation_trajectory(self, start_time, duration, target_com_pos, target_base_quat):  # :35-54
        ws, frame = self._ws()
        B, dev = ws.B, ws.view('com').device
        self._start_time, self._duration = _vecB(start_time, B, dev), _vecB(duration, B, dev)
        tc = torch.as_tensor(target_com_pos, dtype=torch.float64, device=dev).expand(B, 3).contiguous()
        tq = torch.as_tensor(target_base_quat, dtype=torch.float64, device=dev).expand(B, 4).contiguous()
        self._rec = torch.empty((B, FB_REC), dtype=torch.float64, device=dev)
        engine._check(self._lib.pnc_floating_base_init(ws.h, B, frame, _P(tc), _P(tq), _P(self._rec), None), 'pnc_floating_base_init')

    def initialize_floating_base_swaying_trajectory(self, start_time, amp, freq):  # :60-70
        ws, _ = self._ws()
        B, dev = ws.B, ws.view('com').device
        self._start_time = _vecB(start_time, B, dev)
        self._amp, self._freq = np.array(amp, dtype=np.float64), np.array(freq, dtype=np.float64)
        self._ini_com_pos = ws.view('com').clone()

    def update_floating_base_desired(self, current_time):  # :72-115
        if self._b_swaying:
            B, dev = self._ini_com_pos.shape[0], self._ini_com_pos.device
            pos, vel, acc, _, _, _ = _outputs(B, dev)
            dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
            engine._check(self._lib.pnc_sinusoid_eval(B, _P(self._ini_com_pos), _P(self._start_time), dp(self._amp),
                                                      dp(self._freq), float(current_time), _P(pos), _P(vel), _P(acc), None), 'pnc_sinusoid_eval')
            self._com_task.update_desired(pos, vel, acc)
            if self._rec is None:
                return
        B, dev = self._rec.shape[0], self._rec.device
        pos, vel, acc, quat, w, wd = _outputs(B, dev)
        engine._check(self._lib.pnc_floating_base_eval(B, _P(self._rec), _P(self._start_time), _P(self._duration),
                                                       float(current_time), _P(pos), _P(vel), _P(acc), _P(quat), _P(w),
                                                       _P(wd), None), 'pnc_floating_base_eval')
        if not self._b_swaying:
            self._com_task.update_desired(pos, vel, acc)
        self._base_ori_task.update_desired(quat, w, wd)


class DCMEstimator:
    """The DCM part of ``AtlasStateEstimator.update`` (atlas_state_estimator.py:35-44): call after
    ``robot.update_system``.  ``dcm``, ``prev_dcm``, ``dcm_vel`` are [B,3] tensors (the state
    provider's fields)."""

    def __init__(self, robot, dt, alpha_dcm_vel=0.1):
        self._robot, self._dt, self._alpha = robot, float(dt), float(alpha_dcm_vel)
        self._lib = load_library()
        self.dcm = self.prev_dcm = self.dcm_vel = None

    def update(self):
        ws = self._robot.workspace()
        B, dev = ws.B, ws.view('com').device
        if self.dcm is None or self.dcm.shape[0] != B:
            self.dcm = torch.zeros((B, 3), dtype=torch.float64, device=dev)  # AtlasStateProvider starts at zero
            self.prev_dcm, self.dcm_vel = torch.zeros_like(self.dcm), torch.zeros_like(self.dcm)
        engine._check(self._lib.pnc_dcm_update(ws.h, B, self._dt, self._alpha, _P(self.dcm), _P(self.prev_dcm),
                                               _P(self.dcm_vel), None), 'pnc_dcm_update')


class PointFootTrajectoryManager(FootTrajectoryManager):
    """point_foot_trajectory_manager.py:10-99 (Laikago toes): the position half of
    ``FootTrajectoryManager`` -- same swing record (the raised mid foot still uses the interpolated
    orientation, :54-70), only the LINK_XYZ task is written."""

    def __init__(self, pos_task, robot):
        self._pos_task, self._ori_task, self._robot = pos_task, None, robot
        self._target_id = pos_task.target_id
        self._lib = load_library()
        self._swing_start_time = self._swing_duration = self._rec = None
        self._swing_height = 0.05

    def use_current(self):  # :37-45
        ws, frame = self._ws()
        B = ws.B
        pos, vel, acc, quat, w, wd = _outputs(B, ws.view('com').device)
        engine._check(self._lib.pnc_foot_use_current(ws.h, B, frame, _P(pos), _P(vel), _P(acc), _P(quat), _P(w), _P(wd), None),
                      'pnc_foot_use_current')
        self._pos_task.update_desired(pos, vel, acc)

    def update_swing_foot_desired(self, curr_time):  # :73-89
        B, dev = self._rec.shape[0], self._rec.device
        pos, vel, acc, quat, w, wd = _outputs(B, dev)
        engine._check(self._lib.pnc_swing_eval(B, _P(self._rec), _P(self._swing_start_time), _P(self._swing_duration),
                                               float(curr_time), _P(pos), _P(vel), _P(acc), _P(quat), _P(w), _P(wd), None),
                      'pnc_swing_eval')
        self._pos_task.update_desired(pos, vel, acc)


class HandTrajectoryManager:
    """pnc/wbc/manager/hand_trajectory_manager.py:7-178.  Target hand poses are passed as target_pos [B,3] and
    target_quat [B,4] (the reference passes a 4x4 target_hand_iso and takes util.rot_to_quat of its rotation)."""

    def __init__(self, pos_task, ori_task, robot):
        self._pos_task, self._ori_task, self._robot = pos_task, ori_task, robot
        self._target_id = pos_task.target_id
        self._lib = load_library()
        self._start_moving_time = self._moving_duration = self._rec = None

    def _ws(self):
        ws, _ = self._robot._slot(self._target_id)
        return ws, self._robot.model.frame_id(self._target_id)

    def _write(self, pos, vel, acc, quat, w, wd):
        self._pos_task.update_desired(pos, vel, acc)
        if self._ori_task is not None:
            self._ori_task.update_desired(quat, w, wd)

    def update_desired(self, target_pos, target_quat):  # :26-38
        ws, _ = self._ws()
        B, dev = ws.B, ws.view('com').device
        pos, vel, acc, quat, w, wd = _outputs(B, dev)
        pos.copy_(torch.as_tensor(target_pos, dtype=torch.float64, device=dev).expand(B, 3))
        quat.copy_(torch.as_tensor(target_quat, dtype=torch.float64, device=dev).expand(B, 4))
        vel.zero_(); acc.zero_(); w.zero_(); wd.zero_()
        self._write(pos, vel, acc, quat, w, wd)

    def use_current(self):  # :40-55
        ws, frame = self._ws()
        B = ws.B
        pos, vel, acc, quat, w, wd = _outputs(B, ws.view('com').device)
        engine._check(self._lib.pnc_foot_use_current(ws.h, B, frame, _P(pos), _P(vel), _P(acc), _P(quat), _P(w), _P(wd), None),
                      'pnc_foot_use_current')
        self._write(pos, vel, acc, quat, w, wd)

    def _init(self, start_time, duration, target_pos, target_quat, keypoint_pos):
        ws, frame = self._ws()
        B, dev = ws.B, ws.view('com').device
        self._start_moving_time, self._moving_duration = _vecB(start_time, B, dev), _vecB(duration, B, dev)
        tp = torch.as_tensor(target_pos, dtype=torch.float64, device=dev).expand(B, 3).contiguous()
        tq = torch.as_tensor(target_quat, dtype=torch.float64, device=dev).expand(B, 4).contiguous()
        kp = None if keypoint_pos is None else torch.as_tensor(keypoint_pos, dtype=torch.float64, device=dev).expand(B, 3).contiguous()
        self._rec = torch.empty((B, HAND_REC), dtype=torch.float64, device=dev)
        engine._check(self._lib.pnc_hand_init(ws.h, B, frame, _P(tp), _P(tq), None if kp is None else _P(kp), _P(self._rec), None),
                      'pnc_hand_init')

    def initialize_hand_trajectory(self, start_time, duration, target_pos, target_quat):  # :57-77
        self._init(start_time, duration, target_pos, target_quat, None)

    def initialize_keypoint_hand_trajectory(self, start_time, duration, keypoint_pos, target_pos, target_quat):  # :79-102
        self._init(start_time, duration, target_pos, target_quat, keypoint_pos)

    def _update(self, current_time, keypoint):
        B, dev = self._rec.shape[0], self._rec.device
        pos, vel, acc, quat, w, wd = _outputs(B, dev)
        engine._check(self._lib.pnc_hand_eval(B, _P(self._rec), _P(self._start_moving_time), _P(self._moving_duration),
                                              float(current_time), int(keypoint), _P(pos), _P(vel), _P(acc), _P(quat), _P(w),
                                              _P(wd), None), 'pnc_hand_eval')
        self._write(pos, vel, acc, quat, w, wd)

    def update_hand_trajectory(self, current_time):  # :104-129
        self._update(current_time, 0)

    def update_keypoint_hand_trajectory(self, current_time):  # :131-178
        self._update(current_time, 1)


class UpperBodyTrajectoryManager:
    """upper_body_trajectory_manager.py:6-28: hold the SELECTED_JOINT task at the nominal pose."""

    def __init__(self, upper_body_task, robot):
        self._upper_body_task, self._robot = upper_body_task, robot

    task = property(lambda s: s._upper_body_task)

    def use_nominal_upper_body_joint_pos(self, nominal_joint_pos):
        """nominal_joint_pos: {joint name: float or [B] tensor}."""
        vals = [torch.as_tensor(nominal_joint_pos[k], dtype=torch.float64).reshape(-1) for k in self._upper_body_task.target_id]
        B = max(v.numel() for v in vals)
        pos = torch.stack([v.expand(B) for v in vals], dim=-1).cuda()
        self._upper_body_task.update_desired(pos, torch.zeros_like(pos), torch.zeros_like(pos))
