"""Batched ``RobotSystem`` — the reference's rigid-body API with a leading batch dimension.

Mirrors ``RobotSystem`` / ``PinocchioRobotSystem`` (reference pnc/robot_system/robot_system.py:9-393,
pnc/robot_system/pinocchio_robot_system.py:10-278): same constructor arguments, property and
method names and data conventions; every array gains a leading batch dimension ``B`` and lives
on the GPU as a ``torch.float64`` tensor; the per-joint dictionaries of ``update_system`` become
``[B, n_a]`` tensors in ``joint_id`` order (``joint_id[name] -> column`` is preserved).

All numbers come from ``k_update_system`` through the C ABI; there is no CPU path.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import engine
from .urdf_model import RobotModel, build_model_from_urdf


class BatchedRobotSystem:
    def __init__(self, urdf_file, package_name=None, b_fixed_base=False, b_print_robot_info=False, capacity=1024,
                 device=0, links=()):
        """urdf_file: path of a URDF, or an already compiled :class:`RobotModel`.
        capacity: largest batch the HBM workspace holds.  links: frames to materialise up front
        (others are added on first use)."""
        self._b_fixed_base = bool(b_fixed_base)
        self._model = urdf_file if isinstance(urdf_file, RobotModel) else build_model_from_urdf(urdf_file, b_fixed_base)
        self._config_robot()
        self._device = int(device)
        self._capacity = int(capacity)
        self._dev_model = engine.Model(self._model, self._device)
        self._frames = [self._model.frame_id(l) for l in links]
        self._ws = None
        self._inputs = None
        self._joint_positions = None
        self._joint_velocities = None
        if b_print_robot_info:
            print("=" * 80)
            print("PnCRobot (batched, B200)")
            print("nq: ", self.n_q, ", nv: ", self.n_q_dot, ", na: ", self.n_a, ", nvirtual: ", self.n_floating)
            print("Joints:", list(self._joint_id.keys()))
            print("Links:", list(self._link_id.keys()))
            print("=" * 80)

    # ---- pinocchio_robot_system.py:33-93 ------------------------------------------------
    def _config_robot(self):
        m = self._model
        self._n_floating, self._n_q, self._n_q_dot, self._n_a = m.n_floating, m.nq, m.nv, m.na
        self._total_mass = m.total_mass
        self._joint_id = OrderedDict((name, i) for i, name in enumerate(m.joint_names))
        self._link_id = OrderedDict((name, i) for i, name in enumerate(m.frame_names))
        self._joint_pos_limit, self._joint_vel_limit = m.joint_pos_limit, m.joint_vel_limit
        self._joint_trq_limit = m.joint_trq_limit

    n_floating = property(lambda s: s._n_floating)
    n_q = property(lambda s: s._n_q)
    n_q_dot = property(lambda s: s._n_q_dot)
    n_a = property(lambda s: s._n_a)
    total_mass = property(lambda s: s._total_mass)
    joint_pos_limit = property(lambda s: s._joint_pos_limit)
    joint_vel_limit = property(lambda s: s._joint_vel_limit)
    joint_trq_limit = property(lambda s: s._joint_trq_limit)
    joint_id = property(lambda s: s._joint_id)
    link_id = property(lambda s: s._link_id)
    joint_positions = property(lambda s: s._joint_positions)
    joint_velocities = property(lambda s: s._joint_velocities)
    model = property(lambda s: s._model)
    dev_model = property(lambda s: s._dev_model)
    batch = property(lambda s: 0 if s._ws is None else s._ws.B)

    def _cent(self, k):
        assert getattr(self, '_b_cent', False), "call update_system(..., b_cent=True) first"
        return self._ws.centroidal()[k].clone()

    Ag = property(lambda s: s._cent(0))  # [B, 6, nv]  rows [angular; linear]  (:188-190)
    hg = property(lambda s: s._cent(1))  # [B, 6]                              (:184-186)
    Ig = property(lambda s: s._cent(2))  # [B, 6, 6]   diagonal blocks only    (:192-194)

    def get_q_idx(self, joint_id):  # :95-100
        if isinstance(joint_id, (list, tuple)):
            return [self._model.get_q_idx(j) for j in joint_id]
        return self._model.get_q_idx(joint_id)

    def get_q_dot_idx(self, joint_id):  # :102-107
        if isinstance(joint_id, (list, tuple)):
            return [self._model.get_q_dot_idx(j) for j in joint_id]
        return self._model.get_q_dot_idx(joint_id)

    def get_joint_idx(self, joint_id):  # :109-114
        if isinstance(joint_id, (list, tuple)):
            return [self._model.get_joint_idx(j) for j in joint_id]
        return self._model.get_joint_idx(joint_id)

    def create_cmd_ordered_dict(self, joint_pos_cmd, joint_vel_cmd, joint_trq_cmd):  # :116-126
        """[B, n_a] tensors -> {'joint_pos': {name: [B]}, ...} (views, no copies)."""
        cmd = OrderedDict()
        for key, val in (('joint_pos', joint_pos_cmd), ('joint_vel', joint_vel_cmd), ('joint_trq', joint_trq_cmd)):
            cmd[key] = OrderedDict((name, val[..., i]) for name, i in self._joint_id.items())
        return cmd

    # ---- workspace management -----------------------------------------------------------
    def workspace(self, links=()):
        """The HBM workspace, materialising `links` (names) in addition to the current frames."""
        want = list(self._frames)
        for l in links:
            f = l if isinstance(l, (int, np.integer)) else self._model.frame_id(l)
            if f not in want:
                want.append(int(f))
        if self._ws is None or want != self._frames:
            self._frames = want
            self._ws = engine.Workspace(self._dev_model, self._capacity, self._frames)
            if self._inputs is not None:
                self._ws.update_system(*self._inputs, b_cent=getattr(self, '_b_cent', False))
        return self._ws

    def _t(self, a, width):
        if a is None:
            return None
        t = torch.as_tensor(a, dtype=torch.float64)
        if t.dim() == 1:
            t = t[None, :]
        assert t.shape[-1] == width, (t.shape, width)
        return t.to('cuda:%d' % self._device).contiguous()

    # ---- pinocchio_robot_system.py:128-179 ----------------------------------------------
    def update_system(self, base_com_pos, base_com_quat, base_com_lin_vel, base_com_ang_vel, base_joint_pos,
                      base_joint_quat, base_joint_lin_vel, base_joint_ang_vel, joint_pos, joint_vel, b_cent=False):
        """Only base_joint_* and the joint arrays are used, as in the Pinocchio backend (:149-174).
        joint_pos / joint_vel: [B, n_a] in joint_id order, or dicts {name: [B]}."""
        if isinstance(joint_pos, dict):
            joint_pos = torch.stack([torch.as_tensor(joint_pos[k], dtype=torch.float64).reshape(-1)
                                     for k in self._joint_id], dim=-1)
            joint_vel = torch.stack([torch.as_tensor(joint_vel[k], dtype=torch.float64).reshape(-1)
                                     for k in self._joint_id], dim=-1)
        jp, jv = self._t(joint_pos, self.n_a), self._t(joint_vel, self.n_a)
        assert jp.shape[0] <= self._capacity, "batch exceeds the workspace capacity"
        if self._b_fixed_base:
            base = [None] * 4
        else:
            base = [self._t(base_joint_pos, 3), self._t(base_joint_quat, 4), self._t(base_joint_lin_vel, 3),
                    self._t(base_joint_ang_vel, 3)]
        self._inputs = base + [jp, jv]
        self._joint_positions, self._joint_velocities = jp, jv
        self._b_cent = bool(b_cent)
        ws = self.workspace()
        ws.update_system(*self._inputs, b_cent=self._b_cent)

    # ---- getters (:197-278); every one returns a fresh tensor like the reference's np.copy --
    def _get(self, name):
        assert self._ws is not None and self._ws.B > 0, "call update_system first"
        return self._ws.view(name).clone()

    def get_q(self):
        return self._get('q')

    def get_q_dot(self):
        return self._get('qdot')

    def get_mass_matrix(self):
        return self._get('M')

    def get_gravity(self):
        return self._get('grav')

    def get_coriolis(self):
        return self._get('cori')

    def get_com_pos(self):
        return self._get('com')

    def get_com_lin_vel(self):
        return self._get('vcom')

    def get_com_lin_jacobian(self):
        return self._get('jcom')

    def get_com_lin_jacobian_dot(self):
        return self._get('jcomdot')

    def _slot(self, link_id):
        f = self._model.frame_id(link_id)
        ws = self.workspace([f])
        return ws, ws.frames.index(f)

    def get_link_iso(self, link_id):
        """[B, 4, 4] homogeneous transforms."""
        ws, k = self._slot(link_id)
        iso = ws.view('frame_iso')[:, k]
        B = iso.shape[0]
        T = torch.zeros((B, 4, 4), dtype=torch.float64, device=iso.device)
        T[:, :3, :3] = iso[:, :9].reshape(B, 3, 3)
        T[:, :3, 3] = iso[:, 9:12]
        T[:, 3, 3] = 1.0
        return T

    def get_link_vel(self, link_id):
        ws, k = self._slot(link_id)
        return ws.view('frame_vel')[:, k].clone()

    def get_link_jacobian(self, link_id):
        ws, k = self._slot(link_id)
        return ws.view('frame_jac')[:, k].clone()

    def get_link_jacobian_dot_times_qdot(self, link_id):
        ws, k = self._slot(link_id)
        return ws.view('frame_jdqd')[:, k].clone()
