"""Batched whole-body controller — the reference's ``<Robot>Controller`` for B robots at once.

Mirrors ``AtlasController`` pnc/atlas_pnc/atlas_controller.py:9-94 and ``Draco3Controller``
pnc/draco3_pnc/draco3_controller.py:10-103 (which differ only in the passive joints and in mapping
the commands back through ``Sa[:, 6:]^T``): ``__init__`` builds Sa / Sv / Sf, the IHWBC with torque
limits and regularisation weights, and the joint integrator; ``get_command`` runs one control tick
and returns ``robot.create_cmd_ordered_dict(joint_pos_cmd, joint_vel_cmd, joint_trq_cmd)``.

The reference reads its constants from ``config/<robot>_config.py`` (``PnCConfig``, ``WBCConfig``);
here they are constructor arguments with the Atlas defaults.  Everything numeric runs in the CUDA
kernels behind ``IHWBC.solve`` and ``JointIntegrator.integrate``.
"""
from __future__ import annotations

import numpy as np
import torch

from .wbc import IHWBC, JointIntegrator


class TCIContainer:
    """task / contact / internal-constraint lists (pnc/wbc/tci_container.py:4-22)."""

    def __init__(self, task_list, contact_list, internal_constraint_list=()):
        self.task_list = list(task_list)
        self.contact_list = list(contact_list)
        self.internal_constraint_list = list(internal_constraint_list)


class Controller:
    def __init__(self, tci_container, robot, passive_joints=(), b_trq_limit=True, lambda_q_ddot=1e-8, lambda_rf=1e-7,
                 controller_dt=0.001, pos_cutoff_freq=1.0, vel_cutoff_freq=2.0, max_pos_err=0.2):
        """passive_joints: joint names without an actuator (Draco3: the two ``*_knee_fe_jd``,
        draco3_controller.py:17-21); config/atlas_config.py:64-77 supplies the defaults."""
        self._tci_container, self._robot = tci_container, robot
        act_list = [False] * robot.n_floating + [True] * robot.n_a  # atlas_controller.py:15-31
        for j in passive_joints:
            act_list[robot.get_q_dot_idx(j)] = False
        n_q_dot = len(act_list)
        n_active = int(np.count_nonzero(np.array(act_list)))
        n_passive = n_q_dot - n_active - 6
        self._sa, self._sv = np.zeros((n_active, n_q_dot)), np.zeros((n_passive, n_q_dot))
        j = k = 0
        for i in range(6, n_q_dot):
            if act_list[i]:
                self._sa[j, i] = 1.
                j += 1
            else:
                self._sv[k, i] = 1.
                k += 1
        self._sf = np.zeros((6, n_q_dot))
        self._sf[0:6, 0:6] = np.eye(6)
        self._ihwbc = IHWBC(self._sf, self._sa, self._sv)
        if b_trq_limit:  # :34-37
            self._ihwbc.trq_limit = np.dot(self._sa[:, 6:], robot.joint_trq_limit)
        self._ihwbc.lambda_q_ddot = lambda_q_ddot
        self._ihwbc.lambda_rf = lambda_rf
        self._joint_integrator = JointIntegrator(robot.n_a, controller_dt, device=robot._device)  # :41-47
        self._joint_integrator.pos_cutoff_freq = pos_cutoff_freq
        self._joint_integrator.vel_cutoff_freq = vel_cutoff_freq
        self._joint_integrator.max_pos_err = max_pos_err
        self._joint_integrator.joint_pos_limit = robot.joint_pos_limit
        self._joint_integrator.joint_vel_limit = robot.joint_vel_limit
        self._act_cols = torch.as_tensor(np.argmax(self._sa[:, 6:], axis=1), dtype=torch.long)
        self._b_first_visit = True

    ihwbc = property(lambda s: s._ihwbc)
    joint_integrator = property(lambda s: s._joint_integrator)

    def first_visit(self):  # :88-94
        self._joint_integrator.initialize_states(torch.zeros_like(self._robot.joint_positions), self._robot.joint_positions)
        self._b_first_visit = False

    def get_command(self):  # :51-86
        if self._b_first_visit:
            self.first_visit()
        # update_setting(M, Minv, c, g) is a no-op here: the kernels read them from the workspace
        # (task.update_jacobian / update_cmd are fused into the solve: nothing to do per task)
        self._ihwbc.update_setting()
        self._ihwbc.w_hierarchy = [t.w_hierarchy for t in self._tci_container.task_list]
        for contact in self._tci_container.contact_list:
            contact.update_contact()
        for ic in self._tci_container.internal_constraint_list:
            ic.update_internal_constraint()
        trq, acc, rf = self._ihwbc.solve(self._tci_container.task_list, self._tci_container.contact_list,
                                         self._tci_container.internal_constraint_list)
        self.rf_cmd = rf
        # Sa[:, 6:]^T maps the actuated commands back to all joints (draco3_controller.py:88-89);
        # with every joint actuated this is the identity (atlas_controller.py)
        B, na = trq.shape[0], self._robot.n_a
        cols = self._act_cols.to(trq.device)
        joint_trq_cmd = torch.zeros((B, na), dtype=torch.float64, device=trq.device)
        joint_acc_cmd = torch.zeros_like(joint_trq_cmd)
        joint_trq_cmd[:, cols] = trq
        joint_acc_cmd[:, cols] = acc
        joint_vel_cmd, joint_pos_cmd = self._joint_integrator.integrate(joint_acc_cmd, self._robot.joint_velocities,
                                                                        self._robot.joint_positions)
        return self._robot.create_cmd_ordered_dict(joint_pos_cmd, joint_vel_cmd, joint_trq_cmd)
