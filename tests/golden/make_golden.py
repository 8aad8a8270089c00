"""Generate golden input/output vectors by running the REFERENCE's own Python classes.

    python tests/golden/make_golden.py            (needs /root/reference; run in the build container)

What runs from /root/reference, unmodified:
  pnc/wbc/ihwbc/ihwbc.py            IHWBC (cost/constraint assembly, torque recovery)
  pnc/wbc/basic_task.py             BasicTask.update_jacobian / update_cmd (scipy Rotation path included)
  pnc/wbc/basic_contact.py          PointContact / SurfaceContact
  pnc/<robot>_pnc/*_tci_container.py, *_controller.py   task lists, gains, Sa, torque limits
  pnc/draco3*_pnc/*_rolling_joint_constraint.py
  external_source/Goldfarb/QuadProg++.cc  (compiled into oracle/_ref) as the QP behind
                                           `qpsolvers.solve_qp(..., solver="quadprog")`
What is substituted (not installable here):
  pinocchio  -> a RobotSystem subclass whose getters call the C oracle (oracle/pnc_oracle.c);
  qpsolvers  -> a stub module mapping solve_qp onto the reference's QuadProg++ (same
                Goldfarb-Idnani method as quadprog; CE=A^T, ce0=-b, CI=-G^T, ci0=h).

Output: tests/golden/<robot>.npz with the synthetic inputs and the reference outputs.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('PNC_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import oracle as O  # noqa: E402
from pypnc_b200 import robots  # noqa: E402

# ---- qpsolvers stub -> reference QuadProg++ --------------------------------------------
_qps = types.ModuleType('qpsolvers')


def _solve_qp(P, q, G=None, h=None, A=None, b=None, solver=None, verbose=False, **kw):
    return O.solve_qp_qpsolvers(np.array(P), np.array(q), G, h, A, b, use_ref=True)


_qps.solve_qp = _solve_qp
sys.modules['qpsolvers'] = _qps

from pnc.robot_system.robot_system import RobotSystem  # noqa: E402  (reference ABC)


class OracleRobotSystem(RobotSystem):
    """pinocchio_robot_system.py with the pinocchio calls replaced by the C oracle."""

    def __init__(self, model, b_fixed_base):
        self._m = model
        self._o = O.OracleModel(model)
        super().__init__(None, None, b_fixed_base, False)

    def _config_robot(self, urdf_file, package_dir):
        m = self._m
        self._n_floating, self._n_q, self._n_q_dot, self._n_a = m.n_floating, m.nq, m.nv, m.na
        for k, v in m.joint_id.items():
            self._joint_id[k] = v
        for k, v in m.link_id.items():
            self._link_id[k] = v
        self._total_mass = m.total_mass
        self._joint_pos_limit, self._joint_vel_limit = m.joint_pos_limit, m.joint_vel_limit
        self._joint_trq_limit = m.joint_trq_limit

    def get_q_idx(self, j):
        return [self.get_q_idx(x) for x in j] if type(j) is list else self._m.get_q_idx(j)

    def get_q_dot_idx(self, j):
        return [self.get_q_dot_idx(x) for x in j] if type(j) is list else self._m.get_q_dot_idx(j)

    def get_joint_idx(self, j):
        return [self.get_joint_idx(x) for x in j] if type(j) is list else self._m.get_joint_idx(j)

    def create_cmd_ordered_dict(self, *a):
        raise NotImplementedError

    def update_system(self, base_com_pos, base_com_quat, base_com_lin_vel, base_com_ang_vel, base_joint_pos,
                      base_joint_quat, base_joint_lin_vel, base_joint_ang_vel, joint_pos, joint_vel,
                      b_cent=False):
        jp = np.array([joint_pos[k] for k in self._joint_id])
        jv = np.array([joint_vel[k] for k in self._joint_id])
        self._q, self._q_dot = self._o.pack_state(base_joint_pos, base_joint_quat, base_joint_lin_vel,
                                                  base_joint_ang_vel, jp, jv)
        self._joint_positions, self._joint_velocities = jp, jv

    def _update_centroidal_quantities(self):
        self._Ag, self._hg, self._Ig = self._o.centroidal(self._q, self._q_dot)

    def get_q(self): return np.copy(self._q)
    def get_q_dot(self): return np.copy(self._q_dot)
    def get_mass_matrix(self): return self._o.mass_matrix(self._q)
    def get_gravity(self): return self._o.gravity(self._q)
    def get_coriolis(self): return self._o.coriolis(self._q, self._q_dot)
    def get_com_pos(self): return self._o.com(self._q, self._q_dot)[0]
    def get_com_lin_vel(self): return self._o.com(self._q, self._q_dot)[1]
    def get_com_lin_jacobian(self): return self._o.com_jacobian(self._q)
    def get_com_lin_jacobian_dot(self): return self._o.com_jacobian_dot(self._q, self._q_dot)
    def get_link_iso(self, link_id): return self._o.link_iso(self._q, self._m.frame_id(link_id))
    def get_link_vel(self, link_id): return self._o.link_vel(self._q, self._q_dot, self._m.frame_id(link_id))
    def get_link_jacobian(self, link_id): return self._o.link_jacobian(self._q, self._m.frame_id(link_id))

    def get_link_jacobian_dot_times_qdot(self, link_id):
        return self._o.link_jdqd(self._q, self._q_dot, self._m.frame_id(link_id))


def _reference_stack(name, robot, spec):
    """(task_list, contact_list, ic_list, ihwbc, sa) built by the reference's own container/controller."""
    if name == 'atlas':
        from config.atlas_config import PnCConfig
        PnCConfig.SAVE_DATA = False
        from pnc.atlas_pnc.atlas_tci_container import AtlasTCIContainer as TCI
        from pnc.atlas_pnc.atlas_controller import AtlasController as CTRL
    elif name == 'laikago':
        from config.laikago_config import PnCConfig
        PnCConfig.SAVE_DATA = False
        from pnc.laikago_pnc.laikago_tci_container import LaikagoTCIContainer as TCI
        from pnc.laikago_pnc.laikago_controller import LaikagoController as CTRL
    elif name == 'draco3':
        from config.draco3_config import PnCConfig
        PnCConfig.SAVE_DATA = False
        from pnc.draco3_pnc.draco3_tci_container import Draco3TCIContainer as TCI
        from pnc.draco3_pnc.draco3_controller import Draco3Controller as CTRL
    elif name == 'draco3_trans':
        # Draco3 through the reference's IHWBC2 with b_transmission_constraint, set up as
        # draco_manipulation_controller.py:17-51 does (its IHWBC2 line is commented out there): proximal knee
        # joints passive (Sv), Sd on the distal ones; task / contact / constraint lists of the Draco3 container
        from config.draco3_config import PnCConfig
        PnCConfig.SAVE_DATA = False
        from pnc.draco3_pnc.draco3_tci_container import Draco3TCIContainer
        from pnc.wbc.ihwbc.ihwbc2 import IHWBC2
        tci = Draco3TCIContainer(robot)
        l_jp, l_jd, r_jp, r_jd = robot.get_q_dot_idx(['l_knee_fe_jp', 'l_knee_fe_jd', 'r_knee_fe_jp', 'r_knee_fe_jd'])
        act_list = [False] * robot.n_floating + [True] * robot.n_a
        act_list[l_jp] = act_list[r_jp] = False
        nv = len(act_list)
        n_active = int(np.count_nonzero(act_list))
        n_passive = nv - n_active - 6
        sa, sv, sd = np.zeros((n_active, nv)), np.zeros((n_passive, nv)), np.zeros((n_passive, nv))
        j = k = 0
        for i in range(6, nv):
            if act_list[i]:
                sa[j, i] = 1.
                j += 1
            else:
                sv[k, i] = 1.
                k += 1
        sd[0, l_jd] = sd[1, r_jd] = 1.
        sf = np.zeros((6, nv)); sf[:, :6] = np.eye(6)
        ih = IHWBC2(sf, sa, sv, sd, False)
        ih.lambda_q_ddot, ih.lambda_rf = spec.lambda_q_ddot, spec.lambda_rf
        return tci.task_list, tci.contact_list, tci.internal_constraint_list, ih, sa
    elif name == 'draco3_lb':
        from config.draco3_lb_config import PnCConfig
        PnCConfig.SAVE_DATA = False
        from pnc.draco3_lb_pnc.draco3_lb_tci_container import Draco3LBTCIContainer as TCI
        from pnc.draco3_lb_pnc.draco3_lb_controller import Draco3LBController as CTRL
    else:  # valkyrie: no valkyrie_pnc in the reference -> reference classes, our task list
        from pnc.wbc.basic_task import BasicTask
        from pnc.wbc.basic_contact import SurfaceContact
        from pnc.wbc.ihwbc.ihwbc import IHWBC
        inv = {v: k for k, v in robots.TASK_TYPE_NAMES.items()}
        tasks = []
        for t in spec.tasks:
            bt = BasicTask(robot, inv[t['type']], t['dim'], t['target'], False)
            bt.kp = np.array(spec.kp[t['gain_off']:t['gain_off'] + t['dim']])
            bt.kd = np.array(spec.kd[t['gain_off']:t['gain_off'] + t['dim']])
            tasks.append(bt)
        contacts = [SurfaceContact(robot, c['link'], c['x'], c['y'], c['mu'], False) for c in spec.contacts]
        nv = robot.n_q_dot
        sa = np.zeros((robot.n_a, nv)); sa[:, 6:] = np.eye(robot.n_a)
        sf = np.zeros((6, nv)); sf[:, :6] = np.eye(6)
        ih = IHWBC(sf, sa, np.zeros((0, nv)), False)
        ih.trq_limit = np.dot(sa[:, 6:], robot.joint_trq_limit)
        ih.lambda_q_ddot, ih.lambda_rf = spec.lambda_q_ddot, spec.lambda_rf
        return tasks, contacts, [], ih, sa
    tci = TCI(robot)
    ctrl = CTRL(tci, robot)
    return tci.task_list, tci.contact_list, tci.internal_constraint_list, ctrl._ihwbc, ctrl._sa


def generate(name, B=6, seed=0, detail=True, suffix=''):
    """detail=True keeps the per-task / per-contact intermediates (large); detail=False only the inputs and
    the controller outputs trq / acc / rf -- used for the 128-state sets."""
    spec = robots.PROBLEMS[name]()
    model = spec.model
    robot = OracleRobotSystem(model, False)
    tasks, contacts, ics, ihwbc, sa = _reference_stack(name, robot, spec)
    # the reference container must describe the same problem as pypnc_b200/robots.py
    assert len(tasks) == len(spec.tasks) and len(contacts) == len(spec.contacts)
    for t, st in zip(tasks, spec.tasks):
        assert t.dim == st['dim'] and robots.TASK_TYPE_NAMES[t._task_type] == st['type']
        assert t.target_id == st['target'] or st['type'] == robots.TASK_COM
        np.testing.assert_array_equal(t.kp, spec.kp[st['gain_off']:st['gain_off'] + st['dim']])
        np.testing.assert_array_equal(t.kd, spec.kd[st['gain_off']:st['gain_off'] + st['dim']])
    if name != 'valkyrie':
        np.testing.assert_array_equal([t.w_hierarchy for t in tasks], spec.w_default)
    np.testing.assert_array_equal(np.nonzero(sa)[1], spec.act_idx)
    assert (ihwbc.trq_limit is not None) == spec.b_trq_limit
    if spec.b_trq_limit:
        np.testing.assert_array_equal(ihwbc.trq_limit, spec.trq_limit)
    assert ihwbc.lambda_q_ddot == spec.lambda_q_ddot and ihwbc.lambda_rf == spec.lambda_rf

    op = O.OracleProblem(spec)
    states = robots.synth_states(name, spec, B, seed)
    des, w, rfz = robots.synth_desireds(name, spec, states, op.task_actuals(states), seed)
    # cover the degenerate ramp-start value and the <=0 clamp of contact.py:37-41
    for b in range(2, B, 16):
        rfz[b, 0] = 1e-3
    for b in range(3, B, 16):
        rfz[b, -1] = -5.0
    n, p, mi = spec.qp_dims
    out = dict(trq=[], acc=[], rf=[], qddot=[], M=[], grav=[], cori=[], task_jac=[], task_jdqd=[], task_op=[],
               cone_mat=[], cone_vec=[])
    for b in range(B):
        jp = {k: states['joint_pos'][b, i] for k, i in model.joint_id.items()}
        jv = {k: states['joint_vel'][b, i] for k, i in model.joint_id.items()}
        robot.update_system(None, None, None, None, states['base_pos'][b], states['base_quat'][b],
                            states['base_lin_vel'][b], states['base_ang_vel'][b], jp, jv)
        # --- body of <Robot>Controller.get_command (atlas_controller.py:56-77)
        M = robot.get_mass_matrix()
        Minv = np.linalg.inv(M)
        c, g = robot.get_coriolis(), robot.get_gravity()
        ihwbc.update_setting(M, Minv, c, g)
        for i, (t, st) in enumerate(zip(tasks, spec.tasks)):
            o, d = st['des_off'], st['dim']
            npos = 4 if st['type'] == robots.TASK_LINK_ORI else d
            t.update_desired(des[b, o:o + npos].copy(), des[b, o + npos:o + npos + d].copy(),
                             des[b, o + npos + d:o + npos + 2 * d].copy())
            t.w_hierarchy = w[b, i]
            t.update_jacobian()
            t.update_cmd()
        ihwbc.w_hierarchy = np.array([t.w_hierarchy for t in tasks])
        for i, ct in enumerate(contacts):
            ct.rf_z_max = rfz[b, i]
            ct.update_contact()
        for ic in ics:
            ic.update_internal_constraint()
        if name == 'draco3_trans':
            trq, acc, rf = ihwbc.solve(tasks, contacts, ics, b_transmission_constraint=True)
        else:
            trq, acc, rf = ihwbc.solve(tasks, contacts, ics)
        out['trq'].append(trq); out['acc'].append(acc); out['rf'].append(rf)
        if not detail:
            continue
        out['M'].append(M); out['grav'].append(g); out['cori'].append(c)
        out['task_jac'].append(np.concatenate([t.jacobian for t in tasks], 0))
        out['task_jdqd'].append(np.concatenate([t.jacobian_dot_q_dot for t in tasks]))
        out['task_op'].append(np.concatenate([t.op_cmd for t in tasks]))
        out['cone_mat'].append(np.concatenate([ct.cone_constraint_mat.ravel() for ct in contacts]))
        out['cone_vec'].append(np.concatenate([ct.cone_constraint_vec for ct in contacts]))
    arrays = {('in_' + k): v for k, v in states.items()}
    arrays.update(in_des=des, in_w=w, in_rf_z_max=rfz)
    arrays.update({('out_' + k): np.array(v) for k, v in out.items() if len(v)})
    path = os.path.join(HERE, name + suffix + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-10s B=%d  |trq|max=%.1f  sum fz=%s  -> %s' %
          (name, B, np.abs(arrays['out_trq']).max(), np.round(arrays['out_rf'][0], 1)[:6], os.path.basename(path)))


if __name__ == '__main__':
    O.build()
    os.chdir(tempfile.mkdtemp())  # the reference's DataSaver writes ./data when enabled
    for nm in (sys.argv[1:] or ['atlas', 'laikago', 'draco3', 'draco3_lb', 'valkyrie', 'draco3_trans']):
        generate(nm)
        generate(nm, B=128, seed=100, detail=False, suffix='_128')
