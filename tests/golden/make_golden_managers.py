"""Golden vectors for the trajectory / ramp managers, produced by the REFERENCE's own classes.

    python tests/golden/make_golden_managers.py      (needs /root/reference; run in the build container)

Runs unmodified from /root/reference: pnc/wbc/manager/{task_hierarchy_manager, reaction_force_manager,
foot_trajectory_manager, point_foot_trajectory_manager, floating_base_trajectory_manager,
hand_trajectory_manager}.py, util/interpolation.py, util/util.py,
pnc/planner/locomotion/dcm_planner/footstep.py, and the DCM update of
pnc/atlas_pnc/atlas_state_estimator.py:35-44 (re-typed below because the module imports the
simulator config; the arithmetic is the three lines cited).  The robot is a stub that returns the
recorded link transforms / velocities / CoM; tasks and contacts are stubs that record what the
managers write.  Output: tests/golden/managers.npz
"""
import os
import sys

import numpy as np

REF = os.environ.get('PNC_REFERENCE', '/root/reference')
sys.path.insert(0, REF)

from scipy.spatial.transform import Rotation as R  # noqa: E402
from pnc.wbc.manager.task_hierarchy_manager import TaskHierarchyManager  # noqa: E402
from pnc.wbc.manager.reaction_force_manager import ReactionForceManager  # noqa: E402
from pnc.wbc.manager.foot_trajectory_manager import FootTrajectoryManager  # noqa: E402
from pnc.wbc.manager.point_foot_trajectory_manager import PointFootTrajectoryManager  # noqa: E402
from pnc.wbc.manager.floating_base_trajectory_manager import FloatingBaseTrajectoryManager  # noqa: E402
from pnc.wbc.manager.hand_trajectory_manager import HandTrajectoryManager  # noqa: E402
from pnc.planner.locomotion.dcm_planner.footstep import Footstep  # noqa: E402


class StubTask:
    def __init__(self, target_id, w=0.0):
        self.target_id = target_id
        self.w_hierarchy = w
        self.des = None

    def update_desired(self, pos, vel, acc):
        self.des = (np.array(pos, float), np.array(vel, float), np.array(acc, float))


class StubContact:
    def __init__(self, rfz):
        self.rf_z_max = rfz


class StubRobot:
    def __init__(self):
        self.iso, self.vel, self.com, self.vcom = {}, {}, np.zeros(3), np.zeros(3)

    def get_link_iso(self, k):
        return np.copy(self.iso[k])

    def get_link_vel(self, k):
        return np.copy(self.vel[k])

    def get_com_pos(self):
        return np.copy(self.com)

    def get_com_lin_vel(self):
        return np.copy(self.vcom)


def rand_iso(rng, scale=1.0):
    T = np.eye(4)
    T[:3, :3] = R.from_rotvec(rng.normal(0, scale, 3)).as_matrix()
    T[:3, 3] = rng.normal(0, 0.5, 3)
    return T


def main(n=96, seed=7):
    rng = np.random.default_rng(seed)
    out = {}
    # ---- ramps ----
    w0, wmax, wmin = rng.uniform(0, 80, n), 60.0, 40.0
    t0, dur, tc = rng.uniform(0, 2, n), rng.uniform(0.05, 1.0, n), rng.uniform(-0.5, 3.5, n)
    rfz0, rfzmax = rng.uniform(0.001, 2000, n), 1500.0
    r = {k: np.zeros(n) for k in ('w_to_min', 'w_to_max', 'rf_to_min', 'rf_to_max')}
    for i in range(n):
        for key, up in (('w_to_min', False), ('w_to_max', True)):
            task = StubTask('x', w0[i])
            m = TaskHierarchyManager(task, wmax, wmin)
            (m.initialize_ramp_to_max if up else m.initialize_ramp_to_min)(t0[i], dur[i])
            (m.update_ramp_to_max if up else m.update_ramp_to_min)(tc[i])
            r[key][i] = task.w_hierarchy
        for key, up in (('rf_to_min', False), ('rf_to_max', True)):
            c = StubContact(rfz0[i])
            m = ReactionForceManager(c, rfzmax)
            (m.initialize_ramp_to_max if up else m.initialize_ramp_to_min)(t0[i], dur[i])
            (m.update_ramp_to_max if up else m.update_ramp_to_min)(tc[i])
            r[key][i] = c.rf_z_max
    out.update(ramp_w0=w0, ramp_wmax=wmax, ramp_wmin=wmin, ramp_t0=t0, ramp_dur=dur, ramp_t=tc, ramp_rfz0=rfz0,
               ramp_rfzmax=rfzmax, **{'ramp_' + k: v for k, v in r.items()})

    # ---- foot trajectory manager ----
    iso = np.stack([rand_iso(rng, 0.6 if i % 4 else 1e-4) for i in range(n)])
    vel = rng.normal(0, 0.5, (n, 6))
    land_iso = np.stack([rand_iso(rng, 0.6 if i % 5 else 2e-6) for i in range(n)])
    for i in range(0, n, 7):  # a few landing feet with (almost) the initial orientation
        land_iso[i, :3, :3] = iso[i, :3, :3]
    sw_t0, sw_dur = rng.uniform(0, 1, n), rng.uniform(0.2, 1.0, n)
    sw_frac = np.concatenate([rng.uniform(-0.1, 1.1, n - 4), [0.0, 0.5, 1.0, 0.25]])
    height = 0.05
    uc = {k: [] for k in ('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc')}
    sw = {k: [] for k in ('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc')}
    pf = {k: [] for k in ('uc_pos', 'uc_vel', 'uc_acc', 'sw_pos', 'sw_vel', 'sw_acc')}
    land_quat = np.zeros((n, 4))
    for i in range(n):
        robot = StubRobot()
        robot.iso['foot'], robot.vel['foot'] = iso[i], vel[i]
        pt, ot = StubTask('foot'), StubTask('foot')
        m = FootTrajectoryManager(pt, ot, robot)
        m.use_current()
        for k, v in zip(('pos', 'vel', 'acc'), pt.des):
            uc[k].append(v)
        for k, v in zip(('quat', 'angvel', 'angacc'), ot.des):
            uc[k].append(v)
        land = Footstep()
        land.iso = land_iso[i]
        land.side = Footstep.LEFT_SIDE
        land_quat[i] = land.quat
        m.swing_height = height
        m.initialize_swing_foot_trajectory(sw_t0[i], sw_dur[i], land)
        m.update_swing_foot_desired(sw_t0[i] + sw_frac[i] * sw_dur[i])
        for k, v in zip(('pos', 'vel', 'acc'), pt.des):
            sw[k].append(v)
        for k, v in zip(('quat', 'angvel', 'angacc'), ot.des):
            sw[k].append(v)
        # the point-foot manager (Laikago toes): positions only, same landing foot
        pt2 = StubTask('foot')
        m2 = PointFootTrajectoryManager(pt2, robot)
        m2.use_current()
        for k, v in zip(('uc_pos', 'uc_vel', 'uc_acc'), pt2.des):
            pf[k].append(v)
        m2.swing_height = height
        m2.initialize_swing_foot_trajectory(sw_t0[i], sw_dur[i], land)
        m2.update_swing_foot_desired(sw_t0[i] + sw_frac[i] * sw_dur[i])
        for k, v in zip(('sw_pos', 'sw_vel', 'sw_acc'), pt2.des):
            pf[k].append(v)
    out.update(**{'pf_' + k: np.array(v) for k, v in pf.items()})
    out.update(foot_iso=iso, foot_vel=vel, land_pos=land_iso[:, :3, 3], land_quat=land_quat, sw_t0=sw_t0, sw_dur=sw_dur,
               sw_frac=sw_frac, sw_height=height, **{'uc_' + k: np.array(v) for k, v in uc.items()},
               **{'sw_' + k: np.array(v) for k, v in sw.items()})

    # ---- hand trajectory manager (hand_trajectory_manager.py) ----
    hrng = np.random.default_rng(seed + 1000)  # its own stream: the vectors above stay what they were
    h_iso = np.stack([rand_iso(hrng, 0.6 if i % 4 else 1e-4) for i in range(n)])
    h_vel = hrng.normal(0, 0.5, (n, 6))
    h_vel[::9, 0:3] *= 1e-7  # |angular velocity| < 1e-6: the other branch of HermiteCurveQuat.__init__
    h_tgt = np.stack([rand_iso(hrng, 0.6 if i % 5 else 2e-6) for i in range(n)])
    h_key = hrng.normal(0, 0.5, (n, 3))
    h_t0, h_dur = hrng.uniform(0, 1, n), hrng.uniform(0.2, 2.0, n)
    h_frac = np.concatenate([hrng.uniform(-0.1, 1.2, n - 4), [0.0, 0.5, 1.0, 0.25]])
    hd = {k: [] for k in ('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc', 'kp_pos', 'kp_vel', 'kp_acc', 'kp_quat', 'kp_angvel',
                          'kp_angacc', 'uc_pos', 'uc_vel', 'uc_quat', 'uc_angvel')}
    h_tgt_quat = np.zeros((n, 4))
    from util import util as ref_util
    for i in range(n):
        robot = StubRobot()
        robot.iso['hand'], robot.vel['hand'] = h_iso[i], h_vel[i]
        h_tgt_quat[i] = ref_util.rot_to_quat(h_tgt[i, :3, :3])
        tnow = h_t0[i] + h_frac[i] * h_dur[i]
        pt, ot = StubTask('hand'), StubTask('hand')
        m = HandTrajectoryManager(pt, ot, robot)
        m.use_current()
        hd['uc_pos'].append(pt.des[0]); hd['uc_vel'].append(pt.des[1]); hd['uc_quat'].append(ot.des[0]); hd['uc_angvel'].append(ot.des[1])
        m.initialize_hand_trajectory(h_t0[i], h_dur[i], h_tgt[i])
        m.update_hand_trajectory(tnow)
        for k, v in zip(('pos', 'vel', 'acc'), pt.des):
            hd[k].append(v)
        for k, v in zip(('quat', 'angvel', 'angacc'), ot.des):
            hd[k].append(v)
        m.initialize_keypoint_hand_trajectory(h_t0[i], h_dur[i], h_key[i], h_tgt[i])
        m.update_keypoint_hand_trajectory(tnow)
        for k, v in zip(('kp_pos', 'kp_vel', 'kp_acc'), pt.des):
            hd[k].append(v)
        for k, v in zip(('kp_quat', 'kp_angvel', 'kp_angacc'), ot.des):
            hd[k].append(v)
    out.update(hand_iso=h_iso, hand_vel6=h_vel, hand_tgt_pos=h_tgt[:, :3, 3], hand_tgt_quat=h_tgt_quat, hand_key=h_key,
               hand_t0=h_t0, hand_dur=h_dur, hand_frac=h_frac, **{'hand_' + k: np.array(v) for k, v in hd.items()})

    # ---- floating base manager ----
    base_iso = np.stack([rand_iso(rng, 0.3) for _ in range(n)])
    com = rng.normal(0, 0.3, (n, 3)) + [0, 0, 0.9]
    tgt_com = com + rng.normal(0, 0.1, (n, 3))
    tgt_quat = np.stack([(R.from_matrix(base_iso[i, :3, :3]) * R.from_rotvec(rng.normal(0, 0.2 if i % 6 else 1e-6, 3))).as_quat()
                         for i in range(n)])
    fb_t0, fb_dur = rng.uniform(0, 1, n), rng.uniform(0.2, 2.0, n)
    fb_t = fb_t0 + fb_dur * np.concatenate([rng.uniform(0.0, 1.3, n - 2), [0.0, 1.0]])
    amp, freq = np.array([0.02, 0.05, 0.03]), np.array([0.5, 1.0, 0.25])
    fb = {k: [] for k in ('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc', 'sway_pos', 'sway_vel', 'sway_acc')}
    for i in range(n):
        robot = StubRobot()
        robot.iso['base'], robot.com = base_iso[i], com[i]
        ct, ot = StubTask('com'), StubTask('base')
        m = FloatingBaseTrajectoryManager(ct, ot, robot)
        m.initialize_floating_base_interpolation_trajectory(fb_t0[i], fb_dur[i], tgt_com[i], tgt_quat[i])
        m.update_floating_base_desired(fb_t[i])
        for k, v in zip(('pos', 'vel', 'acc'), ct.des):
            fb[k].append(v)
        for k, v in zip(('quat', 'angvel', 'angacc'), ot.des):
            fb[k].append(v)
        m2 = FloatingBaseTrajectoryManager(StubTask('com'), StubTask('base'), robot)
        m2.b_swaying = True
        m2.initialize_floating_base_swaying_trajectory(fb_t0[i], amp, freq)
        m2._duration, m2._exp_error = fb_dur[i], np.zeros(3)  # the orientation part is not what is pinned here
        m2.update_floating_base_desired(fb_t[i])
        for k, v in zip(('sway_pos', 'sway_vel', 'sway_acc'), m2._com_task.des):
            fb[k].append(v)
    out.update(base_iso=base_iso, com=com, tgt_com=tgt_com, tgt_quat=tgt_quat, fb_t0=fb_t0, fb_dur=fb_dur, fb_t=fb_t,
               sway_amp=amp, sway_freq=freq, **{'fb_' + k: np.array(v) for k, v in fb.items()})

    # ---- DCM update (atlas_state_estimator.py:35-44, verbatim arithmetic incl. the dangling `+` line) ----
    vcom = rng.normal(0, 0.3, (n, 3))
    dcm_prev_state = com + rng.normal(0, 0.05, (n, 3))
    dcm_vel_state = rng.normal(0, 0.1, (n, 3))
    dt = 0.001
    dcm, prev, dvel = [], [], []
    for i in range(n):
        com_pos, com_vel = com[i], vcom[i]
        sp_dcm, sp_dcm_vel = np.copy(dcm_prev_state[i]), np.copy(dcm_vel_state[i])
        dcm_omega = np.sqrt(9.81 / com_pos[2])
        sp_prev_dcm = np.copy(sp_dcm)
        sp_dcm = com_pos + com_vel / dcm_omega
        alpha_dcm_vel = 0.1
        sp_dcm_vel = alpha_dcm_vel * (sp_dcm - sp_prev_dcm) / dt
        +(1.0 - alpha_dcm_vel) * sp_dcm_vel  # noqa: B018  (a separate statement in the reference too)
        dcm.append(sp_dcm); prev.append(sp_prev_dcm); dvel.append(sp_dcm_vel)
    out.update(dcm_vcom=vcom, dcm_state=dcm_prev_state, dcm_vel_state=dcm_vel_state, dcm_dt=dt, dcm_out=np.array(dcm),
               dcm_prev_out=np.array(prev), dcm_vel_out=np.array(dvel))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'managers.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: np.shape(v) for k, v in out.items() if np.ndim(v) > 0})


if __name__ == '__main__':
    main()
