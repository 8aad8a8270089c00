"""CUDA managers (csrc/managers.cu, through the C ABI) against the golden vectors produced by the
reference's own manager classes and against the numpy oracle."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'managers.npz'))
N = G['ramp_w0'].shape[0]
TOL = 1e-9  # relative, fp64 (north_star)


def close(a, b, tol=TOL):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = np.asarray(b, float)
    err = np.abs(a - b).max()
    assert err <= tol * max(1.0, np.abs(b).max()), err


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


@pytest.fixture(scope='module')
def rig():
    """Atlas workspace of N states whose frame / CoM arrays are then overwritten with the golden inputs."""
    from pypnc_b200 import robots, engine
    spec = robots.PROBLEMS['atlas']()
    st = robots.synth_states('atlas', spec, N, seed=3)
    dm = engine.Model(spec.model)
    ws = engine.Workspace(dm, N, spec.frames)
    d = {k: dev(v) for k, v in st.items()}
    ws.update_system(d['base_pos'], d['base_quat'], d['base_lin_vel'], d['base_ang_vel'], d['joint_pos'], d['joint_vel'])
    torch.cuda.synchronize()
    return spec, dm, ws


def put_frame(ws, slot, iso, vel):
    fi = np.concatenate([iso[:, :3, :3].reshape(-1, 9), iso[:, :3, 3]], axis=1)
    ws.view('frame_iso')[:, slot] = dev(fi)
    if vel is not None:
        ws.view('frame_vel')[:, slot] = dev(vel)


def P(t):
    return C.c_void_p(t.data_ptr())


def outs():
    mk = lambda w: torch.empty((N, w), dtype=torch.float64, device='cuda')
    return [mk(3), mk(3), mk(3), mk(4), mk(3), mk(3)]


KEYS = ('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc')


def test_ramps_golden():
    from pypnc_b200._cabi import load_library
    lib = load_library()
    out = torch.empty(N, dtype=torch.float64, device='cuda')
    t0, dur = dev(G['ramp_t0']), dev(G['ramp_dur'])
    # the golden cases evaluate every state at its own time: one launch per distinct time is the API, so loop
    for key, start, target in (('ramp_w_to_min', G['ramp_w0'], float(G['ramp_wmin'])),
                               ('ramp_w_to_max', G['ramp_w0'], float(G['ramp_wmax'])),
                               ('ramp_rf_to_min', G['ramp_rfz0'], 0.001),
                               ('ramp_rf_to_max', G['ramp_rfz0'], float(G['ramp_rfzmax']))):
        sv = dev(start)
        got = np.zeros(N)
        for i in range(N):
            assert lib.pnc_ramp_update(N, P(sv), target, P(t0), P(dur), float(G['ramp_t'][i]), P(out), None) == 0
            got[i] = float(out[i])
        close(got, G[key])


def test_foot_use_current_golden(rig):
    from pypnc_b200._cabi import load_library
    spec, dm, ws = rig
    lib = load_library()
    slot = 0
    put_frame(ws, slot, G['foot_iso'], G['foot_vel'])
    o = outs()
    assert lib.pnc_foot_use_current(ws.h, N, spec.frames[slot], *[P(t) for t in o], None) == 0
    for k, t in zip(KEYS, o):
        close(t, G['uc_' + k])


def test_swing_golden(rig):
    from pypnc_b200._cabi import load_library
    from pypnc_b200.managers import SWING_REC
    spec, dm, ws = rig
    lib = load_library()
    slot = 1
    put_frame(ws, slot, G['foot_iso'], G['foot_vel'])
    rec = torch.empty((N, SWING_REC), dtype=torch.float64, device='cuda')
    t0, dur = dev(G['sw_t0']), dev(G['sw_dur'])
    assert lib.pnc_swing_init(ws.h, N, spec.frames[slot], P(dev(G['land_pos'])), P(dev(G['land_quat'])), P(dur),
                              float(G['sw_height']), P(rec), None) == 0
    o = outs()
    got = {k: np.zeros((N, 4 if k == 'quat' else 3)) for k in KEYS}
    for i in range(N):
        t = float(G['sw_t0'][i] + G['sw_frac'][i] * G['sw_dur'][i])
        assert lib.pnc_swing_eval(N, P(rec), P(t0), P(dur), t, *[P(x) for x in o], None) == 0
        for k, x in zip(KEYS, o):
            got[k][i] = x[i].cpu().numpy()
    for k in KEYS:
        close(got[k], G['sw_' + k])


def test_hand_trajectory_golden(rig):
    """pnc_hand_init / pnc_hand_eval against the reference's own HandTrajectoryManager (plain and keypoint
    trajectories, hand_trajectory_manager.py:57-178)."""
    from pypnc_b200._cabi import load_library
    from pypnc_b200.managers import HAND_REC
    spec, dm, ws = rig
    lib = load_library()
    slot = 3
    put_frame(ws, slot, G['hand_iso'], G['hand_vel6'])
    t0, dur = dev(G['hand_t0']), dev(G['hand_dur'])
    tp, tq, kp = dev(G['hand_tgt_pos']), dev(G['hand_tgt_quat']), dev(G['hand_key'])
    o = outs()
    for pre, key in (('', None), ('kp_', kp)):
        rec = torch.empty((N, HAND_REC), dtype=torch.float64, device='cuda')
        assert lib.pnc_hand_init(ws.h, N, spec.frames[slot], P(tp), P(tq), None if key is None else P(key), P(rec), None) == 0
        got = {k: np.zeros((N, 4 if k == 'quat' else 3)) for k in KEYS}
        for i in range(N):
            t = float(G['hand_t0'][i] + G['hand_frac'][i] * G['hand_dur'][i])
            assert lib.pnc_hand_eval(N, P(rec), P(t0), P(dur), t, int(key is not None), *[P(x) for x in o], None) == 0
            for k, x in zip(KEYS, o):
                got[k][i] = x[i].cpu().numpy()
        for k in KEYS:
            close(got[k], G['hand_' + pre + k])


def test_floating_base_golden(rig):
    from pypnc_b200._cabi import load_library
    from pypnc_b200.managers import FB_REC
    spec, dm, ws = rig
    lib = load_library()
    slot = 2
    put_frame(ws, slot, G['base_iso'], None)
    ws.view('com')[:] = dev(G['com'])
    rec = torch.empty((N, FB_REC), dtype=torch.float64, device='cuda')
    t0, dur = dev(G['fb_t0']), dev(G['fb_dur'])
    assert lib.pnc_floating_base_init(ws.h, N, spec.frames[slot], P(dev(G['tgt_com'])), P(dev(G['tgt_quat'])), P(rec), None) == 0
    o = outs()
    got = {k: np.zeros((N, 4 if k == 'quat' else 3)) for k in KEYS}
    sway = {k: np.zeros((N, 3)) for k in ('pos', 'vel', 'acc')}
    amp = np.ascontiguousarray(G['sway_amp'], dtype=np.float64)
    freq = np.ascontiguousarray(G['sway_freq'], dtype=np.float64)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    mid = dev(G['com'])
    for i in range(N):
        t = float(G['fb_t'][i])
        assert lib.pnc_floating_base_eval(N, P(rec), P(t0), P(dur), t, *[P(x) for x in o], None) == 0
        for k, x in zip(KEYS, o):
            got[k][i] = x[i].cpu().numpy()
        assert lib.pnc_sinusoid_eval(N, P(mid), P(t0), dp(amp), dp(freq), t, P(o[0]), P(o[1]), P(o[2]), None) == 0
        for k, x in zip(('pos', 'vel', 'acc'), o[:3]):
            sway[k][i] = x[i].cpu().numpy()
    for k in KEYS:
        close(got[k], G['fb_' + k])
    for k in sway:
        close(sway[k], G['fb_sway_' + k])


def test_dcm_golden(rig):
    from pypnc_b200._cabi import load_library
    spec, dm, ws = rig
    lib = load_library()
    ws.view('com')[:] = dev(G['com'])
    ws.view('vcom')[:] = dev(G['dcm_vcom'])
    dcm, prev, dvel = dev(G['dcm_state']), torch.empty((N, 3), dtype=torch.float64, device='cuda'), dev(G['dcm_vel_state'])
    assert lib.pnc_dcm_update(ws.h, N, float(G['dcm_dt']), 0.1, P(dcm), P(prev), P(dvel), None) == 0
    close(dcm, G['dcm_out']); close(prev, G['dcm_prev_out']); close(dvel, G['dcm_vel_out'])


def test_manager_classes_closed_loop_vs_oracle():
    """The reference-style classes on a BatchedRobotSystem: a stance foot held with use_current, a swing
    foot, ramps and the floating base, fed into IHWBC.solve; every desired is checked against the
    oracle evaluated on the robot's own link transforms."""
    from oracle import managers_oracle as MO
    from pypnc_b200 import robots, managers, wbc
    from pypnc_b200.robot_system import BatchedRobotSystem
    B = 64
    spec = robots.PROBLEMS['atlas']()
    st = robots.synth_states('atlas', spec, B, seed=5)
    robot = BatchedRobotSystem(spec.model, capacity=B, links=['r_sole', 'l_sole', 'pelvis_com'])
    d = {k: dev(v) for k, v in st.items()}
    robot.update_system(None, None, None, None, d['base_pos'], d['base_quat'], d['base_lin_vel'], d['base_ang_vel'],
                        d['joint_pos'], d['joint_vel'])
    rpos, rori = wbc.BasicTask(robot, 'LINK_XYZ', 3, 'r_sole'), wbc.BasicTask(robot, 'LINK_ORI', 3, 'r_sole')
    lpos, lori = wbc.BasicTask(robot, 'LINK_XYZ', 3, 'l_sole'), wbc.BasicTask(robot, 'LINK_ORI', 3, 'l_sole')
    com, pel = wbc.BasicTask(robot, 'COM', 3), wbc.BasicTask(robot, 'LINK_ORI', 3, 'pelvis_com')
    iso_r = robot.get_link_iso('r_sole').cpu().numpy(); vel_r = robot.get_link_vel('r_sole').cpu().numpy()
    iso_l = robot.get_link_iso('l_sole').cpu().numpy()
    iso_p = robot.get_link_iso('pelvis_com').cpu().numpy(); cpos = robot.get_com_pos().cpu().numpy()
    # stance foot
    fr = managers.FootTrajectoryManager(rpos, rori, robot)
    fr.use_current()
    ref = [MO.foot_use_current(iso_r[i], vel_r[i]) for i in range(B)]
    close(rpos.pos_des, np.array([r[0] for r in ref])); close(rori.pos_des, np.array([r[3] for r in ref]))
    # swing foot, every robot at its own phase
    rng = np.random.default_rng(0)
    land_pos = iso_l[:, :3, 3] + rng.normal(0, 0.1, (B, 3))
    land_quat = np.array([MO.q_mul(MO.q_from_matrix(iso_l[i, :3, :3]), MO.q_from_rotvec(rng.normal(0, 0.2, 3))) for i in range(B)])
    t0, dur = rng.uniform(0, 0.3, B), rng.uniform(0.4, 0.8, B)
    fl = managers.FootTrajectoryManager(lpos, lori, robot)
    fl.swing_height = 0.07
    fl.initialize_swing_foot_trajectory(dev(t0), dev(dur), dev(land_pos), dev(land_quat))
    fl.update_swing_foot_desired(0.45)
    ref = [MO.swing_foot_desired(iso_l[i], land_pos[i], land_quat[i], t0[i], dur[i], 0.07, 0.45) for i in range(B)]
    close(lpos.pos_des, np.array([r[0] for r in ref])); close(lpos._vel_des, np.array([r[1] for r in ref]))
    close(lpos._acc_des, np.array([r[2] for r in ref])); close(lori.pos_des, np.array([r[3] for r in ref]))
    close(lori._vel_des, np.array([r[4] for r in ref])); close(lori._acc_des, np.array([r[5] for r in ref]))
    # floating base
    fb = managers.FloatingBaseTrajectoryManager(com, pel, robot)
    tgt_com = cpos + rng.normal(0, 0.05, (B, 3))
    tgt_quat = np.array([MO.q_mul(MO.q_from_matrix(iso_p[i, :3, :3]), MO.q_from_rotvec(rng.normal(0, 0.1, 3))) for i in range(B)])
    fb.initialize_floating_base_interpolation_trajectory(0.1, 1.0, dev(tgt_com), dev(tgt_quat))
    fb.update_floating_base_desired(0.6)
    ref = [MO.floating_base_desired(cpos[i], iso_p[i, :3, :3], tgt_com[i], tgt_quat[i], 0.1, 1.0, 0.6) for i in range(B)]
    close(com.pos_des, np.array([r[0] for r in ref])); close(pel.pos_des, np.array([r[3] for r in ref]))
    close(pel._vel_des, np.array([r[4] for r in ref]))
    # hand trajectory (plain and keypoint) on the right foot frame standing in for a hand
    hpos, hori = wbc.BasicTask(robot, 'LINK_XYZ', 3, 'r_sole'), wbc.BasicTask(robot, 'LINK_ORI', 3, 'r_sole')
    hm = managers.HandTrajectoryManager(hpos, hori, robot)
    h_tp = iso_r[:, :3, 3] + rng.normal(0, 0.1, (B, 3))
    h_tq = np.array([MO.q_mul(MO.q_from_matrix(iso_r[i, :3, :3]), MO.q_from_rotvec(rng.normal(0, 0.3, 3))) for i in range(B)])
    h_kp = iso_r[:, :3, 3] + rng.normal(0, 0.1, (B, 3))
    for key in (None, h_kp):
        if key is None:
            hm.initialize_hand_trajectory(0.1, 0.8, dev(h_tp), dev(h_tq))
            hm.update_hand_trajectory(0.35)
        else:
            hm.initialize_keypoint_hand_trajectory(0.1, 0.8, dev(key), dev(h_tp), dev(h_tq))
            hm.update_keypoint_hand_trajectory(0.65)
        ref = [MO.hand_desired(iso_r[i], vel_r[i], h_tp[i], h_tq[i], 0.1, 0.8, 0.35 if key is None else 0.65,
                               None if key is None else key[i]) for i in range(B)]
        close(hpos.pos_des, np.array([r[0] for r in ref])); close(hpos._vel_des, np.array([r[1] for r in ref]))
        close(hpos._acc_des, np.array([r[2] for r in ref])); close(hori.pos_des, np.array([r[3] for r in ref]))
        close(hori._vel_des, np.array([r[4] for r in ref])); close(hori._acc_des, np.array([r[5] for r in ref]))
    hm.use_current()
    ref = [MO.foot_use_current(iso_r[i], vel_r[i]) for i in range(B)]
    close(hpos.pos_des, np.array([r[0] for r in ref])); close(hori.pos_des, np.array([r[3] for r in ref]))
    # ramps
    lpos.w_hierarchy = dev(rng.uniform(40, 60, B))
    w0 = lpos.w_hierarchy.cpu().numpy()
    th = managers.TaskHierarchyManager(lpos, 60.0, 40.0)
    th.initialize_ramp_to_min(0.2, 0.5)
    th.update_ramp_to_min(0.45)
    close(lpos.w_hierarchy, MO.ramp(w0, 40.0, 0.2, 0.5, 0.45))
    # DCM
    est = managers.DCMEstimator(robot, 0.001)
    est.update()
    vcom = robot.get_com_lin_vel().cpu().numpy()
    ref = [MO.dcm_update(cpos[i], vcom[i], np.zeros(3), 0.001) for i in range(B)]
    close(est.dcm, np.array([r[0] for r in ref])); close(est.dcm_vel, np.array([r[2] for r in ref]), 1e-9)


def test_point_foot_and_upper_body_managers(rig):
    """PointFootTrajectoryManager / UpperBodyTrajectoryManager through the reference-style classes, against
    the golden vectors of the reference's PointFootTrajectoryManager."""
    from pypnc_b200 import managers, wbc

    class Rob:  # just enough of BatchedRobotSystem for the managers: the rigged workspace and its frame ids
        def __init__(self, spec, ws):
            self._ws, self.model, self._spec = ws, spec.model, spec

        def _slot(self, target):
            f = self.model.frame_id(target) if isinstance(target, str) else int(target)
            return self._ws, self._ws.frames.index(f)

    spec, dm, ws = rig
    slot = 3
    put_frame(ws, slot, G['foot_iso'], G['foot_vel'])
    rob = Rob(spec, ws)
    target = spec.frames[slot]

    class T:  # records update_desired like the reference's Task
        target_id = target

        def update_desired(self, p, v, a):
            self.des = (p, v, a)

    class Model2:
        def __init__(self, m):
            self._m = m

        def frame_id(self, t):
            return int(t)

    rob.model = Model2(spec.model)
    task = T()
    m = managers.PointFootTrajectoryManager(task, rob)
    m.use_current()
    for k, t in zip(('uc_pos', 'uc_vel', 'uc_acc'), task.des):
        close(t, G['pf_' + k])
    m.swing_height = float(G['sw_height'])
    m.initialize_swing_foot_trajectory(dev(G['sw_t0']), dev(G['sw_dur']), dev(G['land_pos']), dev(G['land_quat']))
    got = {k: np.zeros((N, 3)) for k in ('sw_pos', 'sw_vel', 'sw_acc')}
    for i in range(N):
        m.update_swing_foot_desired(float(G['sw_t0'][i] + G['sw_frac'][i] * G['sw_dur'][i]))
        for k, t in zip(('sw_pos', 'sw_vel', 'sw_acc'), task.des):
            got[k][i] = t[i].cpu().numpy()
    for k in got:
        close(got[k], G['pf_' + k])
    # upper body: desired = nominal, zero velocity / acceleration
    class JT:
        target_id = ['a', 'b', 'c']

        def update_desired(self, p, v, a):
            self.des = (p, v, a)

    jt = JT()
    ub = managers.UpperBodyTrajectoryManager(jt, rob)
    ub.use_nominal_upper_body_joint_pos({'a': 0.1, 'b': torch.tensor([0.2, 0.3], dtype=torch.float64), 'c': -0.4, 'unused': 9.0})
    assert jt.des[0].shape == (2, 3) and jt.des[0][1].tolist() == [0.1, 0.3, -0.4] and float(jt.des[1].abs().max()) == 0.0
