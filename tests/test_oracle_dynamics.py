"""The rigid-body part of the oracle has no pinocchio to be pinned against (SURVEY.md 8c:
'parity unpinned').  These tests anchor it on the reference's one closed-form fixture
(test/two_link_manipulator_test.py:19-48), on geometry, and on an independent numpy model."""
import os

import numpy as np
import pytest

from pypnc_b200 import robots
from tests import rbd_numpy as N

ROBOTS = ['atlas', 'laikago', 'draco3', 'draco3_lb', 'valkyrie', 'three_link_manipulator']


@pytest.fixture(scope="module")
def om(oracle_mod):
    return {n: oracle_mod.OracleModel(robots.load_model(n)) for n in ROBOTS + ['two_link_manipulator']}


def test_model_tables_match_survey():
    # masses, dof counts, joint order
    exp = {'atlas': (31, 37, 36, 30, 182.4198), 'laikago': (13, 19, 18, 12, 25.5670),
           'draco3': (28, 34, 33, 27, 36.5222), 'valkyrie': (27, 33, 32, 26, 137.2876),
           'three_link_manipulator': (3, 3, 3, 3, 4.001)}
    for n, (nb, nq, nv, na, mass) in exp.items():
        m = robots.load_model(n)
        assert (m.nb, m.nq, m.nv, m.na) == (nb, nq, nv, na)
        assert abs(m.total_mass - mass) < 1e-4
    a = robots.load_model('atlas')
    assert a.joint_names[:4] == ['back_bkz', 'back_bky', 'back_bkx', 'l_arm_shz']
    assert a.joint_names[10] == 'neck_ry' and a.joint_names[18] == 'l_leg_hpz' and a.joint_names[-1] == 'r_leg_akx'
    assert a.joint_id['back_bkz'] == 0 and a.get_q_dot_idx('back_bkz') == 6 and a.get_q_idx('back_bkz') == 7


@pytest.mark.skipif(not os.path.isdir('/root/reference/robot_model'), reason="reference checkout absent")
def test_compiled_tables_match_urdf_compiler():
    from pypnc_b200.models.compile_models import ROBOTS as SRC
    from pypnc_b200.urdf_model import build_model_from_urdf
    for name, (rel, fixed) in SRC.items():
        a = build_model_from_urdf('/root/reference/' + rel, fixed, name=name)
        b = robots.load_model(name)
        assert a.joint_names == b.joint_names and a.frame_names == b.frame_names
        for k in ('parent', 'subtree', 'place_R', 'place_p', 'axis', 'mass', 'com', 'inertia', 'frame_p'):
            np.testing.assert_array_equal(getattr(a, k), getattr(b, k))


def test_two_link_analytic_jacobian(om):
    # reference test/two_link_manipulator_test.py:19-48, [angular; linear] rows
    o = om['two_link_manipulator']
    q = np.array([0.2, 0.1]); qd = np.array([0.012323, 0.01])
    J = np.zeros((6, 2))
    J[3, 0] = -np.sin(q[0]) - np.sin(q[0] + q[1]); J[3, 1] = -np.sin(q[0] + q[1])
    J[4, 0] = np.cos(q[0]) + np.cos(q[0] + q[1]); J[4, 1] = np.cos(q[0] + q[1])
    J[2, :] = 1.
    Jd = np.zeros((6, 2))
    Jd[3, 0] = -np.cos(q[0]) * qd[0] - np.cos(q[0] + q[1]) * (qd[0] + qd[1]); Jd[3, 1] = -np.cos(q[0] + q[1]) * (qd[0] + qd[1])
    Jd[4, 0] = -np.sin(q[0]) * qd[0] - np.sin(q[0] + q[1]) * (qd[0] + qd[1]); Jd[4, 1] = -np.sin(q[0] + q[1]) * (qd[0] + qd[1])
    ee = o.model.frame_id('ee')
    np.testing.assert_allclose(o.link_jacobian(q, ee), J, atol=1e-14)
    np.testing.assert_allclose(o.link_jdqd(q, qd, ee), Jd @ qd, atol=1e-14)
    np.testing.assert_allclose(o.link_vel(q, qd, ee), J @ qd, atol=1e-14)


def test_three_link_geometry(om):
    # reference test/pinocchio_frame_test.py:14-16: q = [pi/2, 0, 0] -> ee at (0, 3, 0)
    o = om['three_link_manipulator']
    T = o.link_iso(np.array([np.pi / 2, 0, 0]), o.model.frame_id('ee'))
    np.testing.assert_allclose(T[:3, 3], [0, 3, 0], atol=1e-14)
    assert abs(o.model.total_mass - 4.001) < 1e-12


@pytest.mark.parametrize("name", ROBOTS)
def test_against_independent_numpy_model(om, name):
    o = om[name]; m = o.model
    rng = np.random.default_rng(11)
    for _ in range(2):
        q, v = N.random_state(m, rng)
        M = o.mass_matrix(q)
        Mn = N.mass_matrix(m, q)
        np.testing.assert_allclose(M, Mn, rtol=0, atol=1e-10 * np.abs(Mn).max())
        np.testing.assert_allclose(M, M.T, atol=0)
        assert np.linalg.eigvalsh(M).min() > 0
        g = o.gravity(q)
        np.testing.assert_allclose(g, N.nle(m, q, 0 * v), atol=1e-9 * np.abs(g).max())
        c = o.coriolis(q, v)
        cn = N.nle(m, q, v, gravity=False)
        np.testing.assert_allclose(c, cn, atol=2e-6 * max(1.0, np.abs(cn).max()))
        # nle(q, 0) == g ;  RNEA(q, v, a) == M a + c + g
        a = rng.normal(size=m.nv)
        np.testing.assert_allclose(o.rnea(q, v, a), M @ a + c + g, atol=1e-9 * np.abs(M).max())
        com, vcom = o.com(q, v)
        np.testing.assert_allclose(com, N.com(m, q), atol=1e-12)
        Jc = o.com_jacobian(q)
        np.testing.assert_allclose(Jc @ v, vcom, atol=1e-11)
        eps = 1e-6
        fd = (N.com(m, N.integrate(m, q, eps * v)) - N.com(m, N.integrate(m, q, -eps * v))) / (2 * eps)
        np.testing.assert_allclose(vcom, fd, atol=1e-7)
        # Jdot_com . v == d/dt (J_com v) at constant v (scaled by mass_dyn/total_mass as the wrapper does)
        dJ = o.com_jacobian_dot(q, v)
        fd2 = (o.com_jacobian(N.integrate(m, q, eps * v)) - o.com_jacobian(N.integrate(m, q, -eps * v))) / (2 * eps)
        np.testing.assert_allclose(dJ * (m.total_mass / m.mass_dyn), fd2, atol=2e-6)


@pytest.mark.parametrize("name,frames", [('atlas', ['pelvis_com', 'r_sole', 'l_hand']),
                                         ('laikago', ['chassis', 'toeFL']),
                                         ('draco3', ['torso_com_link', 'l_foot_contact']),
                                         ('valkyrie', ['leftCOP_Frame', 'rightElbowPitchLink'])])
def test_frame_quantities(om, name, frames):
    o = om[name]; m = o.model
    rng = np.random.default_rng(3)
    q, v = N.random_state(m, rng)
    eps = 1e-6
    for fn in frames:
        f = m.frame_id(fn)
        T = o.link_iso(q, f)
        R, p = N.frame_pose(m, q, f)
        np.testing.assert_allclose(T[:3, :3], R, atol=1e-13)
        np.testing.assert_allclose(T[:3, 3], p, atol=1e-13)
        J = o.link_jacobian(q, f)  # [ang; lin]
        Jn, _, _ = N.point_jacobian(m, q, m.frame_parent[f], m.frame_p[f])
        np.testing.assert_allclose(J[3:], Jn[:3], atol=1e-12)
        np.testing.assert_allclose(J[:3], Jn[3:], atol=1e-12)
        np.testing.assert_allclose(o.link_vel(q, v, f), J @ v, atol=1e-12)
        qp, qm = N.integrate(m, q, eps * v), N.integrate(m, q, -eps * v)
        fd = (o.link_jacobian(qp, f) - o.link_jacobian(qm, f)) @ v / (2 * eps)
        np.testing.assert_allclose(o.link_jdqd(q, v, f), fd, atol=2e-7 * max(1, np.abs(fd).max()))


def test_centroidal(om):
    o = om['atlas']; m = o.model
    q, v = N.random_state(m, np.random.default_rng(2))
    Ag, hg, Ig = o.centroidal(q, v)
    com, vcom = o.com(q, v)
    np.testing.assert_allclose(hg[3:], m.mass_dyn * vcom, atol=1e-10)   # linear momentum
    np.testing.assert_allclose(Ag[3:] / m.mass_dyn, o.com_jacobian(q), atol=1e-12)
    assert abs(Ig[3, 3] - m.mass_dyn) < 1e-12
    # angular momentum about the CoM from per-body sums
    Rw, pw = N.fk(m, q)
    L = np.zeros(3)
    for i in range(m.nb):
        J, pt, R = N.point_jacobian(m, q, i, m.com[i])
        L += np.cross(pt - com, m.mass[i] * (J[:3] @ v)) + R @ m.inertia[i] @ R.T @ (J[3:] @ v)
    np.testing.assert_allclose(hg[:3], L, atol=1e-9)


def test_helpers_vs_scipy(oracle_mod):
    import ctypes as C
    from scipy.spatial.transform import Rotation as Rot
    lib = oracle_mod._load()
    rng = np.random.default_rng(0)
    for k in range(300):
        ang = [1e-6, 5e-5, 1e-3, 0.1, 1.0, 2.0, 3.0, 3.14159][k % 8]
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        Rm = np.ascontiguousarray(Rot.from_rotvec(ax * ang).as_matrix())
        q = np.zeros(4); e = np.zeros(3); R2 = np.zeros(9)
        lib.orc_rot_to_quat(Rm.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
        qs = Rot.from_matrix(Rm).as_quat()
        np.testing.assert_allclose(q, qs, atol=1e-15)
        lib.orc_quat_to_rot(qs.ctypes.data_as(C.c_void_p), R2.ctypes.data_as(C.c_void_p))
        np.testing.assert_allclose(R2.reshape(3, 3), Rot.from_quat(qs).as_matrix(), atol=1e-15)
        lib.orc_quat_to_exp(q.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p))
        # util.quat_to_exp (util/util.py:61-72) restated on scipy's quaternion
        th = 2.0 * np.arcsin(np.sqrt(qs[0] ** 2 + qs[1] ** 2 + qs[2] ** 2))
        exp_ref = np.zeros(3) if abs(th) < 1e-4 else qs[:3] / np.sin(th / 2.0) * th
        np.testing.assert_allclose(e, exp_ref, atol=1e-12)
        if ang < 1e-4:
            assert np.all(e == 0)          # dead zone (util/util.py:68-69)
        elif qs[3] >= 0:
            np.testing.assert_allclose(e, ax * ang, atol=1e-9)


def test_pinv_vs_numpy(oracle_mod):
    import ctypes as C
    lib = oracle_mod._load()
    rng = np.random.default_rng(1)
    for (m, n, rank) in [(3, 3, 3), (3, 3, 2), (25, 25, 25), (2, 2, 2), (3, 7, 3), (3, 3, 1)]:
        A = rng.normal(size=(m, rank)) @ rng.normal(size=(rank, n))
        P = np.zeros((n, m))
        A = np.ascontiguousarray(A)
        lib.orc_pinv(A.ctypes.data_as(C.c_void_p), C.c_int(m), C.c_int(n), C.c_double(1e-10), P.ctypes.data_as(C.c_void_p))
        np.testing.assert_allclose(P, np.linalg.pinv(A, rcond=1e-10), atol=1e-9 * max(1, np.abs(P).max()))
