"""The rigid-body layer against the first-principles referee (tests/first_principles.py; fixture made by
tests/golden/make_first_principles.py from the raw reference URDFs).

pinocchio -- what pnc/robot_system/pinocchio_robot_system.py:181-278 delegates to -- cannot be installed here, so the
oracle's M, g, c, CoM, Jacobians, Jdot*qdot and centroidal quantities have no reference-produced vector to be pinned on.
The referee is the strongest stand-in available offline: it shares NOTHING with the code under test (its only model is the
placement of every unmerged URDF link; inertial forces come from numerical derivatives of those placements projected on
numerical partial velocities), so it pins the conventions the reference relies on -- body-frame base twist as generalized
velocity (:149-162), LOCAL_WORLD_ALIGNED [angular; linear] frame quantities (:231-261), classical acceleration at
vdot = 0 for Jdot*qdot (:263-278), the [angular; linear] block order of Ag / hg / Ig (:181-194), dAg_lin / total_mass for
the CoM Jacobian derivative (:225-228), the welded root link of fixed-base models -- for arbitrary configurations and
velocities, not only the nominal pose.

Tolerance: 1e-6 of the quantity's largest entry (absolute 1e-6 for quantities that vanish); the referee's own
finite-difference error, stored with each state (`fd_err`, evaluation at step H against 2 H), is below 5e-7.
CPU: the oracle.  GPU (pytest -m gpu): k_update_system through the C ABI, fp64.
"""
import os

import numpy as np
import pytest

from pypnc_b200 import robots

HERE = os.path.dirname(os.path.abspath(__file__))
FX = np.load(os.path.join(HERE, 'golden', 'first_principles.npz'))
NAMES = ('atlas', 'laikago', 'draco3', 'draco3_lb', 'valkyrie', 'three_link_manipulator', 'two_link_manipulator')
STATES = 3
TOL = 1e-6
IN_KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')


def close(got, want, what):
    scale = max(float(np.abs(want).max()), 1.0)
    err = float(np.abs(np.asarray(got) - want).max()) / scale
    assert err < TOL, (what, err)
    return err


def fx(name, k, q):
    return FX['%s/%d/%s' % (name, k, q)]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_first_principles(name, oracle_mod):
    m = robots.load_model(name)
    om = oracle_mod.OracleModel(m)
    frames = [str(f) for f in FX[name + '/frames']]
    assert len(frames) >= 4
    worst = 0.0
    for k in range(STATES):
        assert float(fx(name, k, 'fd_err')) < 5e-7
        st = {kk: fx(name, k, 'in_' + kk) for kk in IN_KEYS}
        if m.n_floating:
            q, v = om.pack_state(*[st[kk] for kk in IN_KEYS])
        else:
            q, v = om.pack_state(None, None, None, None, st['joint_pos'], st['joint_vel'])
        tag = '%s state %d ' % (name, k)
        com, vcom = om.com(q, v)
        Ag, hg, Ig = om.centroidal(q, v)
        dJ = om.com_jacobian_dot(q, v)
        pairs = [('qdot', v), ('M', om.mass_matrix(q)), ('grav', om.gravity(q)), ('cori', om.coriolis(q, v)),
                 ('com', com), ('vcom', vcom), ('jcom', om.com_jacobian(q)), ('jcomdot', dJ), ('jcom_jdqd', dJ @ v),
                 ('Ag', Ag), ('hg', hg), ('Ig', Ig)]
        for qn, got in pairs:
            worst = max(worst, close(got, fx(name, k, qn), tag + qn))
        for i, fn in enumerate(frames):
            f = m.frame_id(fn)
            worst = max(worst, close(om.link_iso(q, f), fx(name, k, 'frame_iso')[i], tag + fn + ' iso'))
            worst = max(worst, close(om.link_vel(q, v, f), fx(name, k, 'frame_vel')[i], tag + fn + ' vel'))
            worst = max(worst, close(om.link_jacobian(q, f), fx(name, k, 'frame_jac')[i], tag + fn + ' jac'))
            worst = max(worst, close(om.link_jdqd(q, v, f), fx(name, k, 'frame_jdqd')[i], tag + fn + ' jdqd'))
    print('%s: oracle vs first principles, worst %.1e over %d states x (12 quantities + %d frames x 4)' %
          (name, worst, STATES, len(frames)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_update_system_kernel_matches_first_principles(name):
    """k_update_system (fp64) against the referee directly -- no oracle in between."""
    import torch
    from pypnc_b200 import engine
    m = robots.load_model(name)
    frames = [str(f) for f in FX[name + '/frames']]
    dm = engine.Model(m, 0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ins = [np.stack([fx(name, k, 'in_' + kk) for k in range(STATES)]) for kk in IN_KEYS]
    d_in = [dev(a) for a in ins] if m.n_floating else [None] * 4 + [dev(a) for a in ins[4:]]
    worst = 0.0
    for c0 in range(0, len(frames), 12):  # a workspace materialises at most PNC_MAX_FRAMES = 12 frames
        chunk = frames[c0:c0 + 12]
        ws = engine.Workspace(dm, STATES, [m.frame_id(f) for f in chunk])
        ws.update_system(*d_in, b_cent=True)
        torch.cuda.synchronize()
        g = {f: ws.view(f).cpu().numpy() for f in ('qdot', 'M', 'grav', 'cori', 'com', 'vcom', 'jcom', 'jcomdot',
                                                    'jcom_jdqd', 'frame_iso', 'frame_vel', 'frame_jac', 'frame_jdqd')}
        Ag, hg, Ig = [t.cpu().numpy() for t in ws.centroidal()]
        for k in range(STATES):
            tag = '%s state %d ' % (name, k)
            if c0 == 0:
                for qn in ('qdot', 'M', 'grav', 'cori', 'com', 'vcom', 'jcom', 'jcomdot', 'jcom_jdqd'):
                    worst = max(worst, close(g[qn][k], fx(name, k, qn), tag + qn))
                worst = max(worst, close(Ag[k], fx(name, k, 'Ag'), tag + 'Ag'))
                worst = max(worst, close(hg[k], fx(name, k, 'hg'), tag + 'hg'))
                worst = max(worst, close(Ig[k], fx(name, k, 'Ig'), tag + 'Ig'))
            for i, fn in enumerate(chunk):
                iso = np.eye(4)
                iso[:3, :3] = g['frame_iso'][k, i, :9].reshape(3, 3)
                iso[:3, 3] = g['frame_iso'][k, i, 9:12]
                worst = max(worst, close(iso, fx(name, k, 'frame_iso')[c0 + i], tag + fn + ' iso'))
                worst = max(worst, close(g['frame_vel'][k, i], fx(name, k, 'frame_vel')[c0 + i], tag + fn + ' vel'))
                worst = max(worst, close(g['frame_jac'][k, i], fx(name, k, 'frame_jac')[c0 + i], tag + fn + ' jac'))
                worst = max(worst, close(g['frame_jdqd'][k, i], fx(name, k, 'frame_jdqd')[c0 + i], tag + fn + ' jdqd'))
        del ws
    print('%s: k_update_system vs first principles, worst %.1e' % (name, worst))


def test_referee_motion_is_the_se3_exponential():
    """The referee's motion q (+) t v must be pinocchio's integrate: T_b Exp(t [v; w]) with the twist in the BODY frame.
    Its closed form (first_principles.se3_exp) against scipy's matrix exponential of the 4x4 twist matrix."""
    from scipy.linalg import expm
    from tests.first_principles import se3_exp
    rng = np.random.default_rng(0)
    for _ in range(20):
        v, w, t = rng.normal(size=3), rng.normal(size=3), rng.uniform(-1, 1)
        X = np.zeros((4, 4))
        X[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        X[:3, 3] = v
        E = expm(t * X)
        R, p = se3_exp(v, w, t)
        np.testing.assert_allclose(R, E[:3, :3], atol=1e-13)
        np.testing.assert_allclose(p, E[:3, 3], atol=1e-13)
    R, p = se3_exp(np.array([1., 2., 3.]), np.zeros(3), 0.5)  # pure translation
    np.testing.assert_allclose(R, np.eye(3), atol=0)
    np.testing.assert_allclose(p, [0.5, 1.0, 1.5], atol=1e-15)


def test_referee_reproduces_the_reference_closed_form_two_link_fixture():
    """The referee itself against the one closed form the reference holds (test/two_link_manipulator_test.py:19-48:
    analytic Jacobian and its time derivative of the planar two-link arm, rows [angular; linear])."""
    name = 'two_link_manipulator'
    frames = [str(f) for f in FX[name + '/frames']]
    e = frames.index('ee')
    for k in range(STATES):
        q, qd = fx(name, k, 'in_joint_pos'), fx(name, k, 'in_joint_vel')
        J = np.zeros((6, 2))
        J[3, 0] = -np.sin(q[0]) - np.sin(q[0] + q[1]); J[3, 1] = -np.sin(q[0] + q[1])
        J[4, 0] = np.cos(q[0]) + np.cos(q[0] + q[1]); J[4, 1] = np.cos(q[0] + q[1])
        J[2, :] = 1.
        Jd = np.zeros((6, 2))
        Jd[3, 0] = -np.cos(q[0]) * qd[0] - np.cos(q[0] + q[1]) * (qd[0] + qd[1]); Jd[3, 1] = -np.cos(q[0] + q[1]) * (qd[0] + qd[1])
        Jd[4, 0] = -np.sin(q[0]) * qd[0] - np.sin(q[0] + q[1]) * (qd[0] + qd[1]); Jd[4, 1] = -np.sin(q[0] + q[1]) * (qd[0] + qd[1])
        np.testing.assert_allclose(fx(name, k, 'frame_jac')[e], J, atol=1e-9)
        np.testing.assert_allclose(fx(name, k, 'frame_vel')[e], J @ qd, atol=1e-9)
        np.testing.assert_allclose(fx(name, k, 'frame_jdqd')[e], Jd @ qd, atol=1e-8)


@pytest.mark.skipif(not os.path.isdir('/root/reference/robot_model'), reason="reference checkout absent")
def test_fixture_is_current():
    """A fresh evaluation of one state from the reference URDF equals the committed fixture (guards against a stale file)."""
    from tests.golden.make_first_principles import make_referee
    name, k = 'laikago', 1
    ref, m = make_referee(name)
    r = ref.evaluate(**{kk: fx(name, k, 'in_' + kk) for kk in IN_KEYS})
    for qn in ('M', 'cori', 'jcomdot', 'frame_jdqd', 'Ag'):
        np.testing.assert_allclose(r[qn], fx(name, k, qn), rtol=1e-12, atol=1e-12)
