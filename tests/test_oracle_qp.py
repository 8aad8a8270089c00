"""Pin oracle/pnc_oracle.c:orc_solve_quadprog against (a) the known-answer QPs the reference's
test scripts use and (b) the reference's own QuadProg++.cc compiled into oracle/_ref."""
import numpy as np
import pytest


def test_quadprogpp_demo_qp(oracle_mod):
    # reference test/qp_test.py:14-19 (QuadProg++ demo): f = 12, x = (1, 2)
    G = np.array([[4., -2], [-2, 4]]); g0 = np.array([6., 0])
    CE = np.array([[1.], [1.]]); ce0 = np.array([-3.])
    CI = np.array([[1., 0, 1], [0, 1, 1]]); ci0 = np.array([0., 0, -2])
    r = oracle_mod.solve_quadprog(G, g0, CE, ce0, CI, ci0)
    assert r['status'] == 0 and abs(r['f'] - 12.0) < 1e-12
    np.testing.assert_allclose(r['x'], [1., 2.], atol=1e-12)


def test_qpsolvers_readme_qp(oracle_mod):
    # reference test/qp_solver.py:30-36
    M = np.array([[1., 2., 0.], [-8., 3., 2.], [0., 1., 1.]])
    P = M.T @ M
    q = np.array([3., 2., 3.]) @ M
    G = np.array([[1., 2., 1.], [2., 0., 1.], [-1., 2., -1.]])
    h = np.array([3., 2., -2.])
    A = np.array([[1., 1., 1.]]); b = np.array([1.])
    x = oracle_mod.solve_qp_qpsolvers(P, q, G, h, A, b)
    np.testing.assert_allclose(x, [0.307692307692308, -0.692307692307691, 1.38461538461538], atol=1e-12)
    f = 0.5 * x @ P @ x + q @ x
    assert abs(f - (-2.3076923076923)) < 1e-10


def _random_qp(rng, n, p, m, shift):
    Mx = rng.normal(size=(n, n))
    G = Mx @ Mx.T + 1e-3 * np.eye(n)
    return (G, rng.normal(size=n), rng.normal(size=(n, p)), rng.normal(size=p), rng.normal(size=(n, m)),
            rng.normal(size=m) + shift)


@pytest.mark.parametrize("n,p,m", [(48, 6, 96), (30, 6, 24), (45, 8, 36), (12, 6, 36), (5, 0, 7), (6, 2, 0)])
def test_port_bit_identical_to_reference_solver(oracle_mod, n, p, m):
    if oracle_mod.load_ref_quadprog() is None:
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    rng = np.random.default_rng(n * 1000 + m)
    for trial in range(40):
        qp = _random_qp(rng, n, p, m, shift=2.0 if trial % 2 else 0.0)
        r = oracle_mod.solve_quadprog(*qp)
        x2, f2, st2 = oracle_mod.ref_solve_quadprog(*qp)
        assert r['status'] == st2
        if st2 == 0:
            assert np.array_equal(r['x'], x2), "trial %d" % trial
            assert r['f'] == f2


def test_infeasible_and_not_pd_status(oracle_mod):
    G = np.eye(2); g0 = np.zeros(2)
    CI = np.array([[1., -1.], [0., 0.]]); ci0 = np.array([-1., -1.])  # x0 >= 1 and x0 <= -1
    r = oracle_mod.solve_quadprog(G, g0, None, None, CI, ci0)
    assert r['status'] == 1 and np.isinf(r['f'])
    r = oracle_mod.solve_quadprog(np.array([[1., 2.], [2., 1.]]), g0, None, None, CI, ci0)
    assert r['status'] == 2


def test_kkt_conditions(oracle_mod):
    rng = np.random.default_rng(5)
    G, g0, CE, ce0, CI, ci0 = _random_qp(rng, 20, 3, 30, 0.5)
    r = oracle_mod.solve_quadprog(G, g0, CE, ce0, CI, ci0)
    x = r['x']
    assert r['status'] == 0
    assert np.abs(CE.T @ x + ce0).max() < 1e-9
    s = CI.T @ x + ci0
    assert s.min() > -1e-9
    # stationarity: G x + g0 = CE ue + CI[:,act] ui with ui >= 0
    N = np.concatenate([CE, CI[:, r['active']]], axis=1)
    lam = np.linalg.lstsq(N, G @ x + g0, rcond=None)[0]
    assert np.abs(N @ lam - (G @ x + g0)).max() < 1e-8
    assert lam[3:].min() > -1e-9
    assert np.abs(s[r['active']]).max() < 1e-9
