"""pnc_quadprog_batch (dense batched Goldfarb-Idnani) against the oracle port of the reference's
QuadProg++ -- which tests/test_oracle_qp.py pins bit-for-bit to the reference's own compiled solver --
and, when oracle/_ref travelled to the box, against that compiled reference solver itself."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu
TOL = 1e-9


def random_qps(rng, B, n, p, m, shift):
    Mx = rng.normal(size=(B, n, n))
    G = Mx @ np.transpose(Mx, (0, 2, 1)) + 1e-3 * np.eye(n)
    return (G, rng.normal(size=(B, n)), rng.normal(size=(B, n, p)), rng.normal(size=(B, p)), rng.normal(size=(B, n, m)),
            rng.normal(size=(B, m)) + shift)


def wrench_distribution_qps(rng, B):
    """The QP of DCMPlanner's double-support wrench distribution (towr_plus/src/dcm_planner.cpp:752-855):
    n = 12, p = 6, m = 36; G diagonal, CE = [Ad(lf_T_base)^T | Ad(rf_T_base)^T], CI = U blockdiag(R^T)."""
    from scipy.spatial.transform import Rotation as R

    def adjoint(Rm, pv):
        px = np.array([[0, -pv[2], pv[1]], [pv[2], 0, -pv[0]], [-pv[1], pv[0], 0]])
        A = np.zeros((6, 6)); A[:3, :3] = Rm; A[3:, 3:] = Rm; A[3:, :3] = px @ Rm
        return A

    def U(x, y, mu):  # the 18 x 6 surface wrench cone the planner and SurfaceContact share (basic_contact.py:89-173)
        rows = [[0, 0, 0, 0, 0, 1], [0, 0, 0, 1, 0, mu], [0, 0, 0, -1, 0, mu], [0, 0, 0, 0, 1, mu], [0, 0, 0, 0, -1, mu],
                [1, 0, 0, 0, 0, y], [-1, 0, 0, 0, 0, y], [0, 1, 0, 0, 0, x], [0, -1, 0, 0, 0, x]]
        for sx in (-1, 1):
            for sy in (-1, 1):
                rows.append([sx * mu, sy * mu, 1, sx * y, sy * x, (x + y) * mu])
                rows.append([sx * mu, sy * mu, -1, -sx * y, -sy * x, (x + y) * mu])
        rows.append([0, 0, 0, 0, 0, -1])
        return np.array(rows, float)

    Um = U(0.11, 0.065, 0.3)
    out = [np.zeros((B, 12, 12)), np.zeros((B, 12)), np.zeros((B, 12, 6)), np.zeros((B, 6)), np.zeros((B, 12, 36)),
           np.zeros((B, 36))]
    for b in range(B):
        lf, rf = rng.uniform(0.01, 1.0), rng.uniform(0.01, 1.0)
        g = np.concatenate([np.full(6, lf), np.full(6, rf)]); g[5] = 0.001 * lf; g[11] = 0.001 * rf
        out[0][b] = np.diag(g)
        Rl, Rr = R.from_rotvec(rng.normal(0, 0.1, 3)).as_matrix(), R.from_rotvec(rng.normal(0, 0.1, 3)).as_matrix()
        pl, pr = np.array([0.0, 0.1, -0.9]) + rng.normal(0, 0.02, 3), np.array([0.0, -0.1, -0.9]) + rng.normal(0, 0.02, 3)
        CE = np.zeros((6, 12)); CE[:, :6] = adjoint(Rl.T, -Rl.T @ pl).T; CE[:, 6:] = adjoint(Rr.T, -Rr.T @ pr).T
        out[2][b] = CE.T
        out[3][b] = -(np.array([0, 0, 0, 0, 0, 180 * 9.81]) + rng.normal(0, 20, 6))
        Ut = np.zeros((36, 12)); Ut[:18, :6] = Um; Ut[18:, 6:] = Um
        Rf = np.zeros((12, 12)); Rf[:3, :3] = Rl.T; Rf[3:6, 3:6] = Rl.T; Rf[6:9, 6:9] = Rr.T; Rf[9:, 9:] = Rr.T
        out[4][b] = (Ut @ Rf).T
        out[5][b, 17] = 2000.0 * lf; out[5][b, 35] = 2000.0 * rf
    return tuple(out)


def run_gpu(qp):
    from pypnc_b200 import engine
    d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in qp]
    d = [None if t.numel() == 0 else t for t in d]
    x, f, st, it, act = engine.quadprog_batch(*d, want_active=True)
    torch.cuda.synchronize()
    return x.cpu().numpy(), f.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy(), act.cpu().numpy()


def check(qp, label):
    from oracle import oracle as O
    x, f, st, it, act = run_gpu(qp)
    B, m = qp[0].shape[0], qp[4].shape[2]
    ref_lib = O.load_ref_quadprog()
    nbad_active = 0
    for b in range(B):
        r = O.solve_quadprog(*[a[b] for a in qp])
        assert st[b] == r['status'], (label, b, st[b], r['status'])
        if r['status'] != 0:
            continue
        scale = max(1.0, np.abs(r['x']).max())
        assert np.abs(x[b] - r['x']).max() <= TOL * scale, (label, b, np.abs(x[b] - r['x']).max())
        assert abs(f[b] - r['f']) <= TOL * max(1.0, abs(r['f'])), (label, b, f[b], r['f'])
        got = sorted(i for i in range(m) if (int(act[b, i // 64]) >> (i % 64)) & 1)
        if got != r['active']:
            # a differing set is only acceptable on a degenerate vertex: every disputed constraint is tight
            sl = qp[4][b].T @ r['x'] + qp[5][b]
            assert all(abs(sl[i]) <= 1e-7 * max(1.0, np.abs(qp[5][b]).max()) for i in set(got) ^ set(r['active'])), (label, b)
            nbad_active += 1
        if ref_lib is not None and b < 8:
            x2, f2, st2 = O.ref_solve_quadprog(*[a[b] for a in qp])
            assert st2 == st[b] and np.abs(x[b] - x2).max() <= TOL * scale
    return nbad_active


@pytest.mark.parametrize("n,p,m", [(12, 6, 36), (48, 6, 96), (30, 6, 24), (45, 8, 36), (5, 0, 7), (6, 2, 0), (64, 10, 128)])
def test_random_dense_qps(n, p, m):
    rng = np.random.default_rng(100 * n + m)
    for shift in (0.0, 2.0):
        nb = check(random_qps(rng, 48, n, p, m, shift), 'random %d/%d/%d' % (n, p, m))
        assert nb == 0  # generic position: the active set is unique


def test_wrench_distribution_qp():
    rng = np.random.default_rng(12)
    check(wrench_distribution_qps(rng, 256), 'wrench distribution')


def test_status_codes_and_sizes():
    from pypnc_b200 import engine
    from pypnc_b200._cabi import load_library
    G = np.tile(np.eye(2), (3, 1, 1)); g0 = np.zeros((3, 2))
    CI = np.tile(np.array([[1., -1.], [0., 0.]]), (3, 1, 1)); ci0 = np.tile(np.array([-1., -1.]), (3, 1))
    G[1] = np.array([[1., 2.], [2., 1.]])       # not positive definite
    ci0[2] = np.array([1., 1.])                 # feasible: -1 <= x0 <= 1
    x, f, st, it, act = run_gpu((G, g0, np.zeros((3, 2, 0)), np.zeros((3, 0)), CI, ci0))
    assert list(st) == [1, 2, 0] and np.isinf(f[0]) and np.isnan(x[0]).all() and np.abs(x[2]).max() == 0.0
    lib = load_library()
    assert lib.pnc_quadprog_batch(0, 1, 65, 0, 0, 1, 1, None, None, None, None, 1, None, None, None, None, None) == -3
    assert lib.pnc_quadprog_batch(0, 1, 0, 0, 0, 1, 1, None, None, None, None, 1, None, None, None, None, None) == -1


def test_throughput_property_64k():
    """Full-size batch of the planner's QP: size-independent checks (feasibility, KKT stationarity)."""
    rng = np.random.default_rng(3)
    base = wrench_distribution_qps(rng, 512)
    rep = 128
    qp = tuple(np.tile(a, (rep,) + (1,) * (a.ndim - 1)) for a in base)
    x, f, st, it, act = run_gpu(qp)
    ok = st == 0
    assert set(np.unique(st)) <= {0, 1} and ok.mean() > 0.3  # some random external wrenches exceed the friction cones
    G, g0, CE, ce0, CI, ci0 = qp
    eq = np.einsum('bnp,bn->bp', CE[ok], x[ok]) + ce0[ok]
    ineq = np.einsum('bnm,bn->bm', CI[ok], x[ok]) + ci0[ok]
    assert np.abs(eq).max() <= 1e-8 * 2000 and ineq.min() >= -1e-8 * 2000
    # the same problem gives the same bits (and the same verdict) wherever in the batch it runs
    assert np.array_equal(st[:512], st[512:1024]) and np.array_equal(x[:512][ok[:512]], x[65024:][ok[65024:]])
