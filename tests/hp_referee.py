"""High-precision referee for one IHWBC tick (test infrastructure).

When the CUDA path and the fp64 oracle differ by more than the 1e-9 bar on a state, somebody has to
say which side is off.  Both carry cond(G) * eps of rounding from the factorisation of an
ill-conditioned Hessian (lambda_q_ddot = 1e-8, lambda_rf = 1e-7, config/atlas_config.py:65-66) and, for
robots with an internal constraint, from the two pseudo-inverses of ihwbc.py:130-138.  This module
restates ihwbc.py:124-356 for ONE state in 50-digit arithmetic (mpmath) from fp64 inputs that carry no
amplified error (M, c + g, task Jacobians and commands, contact Jacobians, cone rows -- taken from the
oracle's per-state getters) and solves the KKT system of the final active set exactly.  The result is
the yardstick: a side is "right" when it is as close to it as fp64 allows.
"""
import numpy as np
import mpmath as mp

mp.mp.dps = 50

KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')


def _M(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    return mp.matrix(a.tolist())


def _np(m):
    return np.array([[float(m[i, j]) for j in range(m.cols)] for i in range(m.rows)])


def solve_state(op, spec, st, des, w, rfz, active):
    """op: oracle.OracleProblem; st: dict of ONE state's inputs; active: bool [m] final active set.
    Returns dict(qddot, rf, trq, acc) as float64 arrays computed in 50-digit arithmetic."""
    om = op.om
    q, v = om.pack_state(*[st[k] for k in KEYS])
    nv = spec.model.nv
    M = _M(om.mass_matrix(q))
    cg = _M(om.coriolis(q, v) + om.gravity(q))
    # cost (ihwbc.py:159-175) in high precision from the per-task Jacobians and commands
    P = mp.zeros(nv, nv)
    qv = mp.zeros(nv, 1)
    for it in range(len(spec.tasks)):
        e = op.task_eval(q, v, des, it)
        J = _M(e['jacobian'])
        r = _M(e['jacobian_dot_q_dot'] - e['op_cmd'])
        wi = mp.mpf(float(w[it]))
        P += wi * (J.T * J)
        qv += wi * (J.T * r)
    P += mp.mpf(spec.lambda_q_ddot) * M
    Jc_rows, Uf_blocks, uf = [], [], []
    for ic in range(len(spec.contacts)):
        c = op.contact_eval(q, v, float(rfz[ic]), ic)
        Jc_rows.append(c['jacobian'])
        Uf_blocks.append(c['cone_constraint_mat'])
        uf.append(c['cone_constraint_vec'])
    nc = sum(b.shape[1] for b in Uf_blocks)
    nu = sum(b.shape[0] for b in Uf_blocks)
    Jc = _M(np.concatenate(Jc_rows, 0)) if nc else mp.zeros(0, nv)
    Uf = np.zeros((nu, nc))
    r0 = c0 = 0
    for b in Uf_blocks:
        Uf[r0:r0 + b.shape[0], c0:c0 + b.shape[1]] = b
        r0 += b.shape[0]
        c0 += b.shape[1]
    uf = np.concatenate(uf) if nu else np.zeros(0)
    n = nv + nc
    nact = spec.nact
    Sa = mp.zeros(nact, nv)
    for i, a in enumerate(spec.act_idx):
        Sa[i, a] = 1
    # internal constraints (ihwbc.py:124-146)
    k = spec.n_ic
    I = mp.eye(nv)
    if k > 0:
        Ji = _M(spec.ic_jac)
        Minv = M ** -1
        lmd = (Ji * Minv * Ji.T) ** -1
        Ni = I - Minv * Ji.T * lmd * Ji
        A = (Sa * Ni)[:, 6:nv]
        W = Minv[6:nv, 6:nv]
        Abar = W * A.T * ((A * W * A.T) ** -1)
        B = Abar.T  # nact x (nv - 6)
    else:
        Ni = I
        B = Sa[:, 6:nv]
    S = mp.zeros(nact, nv)
    S[:, 6:nv] = B
    JcNi = Jc * Ni
    # full cost
    G = mp.zeros(n, n)
    G[:nv, :nv] = P
    for i in range(nc):
        G[nv + i, nv + i] = mp.mpf(spec.lambda_rf)
    g = mp.zeros(n, 1)
    g[:nv, 0] = qv
    # equalities (:225-249)
    rows = []
    rhs = []
    Sf_M = M[:6, :]
    Sf_J = -(JcNi.T)[:6, :]
    Ntcg = Ni.T * cg
    for i in range(6):
        rows.append([Sf_M[i, j] for j in range(nv)] + [Sf_J[i, j] for j in range(nc)])
        rhs.append(-Ntcg[i, 0])
    for i in range(k):
        rows.append([mp.mpf(float(spec.ic_jac[i, j])) for j in range(nv)] + [mp.mpf(0)] * nc)
        rhs.append(mp.mpf(0))
    # IHWBC2 transmission rows (ihwbc2.py:249-256): (Sd - Sv) [M | -Jc^T] x = -(Sd - Sv)(c + g)
    trans = list(zip(getattr(spec, 'trans_distal', []), getattr(spec, 'trans_passive', [])))
    JcT = Jc.T
    for vd, vp in trans:
        rows.append([M[vd, j] - M[vp, j] for j in range(nv)] + [-(JcT[vd, j] - JcT[vp, j]) for j in range(nc)])
        rhs.append(-(cg[vd, 0] - cg[vp, 0]))
    k += len(trans)
    # active inequalities as equalities: cone rows  Uf rf = uf ... (:255-317: -Uf rf <= -uf), torque rows
    SM = S * M
    SJ = S * JcNi.T
    Scg = S * Ntcg
    act = np.where(np.asarray(active).astype(bool))[0]
    orient = []  # +1: row written as a.x >= b (cones, lower torque limits); -1: a.x <= b (upper torque limits)
    for i in act:
        orient.append(-1.0 if (i >= nu and (i - nu) >= nact) else 1.0)
        if i < nu:
            rows.append([mp.mpf(0)] * nv + [mp.mpf(float(Uf[i, j])) for j in range(nc)])
            rhs.append(mp.mpf(float(uf[i])))
        else:
            a = (i - nu) % nact
            upper = (i - nu) >= nact
            rows.append([SM[a, j] for j in range(nv)] + [-SJ[a, j] for j in range(nc)])
            lim = spec.trq_limit[a, 1] if upper else spec.trq_limit[a, 0]
            rhs.append(mp.mpf(float(lim)) - Scg[a, 0])
    Aeq = mp.matrix(rows)
    beq = mp.matrix(rhs)
    ne = Aeq.rows
    K = mp.zeros(n + ne, n + ne)
    K[:n, :n] = G
    K[:n, n:] = Aeq.T
    K[n:, :n] = Aeq
    r = mp.zeros(n + ne, 1)
    r[:n, 0] = -g
    r[n:, 0] = beq
    sol = mp.lu_solve(K, r)
    x = sol[:n, 0]
    qdd, rf = x[:nv, 0], x[nv:n, 0]
    gen = M * qdd + Ntcg - (JcNi.T * rf if nc else mp.zeros(nv, 1))
    if trans:
        # ihwbc2.py:131-139,386-403: trq = sa_trc_bar^T Snf (gen - Ji^T f_int), f_int = Sv gen, Ji row = +1 at the
        # passive and -1 at the distal joint, sa_trc_bar = weighted_pinv(Sa[:, 6:], Minv[6:, 6:])
        Minv = M ** -1
        A = Sa[:, 6:nv]
        W = Minv[6:nv, 6:nv]
        S = mp.zeros(nact, nv)
        S[:, 6:nv] = (W * A.T * ((A * W * A.T) ** -1)).T
        for vd, vp in trans:
            f_int = gen[vp, 0]
            gen[vp, 0] -= f_int
            gen[vd, 0] += f_int
    trq = S * gen
    acc = Sa * qdd
    lam = sol[n + 6 + k:, 0]  # multipliers of the active inequalities (sign: all of one sign at an optimum)
    return dict(qddot=_np(qdd)[:, 0], rf=_np(rf)[:, 0] if nc else np.zeros(0), trq=_np(trq)[:, 0], acc=_np(acc)[:, 0],
                multipliers=_np(lam)[:, 0] * np.array(orient) if lam.rows else np.zeros(0))
