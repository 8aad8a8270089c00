"""Diagnostic: quality of the null-space block Z and of the equality-constrained minimiser that
k_ihwbc_setup leaves in the per-state record, against numpy on the oracle's dense QP.
    python scripts/debug_setup_record.py robot B seed"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypnc_b200 import robots, engine  # noqa: E402
from oracle import oracle  # noqa: E402

KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')
name, B, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spec = robots.PROBLEMS[name]()
dm = engine.Model(spec.model, 0)
pb = engine.Problem(dm, spec)
ws = engine.Workspace(dm, B, spec.frames)
st = robots.synth_states(name, spec, B, seed=seed)
ws.update_system(*[torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in KEYS])
op = oracle.OracleProblem(spec, fast=False)
des, w, rfz = robots.synth_desireds(name, spec, st, op.task_actuals(st), seed=seed)
ref = op.tick_batch(st, des, w, rfz, nthreads=os.cpu_count() or 1)
out = engine.ihwbc_solve(pb, ws, *[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)])
torch.cuda.synchronize()
ptr, stride = C.c_void_p(), C.c_int32()
lay = (C.c_int32 * 10)()
assert dm.lib.pnc_debug_setup_record(pb.h, ws.h, C.byref(ptr), C.byref(stride), lay) == 0
n, p, ncol, npc, hw, RJ, RJB, RX, RS, REC = list(lay)
rec = engine._from_ptr(ptr.value, (B, stride.value), 0).cpu().numpy()
err = np.zeros(B)
for k in ('trq', 'acc', 'rf'):
    den = np.maximum(np.abs(ref[k]).max(1), 1e-3 * np.abs(ref[k]).max())
    err = np.maximum(err, np.abs(out[k].cpu().numpy() - ref[k]).max(1) / den)
order = np.argsort(-err)[:4].tolist() + [0, 1]
for b in order:
    r = rec[b]
    Z = np.zeros((n, npc))
    chunks = r[RJ:RJ + 32 * npc].reshape(npc // 2, 32, 2)
    for lane in range(min(n, 32)):
        Z[lane] = chunks[:, lane, :].reshape(-1)
    if n > 32:
        cb = r[RJB:RJB + 32 * hw].reshape(hw // 2, 32, 2)
        for e in range(n - 32):
            Z[32 + e, :hw] = cb[:, 2 * e, :].reshape(-1)
            Z[32 + e, hw:2 * hw] = cb[:, 2 * e + 1, :].reshape(-1)
    Z = Z[:, :ncol]
    x = r[RX:RX + n]
    q, v = op.om.pack_state(*[st[k][b] for k in KEYS])
    qp = op.assemble(q, v, des[b], w[b], rfz[b])
    P, g0, A, bb = qp['P'], qp['q'], qp['A'], qp['b']
    L = np.linalg.cholesky(P)
    J = np.linalg.solve(L, np.eye(n)).T
    x0 = -J @ (J.T @ g0)
    Q, R = np.linalg.qr(J.T @ A.T, mode='complete')
    xe = x0 - (J @ Q)[:, :p] @ np.linalg.solve(R[:p, :p].T, A @ x0 - bb)
    print("state %d err %.2e | Z^T P Z - I %.1e | A Z %.1e (|A| %.1e |Z| %.1e) | x_eq rel err %.1e | A x - b %.1e" %
          (b, err[b], np.abs(Z.T @ P @ Z - np.eye(ncol)).max(), np.abs(A @ Z).max(), np.abs(A).max(), np.abs(Z).max(),
           np.abs(x - xe).max() / np.abs(xe).max(), np.abs(A @ x - bb).max()))
