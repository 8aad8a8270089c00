"""Diagnostic: worst strict states of a full-batch parity run, settled by the 50-digit referee.
    python scripts/debug_fullbatch.py robot B seed"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypnc_b200 import robots, engine, _cabi  # noqa: E402
if os.environ.get('PNC_LIB'):  # A/B against another build of the library (diagnostic only)
    _cabi.LIB_PATH = os.environ['PNC_LIB']
from oracle import oracle  # noqa: E402
from tests import hp_referee  # noqa: E402

KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')
name, B, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spec = robots.PROBLEMS[name]()
dm = engine.Model(spec.model, 0)
pb = engine.Problem(dm, spec)
ws = engine.Workspace(dm, B, spec.frames)
st = robots.synth_states(name, spec, B, seed=seed)
d_st = [torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in KEYS]
ws.update_system(*d_st)
des, w, rfz = robots.synth_desireds(name, spec, st, engine.task_actuals_from_workspace(spec, ws), seed=seed)
d_in = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)]
op = oracle.OracleProblem(spec, fast=False)
ref = op.tick_batch(st, des, w, rfz, nthreads=os.cpu_count() or 1)
out = engine.ihwbc_solve(pb, ws, *d_in)
torch.cuda.synchronize()
o = {k: out[k].cpu().numpy() for k in ('trq', 'acc', 'rf', 'qddot')}
err = np.zeros(B)
for k in ('trq', 'acc', 'rf'):
    den = np.maximum(np.abs(ref[k]).max(1), 1e-6 * np.abs(ref[k]).max())
    err = np.maximum(err, np.abs(o[k] - ref[k]).max(1) / den)
strict = (ref['margin'] > 1e-10) & (ref['psi_rel'] <= 1e-11) & (ref['status'] == 0)
order = np.argsort(-np.where(strict, err, 0))[:6]
a_fin = engine.unpack_active(out['active'], pb.m)
for b in order:
    hp = hp_referee.solve_state(op, spec, {k: st[k][b] for k in KEYS}, des[b], w[b], rfz[b], ref['active'][b])
    eg = max(np.abs(o[k][b] - hp[k]).max() / max(np.abs(hp[k]).max(), 1e-300) for k in ('trq', 'acc', 'rf'))
    eo = max(np.abs(ref[k][b] - hp[k]).max() / max(np.abs(hp[k]).max(), 1e-300) for k in ('trq', 'acc', 'rf'))
    print("state %d err %.2e | cuda vs exact %.2e oracle vs exact %.2e | iters gpu %d ref %d margin %.1e psi_rel %.1e | same set %s | mult %s" %
          (b, err[b], eg, eo, int(out['iters'][b]), ref['iters'][b], ref['margin'][b], ref['psi_rel'][b],
           np.array_equal(a_fin[b], ref['active'][b]), np.round(hp['multipliers'], 3).tolist()))
