"""Diagnostic: the states of a synthetic batch where the CUDA path is furthest from the oracle.
    python scripts/worst_states.py robot N seed [out.npz]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypnc_b200 import robots, engine  # noqa: E402
from oracle import oracle  # noqa: E402

KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')
name, N, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spec = robots.PROBLEMS[name]()
st = robots.synth_states(name, spec, N, seed=seed)
op = oracle.OracleProblem(spec)
des, w, rfz = robots.synth_desireds(name, spec, st, op.task_actuals(st), seed=seed)
ref = op.tick_batch(st, des, w, rfz, nthreads=os.cpu_count() or 1)
dm = engine.Model(spec.model, 0)
pb = engine.Problem(dm, spec)
ws = engine.Workspace(dm, N, spec.frames)
ws.update_system(*[torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in KEYS])
d_in = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)]
for mode in (0, 1):
    engine.set_termination(mode)
    out = engine.ihwbc_solve(pb, ws, *d_in)
    torch.cuda.synchronize()
    err = np.zeros(N)
    for k in ('trq', 'acc', 'rf'):
        err = np.maximum(err, np.abs(out[k].cpu().numpy() - ref[k]).max(1) / np.abs(ref[k]).max(1))
    order = np.argsort(-np.nan_to_num(err, nan=1e9))[:8]
    a_fin = engine.unpack_active(out['active'], pb.m)
    print("mode", mode)
    for b in order:
        print("  state %d err %.3e status gpu %d ref %d iters gpu %d ref %d margin %.2e kind %d | active gpu %s ref %s" %
              (b, err[b], int(out['status'][b]), ref['status'][b], int(out['iters'][b]), ref['iters'][b], ref['margin'][b],
               ref['margin_kind'][b], np.where(a_fin[b])[0].tolist(), np.where(ref['active'][b])[0].tolist()))
    if mode == 0 and len(sys.argv) > 4:
        b = order[0]
        np.savez(sys.argv[4], b=b, **{k: st[k][b] for k in KEYS}, des=des[b], w=w[b], rfz=rfz[b],
                 gpu_trq=out['trq'][b].cpu().numpy(), gpu_rf=out['rf'][b].cpu().numpy(), gpu_qddot=out['qddot'][b].cpu().numpy(),
                 ref_trq=ref['trq'][b], ref_rf=ref['rf'][b], ref_qddot=ref['qddot'][b])
engine.set_termination(0)
