"""GPU: dump the Draco3 transmission-variant states on which the CUDA path and the oracle differ by > 1e-9."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypnc_b200 import robots, engine
from oracle import oracle
name, B, seed = 'draco3_trans', 1024, 5
spec = robots.PROBLEMS[name]()
st = robots.synth_states(name, spec, B, seed=seed)
op = oracle.OracleProblem(spec)
des, w, rfz = robots.synth_desireds(name, spec, st, op.task_actuals(st), seed=seed)
ref = op.tick_batch(st, des, w, rfz)
dm = engine.Model(spec.model, 0); pb = engine.Problem(dm, spec); ws = engine.Workspace(dm, B, spec.frames)
keys = ("base_pos", "base_quat", "base_lin_vel", "base_ang_vel", "joint_pos", "joint_vel")
ws.update_system(*[torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in keys])
res = {}
for mode, label in ((engine.TERM_REFINED, 'refined'), (engine.TERM_REFERENCE, 'reference')):
    engine.set_termination(mode)
    out = engine.ihwbc_solve(pb, ws, *[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)])
    torch.cuda.synchronize()
    err = np.zeros(B)
    for k in ('trq', 'acc', 'rf', 'qddot'):
        g = out[k].cpu().numpy()
        err = np.maximum(err, np.abs(g - ref[k]).max(1) / np.maximum(np.abs(ref[k]).max(1), 1e-3 * np.abs(ref[k]).max()))
        res[label + '_' + k] = g
    res[label + '_active'] = engine.unpack_active(out['active'], pb.m)
    res[label + '_active_ref'] = engine.unpack_active(out['active_ref'], pb.m)
    res[label + '_iters'] = out['iters'].cpu().numpy()
    res[label + '_status'] = out['status'].cpu().numpy()
    bad = np.where(err > 1e-9)[0]
    print(label, 'bad', bad.tolist(), err[bad].tolist(), 'psi_rel', ref['psi_rel'][bad].tolist(), 'margin', ref['margin'][bad].tolist(),
          'iters gpu', res[label + '_iters'][bad].tolist(), 'oracle', ref['iters'][bad].tolist())
np.savez('gpurun_out/dbg_trans.npz', **res)
