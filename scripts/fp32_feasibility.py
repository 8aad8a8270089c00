"""GPU experiment: how far do the IHWBC outputs move when every rigid-body quantity the solve stage reads is
rounded to fp32 (the floor of any fp32 dynamics mode)?  Compared with the fp64 oracle per robot."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypnc_b200 import robots, engine
from oracle import oracle
B, seed = 4096, 21
keys = ("base_pos", "base_quat", "base_lin_vel", "base_ang_vel", "joint_pos", "joint_vel")
for name in ('atlas', 'laikago', 'valkyrie', 'draco3', 'draco3_lb'):
    spec = robots.PROBLEMS[name]()
    st = robots.synth_states(name, spec, B, seed=seed)
    op = oracle.OracleProblem(spec)
    des, w, rfz = robots.synth_desireds(name, spec, st, op.task_actuals(st), seed=seed)
    ref = op.tick_batch(st, des, w, rfz, nthreads=os.cpu_count())
    dm = engine.Model(spec.model, 0); pb = engine.Problem(dm, spec); ws = engine.Workspace(dm, B, spec.frames)
    ws.update_system(*[torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in keys])
    d_in = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)]
    for rounded in (False, True):
        if rounded:
            for f in ('M', 'grav', 'cori', 'jcom', 'jcom_jdqd', 'frame_iso', 'frame_vel', 'frame_jac', 'frame_jdqd', 'com', 'vcom'):
                v = ws.view(f)
                v.copy_(v.float().double())
        out = engine.ihwbc_solve(pb, ws, *d_in)
        torch.cuda.synchronize()
        stt = out['status'].cpu().numpy()
        ok = (stt == 0) & (ref['status'] == 0)
        msg = []
        for k in ('trq', 'acc', 'rf'):
            g = out[k].cpu().numpy()
            den = np.maximum(np.abs(ref[k]).max(1), 1e-3 * np.abs(ref[k]).max())
            e = np.abs(g - ref[k]).max(1) / den
            e = e[ok]
            msg.append('%s max %.2e p99 %.2e med %.2e' % (k, e.max(), np.quantile(e, 0.99), np.median(e)))
        print(name, 'rounded' if rounded else 'fp64   ', 'status mismatch', int((stt != ref['status']).sum()), '|', ' | '.join(msg), flush=True)
