"""Diagnostic: CUDA path vs CPU oracle, printed error table (run on the GPU box)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypnc_b200 import robots, engine
from oracle import oracle


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def main(names, B=128):
    for name in names:
        spec = robots.PROBLEMS[name]()
        m = spec.model
        st = robots.synth_states(name, spec, B, seed=0)
        op = oracle.OracleProblem(spec)
        act = op.task_actuals(st)
        des, w, rfz = robots.synth_desireds(name, spec, st, act, seed=0)
        ref = op.tick_batch(st, des, w, rfz)
        dm = engine.Model(m)
        pb = engine.Problem(dm, spec)
        ws = engine.Workspace(dm, B, spec.frames)
        dev = {k: torch.from_numpy(v).cuda() for k, v in st.items()}
        ws.update_system(dev['base_pos'], dev['base_quat'], dev['base_lin_vel'], dev['base_ang_vel'],
                         dev['joint_pos'], dev['joint_vel'])
        torch.cuda.synchronize()
        # rigid-body quantities
        errs = {}
        M = ws.view('M').cpu().numpy(); g = ws.view('grav').cpu().numpy(); c = ws.view('cori').cpu().numpy()
        com = ws.view('com').cpu().numpy(); vcom = ws.view('vcom').cpu().numpy()
        jcom = ws.view('jcom').cpu().numpy(); jcd = ws.view('jcomdot').cpu().numpy()
        fiso = ws.view('frame_iso').cpu().numpy(); fvel = ws.view('frame_vel').cpu().numpy()
        fjac = ws.view('frame_jac').cpu().numpy(); fjd = ws.view('frame_jdqd').cpu().numpy()
        qg = ws.view('q').cpu().numpy(); vg = ws.view('qdot').cpu().numpy()
        for k in ('M', 'g', 'c', 'com', 'vcom', 'jcom', 'jcomdot', 'iso', 'fvel', 'fjac', 'fjdqd', 'q', 'v'):
            errs[k] = 0.0
        for b in range(min(B, 32)):
            q, v = op.om.pack_state(*[st[k][b] for k in ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')])
            errs['q'] = max(errs['q'], rel(qg[b], q)); errs['v'] = max(errs['v'], rel(vg[b], v))
            errs['M'] = max(errs['M'], rel(M[b], op.om.mass_matrix(q)))
            errs['g'] = max(errs['g'], rel(g[b], op.om.gravity(q)))
            errs['c'] = max(errs['c'], rel(c[b], op.om.coriolis(q, v)))
            cm, vc = op.om.com(q, v)
            errs['com'] = max(errs['com'], rel(com[b], cm)); errs['vcom'] = max(errs['vcom'], rel(vcom[b], vc))
            errs['jcom'] = max(errs['jcom'], rel(jcom[b], op.om.com_jacobian(q)))
            errs['jcomdot'] = max(errs['jcomdot'], rel(jcd[b], op.om.com_jacobian_dot(q, v)))
            for fi, f in enumerate(spec.frames):
                T = op.om.link_iso(q, f)
                iso = np.concatenate([T[:3, :3].ravel(), T[:3, 3]])
                errs['iso'] = max(errs['iso'], rel(fiso[b, fi], iso))
                errs['fvel'] = max(errs['fvel'], rel(fvel[b, fi], op.om.link_vel(q, v, f)))
                errs['fjac'] = max(errs['fjac'], rel(fjac[b, fi], op.om.link_jacobian(q, f)))
                errs['fjdqd'] = max(errs['fjdqd'], rel(fjd[b, fi], op.om.link_jdqd(q, v, f)))
        print(name, 'dyn rel errs:', ' '.join('%s=%.1e' % kv for kv in errs.items()), flush=True)
        d_des, d_w, d_rfz = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)]
        t0 = time.time()
        out = engine.ihwbc_solve(pb, ws, d_des, d_w, d_rfz)
        torch.cuda.synchronize()
        t1 = time.time()
        status = out['status'].cpu().numpy(); iters = out['iters'].cpu().numpy()
        active = engine.unpack_active(out['active'], pb.m)
        print(name, 'solve wall %.3f s  status' % (t1 - t0), np.bincount(status, minlength=5), 'ref status', np.bincount(ref['status'], minlength=5))
        ok = (status == 0) & (ref['status'] == 0)
        per = lambda k: np.abs(out[k].cpu().numpy() - ref[k]).max(1) / np.maximum(1e-300, np.abs(ref[k]).max(1))
        for k in ('trq', 'acc', 'rf', 'qddot'):
            e = per(k)[ok]
            print('   %-6s rel err max %.2e  median %.2e  n>1e-9: %d' % (k, e.max() if e.size else -1, np.median(e) if e.size else -1, int((e > 1e-9).sum())))
        same = (active == ref['active']).all(1)
        print('   active set identical: %d / %d   iters gpu mean %.2f ref mean %.2f, identical iters %d' % (same[ok].sum(), ok.sum(), iters.mean(), ref['iters'].mean(), (iters == ref['iters']).sum()))
        if not same[ok].all():
            b = int(np.where(ok & ~same)[0][0])
            print('   first mismatch state', b, 'gpu', np.where(active[b])[0], 'ref', np.where(ref['active'][b])[0])


if __name__ == '__main__':
    names = sys.argv[1:] or ['atlas', 'laikago', 'valkyrie']
    main(names)
