"""Active-set parity report: per robot, on N fresh states, how often the set
at the reference's termination point (`active_ref`) and the refined final set (`active`) are
identical to the oracle's, split by whether the oracle's own path contains an exact tie
(decision margin <= 1e-10, oracle/pnc_oracle.c: orc_last_qp_margin).  Also runs the batch with
PNC_TERM_REFERENCE and reports the worst output difference against the oracle in both modes.

    python scripts/active_set_report.py [N] [out.json]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypnc_b200 import robots, engine  # noqa: E402
from oracle import oracle  # noqa: E402

KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')
TIE = 1e-10


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rows = {}
    for name in ('atlas', 'laikago', 'valkyrie', 'draco3', 'draco3_lb'):
        spec = robots.PROBLEMS[name]()
        st = robots.synth_states(name, spec, N, seed=21)
        op = oracle.OracleProblem(spec)
        des, w, rfz = robots.synth_desireds(name, spec, st, op.task_actuals(st), seed=21)
        ref = op.tick_batch(st, des, w, rfz, nthreads=os.cpu_count() or 1)
        dm = engine.Model(spec.model, 0)
        pb = engine.Problem(dm, spec)
        ws = engine.Workspace(dm, N, spec.frames)
        ws.update_system(*[torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in KEYS])
        d_in = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)]
        row = {}
        for mode, label in ((engine.TERM_REFINED, 'refined'), (engine.TERM_REFERENCE, 'reference')):
            engine.set_termination(mode)
            out = engine.ihwbc_solve(pb, ws, *d_in)
            torch.cuda.synchronize()
            ok = (out['status'].cpu().numpy() == 0) & (ref['status'] == 0)
            a_ref = engine.unpack_active(out['active_ref'], pb.m)
            a_fin = engine.unpack_active(out['active'], pb.m)
            clear = ref['margin'] > TIE
            same_ref = (a_ref == ref['active']).all(1)
            same_fin = (a_fin == ref['active']).all(1)
            err = max(float((np.abs(out[k].cpu().numpy()[ok] - ref[k][ok]).max(1) / np.abs(ref[k][ok]).max(1)).max())
                      for k in ('trq', 'acc', 'rf'))
            row[label] = dict(states=int(ok.sum()), clear=int((ok & clear).sum()),
                              clear_identical_active_ref=int((ok & clear & same_ref).sum()),
                              tied=int((ok & ~clear).sum()), tied_identical_active_ref=int((ok & ~clear & same_ref).sum()),
                              identical_active_ref=int((ok & same_ref).sum()), identical_final=int((ok & same_fin).sum()),
                              status_equal=bool(np.array_equal(out['status'].cpu().numpy(), ref['status'])),
                              max_rel_err_trq_acc_rf=err)
        engine.set_termination(engine.TERM_REFINED)
        rows[name] = row
        print(name, json.dumps(row))
    if len(sys.argv) > 2:
        with open(sys.argv[2], 'w') as f:
            json.dump(dict(n=N, seed=21, tie_margin=TIE, robots=rows), f, indent=1)


if __name__ == '__main__':
    main()
