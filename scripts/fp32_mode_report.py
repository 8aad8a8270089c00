"""GPU report of the optional fp32 mode (pnc_set_precision(PNC_PREC_FP32_DYNAMICS)): per robot, on B states of the
bench generator, (1) how far the rigid-body quantities of k_update_system<float> are from the fp64 kernel's,
(2) how far trq / acc / rf move against the fp64 path, split by "same active set as fp64" / "different", and
(3) the time of the update in both modes.  Writes one JSON line per robot to stdout."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypnc_b200 import robots, engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
seed = 21
keys = ("base_pos", "base_quat", "base_lin_vel", "base_ang_vel", "joint_pos", "joint_vel")
FIELDS = ('M', 'grav', 'cori', 'com', 'vcom', 'jcom', 'jcomdot', 'jcom_jdqd', 'frame_iso', 'frame_vel', 'frame_jac',
          'frame_jdqd', 'q', 'qdot')


def time_update(ws, d_st, reps=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        ws.update_system(*d_st)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        ws.update_system(*d_st)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


for name in ('atlas', 'laikago', 'valkyrie', 'draco3', 'draco3_lb'):
    spec = robots.PROBLEMS[name]()
    st = robots.synth_states(name, spec, B, seed=seed)
    dm = engine.Model(spec.model, 0)
    pb = engine.Problem(dm, spec)
    ws = engine.Workspace(dm, B, spec.frames)
    d_st = [torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in keys]
    engine.set_precision(engine.PREC_FP64)
    ws.update_system(*d_st)
    des, w, rfz = robots.synth_desireds(name, spec, st, engine.task_actuals_from_workspace(spec, ws), seed=seed)
    d_in = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)]
    res, fields, ms = {}, {}, {}
    for mode, tag in ((engine.PREC_FP64, 'fp64'), (engine.PREC_FP32_DYNAMICS, 'fp32')):
        engine.set_precision(mode)
        ws.update_system(*d_st)
        fields[tag] = {f: ws.view(f).clone() for f in FIELDS}
        out = engine.ihwbc_solve(pb, ws, *d_in)
        torch.cuda.synchronize()
        res[tag] = {k: v.cpu().numpy() for k, v in out.items()}
        ms[tag] = time_update(ws, d_st)
    engine.set_precision(engine.PREC_FP64)
    rec = {'robot': name, 'states': B, 'update_ms': ms, 'fields_rel_err': {}}
    for f in FIELDS:
        a, b = fields['fp64'][f], fields['fp32'][f]
        scale = a.abs().reshape(B, -1).max(1).values.clamp_min(1e-30)
        e = ((a - b).abs().reshape(B, -1).max(1).values / scale)
        rec['fields_rel_err'][f] = float(e.max())
    r64, r32 = res['fp64'], res['fp32']
    ok = (r64['status'] == 0) & (r32['status'] == 0)
    same = ok & (r64['active_ref'] == r32['active_ref']).reshape(B, -1).all(1) & \
        (r64['active'] == r32['active']).reshape(B, -1).all(1)
    err = np.zeros(B)
    for k in ('trq', 'acc', 'rf'):
        den = np.maximum(np.abs(r64[k]).max(1), 1e-3 * np.abs(r64[k]).max())
        err = np.maximum(err, np.abs(r32[k] - r64[k]).max(1) / den)
    rec['status_mismatch'] = int((r64['status'] != r32['status']).sum())
    rec['same_active_set'] = int(same.sum())
    rec['ok'] = int(ok.sum())
    for tag, sel in (('same_set', same), ('other', ok & ~same), ('all', ok)):
        e = err[sel]
        rec[tag] = None if e.size == 0 else {
            'n': int(e.size), 'max': float(e.max()), 'p99': float(np.quantile(e, 0.99)),
            'median': float(np.median(e)), 'over_1e-4': int((e > 1e-4).sum())}
    print(json.dumps(rec), flush=True)
