"""Diagnostic: how far the CUDA path is from the golden vectors of the reference's own Python classes (max relative error per robot)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypnc_b200 import robots, engine
KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')
rel = lambda a, b: (np.abs(a - b).max(axis=1) / np.maximum(1e-300, np.abs(b).max(axis=1))).max()
for name in ('atlas', 'laikago', 'valkyrie', 'draco3', 'draco3_lb'):
    z = np.load(os.path.join(ROOT, 'tests/golden', name + '.npz'))
    spec = robots.PROBLEMS[name]()
    B = z['in_des'].shape[0]
    dm = engine.Model(spec.model, 0); pb = engine.Problem(dm, spec); ws = engine.Workspace(dm, B, spec.frames)
    ws.update_system(*[torch.from_numpy(np.ascontiguousarray(z['in_' + k])).cuda() for k in KEYS])
    out = engine.ihwbc_solve(pb, ws, *[torch.from_numpy(np.ascontiguousarray(z[k])).cuda() for k in ('in_des', 'in_w', 'in_rf_z_max')])
    torch.cuda.synchronize()
    print(name, B, 'states:', ' '.join('%s %.2e' % (k, rel(out[k].cpu().numpy(), z['out_' + k])) for k in ('trq', 'acc', 'rf')))
