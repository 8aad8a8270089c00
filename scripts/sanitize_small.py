"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the path once,
on a few dozen states per robot, plus the managers and the dense QP."""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypnc_b200 import robots, engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
for name in (sys.argv[2:] or ['atlas', 'laikago', 'draco3', 'draco3_lb', 'valkyrie']):
    spec = robots.PROBLEMS[name]()
    st = robots.synth_states(name, spec, B, seed=2)
    dm = engine.Model(spec.model)
    pb = engine.Problem(dm, spec)
    ws = engine.Workspace(dm, B, spec.frames)
    d = {k: torch.from_numpy(v).cuda() for k, v in st.items()}
    ws.update_system(d['base_pos'], d['base_quat'], d['base_lin_vel'], d['base_ang_vel'], d['joint_pos'], d['joint_vel'], b_cent=True)
    des = torch.zeros((B, spec.ndes), dtype=torch.float64, device='cuda')
    # desireds = a plausible constant: unit quaternions where needed
    for t in spec.tasks:
        if t['type'] == 3:
            des[:, t['des_off'] + 3] = 1.0
    w = torch.ones((B, len(spec.tasks)), dtype=torch.float64, device='cuda')
    rfz = torch.full((B, max(1, len(spec.contacts))), 1000.0, dtype=torch.float64, device='cuda')
    out = engine.ihwbc_solve(pb, ws, des, w, rfz)
    torch.cuda.synchronize()
    print(name, 'status', np.bincount(out['status'].cpu().numpy(), minlength=6).tolist())
rng = np.random.default_rng(0)
n, p, m = 12, 6, 36
Mx = rng.normal(size=(32, n, n))
G = torch.from_numpy(Mx @ np.transpose(Mx, (0, 2, 1)) + 1e-3 * np.eye(n)).cuda()
x, f, stt, it = engine.quadprog_batch(G, torch.from_numpy(rng.normal(size=(32, n))).cuda(), torch.from_numpy(rng.normal(size=(32, n, p))).cuda(),
                                      torch.from_numpy(rng.normal(size=(32, p))).cuda(), torch.from_numpy(rng.normal(size=(32, n, m))).cuda(),
                                      torch.from_numpy(rng.normal(size=(32, m)) + 2.0).cuda())
torch.cuda.synchronize()
print('quadprog status', np.bincount(stt.cpu().numpy(), minlength=5).tolist())
