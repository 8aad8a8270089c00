import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypnc_b200 import robots, engine
from oracle import oracle
name = sys.argv[1]; b = int(sys.argv[2])
KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')
z = np.load(os.path.join(ROOT, 'tests/golden', name + '.npz'))
spec = robots.PROBLEMS[name]()
B = z['in_des'].shape[0]
dm = engine.Model(spec.model, 0); pb = engine.Problem(dm, spec); ws = engine.Workspace(dm, B, spec.frames)
ws.update_system(*[torch.from_numpy(np.ascontiguousarray(z['in_' + k])).cuda() for k in KEYS])
out = engine.ihwbc_solve(pb, ws, *[torch.from_numpy(np.ascontiguousarray(z[k])).cuda() for k in ('in_des', 'in_w', 'in_rf_z_max')])
torch.cuda.synchronize()
op = oracle.OracleProblem(spec)
st = {k: z['in_' + k] for k in KEYS}
ref = op.tick_batch(st, z['in_des'], z['in_w'], z['in_rf_z_max'])
act = engine.unpack_active(out['active'], pb.m)
print('gpu iters', out['iters'].cpu().numpy(), 'ref iters', ref['iters'])
print('gpu active', np.where(act[b])[0]); print('ref active', np.where(ref['active'][b])[0])
q, v = op.om.pack_state(*[st[k][b] for k in KEYS])
qp = op.assemble(q, v, z['in_des'][b], z['in_w'][b], z['in_rf_z_max'][b])
for nm, x in (('gpu', np.concatenate([out['qddot'][b].cpu().numpy(), out['rf'][b].cpu().numpy()])), ('ref', np.concatenate([ref['qddot'][b], ref['rf'][b]]))):
    s = qp['h'] - qp['G'] @ x
    print(nm, 'min slack', s.min(), 'argmin', s.argmin(), 'psi', np.minimum(s, 0).sum(), 'obj', 0.5 * x @ qp['P'] @ x + qp['q'] @ x, 'eq res', np.abs(qp['A'] @ x - qp['b']).max())
    print('   near-tight', [(i, '%.2e' % s[i]) for i in np.where(np.abs(s) < 1e-4)[0]])
print('golden trq diff gpu', np.abs(out['trq'][b].cpu().numpy() - z['out_trq'][b]).max(), 'ref', np.abs(ref['trq'][b] - z['out_trq'][b]).max())
c1 = np.trace(qp['P']); L = np.linalg.cholesky(qp['P']); c2 = np.trace(np.linalg.inv(L).T)
print('c1', c1, 'c2', c2, 'psi tol', qp['G'].shape[0] * 2.2e-16 * c1 * c2 * 100)
