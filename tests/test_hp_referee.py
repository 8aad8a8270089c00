"""The fp64 oracle against the 50-digit referee (tests/hp_referee.py): an independent restatement of
ihwbc.py:124-356 plus an exact KKT solve of the oracle's final active set.  Pins the oracle's IHWBC
assembly, internal-constraint projection, QP solution and torque recovery to ~1e-10 on every robot,
and shows that the oracle's final active set is an optimal one (multipliers of one sign)."""
import numpy as np
import pytest

from pypnc_b200 import robots
from tests import hp_referee


@pytest.mark.parametrize("name", ['atlas', 'laikago', 'draco3', 'draco3_lb', 'valkyrie'])
def test_oracle_matches_exact_solution(name, oracle_mod):
    spec = robots.PROBLEMS[name]()
    B = 6
    st = robots.synth_states(name, spec, B, seed=4)
    op = oracle_mod.OracleProblem(spec)
    des, w, rfz = robots.synth_desireds(name, spec, st, op.task_actuals(st), seed=4)
    ref = op.tick_batch(st, des, w, rfz)
    checked = 0
    for b in range(B):
        if ref['status'][b] != 0 or ref['psi_rel'][b] > 1e-11:
            continue
        hp = hp_referee.solve_state(op, spec, {k: st[k][b] for k in hp_referee.KEYS}, des[b], w[b], rfz[b], ref['active'][b])
        for k in ('trq', 'acc', 'rf'):
            den = max(float(np.abs(hp[k]).max()), 1e-6 * float(np.abs(ref[k]).max()))
            assert np.abs(ref[k][b] - hp[k]).max() / den < 1e-9, (name, b, k)
        lam = hp['multipliers']
        if lam.size:  # KKT: the active inequalities' multipliers share one sign (zero at degenerate vertices)
            assert (lam <= 1e-9 * np.abs(lam).max()).all() or (lam >= -1e-9 * np.abs(lam).max()).all(), (name, b, lam)
        checked += 1
    assert checked >= 3
