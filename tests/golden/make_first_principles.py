"""Fixture generator: the rigid-body quantities of random states of every robot, computed by the first-principles referee
tests/first_principles.py from the RAW reference URDFs (fixed links unmerged, no rigid-body algorithm, no compiled table).

    python tests/golden/make_first_principles.py       (needs /root/reference; run in the build container)

Output: tests/golden/first_principles.npz -- per robot and state: the inputs of update_system and M, g, c, CoM, v_CoM,
J_com, d/dt J_com, J_com_dot qd, Ag, hg, Ig, and placement / velocity / Jacobian / Jdot*qdot of every link frame the model
names, plus `fd_err`, the largest relative difference between the evaluation at step H and at step 2 H (the referee's own
finite-difference error).  tests/test_first_principles.py holds the oracle (CPU) and k_update_system (GPU) against it
wherever the tests run, including boxes without the reference checkout."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get('PNC_REFERENCE', '/root/reference')

from pypnc_b200 import robots  # noqa: E402
from pypnc_b200.models.compile_models import ROBOTS as MODELS  # noqa: E402
from tests.first_principles import Referee, H  # noqa: E402

NAMES = ('atlas', 'laikago', 'draco3', 'draco3_lb', 'valkyrie', 'three_link_manipulator', 'two_link_manipulator')
STATES = 3
QUANTS = ('qdot', 'M', 'grav', 'cori', 'com', 'vcom', 'jcom', 'jcomdot', 'jcom_jdqd', 'Ag', 'hg', 'Ig', 'frame_iso',
          'frame_vel', 'frame_jac', 'frame_jdqd')


def make_referee(name):
    rel, fixed = MODELS[name]
    m = robots.load_model(name)
    jn = sorted(m.joint_id, key=lambda k: m.joint_id[k])
    return Referee(os.path.join(REF, rel), jn, not fixed, m.frame_names), m


def random_state(m, rng, k):
    """state k of a robot: k = 0 is at rest in the nominal pose, the others are random (joint velocities ~ 1 rad/s)"""
    if k == 0:
        return dict(base_pos=np.zeros(3), base_quat=np.array([0., 0., 0., 1.]), base_lin_vel=np.zeros(3),
                    base_ang_vel=np.zeros(3), joint_pos=np.zeros(m.na), joint_vel=np.zeros(m.na))
    bq = rng.normal(size=4)
    return dict(base_pos=rng.uniform(-1, 1, 3), base_quat=bq / np.linalg.norm(bq), base_lin_vel=rng.normal(0, 0.5, 3),
                base_ang_vel=rng.normal(0, 0.5, 3), joint_pos=rng.uniform(-0.7, 0.7, m.na),
                joint_vel=rng.normal(0, 1.0, m.na))


def main():
    out = {}
    for name in NAMES:
        ref, m = make_referee(name)
        rng = np.random.default_rng(2026)
        out[name + '/frames'] = np.array(ref.frames)
        worst = 0.0
        for k in range(STATES):
            st = random_state(m, rng, k)
            a = ref.evaluate(**st, h=H)
            b = ref.evaluate(**st, h=2 * H)
            err = 0.0
            for qn in QUANTS:
                scale = max(np.abs(a[qn]).max(), 1.0)  # quantities that vanish (state 0 is at rest): absolute
                err = max(err, float(np.abs(a[qn] - b[qn]).max() / scale))
                out['%s/%d/%s' % (name, k, qn)] = a[qn]
            for kk, v in st.items():
                out['%s/%d/in_%s' % (name, k, kk)] = v
            out['%s/%d/fd_err' % (name, k)] = np.array(err)
            worst = max(worst, err)
        print('%-24s %d states, %d link frames, referee error estimate (H vs 2H) %.1e' % (name, STATES, len(ref.frames), worst))
    np.savez_compressed(os.path.join(HERE, 'first_principles.npz'), **out)


if __name__ == '__main__':
    main()
