"""Fixture generator: invariants of the reference URDFs computed by the independent walker
tests/urdf_walk.py (fixed links unmerged) at two joint configurations per robot.

    python tests/golden/make_urdf_invariants.py       (needs /root/reference; run in the build container)

Output: tests/golden/urdf_invariants.json -- total mass, CoM, 6x6 composite spatial inertia about the
root-link origin, and the placement of every frame the WBC problems use.  tests/test_urdf_independent.py
checks the compiled tables (pypnc_b200/models/*.json) and the oracle's CRBA against them wherever the
tests run, including boxes without the reference checkout."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get('PNC_REFERENCE', '/root/reference')

from pypnc_b200 import robots  # noqa: E402
from pypnc_b200.models.compile_models import ROBOTS as MODELS  # noqa: E402
from tests.urdf_walk import walk  # noqa: E402


def main():
    out = {}
    for name, (rel, fixed) in MODELS.items():
        m = robots.load_model(name)
        rng = np.random.default_rng(42)
        cfgs = []
        for c in range(2):
            q = np.zeros(m.na) if c == 0 else rng.uniform(-0.6, 0.6, m.na)
            angles = {jn: float(q[i]) for jn, i in m.joint_id.items()}
            w = walk(os.path.join(REF, rel), angles)
            frames = {fn: dict(R=w['link_pose'][fn][0].tolist(), p=w['link_pose'][fn][1].tolist())
                      for fn in w['link_pose'] if fn in m.frame_names}
            cfgs.append(dict(joint_pos=q.tolist(), mass=w['mass'], com=w['com'].tolist(), spatial=w['spatial'].tolist(),
                             frames=frames))
        out[name] = dict(urdf=rel, fixed_base=bool(fixed), configs=cfgs)
        print(name, 'mass %.4f' % cfgs[0]['mass'], 'frames pinned', len(cfgs[0]['frames']))
    with open(os.path.join(HERE, 'urdf_invariants.json'), 'w') as f:
        json.dump(out, f)


if __name__ == '__main__':
    main()
