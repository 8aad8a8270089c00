"""The URDF compiler (pypnc_b200/urdf_model.py -> pypnc_b200/models/*.json) against an independent walk
of the raw URDFs with every fixed link kept separate (tests/urdf_walk.py; fixture made by
tests/golden/make_urdf_invariants.py from the reference checkout).  Pins, as far as it can be pinned
without pinocchio, what pin.buildModelsFromUrdf + crba give pinocchio_robot_system.py:33-93,202:
fixed-link inertia merging, joint / frame placement through fixed joints, total mass, and the
floating-base block of the mass matrix."""
import json
import os

import numpy as np
import pytest

from pypnc_b200 import robots

HERE = os.path.dirname(os.path.abspath(__file__))
INV = json.load(open(os.path.join(HERE, 'golden', 'urdf_invariants.json')))
KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')


@pytest.mark.parametrize("name", sorted(INV))
def test_compiled_tables_match_the_unmerged_urdf(name, oracle_mod):
    m = robots.load_model(name)
    om = oracle_mod.OracleModel(m)
    inv = INV[name]
    for cfg in inv['configs']:
        jp = np.array(cfg['joint_pos'])
        if m.n_floating:
            q, v = om.pack_state(np.zeros(3), np.array([0., 0., 0., 1.]), np.zeros(3), np.zeros(3), jp, np.zeros(m.na))
        else:
            q, v = om.pack_state(None, None, None, None, jp, np.zeros(m.na))
        spatial = np.array(cfg['spatial'])
        # the wrapper's total_mass (:73-74) is the sum over every link, merged or not
        assert abs(m.total_mass - cfg['mass']) < 1e-12 * cfg['mass']
        if m.n_floating:
            # base at the origin with identity attitude: root axes = world axes
            com, _ = om.com(q, v)
            np.testing.assert_allclose(com, cfg['com'], rtol=0, atol=1e-12)
            M = om.mass_matrix(q)
            np.testing.assert_allclose(M[:6, :6], spatial, rtol=0, atol=1e-10 * np.abs(spatial).max())
        for fn, fp in cfg['frames'].items():
            T = om.link_iso(q, m.frame_id(fn))
            np.testing.assert_allclose(T[:3, :3], np.array(fp['R']), rtol=0, atol=1e-12, err_msg=fn)
            np.testing.assert_allclose(T[:3, 3], np.array(fp['p']), rtol=0, atol=1e-12, err_msg=fn)


@pytest.mark.skipif(not os.path.exists('/root/reference/robot_model'), reason="reference checkout absent")
def test_fixture_is_current():
    """The committed fixture equals a fresh walk of the reference URDFs (guards against a stale file)."""
    from pypnc_b200.models.compile_models import ROBOTS
    from tests.urdf_walk import walk
    for name, (rel, _) in ROBOTS.items():
        m = robots.load_model(name)
        cfg = INV[name]['configs'][1]
        w = walk(os.path.join('/root/reference', rel), {jn: cfg['joint_pos'][i] for jn, i in m.joint_id.items()})
        np.testing.assert_allclose(w['spatial'], np.array(cfg['spatial']), rtol=1e-13, atol=1e-13)
