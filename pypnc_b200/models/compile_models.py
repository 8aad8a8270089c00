"""Compile the reference's URDFs into the flat model tables shipped in this directory.

    python -m pypnc_b200.models.compile_models [/root/reference]

The tables (``<robot>.json``) are what travels to the GPU box; the reference
checkout (URDF text, meshes) does not.  Source URDFs: robot_model/<robot>/*.urdf of
junhyeokahn/PyPnC.
"""
import os
import sys

from pypnc_b200.urdf_model import build_model_from_urdf

ROBOTS = {
    'atlas': ('robot_model/atlas/atlas.urdf', False),
    'laikago': ('robot_model/laikago/laikago_toes.urdf', False),
    'draco3': ('robot_model/draco3/draco3_pin_model.urdf', False),
    'draco3_lb': ('robot_model/draco3/draco3_lb.urdf', False),
    'valkyrie': ('robot_model/valkyrie/valkyrie.urdf', False),
    'three_link_manipulator': ('robot_model/manipulator/three_link_manipulator.urdf', True),
    'two_link_manipulator': ('robot_model/manipulator/two_link_manipulator.urdf', True),
}


def main(ref='/root/reference'):
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (rel, fixed) in ROBOTS.items():
        m = build_model_from_urdf(os.path.join(ref, rel), fixed, name=name)
        m.save(os.path.join(here, name + '.json'))
        print('%-24s nb=%2d nq=%2d nv=%2d na=%2d mass=%.4f frames=%d' %
              (name, m.nb, m.nq, m.nv, m.na, m.total_mass, len(m.frame_names)))


if __name__ == '__main__':
    main(*sys.argv[1:])
