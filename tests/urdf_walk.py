"""Independent check of the URDF compiler (test infrastructure).

Everything downstream of pypnc_b200/urdf_model.py (oracle, kernels, numpy cross-checks) consumes ITS
tables, so a bug in how fixed links are merged or placed would pass every other test.  This module
walks the RAW URDF on its own -- own XML reading, own roll/pitch/yaw (Rodrigues products, not the
closed form of urdf_model.rpy_to_rot), every link kept separate (fixed links NOT merged), rotation
about an arbitrary axis by Rodrigues' formula -- and returns, for a given joint configuration:

  total mass, centre of mass and composite rotational inertia about the base origin (base axes),
  i.e. the 6 x 6 spatial inertia the floating-base block M[0:6, 0:6] of CRBA must equal
  ([lin; ang] base twist in the base frame, pinocchio_robot_system.py:152-162);
  the world placement of every link (to pin frame placements through fixed joints).

It restates what pin.buildModelsFromUrdf + crba must produce for pinocchio_robot_system.py:33-93,202
as far as that can be pinned without pinocchio.
"""
import xml.etree.ElementTree as ET

import numpy as np


def _vec(s, default):
    return np.array([float(t) for t in s.split()]) if s else np.array(default, dtype=float)


def _axis_rot(axis, angle):
    a = np.asarray(axis, dtype=float)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def _rpy(rpy):
    # URDF: fixed-axis roll (x), pitch (y), yaw (z): R = Rz(yaw) Ry(pitch) Rx(roll)
    return _axis_rot([0, 0, 1], rpy[2]) @ _axis_rot([0, 1, 0], rpy[1]) @ _axis_rot([1, 0, 0], rpy[0])


def _origin(e):
    o = e.find('origin') if e is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    return _rpy(_vec(o.get('rpy'), [0, 0, 0])), _vec(o.get('xyz'), [0, 0, 0])


def walk(urdf_path, joint_angles):
    """joint_angles: {joint name: angle}.  Returns dict(mass, com, inertia_o (3x3 about the root link origin,
    root axes), spatial (6x6 [lin; ang]), link_pose {name: (R, p)} relative to the root link)."""
    root = ET.parse(urdf_path).getroot()
    links = {l.get('name'): l for l in root.findall('link')}
    children = {}
    has_parent = set()
    for j in root.findall('joint'):
        children.setdefault(j.find('parent').get('link'), []).append(j)
        has_parent.add(j.find('child').get('link'))
    roots = [n for n in links if n not in has_parent]
    assert len(roots) == 1
    pose = {}
    mass, first, second = 0.0, np.zeros(3), np.zeros((3, 3))

    def add_link(name, R, p):
        nonlocal mass, first, second
        pose[name] = (R, p)
        inert = links[name].find('inertial')
        if inert is None:
            return
        m = float(inert.find('mass').get('value'))
        Ri, pi = _origin(inert)
        it = inert.find('inertia')
        I = np.zeros((3, 3))
        if it is not None:
            g = lambda k: float(it.get(k, 0.0))
            I = np.array([[g('ixx'), g('ixy'), g('ixz')], [g('ixy'), g('iyy'), g('iyz')], [g('ixz'), g('iyz'), g('izz')]])
        c = R @ pi + p                       # centre of mass of this link, root axes
        Ic = (R @ Ri) @ I @ (R @ Ri).T       # about its own centre of mass, root axes
        mass += m
        first += m * c
        second += Ic + m * (np.dot(c, c) * np.eye(3) - np.outer(c, c))  # parallel axes to the root origin

    stack = [(roots[0], np.eye(3), np.zeros(3))]
    while stack:
        name, R, p = stack.pop()
        add_link(name, R, p)
        for j in children.get(name, []):
            Rj, pj = _origin(j)
            Rc, pc = R @ Rj, R @ pj + p
            if j.get('type') in ('revolute', 'continuous'):
                ax = j.find('axis')
                axis = _vec(ax.get('xyz') if ax is not None else None, [1, 0, 0])
                Rc = Rc @ _axis_rot(axis, joint_angles.get(j.get('name'), 0.0))
            elif j.get('type') != 'fixed':
                raise NotImplementedError(j.get('type'))
            stack.append((j.find('child').get('link'), Rc, pc))
    com = first / mass
    cx = np.array([[0, -com[2], com[1]], [com[2], 0, -com[0]], [-com[1], com[0], 0]])
    spatial = np.zeros((6, 6))
    spatial[:3, :3] = mass * np.eye(3)
    spatial[:3, 3:] = -mass * cx
    spatial[3:, :3] = mass * cx
    spatial[3:, 3:] = second
    return dict(mass=mass, com=com, inertia_o=second, spatial=spatial, link_pose=pose)


class Walker:
    """The same walk with the XML parsed once: per-link placement and inertial data for many configurations
    (tests/first_principles.py differentiates them numerically).  links(joint_angles) returns, in a fixed link order,
    R [L,3,3], p [L,3] (link frames relative to the root link), and for the links with an <inertial>: m [L], c [L,3]
    (centre of mass, root axes, relative to the root origin) and Ic [L,3,3] (rotational inertia about that centre of
    mass, root axes); massless links carry zeros."""

    def __init__(self, urdf_path):
        root = ET.parse(urdf_path).getroot()
        self.link_el = {l.get('name'): l for l in root.findall('link')}
        self.children, has_parent = {}, set()
        for j in root.findall('joint'):
            Rj, pj = _origin(j)
            axis = None
            if j.get('type') in ('revolute', 'continuous'):
                ax = j.find('axis')
                axis = _vec(ax.get('xyz') if ax is not None else None, [1, 0, 0])
            elif j.get('type') != 'fixed':
                raise NotImplementedError(j.get('type'))
            self.children.setdefault(j.find('parent').get('link'), []).append(
                (j.get('name'), j.find('child').get('link'), Rj, pj, axis))
            has_parent.add(j.find('child').get('link'))
        roots = [n for n in self.link_el if n not in has_parent]
        assert len(roots) == 1
        self.root = roots[0]
        self.inertial = {}
        for name, l in self.link_el.items():
            inert = l.find('inertial')
            if inert is None:
                continue
            m = float(inert.find('mass').get('value'))
            Ri, pi = _origin(inert)
            it = inert.find('inertia')
            I = np.zeros((3, 3))
            if it is not None:
                g = lambda k: float(it.get(k, 0.0))
                I = np.array([[g('ixx'), g('ixy'), g('ixz')], [g('ixy'), g('iyy'), g('iyz')], [g('ixz'), g('iyz'), g('izz')]])
            self.inertial[name] = (m, Ri, pi, I)
        # fixed traversal order
        self.names, stack = [], [self.root]
        while stack:
            n = stack.pop()
            self.names.append(n)
            for (_, child, _, _, _) in self.children.get(n, []):
                stack.append(child)
        self.index = {n: i for i, n in enumerate(self.names)}

    def links(self, joint_angles):
        L = len(self.names)
        R, p = np.zeros((L, 3, 3)), np.zeros((L, 3))
        m, c, Ic = np.zeros(L), np.zeros((L, 3)), np.zeros((L, 3, 3))
        stack = [(self.root, np.eye(3), np.zeros(3))]
        while stack:
            name, Rl, pl = stack.pop()
            k = self.index[name]
            R[k], p[k] = Rl, pl
            if name in self.inertial:
                mi, Ri, pi, I = self.inertial[name]
                m[k] = mi
                c[k] = Rl @ pi + pl
                Ic[k] = (Rl @ Ri) @ I @ (Rl @ Ri).T
            for (jn, child, Rj, pj, axis) in self.children.get(name, []):
                Rc, pc = Rl @ Rj, Rl @ pj + pl
                if axis is not None:
                    Rc = Rc @ _axis_rot(axis, joint_angles.get(jn, 0.0))
                stack.append((child, Rc, pc))
        return R, p, m, c, Ic
