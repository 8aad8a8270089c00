"""URDF -> flat rigid-body tables ("model compiler"), host side, numpy only.

Replaces what the reference gets from ``pin.buildModelsFromUrdf`` inside
``PinocchioRobotSystem._config_robot``
(reference: pnc/robot_system/pinocchio_robot_system.py:33-93).  The semantics
restated here are the ones that wrapper observes:

* depth-first traversal from the URDF root link; the children of a link are
  visited in alphabetical order of the *connecting joint's name*;
* ``revolute``/``continuous`` joints become 1-dof bodies, ``fixed`` joints merge
  the child link's inertia into the nearest movable ancestor and leave a frame;
* a floating base adds a free-flyer ``root_joint`` (q: xyz + quat xyzw, v:
  [lin; ang] in the base frame);
* ``joint_id[name] = movable-joint index`` (idx_v - 6 on a floating base),
* limits come from ``<limit lower upper velocity effort>``.

The result (:class:`RobotModel`) is a set of flat numpy arrays that is uploaded
once to the GPU (pypnc_b200/csrc) and is also what the CPU oracle consumes.
"""
from __future__ import annotations

import json
import xml.etree.ElementTree as ET
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np

JOINT_FREEFLYER = 0
JOINT_REVOLUTE = 1

FRAME_JOINT = 0
FRAME_FIXED_JOINT = 1
FRAME_BODY = 2


def _floats(s, n, default):
    if s is None:
        return np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    assert v.shape[0] == n, s
    return v


def rpy_to_rot(rpy):
    """URDF fixed-axis roll/pitch/yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _origin(elem):
    o = elem.find('origin') if elem is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    return rpy_to_rot(_floats(o.get('rpy'), 3, [0, 0, 0])), _floats(o.get('xyz'), 3, [0, 0, 0])


def _link_inertia(link):
    """(mass, com, I about com) of a URDF link, expressed in the link frame."""
    inertial = link.find('inertial')
    if inertial is None:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    R, c = _origin(inertial)
    mass = float(inertial.find('mass').get('value'))
    it = inertial.find('inertia')
    g = lambda k: float(it.get(k, 0.0))
    I = np.array([[g('ixx'), g('ixy'), g('ixz')], [g('ixy'), g('iyy'), g('iyz')],
                  [g('ixz'), g('iyz'), g('izz')]])
    return mass, c, R @ I @ R.T


class _SpatialInertia:
    """mass, first moment and rotational inertia about the frame origin."""

    def __init__(self):
        self.m = 0.0
        self.h = np.zeros(3)  # m * com
        self.Io = np.zeros((3, 3))  # about the origin of the owning joint frame

    def add(self, mass, com, Ic, R, p):
        """Add a body given in a child frame placed at (R, p) in this frame."""
        c = p + R @ com
        self.m += mass
        self.h += mass * c
        self.Io += R @ Ic @ R.T + mass * (np.dot(c, c) * np.eye(3) - np.outer(c, c))

    def com(self):
        return self.h / self.m if self.m > 0 else np.zeros(3)

    def Ic(self):
        c = self.com()
        return self.Io - self.m * (np.dot(c, c) * np.eye(3) - np.outer(c, c))


@dataclass
class RobotModel:
    name: str
    fixed_base: bool
    nb: int  # movable bodies (free-flyer root counts as one)
    nq: int
    nv: int
    na: int
    n_floating: int
    total_mass: float  # sum of every link mass, as the wrapper computes it (:73-74)
    mass_dyn: float  # mass of the moving bodies (what centre-of-mass algorithms see)
    parent: np.ndarray  # [nb] int32, -1 = universe
    jtype: np.ndarray  # [nb] int32
    idx_v: np.ndarray  # [nb] int32
    idx_q: np.ndarray  # [nb] int32
    subtree: np.ndarray  # [nb] int32: bodies i..i+subtree[i]-1 are the subtree of i
    depth: np.ndarray  # [nb] int32
    place_R: np.ndarray  # [nb,3,3] joint placement in the parent joint frame
    place_p: np.ndarray  # [nb,3]
    axis: np.ndarray  # [nb,3]
    mass: np.ndarray  # [nb]
    com: np.ndarray  # [nb,3] in the joint frame
    inertia: np.ndarray  # [nb,3,3] about com, joint-frame axes
    joint_names: list  # movable 1-dof joints, in idx_v order (the reference's joint_id keys)
    joint_pos_limit: np.ndarray  # [na,2]
    joint_vel_limit: np.ndarray  # [na,2]
    joint_trq_limit: np.ndarray  # [na,2]
    frame_names: list
    frame_type: np.ndarray  # [nf]
    frame_parent: np.ndarray  # [nf] body index (-1 = universe)
    frame_R: np.ndarray  # [nf,3,3] placement in the parent joint frame
    frame_p: np.ndarray  # [nf,3]
    universe_mass: float = 0.0
    joint_id: OrderedDict = field(default_factory=OrderedDict)
    link_id: OrderedDict = field(default_factory=OrderedDict)

    # ---- lookups mirroring pinocchio_robot_system.py:95-112 -------------
    def frame_id(self, name):
        try:
            return self.frame_names.index(name)
        except ValueError:
            raise KeyError("no frame named %r" % (name,))

    def get_q_idx(self, joint_name):
        return int(self.idx_q[self._body_of(joint_name)])

    def get_q_dot_idx(self, joint_name):
        return int(self.idx_v[self._body_of(joint_name)])

    def get_joint_idx(self, joint_name):
        return self.get_q_dot_idx(joint_name) - self.n_floating

    def _body_of(self, joint_name):
        off = 0 if self.fixed_base else 1
        return self.joint_id[joint_name] + off

    # ---- (de)serialisation: compiled tables travel, URDFs need not ------
    def to_dict(self):
        d = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                d[k] = {'dtype': str(v.dtype), 'shape': list(v.shape), 'data': v.ravel().tolist()}
            elif isinstance(v, OrderedDict):
                d[k] = {'odict': list(v.items())}
            else:
                d[k] = v
        return d

    @staticmethod
    def from_dict(d):
        kw = {}
        for k, v in d.items():
            if isinstance(v, dict) and 'dtype' in v:
                kw[k] = np.array(v['data'], dtype=v['dtype']).reshape(v['shape'])
            elif isinstance(v, dict) and 'odict' in v:
                kw[k] = OrderedDict((a, b) for a, b in v['odict'])
            else:
                kw[k] = v
        return RobotModel(**kw)

    def save(self, path):
        with open(path, 'w') as f:
            json.dump(self.to_dict(), f)

    @staticmethod
    def load(path):
        with open(path) as f:
            return RobotModel.from_dict(json.load(f))


def build_model_from_urdf(urdf_file, fixed_base, name=None):
    """Compile a URDF into :class:`RobotModel` (see module docstring)."""
    root = ET.parse(urdf_file).getroot()
    return build_model_from_xml(root, fixed_base, name or root.get('name', 'robot'))


def build_model_from_string(urdf_text, fixed_base, name=None):
    root = ET.fromstring(urdf_text)
    return build_model_from_xml(root, fixed_base, name or root.get('name', 'robot'))


def build_model_from_xml(root, fixed_base, name):
    links = OrderedDict((l.get('name'), l) for l in root.findall('link'))
    joints = {}
    children = {ln: [] for ln in links}
    has_parent = set()
    for j in root.findall('joint'):
        jn = j.get('name')
        jt = j.get('type')
        if jt == 'continuous':
            # pinocchio maps a continuous joint to JointModelRUB* (nq = 2: cos/sin, nv = 1); the tables
            # and kernels here assume nq = 1 per joint, so q indexing would silently diverge
            raise NotImplementedError("joint %s: 'continuous' joints (pinocchio nq = 2) are not supported; none of "
                                      "the reference's hot-path URDFs has one" % jn)
        if jt not in ('revolute', 'fixed'):
            raise NotImplementedError("joint %s: type %s is not on the WBC hot path" % (jn, jt))
        joints[jn] = j
        p, c = j.find('parent').get('link'), j.find('child').get('link')
        children[p].append(jn)
        has_parent.add(c)
    roots = [ln for ln in links if ln not in has_parent]
    assert len(roots) == 1, "URDF must have exactly one root link, got %s" % roots
    root_link = roots[0]
    for ln in children:
        children[ln].sort()  # urdfdom keeps joints in a std::map -> alphabetical

    bodies = []  # dicts
    frames = []  # (name, type, parent_body, R, p)
    joint_id = OrderedDict()
    link_id = OrderedDict()
    universe = _SpatialInertia()

    def new_body(parent, jtype, R, p, axis, jname, lim):
        bodies.append(dict(parent=parent, jtype=jtype, R=R, p=p, axis=axis, Y=_SpatialInertia(),
                           name=jname, lim=lim, depth=0 if parent < 0 else bodies[parent]['depth'] + 1))
        return len(bodies) - 1

    frames.append(('universe', FRAME_FIXED_JOINT, -1, np.eye(3), np.zeros(3)))
    if fixed_base:
        cur_body = -1
    else:
        cur_body = new_body(-1, JOINT_FREEFLYER, np.eye(3), np.zeros(3), np.zeros(3), 'root_joint', None)
        frames.append(('root_joint', FRAME_JOINT, cur_body, np.eye(3), np.zeros(3)))

    def visit(link_name, body, R, p):
        """link frame sits at (R, p) in the joint frame of `body` (or universe)."""
        m, c, Ic = _link_inertia(links[link_name])
        if body >= 0:
            bodies[body]['Y'].add(m, c, Ic, R, p)
        else:
            universe.add(m, c, Ic, R, p)
        frames.append((link_name, FRAME_BODY, body, R, p))
        link_id[link_name] = len(link_id)
        for jn in children[link_name]:
            j = joints[jn]
            Rj, pj = _origin(j)
            Rc, pc = R @ Rj, p + R @ pj  # child joint frame in current body's joint frame
            child = j.find('child').get('link')
            if j.get('type') == 'fixed':
                frames.append((jn, FRAME_FIXED_JOINT, body, Rc, pc))
                visit(child, body, Rc, pc)
            else:
                ax = j.find('axis')
                axis = _floats(ax.get('xyz') if ax is not None else None, 3, [1, 0, 0])
                axis = axis / np.linalg.norm(axis)
                lim = j.find('limit')
                lo = float(lim.get('lower', 0.0)) if lim is not None else 0.0
                hi = float(lim.get('upper', 0.0)) if lim is not None else 0.0
                if j.get('type') == 'continuous':
                    lo, hi = -np.inf, np.inf
                vel = float(lim.get('velocity', 0.0)) if lim is not None else 0.0
                eff = float(lim.get('effort', 0.0)) if lim is not None else 0.0
                b = new_body(body, JOINT_REVOLUTE, Rc, pc, axis, jn, (lo, hi, vel, eff))
                joint_id[jn] = len(joint_id)
                frames.append((jn, FRAME_JOINT, b, np.eye(3), np.zeros(3)))
                visit(child, b, np.eye(3), np.zeros(3))

    visit(root_link, cur_body, np.eye(3), np.zeros(3))

    nb = len(bodies)
    nfl = 0 if fixed_base else 6
    na = len(joint_id)
    nv = nfl + na
    nq = (7 if not fixed_base else 0) + na
    parent = np.array([b['parent'] for b in bodies], dtype=np.int32)
    subtree = np.ones(nb, dtype=np.int32)
    for i in range(nb - 1, 0, -1):
        if parent[i] >= 0:
            subtree[parent[i]] += subtree[i]
    idx_v = np.zeros(nb, dtype=np.int32)
    idx_q = np.zeros(nb, dtype=np.int32)
    k = 0
    for i, b in enumerate(bodies):
        if b['jtype'] == JOINT_FREEFLYER:
            idx_v[i], idx_q[i] = 0, 0
        else:
            idx_v[i], idx_q[i] = nfl + k, (7 if not fixed_base else 0) + k
            k += 1
    lims = np.array([b['lim'] for b in bodies if b['lim'] is not None], dtype=np.float64).reshape(na, 4)
    model = RobotModel(
        name=name, fixed_base=bool(fixed_base), nb=nb, nq=nq, nv=nv, na=na, n_floating=nfl,
        total_mass=float(universe.m + sum(b['Y'].m for b in bodies)),
        mass_dyn=float(sum(b['Y'].m for b in bodies)),
        universe_mass=float(universe.m),
        parent=parent,
        jtype=np.array([b['jtype'] for b in bodies], dtype=np.int32),
        idx_v=idx_v, idx_q=idx_q, subtree=subtree,
        depth=np.array([b['depth'] for b in bodies], dtype=np.int32),
        place_R=np.array([b['R'] for b in bodies]).reshape(nb, 3, 3),
        place_p=np.array([b['p'] for b in bodies]).reshape(nb, 3),
        axis=np.array([b['axis'] for b in bodies]).reshape(nb, 3),
        mass=np.array([b['Y'].m for b in bodies]),
        com=np.array([b['Y'].com() for b in bodies]).reshape(nb, 3),
        inertia=np.array([b['Y'].Ic() for b in bodies]).reshape(nb, 3, 3),
        joint_names=list(joint_id.keys()),
        joint_pos_limit=np.stack([lims[:, 0], lims[:, 1]], axis=1),
        joint_vel_limit=np.stack([-lims[:, 2], lims[:, 2]], axis=1),
        joint_trq_limit=np.stack([-lims[:, 3], lims[:, 3]], axis=1),
        frame_names=[f[0] for f in frames],
        frame_type=np.array([f[1] for f in frames], dtype=np.int32),
        frame_parent=np.array([f[2] for f in frames], dtype=np.int32),
        frame_R=np.array([f[3] for f in frames]).reshape(len(frames), 3, 3),
        frame_p=np.array([f[4] for f in frames]).reshape(len(frames), 3),
        joint_id=joint_id, link_id=link_id)
    # DFS numbering is what the kernels rely on: subtree(i) is a contiguous index range
    for i in range(nb):
        assert parent[i] < i
    return model
