"""First-principles referee for the rigid-body layer (test infrastructure).

The reference delegates M, g, c, CoM, Jacobians, Jdot*qdot and the centroidal quantities to pinocchio
(pnc/robot_system/pinocchio_robot_system.py:181-278), which cannot be installed here.  This module computes the same
quantities WITHOUT any rigid-body algorithm and without the compiled model tables: the only model it has is
tests/urdf_walk.Walker -- the placement of every link of the RAW URDF (fixed links unmerged) for a given configuration --
and everything else is mechanics applied to numerical derivatives of those placements (Kane's / d'Alembert's form of the
equations of motion, valid for the body-frame base twist pinocchio uses as generalized velocity):

  motion            q(t) = q (+) t v : base T_b Exp(t [v_b; w_b]) (constant body twist, i.e. vdot = 0), joints q_j + t qd_j
  partial velocity  Jv_l[:, i], Jw_l[:, i] = d/ds of link l's CoM / attitude along the basis direction e_i
  M[i, j]           sum_l  m_l Jv_l[:, i] . Jv_l[:, j] + Jw_l[:, i] . I_l Jw_l[:, j]
  (c + g)[i]        sum_l  m_l (a_l + 9.81 z) . Jv_l[:, i] + (I_l alpha_l + w_l x I_l w_l) . Jw_l[:, i]     (vdot = 0)
  g[i]              sum_l  m_l 9.81 z . Jv_l[:, i]
  CoM, J_com, d/dt J_com, J_com_dot qd = CoM acceleration at vdot = 0
  frames            iso; [w; v] ; Jacobian columns [Jw; Jv] (LOCAL_WORLD_ALIGNED, [angular; linear] as :259-261);
                    classical acceleration at vdot = 0  (= Jdot qdot, :263-278)
  centroidal        hg = [sum I_l w_l + (c_l - c) x m_l v_l ; sum m_l v_l], Ag column by column the same with the partial
                    velocities, Ig = blockdiag(rotational inertia about the CoM, m I3)   (block order of :181-194)

Derivatives are central differences of 4th order (step H); the fixture generator evaluates every state at H and at 2 H
and stores the difference as the referee's own error estimate.
"""
import numpy as np

from tests.urdf_walk import Walker, _axis_rot

GZ = 9.81
H = 4e-3


def quat_to_rot(q):  # scalar-last, normalised
    x, y, z, w = np.asarray(q, dtype=float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _hat(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


def _vee_antisym(A):
    """axis vector of the antisymmetric part of a batch of 3x3 matrices [..., 3, 3]"""
    S = 0.5 * (A - np.swapaxes(A, -1, -2))
    return np.stack([S[..., 2, 1], S[..., 0, 2], S[..., 1, 0]], axis=-1)


def se3_exp(v, w, t):
    """Exp(t [v; w]): rotation and translation of the constant-body-twist motion"""
    th = np.linalg.norm(w) * t
    if abs(th) < 1e-12:
        return np.eye(3), np.asarray(v, dtype=float) * t
    a = np.asarray(w, dtype=float) / np.linalg.norm(w)
    R = _axis_rot(a, th)
    K = _hat(a)
    wn = np.linalg.norm(w)
    V = np.eye(3) * t + (1 - np.cos(th)) / wn * K + (th - np.sin(th)) / wn * (K @ K)
    return R, V @ np.asarray(v, dtype=float)


class Referee:
    def __init__(self, urdf_path, joint_names, floating, frame_names=()):
        """joint_names: actuated joints in the order of the velocity vector (after the 6 base entries when floating)"""
        self.wk = Walker(urdf_path)
        self.joint_names = list(joint_names)
        self.floating = bool(floating)
        self.na = len(self.joint_names)
        self.nv = self.na + (6 if floating else 0)
        self.frames = [f for f in frame_names if f in self.wk.index]
        self.fidx = np.array([self.wk.index[f] for f in self.frames], dtype=int)
        self.welded_mass = 0.0

    # -- configuration = (R_b, p_b, jp); tangent = nv-vector [v_b (base frame); w_b (base frame); qd] ----------------
    def plus(self, cfg, u, s):
        Rb, pb, jp = cfg
        if self.floating:
            dR, dp = se3_exp(u[0:3], u[3:6], s)
            return Rb @ dR, pb + Rb @ dp, jp + s * u[6:]
        return Rb, pb, jp + s * u

    def world(self, cfg):
        """per link, in the world: R, p, m, c (CoM), Iw (about the CoM)"""
        Rb, pb, jp = cfg
        R, p, m, c, Ic = self.wk.links({jn: float(jp[i]) for i, jn in enumerate(self.joint_names)})
        if not self.floating:
            # fixed base: the root link is welded to the universe.  pinocchio keeps its inertia on joint 0, outside every
            # algorithm (CoM included); only the wrapper's total_mass (:73-74) counts it.
            m = m.copy()
            self.welded_mass = float(m[self.wk.index[self.wk.root]])
            m[self.wk.index[self.wk.root]] = 0.0
            Ic = Ic.copy()
            Ic[self.wk.index[self.wk.root]] = 0.0
        Rw = np.einsum('ij,ljk->lik', Rb, R)
        pw = p @ Rb.T + pb
        cw = c @ Rb.T + pb
        Iw = np.einsum('ij,ljk,mk->lim', Rb, Ic, Rb)
        return Rw, pw, m, cw, Iw

    def derivs(self, cfg, u, h=H, second=False):
        """d/ds (and d2/ds2) at s = 0 of (R, p, c) of every link along cfg (+) s u; 4th-order central stencils"""
        f = {k: self.world(self.plus(cfg, u, k * h)) for k in (-2, -1, 1, 2)}
        out = []
        for idx in (0, 1, 3):  # R, p, c
            d1 = (-f[2][idx] + 8 * f[1][idx] - 8 * f[-1][idx] + f[-2][idx]) / (12 * h)
            out.append(d1)
        if second:
            f0 = self.world(cfg)
            for idx in (0, 1, 3):
                d2 = (-f[2][idx] + 16 * f[1][idx] - 30 * f0[idx] + 16 * f[-1][idx] - f[-2][idx]) / (12 * h * h)
                out.append(d2)
        return out

    def partials(self, cfg, h=H):
        """Jv [L,3,nv] (CoM), Jp [L,3,nv] (link origin), Jw [L,3,nv]"""
        R0 = self.world(cfg)[0]
        L = R0.shape[0]
        Jv, Jp, Jw = np.zeros((L, 3, self.nv)), np.zeros((L, 3, self.nv)), np.zeros((L, 3, self.nv))
        for i in range(self.nv):
            e = np.zeros(self.nv)
            e[i] = 1.0
            dR, dp, dc = self.derivs(cfg, e, h)
            Jv[:, :, i], Jp[:, :, i] = dc, dp
            Jw[:, :, i] = _vee_antisym(np.einsum('lij,lkj->lik', dR, R0))
        return Jv, Jp, Jw

    def com_jacobian(self, cfg, h=H):
        m = self.world(cfg)[2]
        Jv = self.partials(cfg, h)[0]
        return np.einsum('l,lij->ij', m, Jv) / m.sum()

    def evaluate(self, base_pos, base_quat, base_lin_vel, base_ang_vel, joint_pos, joint_vel, h=H, with_jcomdot=True):
        """Inputs as RobotSystem.update_system takes them (base twist in WORLD axes, :149-162)."""
        jp, jv = np.asarray(joint_pos, dtype=float), np.asarray(joint_vel, dtype=float)
        if self.floating:
            Rb = quat_to_rot(base_quat)
            cfg = (Rb, np.asarray(base_pos, dtype=float), jp)
            v = np.concatenate([Rb.T @ np.asarray(base_lin_vel), Rb.T @ np.asarray(base_ang_vel), jv])
        else:
            cfg = (np.eye(3), np.zeros(3), jp)
            v = jv.copy()
        R0, p0, m, c0, Iw = self.world(cfg)
        mt = m.sum()
        Jv, Jp, Jw = self.partials(cfg, h)
        dR, dp, dc, ddR, ddp, ddc = self.derivs(cfg, v, h, second=True)
        w = _vee_antisym(np.einsum('lij,lkj->lik', dR, R0))
        al = _vee_antisym(np.einsum('lij,lkj->lik', ddR, R0))
        out = dict(qdot=v)
        out['M'] = np.einsum('l,lki,lkj->ij', m, Jv, Jv) + np.einsum('lki,lkm,lmj->ij', Jw, Iw, Jw)
        zhat = np.array([0.0, 0.0, GZ])
        out['grav'] = np.einsum('l,k,lki->i', m, zhat, Jv)
        Iw_w = np.einsum('lij,lj->li', Iw, w)
        torque = np.einsum('lij,lj->li', Iw, al) + np.cross(w, Iw_w)
        nle = np.einsum('l,lk,lki->i', m, ddc + zhat, Jv) + np.einsum('lk,lki->i', torque, Jw)
        out['cori'] = nle - out['grav']
        com = (m[:, None] * c0).sum(0) / mt
        out['com'] = com
        out['jcom'] = np.einsum('l,lij->ij', m, Jv) / mt
        out['vcom'] = (m[:, None] * dc).sum(0) / mt
        # the reference's J_com_dot is dAg_lin / total_mass with the wrapper's total mass (:229, :73-74), welded root included
        total_mass = mt + self.welded_mass
        out['total_mass'] = total_mass
        out['jcom_jdqd'] = (m[:, None] * ddc).sum(0) / total_mass
        if with_jcomdot:
            J = {k: self.com_jacobian(self.plus(cfg, v, k * h), h) for k in (-2, -1, 1, 2)}
            out['jcomdot'] = (-J[2] + 8 * J[1] - 8 * J[-1] + J[-2]) / (12 * h) * (mt / total_mass)
        # centroidal quantities, rows [angular; linear]
        d = c0 - com
        out['hg'] = np.concatenate([(Iw_w + np.cross(d, m[:, None] * dc)).sum(0), (m[:, None] * dc).sum(0)])
        Ag = np.zeros((6, self.nv))
        Ag[0:3] = np.einsum('lkm,lmi->ki', Iw, Jw) + np.einsum('lki->ki', np.cross(d[:, :, None], m[:, None, None] * Jv, axis=1))
        Ag[3:6] = np.einsum('l,lki->ki', m, Jv)
        out['Ag'] = Ag
        Ig = np.zeros((6, 6))
        Ig[0:3, 0:3] = Iw.sum(0) + np.einsum('l,lij->ij', m, (d * d).sum(1)[:, None, None] * np.eye(3) - np.einsum('li,lj->lij', d, d))
        Ig[3:6, 3:6] = mt * np.eye(3)
        out['Ig'] = Ig
        # frames (link frames), [angular; linear]
        F = self.fidx
        nf = len(F)
        iso = np.zeros((nf, 4, 4))
        iso[:, 3, 3] = 1.0
        iso[:, :3, :3], iso[:, :3, 3] = R0[F], p0[F]
        out['frame_iso'] = iso
        out['frame_vel'] = np.concatenate([w[F], dp[F]], axis=1)
        out['frame_jac'] = np.concatenate([Jw[F], Jp[F]], axis=1)
        out['frame_jdqd'] = np.concatenate([al[F], ddp[F]], axis=1)
        return out
