"""Independent numpy 'textbook' rigid-body model used to cross-check the C oracle.

World-frame geometric Jacobians + finite differences in time; no recursion shared with
oracle/pnc_oracle.c (body-frame RNEA/CRBA) or the CUDA kernels (world-frame prefix sums)."""
import numpy as np
from scipy.spatial.transform import Rotation as Rot


def fk(model, q):
    """world (R, p) of every movable body's joint frame"""
    Rw, pw = [], []
    for i in range(model.nb):
        if model.jtype[i] == 0:
            Rj, pj = Rot.from_quat(q[3:7]).as_matrix(), q[0:3]
        else:
            iq = model.idx_q[i]
            Rj, pj = Rot.from_rotvec(model.axis[i] * q[iq]).as_matrix(), np.zeros(3)
        Rl = model.place_R[i] @ Rj
        pl = model.place_p[i] + model.place_R[i] @ pj
        par = model.parent[i]
        if par < 0:
            Rw.append(Rl); pw.append(pl)
        else:
            Rw.append(Rw[par] @ Rl); pw.append(pw[par] + Rw[par] @ pl)
    return Rw, pw


def integrate(model, q, dv):
    """q (+) dv with the base twist expressed in the base frame (pinocchio convention)"""
    q2 = q.copy()
    if model.n_floating:
        R = Rot.from_quat(q[3:7])
        q2[0:3] = q[0:3] + R.apply(dv[0:3])
        q2[3:7] = (R * Rot.from_rotvec(dv[3:6])).as_quat()
        q2[7:] = q[7:] + dv[6:]
    else:
        q2 = q + dv
    return q2


def point_jacobian(model, q, body, p_local):
    """6 x nv [linear; angular] world-axes Jacobian of a point fixed in `body`"""
    Rw, pw = fk(model, q)
    nv = model.nv
    J = np.zeros((6, nv))
    pt = pw[body] + Rw[body] @ p_local
    j = body
    while j >= 0:
        iv = model.idx_v[j]
        if model.jtype[j] == 0:
            R = Rw[j]
            J[0:3, iv:iv + 3] = R
            for k in range(3):
                J[0:3, iv + 3 + k] = np.cross(R[:, k], pt - pw[j])
                J[3:6, iv + 3 + k] = R[:, k]
        else:
            a = Rw[j] @ model.axis[j]
            J[0:3, iv] = np.cross(a, pt - pw[j])
            J[3:6, iv] = a
        j = model.parent[j]
    return J, pt, Rw[body]


def mass_matrix(model, q):
    M = np.zeros((model.nv, model.nv))
    for i in range(model.nb):
        J, _, R = point_jacobian(model, q, i, model.com[i])
        Iw = R @ model.inertia[i] @ R.T
        M += model.mass[i] * J[0:3].T @ J[0:3] + J[3:6].T @ Iw @ J[3:6]
    return M


def nle(model, q, v, gravity=True, eps=1e-6):
    """C(q,v) v + g(q) from sum_i J_i^T [m a_i; I wdot + w x I w] with Jdot v by central differences"""
    tau = np.zeros(model.nv)
    qp, qm = integrate(model, q, eps * v), integrate(model, q, -eps * v)
    g = np.array([0, 0, 9.81]) if gravity else np.zeros(3)
    for i in range(model.nb):
        J, _, R = point_jacobian(model, q, i, model.com[i])
        Jp, _, _ = point_jacobian(model, qp, i, model.com[i])
        Jm, _, _ = point_jacobian(model, qm, i, model.com[i])
        Jdv = (Jp - Jm) @ v / (2 * eps)
        Iw = R @ model.inertia[i] @ R.T
        w = J[3:6] @ v
        tau += J[0:3].T @ (model.mass[i] * (Jdv[0:3] + g)) + J[3:6].T @ (Iw @ Jdv[3:6] + np.cross(w, Iw @ w))
    return tau


def com(model, q):
    Rw, pw = fk(model, q)
    return sum(model.mass[i] * (pw[i] + Rw[i] @ model.com[i]) for i in range(model.nb)) / model.mass_dyn


def frame_pose(model, q, frame):
    Rw, pw = fk(model, q)
    b = model.frame_parent[frame]
    if b < 0:
        return model.frame_R[frame], model.frame_p[frame]
    return Rw[b] @ model.frame_R[frame], pw[b] + Rw[b] @ model.frame_p[frame]


def random_state(model, rng):
    q = np.zeros(model.nq)
    if model.n_floating:
        q[0:3] = rng.normal(0, 0.3, 3)
        q[3:7] = Rot.from_rotvec(rng.normal(0, 0.8, 3)).as_quat()
        q[7:] = rng.uniform(-1, 1, model.na)
    else:
        q[:] = rng.uniform(-1.5, 1.5, model.na)
    v = rng.normal(0, 1.0, model.nv)
    return q, v
