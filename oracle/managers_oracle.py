"""CPU restatement of the reference's ramp / trajectory managers and the DCM update.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu legs): nothing under
pypnc_b200/ may import this module.  Pinned against the reference's own classes through
tests/golden/managers.npz (tests/golden/make_golden_managers.py, tests/test_golden_oracle.py).

Follows (relative to /root/reference):
  pnc/wbc/manager/task_hierarchy_manager.py:23-35, reaction_force_manager.py:23-35       ramps
  pnc/wbc/manager/foot_trajectory_manager.py:41-110                                       use_current, swing
  pnc/planner/locomotion/dcm_planner/footstep.py:64-69                                    interpolate
  util/interpolation.py:8-32 (smooth_changing*), :46-101 (Hermite), :104-182 (HermiteCurveQuat)
  pnc/wbc/manager/floating_base_trajectory_manager.py:35-115, util/util.py:61-88, :98-107
  pnc/atlas_pnc/atlas_state_estimator.py:35-44                                            DCM
scipy.spatial.transform.Rotation (unpinned dependency of the reference; 1.18 here) is restated
from its documented formulas: from_quat normalises, from_rotvec / as_rotvec switch to a Taylor
series at 1e-3 rad, as_rotvec canonicalises w >= 0, from_matrix picks the largest of
(m00, m11, m22, trace), `a * b` is the Hamilton product, Slerp is R0 * exp(alpha log(R0^-1 R1)).
Quaternions are scalar-last (x, y, z, w).
"""
import numpy as np


# ---- scipy Rotation, restated ------------------------------------------------------------
def q_normalize(q):
    q = np.asarray(q, float)
    return q / np.linalg.norm(q)


def q_mul(p, q):
    c = np.cross(p[:3], q[:3])
    return np.array([p[3] * q[0] + q[3] * p[0] + c[0], p[3] * q[1] + q[3] * p[1] + c[1],
                     p[3] * q[2] + q[3] * p[2] + c[2], p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2]])


def q_inv(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def q_from_rotvec(v):
    v = np.asarray(v, float)
    angle = np.linalg.norm(v)
    if angle <= 1e-3:
        a2 = angle * angle
        scale = 0.5 - a2 / 48 + a2 * a2 / 3840
    else:
        scale = np.sin(angle / 2) / angle
    return np.array([scale * v[0], scale * v[1], scale * v[2], np.cos(angle / 2)])


def q_canonical(q):
    neg = q[3] < 0 or (q[3] == 0 and (q[0] < 0 or (q[0] == 0 and (q[1] < 0 or (q[1] == 0 and q[2] < 0)))))
    return -q if neg else q


def q_as_rotvec(q):
    q = q_canonical(np.asarray(q, float))
    axn = np.linalg.norm(q[:3])
    angle = 2 * np.arctan2(axn, q[3])
    if angle <= 1e-3:
        a2 = angle * angle
        scale = 2 + a2 / 12 + 7 * a2 * a2 / 2880
    else:
        scale = angle / np.sin(angle / 2)
    return scale * q[:3]


def q_to_matrix(q):
    x, y, z, w = q
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    return np.array([[x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw)],
                     [2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw)],
                     [2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2]])


def q_from_matrix(m):
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    dec = [m[0, 0], m[1, 1], m[2, 2], tr]
    c = int(np.argmax(dec))
    if c == 3:
        q = np.array([m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1], 1 + tr])
    else:
        i, j, k = c, (c + 1) % 3, (c + 2) % 3
        q = np.zeros(4)
        q[i] = 1 - tr + 2 * m[i, i]
        q[j] = m[j, i] + m[i, j]
        q[k] = m[k, i] + m[i, k]
        q[3] = m[k, j] - m[j, k]
    return q / np.linalg.norm(q)


def slerp01(q0, q1, alpha):
    """Slerp([0, 1], [q0, q1])(alpha)."""
    q0, q1 = q_normalize(q0), q_normalize(q1)
    rv = q_as_rotvec(q_mul(q_inv(q0), q1))
    return q_mul(q0, q_from_rotvec(alpha * rv))


# ---- util/util.py ------------------------------------------------------------------------
def quat_to_exp(q):  # :61-72
    nv = np.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2])
    theta = 2.0 * np.arcsin(nv)
    if abs(theta) < 1e-4:
        return np.zeros(3)
    return np.asarray(q[:3]) / np.sin(theta / 2.0) * theta


def exp_to_quat(e):  # :75-88
    theta = np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
    if theta > 1e-4:
        s = np.sin(theta / 2.0)
        return np.array([s * e[0] / theta, s * e[1] / theta, s * e[2] / theta, np.cos(theta / 2.0)])
    return np.array([0.5 * e[0], 0.5 * e[1], 0.5 * e[2], 1.0])


# ---- ramps (task_hierarchy_manager.py:23-35, reaction_force_manager.py:23-35) --------------
def ramp(start_value, target, start_time, duration, current_time):
    t = np.clip(current_time, start_time, start_time + duration)
    return (target - start_value) / duration * (t - start_time) + start_value


# ---- util/interpolation.py ---------------------------------------------------------------
def smooth_changing(ini, end, dur, t):  # :8-13
    return end if t > dur else ini + (end - ini) * 0.5 * (1 - np.cos(t / dur * np.pi))


def smooth_changing_vel(ini, end, dur, t):  # :16-21
    return 0.0 if t > dur else (end - ini) * 0.5 * (np.pi / dur) * np.sin(t / dur * np.pi)


def smooth_changing_acc(ini, end, dur, t):  # :24-30
    return 0.0 if t > dur else (end - ini) * 0.5 * (np.pi / dur) * (np.pi / dur) * np.cos(t / dur * np.pi)


def hermite(p1, v1, p2, v2, s_in):  # :46-72, value / first / second derivative with respect to s
    s = np.clip(s_in, 0.0, 1.0)
    p = p1 * (2 * s**3 - 3 * s**2 + 1) + p2 * (-2 * s**3 + 3 * s**2) + v1 * (s**3 - 2 * s**2 + s) + v2 * (s**3 - s**2)
    d = p1 * (6 * s**2 - 6 * s) + p2 * (-6 * s**2 + 6 * s) + v1 * (3 * s**2 - 4 * s + 1) + v2 * (3 * s**2 - 2 * s)
    dd = p1 * (12 * s - 6) + p2 * (-12 * s + 6) + v1 * (6 * s - 4) + v2 * (6 * s - 2)
    return p, d, dd


class HermiteQuat:  # :104-182
    def __init__(self, quat_start, w_start, quat_end, w_end):
        qa, qb = q_normalize(quat_start), q_normalize(quat_end)
        self.q0 = qa
        n = np.linalg.norm(w_start)
        q1 = q_mul(qa, np.array([0, 0, 0, 1.0])) if n < 1e-6 else q_mul(qa, q_from_rotvec((n / 3.0) * (w_start / n)))
        n = np.linalg.norm(w_end)
        q2 = q_mul(qb, np.array([0, 0, 0, 1.0])) if n < 1e-6 else q_mul(qb, q_from_rotvec((n / 3.0) * (w_end / n)))
        q3 = qb
        self.w1 = q_as_rotvec(q_mul(q1, q_inv(self.q0)))
        self.w2 = q_as_rotvec(q_mul(q2, q_inv(q1)))
        self.w3 = q_as_rotvec(q_mul(q3, q_inv(q2)))

    def evaluate(self, s_in):
        s = np.clip(s_in, 0.0, 1.0)
        b = (1 - (1 - s)**3, 3 * s**2 - 2 * s**3, s**3)
        bd = (3 * (1 - s)**2, 6 * s - 6 * s**2, 3 * s**2)
        bdd = (-6 * (1 - s), 6 - 12 * s, 6 * s)
        qs = []
        for w, bi in zip((self.w1, self.w2, self.w3), b):
            n = np.linalg.norm(w)
            qs.append(q_from_rotvec((n * bi) * (w / n)) if n > 1e-5 else np.array([0, 0, 0, 1.0]))
        quat = q_mul(q_mul(q_mul(qs[2], qs[1]), qs[0]), self.q0)
        return quat, self.w1 * bd[0] + self.w2 * bd[1] + self.w3 * bd[2], self.w1 * bdd[0] + self.w2 * bdd[1] + self.w3 * bdd[2]


# ---- FootTrajectoryManager ----------------------------------------------------------------
def foot_use_current(iso, vel):  # foot_trajectory_manager.py:41-52
    """iso 4x4, vel [angular; linear] -> (pos, lin_vel, 0, quat, ang_vel, 0)."""
    return iso[:3, 3].copy(), vel[3:6].copy(), np.zeros(3), q_from_matrix(iso[:3, :3]), vel[0:3].copy(), np.zeros(3)


def swing_foot_desired(init_iso, land_pos, land_quat, start_time, duration, height, t):  # :54-110
    p0, q0 = init_iso[:3, 3], q_from_matrix(init_iso[:3, :3])
    # footstep.interpolate(init, land, 0.5): pos = a p1 + (1-a) p2, quat = Slerp(alpha)
    mid_pos = 0.5 * p0 + 0.5 * land_pos
    mid_rot = q_to_matrix(q_normalize(slerp01(q0, land_quat, 0.5)))
    mid_swing_pos = mid_pos + mid_rot @ np.array([0.0, 0.0, height])
    mid_swing_vel = (land_pos - p0) / duration
    s = (t - start_time) / duration
    if s <= 0.5:
        pos, vel, acc = hermite(p0, np.zeros(3), mid_swing_pos, mid_swing_vel, 2.0 * s)
    else:
        pos, vel, acc = hermite(mid_swing_pos, mid_swing_vel, land_pos, np.zeros(3), 2.0 * (s - 0.5))
    quat, w, wd = HermiteQuat(q0, np.zeros(3), land_quat, np.zeros(3)).evaluate(s)
    return pos, vel, acc, quat, w, wd


# ---- HandTrajectoryManager -----------------------------------------------------------------
def hand_desired(init_iso, init_vel, target_pos, target_quat, start_time, duration, t, keypoint=None):
    """hand_trajectory_manager.py:57-163: initialize_hand_trajectory (or initialize_keypoint_hand_trajectory when
    `keypoint` is given) from the current hand placement / velocity ([angular; linear]), then
    update_hand_trajectory / update_keypoint_hand_trajectory at time t."""
    p0, q0 = init_iso[:3, 3], q_from_matrix(init_iso[:3, :3])
    curve = HermiteQuat(q0, init_vel[0:3], target_quat, np.zeros(3))
    pos, vel, acc = np.zeros(3), np.zeros(3), np.zeros(3)
    if keypoint is None:
        a, b, dur, tt = p0, target_pos, duration, t - start_time
        s = (t - start_time) / duration
    else:
        if t <= start_time + duration * 0.5:
            a, b, dur, tt = p0, keypoint, duration * 0.5, t - start_time
        else:
            a, b, dur, tt = keypoint, target_pos, duration * 0.5, t - start_time - duration * 0.5
        s = (t - start_time) / (0.5 * duration)
    for i in range(3):
        pos[i] = smooth_changing(a[i], b[i], dur, tt)
        vel[i] = smooth_changing_vel(a[i], b[i], dur, tt)
        acc[i] = smooth_changing_acc(a[i], b[i], dur, tt)
    quat, w, wd = curve.evaluate(s)
    return pos, vel, acc, quat, w, wd


# ---- FloatingBaseTrajectoryManager --------------------------------------------------------
def floating_base_desired(ini_com, base_rot, target_com, target_quat, start_time, duration, t):  # :35-54, :72-115
    ini_quat = q_from_matrix(base_rot)
    quat_error = q_from_matrix(q_to_matrix(q_normalize(target_quat)) @ q_to_matrix(ini_quat).T)
    exp_error = quat_to_exp(quat_error)
    dt = t - start_time
    pos = np.array([smooth_changing(ini_com[i], target_com[i], duration, dt) for i in range(3)])
    vel = np.array([smooth_changing_vel(ini_com[i], target_com[i], duration, dt) for i in range(3)])
    acc = np.array([smooth_changing_acc(ini_com[i], target_com[i], duration, dt) for i in range(3)])
    st, sd, sdd = smooth_changing(0, 1, duration, dt), smooth_changing_vel(0, 1, duration, dt), smooth_changing_acc(0, 1, duration, dt)
    quat_inc = exp_to_quat(exp_error * st)
    quat_des = q_from_matrix(q_to_matrix(ini_quat) @ q_to_matrix(q_normalize(quat_inc)))
    return pos, vel, acc, quat_des, exp_error * sd, exp_error * sdd


def sinusoid(start_time, mid, amp, freq, t):  # util.py:98-107
    ph = 2 * np.pi * freq * (t - start_time)
    return amp * np.sin(ph) + mid, amp * 2 * np.pi * freq * np.cos(ph), -amp * (2 * np.pi * freq)**2 * np.sin(ph)


# ---- DCM (atlas_state_estimator.py:35-44) ---------------------------------------------------
def dcm_update(com, vcom, dcm_state, dt, alpha=0.1):
    """Returns (dcm, prev_dcm, dcm_vel).  The reference's low-pass term sits on a line of its own
    (`+(1.0 - alpha) * dcm_vel` is a separate expression statement), so dcm_vel is just the scaled
    finite difference -- restated as executed, not as intended."""
    omega = np.sqrt(9.81 / com[2])
    dcm = com + vcom / omega
    return dcm, dcm_state.copy(), alpha * (dcm - dcm_state) / dt
