// managers.cu -- the reference generators that feed the tasks each tick, one thread per state.
//
// Reference (relative to /root/reference):
//   pnc/wbc/manager/task_hierarchy_manager.py:23-35, reaction_force_manager.py:23-35   linear ramps
//   pnc/wbc/manager/foot_trajectory_manager.py:41-110                                   use_current, swing foot
//   pnc/planner/locomotion/dcm_planner/footstep.py:64-69                                mid-foot interpolation
//   util/interpolation.py:8-32, :46-101, :104-182                                       smooth_changing*, Hermite curves
//   pnc/wbc/manager/floating_base_trajectory_manager.py:35-115                          CoM / base orientation targets
//   pnc/atlas_pnc/atlas_state_estimator.py:35-44                                        DCM update
// scipy.spatial.transform.Rotation as the reference uses it: from_quat normalises, from_rotvec /
// as_rotvec switch to Taylor series at 1e-3 rad, as_rotvec canonicalises w >= 0, `a * b` is the
// Hamilton product, Slerp(alpha) = R0 exp(alpha log(R0^-1 R1)); quaternions scalar-last.
//
// These are elementwise (tens of flops per state, all inputs already in the HBM workspace):
// bandwidth trivial, one launch each; they exist so that a closed-loop many-robot controller never
// leaves the device between update_system and the solve.
#include <math.h>

#include "pnc_launch.h"

namespace pnc {

namespace {

__device__ __forceinline__ void q_normalize(double* q) {
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

__device__ __forceinline__ void q_mul(const double* p, const double* q, double* r) {
  const double cx = p[1] * q[2] - p[2] * q[1], cy = p[2] * q[0] - p[0] * q[2], cz = p[0] * q[1] - p[1] * q[0];
  const double x = p[3] * q[0] + q[3] * p[0] + cx, y = p[3] * q[1] + q[3] * p[1] + cy, z = p[3] * q[2] + q[3] * p[2] + cz;
  const double w = p[3] * q[3] - p[0] * q[0] - p[1] * q[1] - p[2] * q[2];
  r[0] = x; r[1] = y; r[2] = z; r[3] = w;
}

__device__ __forceinline__ void q_inv(const double* q, double* r) { r[0] = -q[0]; r[1] = -q[1]; r[2] = -q[2]; r[3] = q[3]; }

__device__ __forceinline__ void q_from_rotvec(const double* v, double* q) {
  const double angle = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double scale;
  if (angle <= 1e-3) {
    const double a2 = angle * angle;
    scale = 0.5 - a2 / 48 + a2 * a2 / 3840;
  } else {
    scale = sin(angle / 2) / angle;
  }
  q[0] = scale * v[0]; q[1] = scale * v[1]; q[2] = scale * v[2]; q[3] = cos(angle / 2);
}

__device__ __forceinline__ void q_as_rotvec(const double* qin, double* v) {
  double q[4] = {qin[0], qin[1], qin[2], qin[3]};
  const bool neg = q[3] < 0 || (q[3] == 0 && (q[0] < 0 || (q[0] == 0 && (q[1] < 0 || (q[1] == 0 && q[2] < 0)))));
  if (neg) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double axn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  const double angle = 2 * atan2(axn, q[3]);
  double scale;
  if (angle <= 1e-3) {
    const double a2 = angle * angle;
    scale = 2 + a2 / 12 + 7 * a2 * a2 / 2880;
  } else {
    scale = angle / sin(angle / 2);
  }
  v[0] = scale * q[0]; v[1] = scale * q[1]; v[2] = scale * q[2];
}

// value, d/ds and d2/ds2 of the cubic Hermite segment (util/interpolation.py:46-72)
__device__ __forceinline__ void hermite(double p1, double v1, double p2, double v2, double s_in, double& p, double& d,
                                        double& dd) {
  const double s = fmin(fmax(s_in, 0.0), 1.0), s2 = s * s, s3 = s2 * s;
  p = p1 * (2 * s3 - 3 * s2 + 1) + p2 * (-2 * s3 + 3 * s2) + v1 * (s3 - 2 * s2 + s) + v2 * (s3 - s2);
  d = p1 * (6 * s2 - 6 * s) + p2 * (-6 * s2 + 6 * s) + v1 * (3 * s2 - 4 * s + 1) + v2 * (3 * s2 - 2 * s);
  dd = p1 * (12 * s - 6) + p2 * (-12 * s + 6) + v1 * (6 * s - 4) + v2 * (6 * s - 2);
}

// util/util.py:75-88
__device__ __forceinline__ void exp_to_quat(const double* e, double* q) {
  const double theta = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  if (theta > 1e-4) {
    const double s = sin(theta / 2.0);
    q[0] = s * e[0] / theta; q[1] = s * e[1] / theta; q[2] = s * e[2] / theta; q[3] = cos(theta / 2.0);
  } else {
    q[0] = 0.5 * e[0]; q[1] = 0.5 * e[1]; q[2] = 0.5 * e[2]; q[3] = 1.0;
  }
}

// util/util.py:61-72
__device__ __forceinline__ void quat_to_exp_m(const double* q, double* e) {
  const double theta = 2.0 * asin(sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]));
  if (fabs(theta) < 1e-4) { e[0] = e[1] = e[2] = 0.0; return; }
  const double sn = sin(theta / 2.0);
  e[0] = q[0] / sn * theta; e[1] = q[1] / sn * theta; e[2] = q[2] / sn * theta;
}

constexpr double kPi = 3.141592653589793238462643383279502884;

}  // namespace

// ---- TaskHierarchyManager / ReactionForceManager.update_ramp_to_{min,max} ------------------
__global__ void k_ramp_update(int B, const double* __restrict__ start_value, double target,
                              const double* __restrict__ start_time, const double* __restrict__ duration, double t_now,
                              double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const double t0 = start_time[s], dur = duration[s], v0 = start_value[s];
  const double t = fmin(fmax(t_now, t0), t0 + dur);
  out[s] = (target - v0) / dur * (t - t0) + v0;
}

// ---- FootTrajectoryManager.use_current (:41-52) ---------------------------------------------
__global__ void k_foot_use_current(WsPtrs ws, int nfr, int slot, int B, double* __restrict__ pos, double* __restrict__ vel,
                                   double* __restrict__ acc, double* __restrict__ quat, double* __restrict__ angvel,
                                   double* __restrict__ angacc) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const double* iso = ws.fr_iso + ((size_t)s * nfr + slot) * 12;
  const double* fv = ws.fr_vel + ((size_t)s * nfr + slot) * 6;
  double q[4];
  rot_to_quat(iso, q);
  for (int i = 0; i < 3; i++) {
    pos[3 * s + i] = iso[9 + i];
    vel[3 * s + i] = fv[3 + i];
    acc[3 * s + i] = 0.0;
    angvel[3 * s + i] = fv[i];
    angacc[3 * s + i] = 0.0;
  }
  for (int i = 0; i < 4; i++) quat[4 * s + i] = q[i];
}

// ---- FootTrajectoryManager.initialize_swing_foot_trajectory (:54-84) ------------------------
// record (PNC_SWING_REC doubles): p_init 3, p_mid 3, v_mid 3, p_land 3, q0 4, omega1 3, omega2 3, omega3 3
__global__ void k_swing_init(WsPtrs ws, int nfr, int slot, int B, const double* __restrict__ land_pos,
                             const double* __restrict__ land_quat, const double* __restrict__ duration, double height,
                             double* __restrict__ rec_all) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const double* iso = ws.fr_iso + ((size_t)s * nfr + slot) * 12;
  double q0[4], ql[4] = {land_quat[4 * s], land_quat[4 * s + 1], land_quat[4 * s + 2], land_quat[4 * s + 3]};
  rot_to_quat(iso, q0);  // Footstep.iso setter: R.from_matrix(...).as_quat()
  q_normalize(ql);
  // mid foot: interpolate(init, land, 0.5) -- pos = 0.5 p0 + 0.5 pl, quat = Slerp(0.5)
  double qi[4], rel[4], rv[3], half[4], qm[4], Rm[9];
  q_inv(q0, qi);
  q_mul(qi, ql, rel);
  q_as_rotvec(rel, rv);
  rv[0] *= 0.5; rv[1] *= 0.5; rv[2] *= 0.5;
  q_from_rotvec(rv, half);
  q_mul(q0, half, qm);
  q_normalize(qm);  // Footstep.quat setter goes through R.from_quat
  quat_to_rot(qm, Rm);
  double* rec = rec_all + (size_t)s * PNC_SWING_REC;
  const double dur = duration[s];
  for (int i = 0; i < 3; i++) {
    const double p0 = iso[9 + i], pl = land_pos[3 * s + i];
    rec[i] = p0;
    rec[3 + i] = 0.5 * p0 + 0.5 * pl + Rm[3 * i + 2] * height;  // mid pos + R_mid [0, 0, h]
    rec[6 + i] = (pl - p0) / dur;
    rec[9 + i] = pl;
  }
  for (int i = 0; i < 4; i++) rec[12 + i] = q0[i];
  // HermiteCurveQuat(q0, 0, ql, 0): q1 = q0 * identity, q2 = ql * identity, q3 = ql (:104-134)
  const double ident[4] = {0.0, 0.0, 0.0, 1.0};
  double q1[4], q2[4], t4[4], w[3];
  q_mul(q0, ident, q1);
  q_mul(ql, ident, q2);
  q_mul(q1, qi, t4); q_as_rotvec(t4, w);
  rec[16] = w[0]; rec[17] = w[1]; rec[18] = w[2];
  q_inv(q1, t4); { double u[4]; q_mul(q2, t4, u); q_as_rotvec(u, w); }
  rec[19] = w[0]; rec[20] = w[1]; rec[21] = w[2];
  q_inv(q2, t4); { double u[4]; q_mul(ql, t4, u); q_as_rotvec(u, w); }
  rec[22] = w[0]; rec[23] = w[1]; rec[24] = w[2];
  rec[25] = 0.0;
}

// ---- FootTrajectoryManager.update_swing_foot_desired (:86-110) ------------------------------
__global__ void k_swing_eval(int B, const double* __restrict__ rec_all, const double* __restrict__ start_time,
                             const double* __restrict__ duration, double t_now, double* __restrict__ pos,
                             double* __restrict__ vel, double* __restrict__ acc, double* __restrict__ quat,
                             double* __restrict__ angvel, double* __restrict__ angacc) {
  const int st = blockIdx.x * blockDim.x + threadIdx.x;
  if (st >= B) return;
  const double* rec = rec_all + (size_t)st * PNC_SWING_REC;
  const double s = (t_now - start_time[st]) / duration[st];
  for (int i = 0; i < 3; i++) {
    double p, d, dd;
    if (s <= 0.5) hermite(rec[i], 0.0, rec[3 + i], rec[6 + i], 2.0 * s, p, d, dd);
    else hermite(rec[3 + i], rec[6 + i], rec[9 + i], 0.0, 2.0 * (s - 0.5), p, d, dd);
    pos[3 * st + i] = p; vel[3 * st + i] = d; acc[3 * st + i] = dd;
  }
  // HermiteCurveQuat.evaluate / evaluate_ang_vel / evaluate_ang_acc (:136-182)
  const double sc = fmin(fmax(s, 0.0), 1.0), om = 1.0 - sc;
  const double b[3] = {1 - om * om * om, 3 * sc * sc - 2 * sc * sc * sc, sc * sc * sc};
  const double bd[3] = {3 * om * om, 6 * sc - 6 * sc * sc, 3 * sc * sc};
  const double bdd[3] = {-6 * om, 6 - 12 * sc, 6 * sc};
  double qacc[4] = {rec[12], rec[13], rec[14], rec[15]};  // q0, then qtmp1, qtmp2, qtmp3 from the left
  for (int k = 0; k < 3; k++) {
    const double* w = rec + 16 + 3 * k;
    const double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double qt[4] = {0.0, 0.0, 0.0, 1.0};
    if (n > 1e-5) {
      const double v[3] = {(n * b[k]) * (w[0] / n), (n * b[k]) * (w[1] / n), (n * b[k]) * (w[2] / n)};
      q_from_rotvec(v, qt);
    }
    double r[4];
    q_mul(qt, qacc, r);
    qacc[0] = r[0]; qacc[1] = r[1]; qacc[2] = r[2]; qacc[3] = r[3];
  }
  for (int i = 0; i < 4; i++) quat[4 * st + i] = qacc[i];
  for (int i = 0; i < 3; i++) {
    angvel[3 * st + i] = rec[16 + i] * bd[0] + rec[19 + i] * bd[1] + rec[22 + i] * bd[2];
    angacc[3 * st + i] = rec[16 + i] * bdd[0] + rec[19 + i] * bdd[1] + rec[22 + i] * bdd[2];
  }
}

// ---- HandTrajectoryManager (pnc/wbc/manager/hand_trajectory_manager.py:57-163) ---------------
// record (PNC_HAND_REC doubles): p_init 3, p_keypoint 3, p_target 3, q0 4, omega1 3, omega2 3, omega3 3, pad 2
__global__ void k_hand_init(WsPtrs ws, int nfr, int slot, int B, const double* __restrict__ target_pos,
                            const double* __restrict__ target_quat, const double* __restrict__ keypoint,
                            double* __restrict__ rec_all) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const double* iso = ws.fr_iso + ((size_t)s * nfr + slot) * 12;
  const double* fv = ws.fr_vel + ((size_t)s * nfr + slot) * 6;
  double q0[4], qt[4] = {target_quat[4 * s], target_quat[4 * s + 1], target_quat[4 * s + 2], target_quat[4 * s + 3]};
  rot_to_quat(iso, q0);
  q_normalize(qt);
  double* rec = rec_all + (size_t)s * PNC_HAND_REC;
  for (int i = 0; i < 3; i++) {
    rec[i] = iso[9 + i];
    rec[3 + i] = keypoint ? keypoint[3 * s + i] : 0.0;
    rec[6 + i] = target_pos[3 * s + i];
  }
  for (int i = 0; i < 4; i++) rec[9 + i] = q0[i];
  // HermiteCurveQuat(q0, omega_a, qt, 0) (util/interpolation.py:104-134): q1 = q0 * exp(omega_a / 3) unless
  // |omega_a| < 1e-6, q2 = qt * identity, q3 = qt
  const double wa[3] = {fv[0], fv[1], fv[2]};
  const double na = sqrt(wa[0] * wa[0] + wa[1] * wa[1] + wa[2] * wa[2]);
  double step[4] = {0.0, 0.0, 0.0, 1.0};
  if (!(na < 1e-6)) {
    const double v[3] = {(na / 3.0) * (wa[0] / na), (na / 3.0) * (wa[1] / na), (na / 3.0) * (wa[2] / na)};
    q_from_rotvec(v, step);
  }
  const double ident[4] = {0.0, 0.0, 0.0, 1.0};
  double q1[4], q2[4], qi[4], t4[4], u[4], w[3];
  q_mul(q0, step, q1);
  q_mul(qt, ident, q2);
  q_inv(q0, qi); q_mul(q1, qi, t4); q_as_rotvec(t4, w);
  rec[13] = w[0]; rec[14] = w[1]; rec[15] = w[2];
  q_inv(q1, t4); q_mul(q2, t4, u); q_as_rotvec(u, w);
  rec[16] = w[0]; rec[17] = w[1]; rec[18] = w[2];
  q_inv(q2, t4); q_mul(qt, t4, u); q_as_rotvec(u, w);
  rec[19] = w[0]; rec[20] = w[1]; rec[21] = w[2];
  rec[22] = 0.0; rec[23] = 0.0;
}

// update_hand_trajectory (:104-129) / update_keypoint_hand_trajectory (:131-178): cosine blend of the position
// (util/interpolation.py:8-30, no clipping below zero, held after the end), HermiteCurveQuat for the orientation
__global__ void k_hand_eval(int B, const double* __restrict__ rec_all, const double* __restrict__ start_time,
                            const double* __restrict__ duration, double t_now, int keypoint, double* __restrict__ pos,
                            double* __restrict__ vel, double* __restrict__ acc, double* __restrict__ quat,
                            double* __restrict__ angvel, double* __restrict__ angacc) {
  const int st = blockIdx.x * blockDim.x + threadIdx.x;
  if (st >= B) return;
  const double* rec = rec_all + (size_t)st * PNC_HAND_REC;
  const double t0 = start_time[st], dur = duration[st];
  const double* a = rec;
  const double* b = rec + 6;
  double d = dur, tt = t_now - t0, s = (t_now - t0) / dur;
  if (keypoint) {
    d = dur * 0.5;
    if (t_now <= t0 + dur * 0.5) { b = rec + 3; }
    else { a = rec + 3; tt = t_now - t0 - dur * 0.5; }
    s = (t_now - t0) / (0.5 * dur);
  }
  const double PI = 3.141592653589793;
  for (int i = 0; i < 3; i++) {
    const double ph = tt / d * PI;
    double p = a[i] + (b[i] - a[i]) * 0.5 * (1.0 - cos(ph));
    double v = (b[i] - a[i]) * 0.5 * (PI / d) * sin(ph);
    double c = (b[i] - a[i]) * 0.5 * (PI / d) * (PI / d) * cos(ph);
    if (tt > d) { p = b[i]; v = 0.0; c = 0.0; }
    pos[3 * st + i] = p; vel[3 * st + i] = v; acc[3 * st + i] = c;
  }
  const double sc = fmin(fmax(s, 0.0), 1.0), om = 1.0 - sc;
  const double bb[3] = {1 - om * om * om, 3 * sc * sc - 2 * sc * sc * sc, sc * sc * sc};
  const double bd[3] = {3 * om * om, 6 * sc - 6 * sc * sc, 3 * sc * sc};
  const double bdd[3] = {-6 * om, 6 - 12 * sc, 6 * sc};
  double qacc[4] = {rec[9], rec[10], rec[11], rec[12]};
  for (int k = 0; k < 3; k++) {
    const double* w = rec + 13 + 3 * k;
    const double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double qt[4] = {0.0, 0.0, 0.0, 1.0};
    if (n > 1e-5) {
      const double v[3] = {(n * bb[k]) * (w[0] / n), (n * bb[k]) * (w[1] / n), (n * bb[k]) * (w[2] / n)};
      q_from_rotvec(v, qt);
    }
    double r[4];
    q_mul(qt, qacc, r);
    qacc[0] = r[0]; qacc[1] = r[1]; qacc[2] = r[2]; qacc[3] = r[3];
  }
  for (int i = 0; i < 4; i++) quat[4 * st + i] = qacc[i];
  for (int i = 0; i < 3; i++) {
    angvel[3 * st + i] = rec[13 + i] * bd[0] + rec[16 + i] * bd[1] + rec[19 + i] * bd[2];
    angacc[3 * st + i] = rec[13 + i] * bdd[0] + rec[16 + i] * bdd[1] + rec[19 + i] * bdd[2];
  }
}

// ---- FloatingBaseTrajectoryManager.initialize_floating_base_interpolation_trajectory (:35-54) -
// record (PNC_FB_REC doubles): ini_com 3, target_com 3, ini_quat 4, exp_error 3, pad
__global__ void k_floating_base_init(WsPtrs ws, int nfr, int slot, int B, const double* __restrict__ target_com,
                                     const double* __restrict__ target_quat, double* __restrict__ rec_all) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const double* iso = ws.fr_iso + ((size_t)s * nfr + slot) * 12;
  double* rec = rec_all + (size_t)s * PNC_FB_REC;
  double qi[4], qt[4] = {target_quat[4 * s], target_quat[4 * s + 1], target_quat[4 * s + 2], target_quat[4 * s + 3]};
  rot_to_quat(iso, qi);  // util.rot_to_quat
  q_normalize(qt);
  double Rt[9], Ri[9], Re[9], qe[4], ee[3];
  quat_to_rot(qt, Rt);
  quat_to_rot(qi, Ri);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      Re[3 * i + j] = Rt[3 * i] * Ri[3 * j] + Rt[3 * i + 1] * Ri[3 * j + 1] + Rt[3 * i + 2] * Ri[3 * j + 2];
  rot_to_quat(Re, qe);
  quat_to_exp_m(qe, ee);
  for (int i = 0; i < 3; i++) {
    rec[i] = ws.com[3 * (size_t)s + i];
    rec[3 + i] = target_com[3 * s + i];
    rec[10 + i] = ee[i];
  }
  for (int i = 0; i < 4; i++) rec[6 + i] = qi[i];
  rec[13] = 0.0;
}

// ---- FloatingBaseTrajectoryManager.update_floating_base_desired, interpolation branch (:72-115) -
__global__ void k_floating_base_eval(int B, const double* __restrict__ rec_all, const double* __restrict__ start_time,
                                     const double* __restrict__ duration, double t_now, double* __restrict__ pos,
                                     double* __restrict__ vel, double* __restrict__ acc, double* __restrict__ quat,
                                     double* __restrict__ angvel, double* __restrict__ angacc) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const double* rec = rec_all + (size_t)s * PNC_FB_REC;
  const double dur = duration[s], t = t_now - start_time[s];
  const bool past = t > dur;
  const double ph = t / dur * kPi, w = kPi / dur;
  // smooth_changing / _vel / _acc of (0 -> 1) (util/interpolation.py:8-32)
  const double st = past ? 1.0 : 0.0 + (1.0 - 0.0) * 0.5 * (1 - cos(ph));
  const double sd = past ? 0.0 : (1.0 - 0.0) * 0.5 * w * sin(ph);
  const double sdd = past ? 0.0 : (1.0 - 0.0) * 0.5 * w * w * cos(ph);
  for (int i = 0; i < 3; i++) {
    const double a = rec[i], b = rec[3 + i];
    pos[3 * s + i] = past ? b : a + (b - a) * 0.5 * (1 - cos(ph));
    vel[3 * s + i] = past ? 0.0 : (b - a) * 0.5 * w * sin(ph);
    acc[3 * s + i] = past ? 0.0 : (b - a) * 0.5 * w * w * cos(ph);
    angvel[3 * s + i] = rec[10 + i] * sd;
    angacc[3 * s + i] = rec[10 + i] * sdd;
  }
  const double einc[3] = {rec[10] * st, rec[11] * st, rec[12] * st};
  double qinc[4], Ri[9], Rn[9], Rd[9], qd[4];
  exp_to_quat(einc, qinc);
  q_normalize(qinc);  // R.from_quat
  quat_to_rot(rec + 6, Ri);
  quat_to_rot(qinc, Rn);
  mul33(Ri, Rn, Rd);
  rot_to_quat(Rd, qd);
  for (int i = 0; i < 4; i++) quat[4 * s + i] = qd[i];
}

// ---- util.get_sinusoid_trajectory (util/util.py:98-107), the swaying branch (:76-80) ----------
__global__ void k_sinusoid_eval(int B, const double* __restrict__ mid, const double* __restrict__ start_time, double a0,
                                double a1, double a2, double f0, double f1, double f2, double t_now,
                                double* __restrict__ pos, double* __restrict__ vel, double* __restrict__ acc) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const double amp[3] = {a0, a1, a2}, freq[3] = {f0, f1, f2};
  const double dt = t_now - start_time[s];
  for (int i = 0; i < 3; i++) {
    const double ph = 2 * kPi * freq[i] * dt;
    pos[3 * s + i] = amp[i] * sin(ph) + mid[3 * s + i];
    vel[3 * s + i] = amp[i] * 2 * kPi * freq[i] * cos(ph);
    const double w2 = (2 * kPi * freq[i]) * (2 * kPi * freq[i]);
    acc[3 * s + i] = -amp[i] * w2 * sin(ph);
  }
}

// ---- AtlasStateEstimator._update_dcm (:35-44), as executed: the low-pass term of dcm_vel sits on
//      a statement of its own in the reference, so dcm_vel = alpha (dcm - prev) / dt --------------
__global__ void k_dcm_update(WsPtrs ws, int B, double dt, double alpha, double* __restrict__ dcm, double* __restrict__ prev,
                             double* __restrict__ dcm_vel) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const double* c = ws.com + 3 * (size_t)s;
  const double* v = ws.vcom + 3 * (size_t)s;
  const double omega = sqrt(9.81 / c[2]);
  for (int i = 0; i < 3; i++) {
    const double old = dcm[3 * s + i], now = c[i] + v[i] / omega;
    prev[3 * s + i] = old;
    dcm[3 * s + i] = now;
    dcm_vel[3 * s + i] = alpha * (now - old) / dt;
  }
}

static inline int grid_for(int B) { return (B + 127) / 128; }

void launch_ramp_update(int B, const double* sv, double target, const double* st, const double* dur, double t, double* out,
                        cudaStream_t stream) {
  k_ramp_update<<<grid_for(B), 128, 0, stream>>>(B, sv, target, st, dur, t, out);
  g_launch_count++;
}
void launch_foot_use_current(const WsPtrs& ws, int nfr, int slot, int B, double* pos, double* vel, double* acc, double* quat,
                             double* angvel, double* angacc, cudaStream_t stream) {
  k_foot_use_current<<<grid_for(B), 128, 0, stream>>>(ws, nfr, slot, B, pos, vel, acc, quat, angvel, angacc);
  g_launch_count++;
}
void launch_swing_init(const WsPtrs& ws, int nfr, int slot, int B, const double* land_pos, const double* land_quat,
                       const double* duration, double height, double* rec, cudaStream_t stream) {
  k_swing_init<<<grid_for(B), 128, 0, stream>>>(ws, nfr, slot, B, land_pos, land_quat, duration, height, rec);
  g_launch_count++;
}
void launch_swing_eval(int B, const double* rec, const double* st, const double* dur, double t, double* pos, double* vel,
                       double* acc, double* quat, double* angvel, double* angacc, cudaStream_t stream) {
  k_swing_eval<<<grid_for(B), 128, 0, stream>>>(B, rec, st, dur, t, pos, vel, acc, quat, angvel, angacc);
  g_launch_count++;
}
void launch_hand_init(const WsPtrs& ws, int nfr, int slot, int B, const double* tpos, const double* tquat, const double* key,
                      double* rec, cudaStream_t stream) {
  k_hand_init<<<grid_for(B), 128, 0, stream>>>(ws, nfr, slot, B, tpos, tquat, key, rec);
  g_launch_count++;
}
void launch_hand_eval(int B, const double* rec, const double* st, const double* dur, double t, int keypoint, double* pos,
                      double* vel, double* acc, double* quat, double* angvel, double* angacc, cudaStream_t stream) {
  k_hand_eval<<<grid_for(B), 128, 0, stream>>>(B, rec, st, dur, t, keypoint, pos, vel, acc, quat, angvel, angacc);
  g_launch_count++;
}
void launch_floating_base_init(const WsPtrs& ws, int nfr, int slot, int B, const double* tcom, const double* tquat,
                               double* rec, cudaStream_t stream) {
  k_floating_base_init<<<grid_for(B), 128, 0, stream>>>(ws, nfr, slot, B, tcom, tquat, rec);
  g_launch_count++;
}
void launch_floating_base_eval(int B, const double* rec, const double* st, const double* dur, double t, double* pos,
                               double* vel, double* acc, double* quat, double* angvel, double* angacc,
                               cudaStream_t stream) {
  k_floating_base_eval<<<grid_for(B), 128, 0, stream>>>(B, rec, st, dur, t, pos, vel, acc, quat, angvel, angacc);
  g_launch_count++;
}
void launch_sinusoid_eval(int B, const double* mid, const double* st, const double* amp3, const double* freq3, double t,
                          double* pos, double* vel, double* acc, cudaStream_t stream) {
  k_sinusoid_eval<<<grid_for(B), 128, 0, stream>>>(B, mid, st, amp3[0], amp3[1], amp3[2], freq3[0], freq3[1], freq3[2], t,
                                                   pos, vel, acc);
  g_launch_count++;
}
void launch_dcm_update(const WsPtrs& ws, int B, double dt, double alpha, double* dcm, double* prev, double* dcm_vel,
                       cudaStream_t stream) {
  k_dcm_update<<<grid_for(B), 128, 0, stream>>>(ws, B, dt, alpha, dcm, prev, dcm_vel);
  g_launch_count++;
}

}  // namespace pnc
