// ihwbc_common.cuh -- device helpers shared by the IHWBC kernels: warp reductions, task and
// contact evaluation (reference pnc/wbc/basic_task.py:25-137, pnc/wbc/basic_contact.py:24-173).
#pragma once
#include <math.h>

#include "pnc_launch.h"

namespace pnc {

#define FULL 0xffffffffu
#define QP_EPS 2.220446049250313e-16

static __device__ __noinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
// lexicographic (value, index) minimum over the warp; every lane gets the result.
// Three REDUX (integer warp reductions) on an order-preserving 64-bit key instead of five
// shuffle rounds: the latency of this chain sits on the critical path of every iteration.
static __device__ __forceinline__ void wargmin(double& v, int& idx) {
  const long long bits = __double_as_longlong(v);
  const unsigned long long key = bits < 0 ? ~(unsigned long long)bits : ((unsigned long long)bits | 0x8000000000000000ull);
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_min_sync(FULL, hi);
  const unsigned ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
  const bool win = hi == mh && lo == ml;
  const unsigned mi = __reduce_min_sync(FULL, win ? (unsigned)idx : 0xffffffffu);
  const unsigned long long mk = ((unsigned long long)mh << 32) | ml;
  const long long mb = (mk & 0x8000000000000000ull) ? (long long)(mk & 0x7fffffffffffffffull) : (long long)~mk;
  v = __longlong_as_double(mb);
  idx = (int)mi;
}

// 1 / x without the IEEE corner-case path of the compiler's division sequence (a slow-path call the hot loops pay
// for in instructions and in instruction-cache footprint): hardware seed (2^-23), two Newton steps, <= 2 ulp.
// For finite, normal, non-zero x -- step-length and reflection denominators, pivots.
__device__ __forceinline__ double fast_rcp(double x) {
#ifdef PNC_IEEE_DIV
  return 1.0 / x;
#endif
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

__device__ __forceinline__ double qp_distance(double a, double b) {  // QuadProg++.cc:681-692
  double a1 = fabs(a), b1 = fabs(b);
  if (a1 > b1) { double t = b1 / a1; return a1 * sqrt(1.0 + t * t); }
  if (b1 > a1) { double t = a1 / b1; return b1 * sqrt(1.0 + t * t); }
  return a1 * sqrt(2.0);
}

// util/util.py:61-72
__device__ __forceinline__ void quat_to_exp(const double* q, double* e) {
  double theta = 2.0 * asin(sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]));
  if (fabs(theta) < 1e-4) { e[0] = e[1] = e[2] = 0.0; return; }
  double sn = sin(theta / 2.0);
  e[0] = q[0] / sn * theta; e[1] = q[1] / sn * theta; e[2] = q[2] / sn * theta;
}

// ---------------------------------------------------------------------------------------
// BasicTask.update_cmd for the 3-row tasks (COM, LINK_XYZ, LINK_ORI), basic_task.py:25-109
struct TaskVals {
  double perr[3], op[3], jd[3];
};

__device__ inline void eval_dense_task(const DevProblem& P, const DevTask& T, int wsf, const WsPtrs& ws, size_t s, int nfr,
                                const double* __restrict__ des, TaskVals& o) {
  const double* pos_des = des + T.des_off;
  const int npos = T.type == PNC_TASK_LINK_ORI ? 4 : 3;
  const double* vel_des = pos_des + npos;
  const double* acc_des = vel_des + 3;
  double vact[3];
  if (T.type == PNC_TASK_COM) {
    for (int i = 0; i < 3; i++) {
      o.perr[i] = pos_des[i] - ws.com[3 * s + i];
      vact[i] = ws.vcom[3 * s + i];
      o.jd[i] = ws.jcomdot_qdot[3 * s + i];
    }
  } else {
    const double* iso = ws.fr_iso + (s * nfr + wsf) * 12;
    const double* fv = ws.fr_vel + (s * nfr + wsf) * 6;
    const double* fa = ws.fr_jdqd + (s * nfr + wsf) * 6;
    if (T.type == PNC_TASK_LINK_XYZ) {
      for (int i = 0; i < 3; i++) { o.perr[i] = pos_des[i] - iso[9 + i]; vact[i] = fv[3 + i]; o.jd[i] = fa[3 + i]; }
    } else {  // LINK_ORI, basic_task.py:62-78
      double Rdes[9], Ract[9], Rerr[9], qa[4], qe[4];
      quat_to_rot(pos_des, Rdes);
      double Rl[9];
      for (int k = 0; k < 9; k++) Rl[k] = iso[k];
      rot_to_quat(Rl, qa);  // R.from_matrix(rot) -> quaternion -> back to a rotation
      quat_to_rot(qa, Ract);
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          Rerr[3 * i + j] = Rdes[3 * i] * Ract[3 * j] + Rdes[3 * i + 1] * Ract[3 * j + 1] + Rdes[3 * i + 2] * Ract[3 * j + 2];
      rot_to_quat(Rerr, qe);
      quat_to_exp(qe, o.perr);
      for (int i = 0; i < 3; i++) { vact[i] = fv[i]; o.jd[i] = fa[i]; }
    }
  }
  for (int i = 0; i < 3; i++)
    o.op[i] = acc_des[i] + P.kp[T.gain_off + i] * o.perr[i] + P.kd[T.gain_off + i] * (vel_des[i] - vact[i]);
}

// JOINT / SELECTED_JOINT row i, basic_task.py:27-47
__device__ __forceinline__ void eval_joint_row(const DevProblem& P, const DevTask& T, int i, const WsPtrs& ws, size_t s,
                                               int nq, const double* __restrict__ des, int& vi, double& perr,
                                               double& op) {
  const double* pos_des = des + T.des_off;
  const double* vel_des = pos_des + T.dim;
  const double* acc_des = vel_des + T.dim;
  vi = T.type == PNC_TASK_JOINT ? 6 + i : P.joint_sel[T.joint_off + i];
  double qj = ws.q[s * nq + vi + 1];
  double qd = ws.qdot[s * P.nv + vi];
  perr = pos_des[i] - qj;
  op = acc_des[i] + P.kp[T.gain_off + i] * perr + P.kd[T.gain_off + i] * (vel_des[i] - qd);
}

__device__ __forceinline__ const double* dense_row_ptr(const DevTask& T, int wsf, const WsPtrs& ws, size_t s, int nfr,
                                                       int nv, int i) {
  if (T.type == PNC_TASK_COM) return ws.jcom + (s * 3 + i) * nv;
  const double* J = ws.fr_jac + (s * nfr + wsf) * 6 * nv;
  return J + (T.type == PNC_TASK_LINK_XYZ ? 3 + i : i) * nv;
}

// wrench-cone row `lr` of a contact in its own frame, basic_contact.py:24-51 / 89-173
__device__ __forceinline__ void cone_row_local(const DevContact& C, int lr, double* u6) {
#pragma unroll
  for (int k = 0; k < 6; k++) u6[k] = 0.0;
  const double mu = C.mu;
  if (C.type == PNC_CONTACT_POINT) {
    switch (lr) {
      case 0: u6[2] = 1.0; break;
      case 1: u6[0] = 1.0; u6[2] = mu; break;
      case 2: u6[0] = -1.0; u6[2] = mu; break;
      case 3: u6[1] = 1.0; u6[2] = mu; break;
      case 4: u6[1] = -1.0; u6[2] = mu; break;
      default: u6[2] = -1.0; break;
    }
    return;
  }
  const double x = C.x, y = C.y;
  if (lr == 0) u6[5] = 1.0;
  else if (lr == 1) { u6[3] = 1.0; u6[5] = mu; }
  else if (lr == 2) { u6[3] = -1.0; u6[5] = mu; }
  else if (lr == 3) { u6[4] = 1.0; u6[5] = mu; }
  else if (lr == 4) { u6[4] = -1.0; u6[5] = mu; }
  else if (lr == 5) { u6[0] = 1.0; u6[5] = y; }
  else if (lr == 6) { u6[0] = -1.0; u6[5] = y; }
  else if (lr == 7) { u6[1] = 1.0; u6[5] = x; }
  else if (lr == 8) { u6[1] = -1.0; u6[5] = x; }
  else if (lr == 17) u6[5] = -1.0;
  else {
    // rows 9..16: signs of the (mu, mu, 1, y, x) columns, basic_contact.py:121-169
    const int r = lr - 9;
    const double s0 = (r & 2) ? 1.0 : -1.0;            // -,-,+,+,-,-,+,+
    const double s1 = (r & 1) ? 1.0 : -1.0;            // -,+,-,+,-,+,-,+
    const double s2 = (r & 4) ? -1.0 : 1.0;            // +,+,+,+,-,-,-,-
    const double s3 = ((r & 2) ? -1.0 : 1.0) * s2;     // +,+,-,-,-,-,+,+
    const double s4 = ((r & 1) ? -1.0 : 1.0) * s2;     // +,-,+,-,-,+,-,+
    u6[0] = s0 * mu; u6[1] = s1 * mu; u6[2] = s2; u6[3] = s3 * y; u6[4] = s4 * x; u6[5] = (x + y) * mu;
  }
}

// cone row in world axes: Uf = U blockdiag(R^T, R^T)  (surface, :72-87); point cones are not rotated (:31 unused)
__device__ __forceinline__ void cone_row_world(const DevContact& C, int lr, const double* __restrict__ iso, double* u6) {
  double l[6];
  cone_row_local(C, lr, l);
  if (C.type == PNC_CONTACT_POINT) {
#pragma unroll
    for (int k = 0; k < 6; k++) u6[k] = l[k];
    return;
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    u6[j] = l[0] * iso[3 * j] + l[1] * iso[3 * j + 1] + l[2] * iso[3 * j + 2];
    u6[3 + j] = l[3] * iso[3 * j] + l[4] * iso[3 * j + 1] + l[5] * iso[3 * j + 2];
  }
}

__device__ __forceinline__ const double* contact_jac_row(const DevContact& C, int wsf, const WsPtrs& ws, size_t s, int nfr,
                                                         int nv, int k) {
  const double* J = ws.fr_jac + (s * nfr + wsf) * 6 * nv;
  return J + (C.type == PNC_CONTACT_POINT ? 3 + k : k) * nv;
}


}  // namespace pnc
