// pnc_dev.cuh -- device-side tables and helpers shared by the kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pnc_b200.h"

#define PNC_QP_OVERFLOW 5  /* internal status: the register kernel hands the state to the generic kernel */

namespace pnc {

// ---- model tables (one copy in HBM, staged into shared memory by each CTA) ----------
struct DevModel {
  int nb, nq, nv, na, nfl, nf, maxdepth, pad0;
  double total_mass, mass_dyn;
  int parent[PNC_MAX_BODIES], jtype[PNC_MAX_BODIES], idx_v[PNC_MAX_BODIES], subtree[PNC_MAX_BODIES],
      depth[PNC_MAX_BODIES];
  double place_R[PNC_MAX_BODIES][9], place_p[PNC_MAX_BODIES][3], axis[PNC_MAX_BODIES][3];
  double mass[PNC_MAX_BODIES], com[PNC_MAX_BODIES][3], inertia[PNC_MAX_BODIES][6];  // xx xy xz yy yz zz
};

// frames materialised by a workspace
struct DevFrames {
  int nfr, pad;
  int body[PNC_MAX_FRAMES];      // parent body (-1 = universe)
  int model_frame[PNC_MAX_FRAMES];
  double R[PNC_MAX_FRAMES][9];   // placement in the parent joint frame
  double p[PNC_MAX_FRAMES][3];
};

// per-state rigid-body quantities in HBM, batch-major (SURVEY.md 8b / pnc_b200.h)
struct WsPtrs {
  double *q, *qdot, *M, *grav, *cori, *com, *vcom, *jcom, *jcomdot, *jcomdot_qdot;
  double *fr_iso, *fr_vel, *fr_jac, *fr_jdqd;
  double *cent_ag, *cent_hg, *cent_ig;  // centroidal momentum matrix [B,6,nv], momentum [B,6], inertia [B,36]
  // internal-constraint projection products (only when the problem has n_ic > 0)
  double *ic_cg;    // [B, nv]        Ni^T (c+g)
  double *ic_jcni;  // [B, nc, nv]    Jc Ni
  double *ic_tq;    // [B, n, nact]   (S [M | -(Jc Ni)^T])^T, transposed for coalesced per-row reads
  double *ic_tq0;   // [B, nact]      S Ni^T (c+g)
  int* ic_status;   // [B]
  double* task_e;   // [B, ntaskrows]  Jdot qdot - xddot_cmd of every task row (k_task_prepare)
  double* setup_rec;  // [B, setup_stride]  per-state record of k_ihwbc_setup (factor block, x0, D, scalars)
  int setup_stride;
};

// ---- WBC problem constants ----------------------------------------------------------
struct DevTask {
  int type, dim, ws_frame, joint_off, des_off, gain_off, row_off, dense_off;
};
// workspace frame slot of every task / contact (resolved per launch: a problem can be
// solved on any workspace that materialises its frames)
struct FrameMap {
  int task_f[PNC_MAX_TASKS];
  int contact_f[PNC_MAX_CONTACTS];
};
struct DevContact {
  int type, ws_frame, dim, rows, rf_off, cone_off;
  double mu, x, y;
};
constexpr int PNC_MAX_N = PNC_MAX_NV + 6 * PNC_MAX_CONTACTS;            // QP variables
constexpr int PNC_MAX_CONE = 18 * PNC_MAX_CONTACTS;                     // cone rows
constexpr int PNC_MAX_M = PNC_MAX_CONE + 2 * PNC_MAX_NV;                // QP inequalities
constexpr int PNC_MAX_TASK_ROWS = 96;

struct DevProblem {
  int ntask, ncontact, n_ic, nact, b_trq_limit, ndes;
  int nv, nc, nu, n, p, m;  // QP dims: n = nv + nc, p = 6 + n_ic, m = nu + 2 nact (if trq limits)
  int ndense;               // dense task rows (LINK_*, COM)
  int ntaskrows;            // sum of task dims
  double lambda_q_ddot, lambda_rf;
  DevTask tasks[PNC_MAX_TASKS];
  DevContact contacts[PNC_MAX_CONTACTS];
  int cone_contact[PNC_MAX_CONE];  // cone row -> contact
  int joint_sel[PNC_MAX_NV * 2];
  double kp[64], kd[64];
  int act_idx[PNC_MAX_NV];
  int act_of_joint[PNC_MAX_NV];  // joint (0-based, v index - 6) -> actuated row, -1 for a passive joint
  int n_trans;                   // IHWBC2 transmission rows (0 = none)
  int trans_d[PNC_MAX_IC], trans_p[PNC_MAX_IC];  // v index of the distal / passive joint of each row
  double trq_lo[PNC_MAX_NV], trq_hi[PNC_MAX_NV];
  double ic_jac[PNC_MAX_IC][PNC_MAX_NV];
};

// ---- small vector helpers -----------------------------------------------------------
struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 mk(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// row-major 3x3
__device__ __forceinline__ V3 mulR(const double* R, V3 a) {
  return V3{R[0] * a.x + R[1] * a.y + R[2] * a.z, R[3] * a.x + R[4] * a.y + R[5] * a.z,
            R[6] * a.x + R[7] * a.y + R[8] * a.z};
}
__device__ __forceinline__ V3 mulRt(const double* R, V3 a) {
  return V3{R[0] * a.x + R[3] * a.y + R[6] * a.z, R[1] * a.x + R[4] * a.y + R[7] * a.z,
            R[2] * a.x + R[5] * a.y + R[8] * a.z};
}
__device__ __forceinline__ void mul33(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// symmetric 3x3 stored xx xy xz yy yz zz
__device__ __forceinline__ V3 mulS(const double* S, V3 a) {
  return V3{S[0] * a.x + S[1] * a.y + S[2] * a.z, S[1] * a.x + S[3] * a.y + S[4] * a.z,
            S[2] * a.x + S[4] * a.y + S[5] * a.z};
}

// scalar-last quaternion -> rotation, normalising like scipy Rotation.from_quat
__device__ __forceinline__ void quat_to_rot(const double* qq, double* R) {
  double nrm = sqrt(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
  double x = qq[0] / nrm, y = qq[1] / nrm, z = qq[2] / nrm, w = qq[3] / nrm;
  double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);      R[2] = 2 * (xz + yw);
  R[3] = 2 * (xy + zw);     R[4] = -x2 + y2 - z2 + w2; R[5] = 2 * (yz - xw);
  R[6] = 2 * (xz - yw);     R[7] = 2 * (yz + xw);      R[8] = -x2 - y2 + z2 + w2;
}
// scipy Rotation.from_matrix(...).as_quat(): pick the largest of (diag, trace)
__device__ __forceinline__ void rot_to_quat(const double* m, double* q) {
  double tr = m[0] + m[4] + m[8];
  double best = m[0];
  int choice = 0;
  if (m[4] > best) { best = m[4]; choice = 1; }
  if (m[8] > best) { best = m[8]; choice = 2; }
  if (tr > best) choice = 3;
  if (choice == 3) {
    q[0] = m[7] - m[5]; q[1] = m[2] - m[6]; q[2] = m[3] - m[1]; q[3] = 1 + tr;
  } else {
    int i = choice, j = (i + 1) % 3, k = (j + 1) % 3;
    double t[4];
    t[i] = 1 - tr + 2 * m[3 * i + i];
    t[j] = m[3 * j + i] + m[3 * i + j];
    t[k] = m[3 * k + i] + m[3 * i + k];
    t[3] = m[3 * k + j] - m[3 * j + k];
    q[0] = t[0]; q[1] = t[1]; q[2] = t[2]; q[3] = t[3];
  }
  double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= nrm; q[1] /= nrm; q[2] /= nrm; q[3] /= nrm;
}

extern unsigned long long g_launch_count;  // host-side counter (capi.cu)

}  // namespace pnc
