// pnc_launch.h -- host-side launch interface between capi.cu and the kernel files.
#pragma once
#include <mutex>

#include "pnc_dev.cuh"

namespace pnc {

// shared-memory layout of ONE state's QP in the solve kernel (offsets in doubles)
struct IhwbcLayout {
  int n, ld, ldv, nrow;
  int oJ, oR, oA, oa0, oUf, ouf, ox, oz, od, onp, ohw, oxold, oxacc, og0, orr, ou, ouold, os, oe, owt, oint;
  int per_state;  // doubles, even
};

__host__ __device__ inline IhwbcLayout ihwbc_layout(const DevProblem& P) {
  IhwbcLayout L;
  const int n = P.n, m = P.m > 0 ? P.m : 1;
  L.n = n;
  L.ld = n | 1;
  L.ldv = P.nv | 1;
  L.nrow = P.p + (P.b_trq_limit ? P.nact : 0);
  int o = 0;
  L.oJ = o; o += n * L.ld;
  int rsz = n * (n + 3) / 2, stg = P.ndense * L.ldv;
  L.oR = o; o += (rsz > stg ? rsz : stg);
  L.oA = o; o += L.nrow * L.ld;
  L.oa0 = o; o += L.nrow;
  L.oUf = o; o += (P.nu > 0 ? P.nu : 1) * 6;
  L.ouf = o; o += (P.nu > 0 ? P.nu : 1);
  L.ox = o; o += n;
  L.oz = o; o += n;
  L.od = o; o += n;
  L.onp = o; o += n;
  L.ohw = o; o += n;
  L.oxold = o; o += n;
  L.oxacc = o; o += n;
  L.og0 = o; o += n;
  L.orr = o; o += n + 1;
  L.ou = o; o += n + 1;
  L.ouold = o; o += n + 1;
  L.os = o; o += m;
  L.oe = o; o += P.ntaskrows > 0 ? P.ntaskrows : 1;
  L.owt = o; o += P.ntaskrows > 0 ? P.ntaskrows : 1;
  L.oint = o; o += (3 * (n + 1) + 2 * m + 1) / 2 + 1;
  L.per_state = (o + 1) & ~1;
  return L;
}

struct SolveIO {
  const double *des, *w, *rfz;
  double *trq, *acc, *rf, *qddot;
  int *status, *iters;
  unsigned long long* active;
  unsigned long long* active_ref;  // active set at the first iterate that meets QuadProg++.cc:306-312 (may be NULL)
  double psi_refine;               // termination bound as a factor of the reference's: 1e-6 (default) or 1 (PNC_TERM_REFERENCE)
  int* counter;
  int* overflow;  // set by the register kernel when a state needs the generic kernel
};

// workspace pointers advanced by `off` states
inline WsPtrs ws_offset(const WsPtrs& a, size_t off, int nq, int nv, int nfr, int nc, int nact, int n, int ntaskrows = 0) {
  WsPtrs w = a;
  w.q += off * nq; w.qdot += off * nv; w.M += off * nv * nv; w.grav += off * nv; w.cori += off * nv;
  w.com += off * 3; w.vcom += off * 3; w.jcom += off * 3 * nv; w.jcomdot += off * 3 * nv; w.jcomdot_qdot += off * 3;
  w.cent_ag += off * 6 * nv; w.cent_hg += off * 6; w.cent_ig += off * 36;
  w.fr_iso += off * nfr * 12; w.fr_vel += off * nfr * 6; w.fr_jac += off * nfr * 6 * nv; w.fr_jdqd += off * nfr * 6;
  if (w.ic_cg) {
    w.ic_cg += off * nv; w.ic_jcni += off * nc * nv; w.ic_tq += off * nact * n; w.ic_tq0 += off * nact;
    w.ic_status += off;
  }
  if (w.task_e) w.task_e += off * ntaskrows;
  if (w.setup_rec) w.setup_rec += off * (size_t)w.setup_stride;
  return w;
}

// (dof, ancestor-or-self dof) pairs of the mass matrix for k_update_system's CRBA phase: x = offset of the
// descendant dof's force vector | offset of the ancestor dof's motion vector << 16 (elements, in the warp's
// shared-memory region), y = row | column << 8.  Returns the number of pairs written (<= max_pairs).
int build_crba_pairs(const DevModel& h, int2* pairs, int max_pairs);
// per dof (nv entries): x = offset of its motion vector in the warp's shared-memory region, [y, z) = bodies of its subtree
void build_dof_table(const DevModel& h, int4* tab);
void launch_update_system(const DevModel* dm, const DevFrames* dfr, const int4* dof_tab, const int2* crba_pairs, int n_pairs,
                          const WsPtrs& ws, int B, int nv, const double* base_pos, const double* base_quat, const double* base_lin_vel,
                          const double* base_ang_vel, const double* joint_pos, const double* joint_vel,
                          int num_sms, cudaStream_t stream, int b_cent = 0, int fp32 = 0);
void launch_ihwbc_solve(const DevProblem* dp, const DevProblem& hp, const IhwbcLayout& L, const WsPtrs& ws,
                        const FrameMap& fm, int nfr, int B, const SolveIO& io, int num_sms, cudaStream_t stream);
bool launch_ihwbc_solve_reg(const DevProblem* dp, const DevProblem& hp, const WsPtrs& ws, const FrameMap& fm, int nfr,
                            int B, const SolveIO& io, int num_sms, cudaStream_t stream);
void launch_task_prepare(const DevProblem* dp, const DevProblem& hp, const WsPtrs& ws, const FrameMap& fm, int nfr, int B,
                         const double* des, cudaStream_t stream);
bool reg_kernel_available(const DevProblem& hp);
int reg_setup_record_doubles(const DevProblem& hp);
int reg_shape_id(const DevProblem& hp);
bool reg_record_layout(const DevProblem& hp, int* layout10);
void launch_ic_prepare(const DevProblem* dp, const DevProblem& hp, const WsPtrs& ws, const FrameMap& fm, int nfr,
                       int B, int num_sms, cudaStream_t stream, int* counter);
void launch_task_eval(const DevProblem* dp, const WsPtrs& ws, const FrameMap& fm, int nfr, int B, int itask, int dim,
                      int nv, const double* des, double* jac, double* jdqd, double* op_cmd, double* pos_err,
                      cudaStream_t stream);
void launch_contact_eval(const DevProblem* dp, const WsPtrs& ws, const FrameMap& fm, int nfr, int B, int icontact,
                         int nv, const double* rfz, double* jac, double* cone_mat, double* cone_vec,
                         cudaStream_t stream);
void launch_joint_integrate(int B, int na, const double* acc, const double* joint_pos, double* vel_state,
                            double* pos_state, const double* pos_limit, const double* vel_limit, double alpha_vel,
                            double alpha_pos, double dt, cudaStream_t stream);
void launch_command_epilogue(const DevProblem* dp, int B, int na, const double* trq, const double* acc,
                             const double* joint_pos, double* vel_state, double* pos_state, const double* pos_limit,
                             const double* vel_limit, double alpha_vel, double alpha_pos, double dt,
                             double* joint_trq_cmd, double* joint_vel_cmd, double* joint_pos_cmd, cudaStream_t stream);
bool launch_quadprog_dense(int n, int p, int m, int B, const double* G, const double* g0, const double* CE, const double* ce0,
                           const double* CI, const double* ci0, double* x, double* f, int* status, int* iters,
                           unsigned long long* active, int num_sms, cudaStream_t stream);
// The opt-in dynamic shared-memory limit is an attribute of (kernel function, device): remember the
// high-water mark per pair.  Keyed by the function address, not by the static object at the call
// site: template instantiations of one kernel share a pointer type and would share that object.
struct SmemOptIn {
  template <class K>
  bool ensure(K kernel, size_t smem) {
    struct Entry { const void* fn; int dev; size_t bytes; };
    static Entry table[256];
    static int used = 0;
    static std::mutex mu;  // host threads driving different handles (one per GPU) may launch at the same time
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const void* fn = reinterpret_cast<const void*>(kernel);
    Entry* e = nullptr;
    for (int i = 0; i < used; i++)
      if (table[i].fn == fn && table[i].dev == dev) { e = &table[i]; break; }
    if (!e) {
      if (used >= 256) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
      e = &table[used++];
      e->fn = fn; e->dev = dev; e->bytes = 0;
    }
    if (smem > e->bytes) {
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
      e->bytes = smem;
    }
    return true;
  }
};

// managers.cu
void launch_ramp_update(int B, const double* sv, double target, const double* st, const double* dur, double t, double* out,
                        cudaStream_t stream);
void launch_foot_use_current(const WsPtrs& ws, int nfr, int slot, int B, double* pos, double* vel, double* acc, double* quat,
                             double* angvel, double* angacc, cudaStream_t stream);
void launch_hand_init(const WsPtrs& ws, int nfr, int slot, int B, const double* tpos, const double* tquat, const double* key,
                      double* rec, cudaStream_t stream);
void launch_hand_eval(int B, const double* rec, const double* st, const double* dur, double t, int keypoint, double* pos,
                      double* vel, double* acc, double* quat, double* angvel, double* angacc, cudaStream_t stream);
void launch_swing_init(const WsPtrs& ws, int nfr, int slot, int B, const double* land_pos, const double* land_quat,
                       const double* duration, double height, double* rec, cudaStream_t stream);
void launch_swing_eval(int B, const double* rec, const double* st, const double* dur, double t, double* pos, double* vel,
                       double* acc, double* quat, double* angvel, double* angacc, cudaStream_t stream);
void launch_floating_base_init(const WsPtrs& ws, int nfr, int slot, int B, const double* tcom, const double* tquat,
                               double* rec, cudaStream_t stream);
void launch_floating_base_eval(int B, const double* rec, const double* st, const double* dur, double t, double* pos,
                               double* vel, double* acc, double* quat, double* angvel, double* angacc,
                               cudaStream_t stream);
void launch_sinusoid_eval(int B, const double* mid, const double* st, const double* amp3, const double* freq3, double t,
                          double* pos, double* vel, double* acc, cudaStream_t stream);
void launch_dcm_update(const WsPtrs& ws, int B, double dt, double alpha, double* dcm, double* prev, double* dcm_vel,
                       cudaStream_t stream);
void launch_osc_step(const WsPtrs& ws, int nfr, int slot, int nv, int B, const double* pos_des, const double* vel_des,
                     const double* kp, const double* kd, double smooth, double rcond, double* trq, double* qddot,
                     cudaStream_t stream);
int measure_fp64_peak(int num_sms, double* tflops);

}  // namespace pnc
