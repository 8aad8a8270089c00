// capi.cu -- the C ABI of include/pnc_b200.h: handles, HBM workspace, kernel launches.
//
// Host-side only plumbing; the arithmetic lives in dynamics.cu / ihwbc.cu / icprep.cu.
// There is no CPU fallback: every compute entry point launches a CUDA kernel or fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "pnc_dev.cuh"
#include "pnc_launch.h"

namespace pnc {
unsigned long long g_launch_count = 0;
}
using namespace pnc;

// PNC_TERM_REFINED (default): keep iterating to 1e-6 x the bound of QuadProg++.cc:306-312;
// PNC_TERM_REFERENCE: stop at the reference's own bound.  PNC_TERMINATION=reference sets the default.
// PNC_PREC_FP64 (default) / PNC_PREC_FP32_DYNAMICS: the optional fp32 mode of the north star -- k_update_system in
// fp32 arithmetic, factorisations and active-set iterations fp64.  PNC_PRECISION=fp32 sets the default.
static int g_precision = [] {
  const char* e = getenv("PNC_PRECISION");
  return (e && (e[0] == 'f' || e[0] == 'F') && e[1] && e[2] == '3' && e[3] == '2') ? PNC_PREC_FP32_DYNAMICS : PNC_PREC_FP64;
}();
static int g_termination = [] {
  const char* e = getenv("PNC_TERMINATION");
  return (e && (e[0] == 'r' || e[0] == 'R') && (e[2] == 'f' || e[2] == 'F') && (e[3] == 'e' || e[3] == 'E')) ? PNC_TERM_REFERENCE
                                                                                                             : PNC_TERM_REFINED;
}();

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      fprintf(stderr, "pnc_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return PNC_E_CUDA;                                                                          \
    }                                                                                             \
  } while (0)

// like CK, but runs `cleanup` first (frees a half-built handle)
#define CKF(call, cleanup)                                                                        \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      fprintf(stderr, "pnc_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      cleanup;                                                                                    \
      return PNC_E_CUDA;                                                                          \
    }                                                                                             \
  } while (0)

// Every entry point runs on the device of its handle (or of its first device pointer) and leaves the caller's
// current device as it found it: a host program that holds models on several GPUs -- or a torch program, whose
// allocations follow the current device -- must not find it changed behind its back.
struct DevGuard {
  int prev = -1;
  bool ok = false;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
    if (prev == dev) { prev = -1; ok = true; return; }
    ok = cudaSetDevice(dev) == cudaSuccess;
    if (!ok) prev = -1;
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DevGuard(const DevGuard&) = delete;
  DevGuard& operator=(const DevGuard&) = delete;
};
#define ON_DEVICE(dev)        \
  DevGuard dev_guard_(dev);   \
  if (!dev_guard_.ok) return PNC_E_CUDA

// The stateless entry points (no handle) launch on the device that owns their first pointer.
static int device_of(const void* dptr, int* dev) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, dptr) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
    cudaGetLastError();
    return PNC_E_ARG;  // not a device pointer
  }
  *dev = at.device;
  return PNC_OK;
}
#define SETDEV(ptr)                         \
  int dev_of_ptr_ = 0;                      \
  {                                         \
    const int rc_ = device_of(ptr, &dev_of_ptr_); \
    if (rc_ != PNC_OK) return rc_;          \
  }                                         \
  ON_DEVICE(dev_of_ptr_)

struct pnc_model {
  int device, num_sms;
  DevModel h;
  DevModel* d;
  int4* d_dofs;   // per-dof table and CRBA (dof, ancestor dof) pair table of k_update_system (one allocation)
  int2* d_pairs;
  int n_pairs;
  int nf;
  std::vector<int> frame_parent;
  std::vector<double> frame_R, frame_p;
};

struct pnc_problem {
  const pnc_model* m;
  int device;  // copy of m->device (see pnc_workspace)
  DevProblem h;
  DevProblem* d;
  int task_model_frame[PNC_MAX_TASKS];
  int contact_model_frame[PNC_MAX_CONTACTS];
  IhwbcLayout lay;
};

struct pnc_workspace {
  const pnc_model* m;
  int device;  // copy of m->device: destroy must not touch the model, which may already be gone
  int cap, nfr;
  DevFrames hfr;
  DevFrames* dfr;
  WsPtrs p;
  char* slab;
  size_t bytes;
  int* counter;  // work-distribution counter of the solve kernel
  // internal-constraint scratch (allocated lazily for problems with n_ic > 0)
  char* ic_slab;
  size_t ic_bytes;
  int ic_nc, ic_nact;
  // task error terms written by k_task_prepare (allocated lazily, sized by the largest problem seen)
  double* task_slab;
  int task_rows;
  double* setup_slab;
  int setup_stride, setup_shape;
  // staging buffers of pnc_wbc_tick_host
  char* stage;
  size_t stage_bytes;
  cudaStream_t tick_stream[2];          // compute, chunks alternate
  cudaStream_t tick_in, tick_out;       // host -> device and device -> host copies
  cudaEvent_t tick_ev_in[16], tick_ev_done[16];  // per chunk: inputs landed, kernels finished
  cudaEvent_t tick_ev_ready;                     // producer of device-resident desireds done
  cudaEvent_t tick_ev_start;                     // first input copy of the tick enqueued (timing)
  int tick_last;                                 // index of the last chunk of the tick in flight (-1: none)
  double tick_growth;                            // chunk growth factor measured on the previous tick (0: not yet)
  bool tick_init;
};

extern "C" {

const char* pnc_version(void) { return "pypnc_b200 0.1 (sm_100a)"; }
int64_t pnc_launch_count(void) { return (int64_t)g_launch_count; }

int pnc_model_create(const pnc_model_desc* dsc, int device, pnc_model** out) {
  if (!dsc || !out) return PNC_E_ARG;
  if (dsc->nb < 1 || dsc->nb > PNC_MAX_BODIES || dsc->nv > PNC_MAX_NV) return PNC_E_SIZE;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PNC_E_NOGPU;
  if (device < 0 || device >= ndev) return PNC_E_ARG;
  ON_DEVICE(device);
  pnc_model* m = new pnc_model();
  m->device = device;
  cudaDeviceProp prop;
  CKF(cudaGetDeviceProperties(&prop, device), delete m);
  m->num_sms = prop.multiProcessorCount;
  DevModel& h = m->h;
  memset(&h, 0, sizeof(h));
  h.nb = dsc->nb; h.nq = dsc->nq; h.nv = dsc->nv; h.na = dsc->na; h.nfl = dsc->n_floating; h.nf = dsc->nf;
  h.total_mass = dsc->total_mass; h.mass_dyn = dsc->mass_dyn;
  int maxdepth = 0;
  for (int i = 0; i < h.nb; i++) {
    h.parent[i] = dsc->parent[i];
    if (h.parent[i] >= i) { delete m; return PNC_E_ARG; }
    h.jtype[i] = dsc->jtype[i];
    h.idx_v[i] = dsc->idx_v[i];
    h.subtree[i] = dsc->subtree[i];
    h.depth[i] = h.parent[i] < 0 ? 0 : h.depth[h.parent[i]] + 1;
    if (h.depth[i] > maxdepth) maxdepth = h.depth[i];
    for (int k = 0; k < 9; k++) h.place_R[i][k] = dsc->place_R[9 * i + k];
    for (int k = 0; k < 3; k++) {
      h.place_p[i][k] = dsc->place_p[3 * i + k];
      h.axis[i][k] = dsc->axis[3 * i + k];
      h.com[i][k] = dsc->com[3 * i + k];
    }
    h.mass[i] = dsc->mass[i];
    const double* I = dsc->inertia + 9 * i;
    h.inertia[i][0] = I[0]; h.inertia[i][1] = I[1]; h.inertia[i][2] = I[2];
    h.inertia[i][3] = I[4]; h.inertia[i][4] = I[5]; h.inertia[i][5] = I[8];
  }
  h.maxdepth = maxdepth;
  m->nf = dsc->nf;
  m->frame_parent.assign(dsc->frame_parent, dsc->frame_parent + dsc->nf);
  m->frame_R.assign(dsc->frame_R, dsc->frame_R + 9 * (size_t)dsc->nf);
  m->frame_p.assign(dsc->frame_p, dsc->frame_p + 3 * (size_t)dsc->nf);
  m->d = nullptr;
  CKF(cudaMalloc(&m->d, sizeof(DevModel)), delete m);
  CKF(cudaMemcpy(m->d, &h, sizeof(DevModel), cudaMemcpyHostToDevice), (cudaFree(m->d), delete m));
  {
    std::vector<int2> pairs((size_t)h.nv * (h.nv + 1) / 2 + 64);
    std::vector<int4> dofs((size_t)h.nv);
    m->n_pairs = build_crba_pairs(h, pairs.data(), (int)pairs.size());
    build_dof_table(h, dofs.data());
    if (m->n_pairs < 0) { cudaFree(m->d); delete m; return PNC_E_SIZE; }
    const size_t bd = sizeof(int4) * dofs.size(), bp = sizeof(int2) * (size_t)(m->n_pairs > 0 ? m->n_pairs : 1);
    m->d_dofs = nullptr;
    CKF(cudaMalloc(&m->d_dofs, bd + bp), (cudaFree(m->d), delete m));
    m->d_pairs = reinterpret_cast<int2*>(reinterpret_cast<char*>(m->d_dofs) + bd);
    CKF(cudaMemcpy(m->d_dofs, dofs.data(), bd, cudaMemcpyHostToDevice), (cudaFree(m->d_dofs), cudaFree(m->d), delete m));
    CKF(cudaMemcpy(m->d_pairs, pairs.data(), sizeof(int2) * (size_t)m->n_pairs, cudaMemcpyHostToDevice),
        (cudaFree(m->d_dofs), cudaFree(m->d), delete m));
  }
  *out = m;
  return PNC_OK;
}

void pnc_model_destroy(pnc_model* m) {
  if (!m) return;
  DevGuard dev_guard_(m->device);
  cudaFree(m->d);
  cudaFree(m->d_dofs);
  cudaGetLastError();
  delete m;
}

// ---------------------------------------------------------------------------------------
int pnc_problem_create(const pnc_model* m, const pnc_problem_desc* dsc, pnc_problem** out) {
  if (!m || !dsc || !out) return PNC_E_ARG;
  if (dsc->ntask > PNC_MAX_TASKS || dsc->ncontact > PNC_MAX_CONTACTS || dsc->n_ic > PNC_MAX_IC ||
      dsc->nact > PNC_MAX_NV || dsc->n_gain > 64 || dsc->n_joint_sel > 2 * PNC_MAX_NV)
    return PNC_E_SIZE;
  if (m->h.nfl != 6) return PNC_E_ARG;  // IHWBC needs the floating base rows (Sf), ihwbc.py:225
  ON_DEVICE(m->device);
  pnc_problem* p = new pnc_problem();
  p->m = m;
  p->device = m->device;
  DevProblem& h = p->h;
  memset(&h, 0, sizeof(h));
  h.ntask = dsc->ntask; h.ncontact = dsc->ncontact; h.n_ic = dsc->n_ic; h.nact = dsc->nact;
  h.b_trq_limit = dsc->b_trq_limit ? 1 : 0; h.ndes = dsc->ndes;
  h.nv = m->h.nv;
  h.lambda_q_ddot = dsc->lambda_q_ddot; h.lambda_rf = dsc->lambda_rf;
  int row = 0, dense = 0;
  for (int i = 0; i < dsc->ntask; i++) {
    const pnc_task_desc& t = dsc->tasks[i];
    DevTask& o = h.tasks[i];
    o.type = t.type; o.dim = t.dim; o.ws_frame = -1; o.joint_off = t.joint_off; o.des_off = t.des_off;
    o.gain_off = t.gain_off; o.row_off = row; o.dense_off = -1;
    p->task_model_frame[i] = t.frame;
    bool is_dense = t.type == PNC_TASK_LINK_XYZ || t.type == PNC_TASK_LINK_ORI || t.type == PNC_TASK_COM;
    if (is_dense) {
      if (t.dim != 3) { delete p; return PNC_E_ARG; }
      if (t.type != PNC_TASK_COM && (t.frame < 0 || t.frame >= m->nf)) { delete p; return PNC_E_ARG; }
      o.dense_off = dense;
      dense += 3;
    } else if (t.type == PNC_TASK_JOINT) {
      if (t.dim != m->h.na) { delete p; return PNC_E_ARG; }
    } else if (t.type == PNC_TASK_SELECTED_JOINT) {
      if (t.joint_off < 0 || t.joint_off + t.dim > dsc->n_joint_sel) { delete p; return PNC_E_ARG; }
    } else {
      delete p;
      return PNC_E_ARG;  // basic_task.py:105 raises ValueError on an unknown task type
    }
    row += t.dim;
  }
  if (row > PNC_MAX_TASK_ROWS) { delete p; return PNC_E_SIZE; }
  h.ntaskrows = row;
  h.ndense = dense;
  int nc = 0, nu = 0;
  for (int i = 0; i < dsc->ncontact; i++) {
    const pnc_contact_desc& c = dsc->contacts[i];
    DevContact& o = h.contacts[i];
    if (c.frame < 0 || c.frame >= m->nf) { delete p; return PNC_E_ARG; }
    o.type = c.type; o.ws_frame = -1;
    o.dim = c.type == PNC_CONTACT_SURFACE ? 6 : 3;
    o.rows = c.type == PNC_CONTACT_SURFACE ? 18 : 6;
    o.rf_off = nc; o.cone_off = nu;
    o.mu = c.mu; o.x = c.x; o.y = c.y;
    p->contact_model_frame[i] = c.frame;
    for (int r = 0; r < o.rows; r++) h.cone_contact[nu + r] = i;
    nc += o.dim; nu += o.rows;
  }
  h.nc = nc; h.nu = nu;
  h.n = h.nv + nc;
  h.n_trans = dsc->n_trans;
  if (h.n_trans < 0 || h.n_trans > PNC_MAX_IC || (h.n_trans > 0 && (h.n_ic > 0 || h.b_trq_limit))) { delete p; return PNC_E_ARG; }
  for (int i = 0; i < h.n_trans; i++) {
    h.trans_d[i] = dsc->trans_distal[i];
    h.trans_p[i] = dsc->trans_passive[i];
    if (h.trans_d[i] < 6 || h.trans_d[i] >= h.nv || h.trans_p[i] < 6 || h.trans_p[i] >= h.nv) { delete p; return PNC_E_ARG; }
  }
  h.p = 6 + h.n_ic + h.n_trans;
  h.m = nu + (h.b_trq_limit ? 2 * h.nact : 0);
  if (h.n > PNC_MAX_N || h.m > PNC_MAX_M) { delete p; return PNC_E_SIZE; }
  for (int i = 0; i < dsc->n_joint_sel; i++) h.joint_sel[i] = dsc->joint_sel[i];
  for (int i = 0; i < dsc->n_gain; i++) { h.kp[i] = dsc->kp[i]; h.kd[i] = dsc->kd[i]; }
  for (int i = 0; i < dsc->nact; i++) {
    h.act_idx[i] = dsc->act_idx[i];
    if (h.act_idx[i] < 6 || h.act_idx[i] >= h.nv) { delete p; return PNC_E_ARG; }
    if (dsc->b_trq_limit) { h.trq_lo[i] = dsc->trq_limit[2 * i]; h.trq_hi[i] = dsc->trq_limit[2 * i + 1]; }
  }
  for (int j = 0; j < PNC_MAX_NV; j++) h.act_of_joint[j] = -1;
  for (int i = 0; i < dsc->nact; i++) h.act_of_joint[h.act_idx[i] - 6] = i;
  for (int i = 0; i < dsc->n_ic; i++)
    for (int j = 0; j < h.nv; j++) h.ic_jac[i][j] = dsc->ic_jac[i * h.nv + j];
  p->lay = ihwbc_layout(h);
  p->d = nullptr;
  CKF(cudaMalloc(&p->d, sizeof(DevProblem)), delete p);
  CKF(cudaMemcpy(p->d, &h, sizeof(DevProblem), cudaMemcpyHostToDevice), (cudaFree(p->d), delete p));
  *out = p;
  return PNC_OK;
}

void pnc_problem_destroy(pnc_problem* p) {
  if (!p) return;
  DevGuard dev_guard_(p->device);
  cudaFree(p->d);
  cudaGetLastError();  // a destroy during interpreter teardown must not leave a stale error behind
  delete p;
}

int pnc_problem_dims(const pnc_problem* p, int32_t* nc, int32_t* ncone, int32_t* n, int32_t* p_eq, int32_t* m_ineq) {
  if (!p) return PNC_E_ARG;
  if (nc) *nc = p->h.nc;
  if (ncone) *ncone = p->h.nu;
  if (n) *n = p->h.n;
  if (p_eq) *p_eq = p->h.p;
  if (m_ineq) *m_ineq = p->h.m;
  return PNC_OK;
}

// ---------------------------------------------------------------------------------------
static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

int pnc_workspace_create(const pnc_model* m, int32_t capacity, const int32_t* frames, int32_t nframes,
                         pnc_workspace** out) {
  if (!m || !out || capacity < 1 || nframes < 0 || (nframes > 0 && !frames)) return PNC_E_ARG;
  if (nframes > PNC_MAX_FRAMES) return PNC_E_SIZE;
  ON_DEVICE(m->device);
  pnc_workspace* ws = new pnc_workspace();
  memset(&ws->p, 0, sizeof(ws->p));
  ws->m = m; ws->device = m->device; ws->cap = capacity; ws->nfr = nframes;
  ws->ic_slab = nullptr; ws->ic_bytes = 0; ws->ic_nc = ws->ic_nact = 0;
  ws->task_slab = nullptr; ws->task_rows = 0;
  ws->setup_slab = nullptr; ws->setup_stride = 0; ws->setup_shape = 0;
  ws->stage = nullptr; ws->stage_bytes = 0; ws->tick_init = false; ws->tick_last = -1; ws->tick_growth = 0.0;
  memset(&ws->hfr, 0, sizeof(ws->hfr));
  ws->hfr.nfr = nframes;
  for (int f = 0; f < nframes; f++) {
    int mf = frames[f];
    if (mf < 0 || mf >= m->nf) { delete ws; return PNC_E_ARG; }
    ws->hfr.model_frame[f] = mf;
    ws->hfr.body[f] = m->frame_parent[mf];
    for (int k = 0; k < 9; k++) ws->hfr.R[f][k] = m->frame_R[9 * (size_t)mf + k];
    for (int k = 0; k < 3; k++) ws->hfr.p[f][k] = m->frame_p[3 * (size_t)mf + k];
  }
  const size_t B = capacity, nv = m->h.nv, nq = m->h.nq, nfr = nframes;
  size_t sizes[17] = {B * nq, B * nv, B * nv * nv, B * nv, B * nv, B * 3, B * 3, B * 3 * nv, B * 3 * nv, B * 3,
                      B * nfr * 12, B * nfr * 6, B * nfr * 6 * nv, B * nfr * 6, B * 6 * nv, B * 6, B * 36};
  size_t total = 0, off[17];
  for (int i = 0; i < 17; i++) { off[i] = total; total += align_up(sizes[i] * sizeof(double)); }
  total += 256;  // counter
  ws->slab = nullptr; ws->dfr = nullptr;
  CKF(cudaMalloc(&ws->slab, total), delete ws);
  CKF(cudaMemset(ws->slab, 0, total), (cudaFree(ws->slab), delete ws));
  ws->bytes = total;
  double** dst[17] = {&ws->p.q, &ws->p.qdot, &ws->p.M, &ws->p.grav, &ws->p.cori, &ws->p.com, &ws->p.vcom,
                      &ws->p.jcom, &ws->p.jcomdot, &ws->p.jcomdot_qdot, &ws->p.fr_iso, &ws->p.fr_vel,
                      &ws->p.fr_jac, &ws->p.fr_jdqd, &ws->p.cent_ag, &ws->p.cent_hg, &ws->p.cent_ig};
  for (int i = 0; i < 17; i++) *dst[i] = reinterpret_cast<double*>(ws->slab + off[i]);
  ws->counter = reinterpret_cast<int*>(ws->slab + total - 256);
  CKF(cudaMalloc(&ws->dfr, sizeof(DevFrames)), (cudaFree(ws->slab), delete ws));
  CKF(cudaMemcpy(ws->dfr, &ws->hfr, sizeof(DevFrames), cudaMemcpyHostToDevice),
      (cudaFree(ws->dfr), cudaFree(ws->slab), delete ws));
  *out = ws;
  return PNC_OK;
}

void pnc_workspace_destroy(pnc_workspace* ws) {
  if (!ws) return;
  DevGuard dev_guard_(ws->device);
  cudaFree(ws->slab);
  cudaFree(ws->dfr);
  if (ws->ic_slab) cudaFree(ws->ic_slab);
  if (ws->task_slab) cudaFree(ws->task_slab);
  if (ws->setup_slab) cudaFree(ws->setup_slab);
  if (ws->stage) cudaFree(ws->stage);
  if (ws->tick_init) {
    for (int i = 0; i < 2; i++) cudaStreamDestroy(ws->tick_stream[i]);
    cudaStreamDestroy(ws->tick_in); cudaStreamDestroy(ws->tick_out);
    for (int i = 0; i < 16; i++) { cudaEventDestroy(ws->tick_ev_in[i]); cudaEventDestroy(ws->tick_ev_done[i]); }
    cudaEventDestroy(ws->tick_ev_ready); cudaEventDestroy(ws->tick_ev_start);
  }
  cudaGetLastError();
  delete ws;
}

int64_t pnc_workspace_bytes(const pnc_workspace* ws) {
  return ws ? (int64_t)(ws->bytes + ws->ic_bytes + ws->stage_bytes + (size_t)ws->cap * ((size_t)ws->task_rows + ws->setup_stride) * 8) : 0;
}

static int update_system_at(pnc_workspace* ws, int off, int32_t B, const double* bp, const double* bq,
                            const double* blv, const double* bav, const double* jp, const double* jv,
                            cudaStream_t st, int b_cent = 0) {
  const pnc_model* m = ws->m;
  if (m->h.nfl && (!bp || !bq || !blv || !bav)) return PNC_E_ARG;
  WsPtrs w = ws_offset(ws->p, off, m->h.nq, m->h.nv, ws->nfr, 0, 0, 0);
  launch_update_system(m->d, ws->dfr, m->d_dofs, m->d_pairs, m->n_pairs, w, B, m->h.nv, bp, bq, blv, bav, jp, jv, m->num_sms, st, b_cent,
                       g_precision == PNC_PREC_FP32_DYNAMICS);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_update_system(pnc_workspace* ws, int32_t B, const double* d_base_pos, const double* d_base_quat,
                      const double* d_base_lin_vel, const double* d_base_ang_vel, const double* d_joint_pos,
                      const double* d_joint_vel, void* stream) {
  if (!ws || B < 0 || B > ws->cap || !d_joint_pos || !d_joint_vel) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  return update_system_at(ws, 0, B, d_base_pos, d_base_quat, d_base_lin_vel, d_base_ang_vel, d_joint_pos,
                          d_joint_vel, (cudaStream_t)stream);
}

int pnc_update_system_cent(pnc_workspace* ws, int32_t B, const double* d_base_pos, const double* d_base_quat,
                           const double* d_base_lin_vel, const double* d_base_ang_vel, const double* d_joint_pos,
                           const double* d_joint_vel, int32_t b_cent, void* stream) {
  if (!ws || B < 0 || B > ws->cap || !d_joint_pos || !d_joint_vel) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  return update_system_at(ws, 0, B, d_base_pos, d_base_quat, d_base_lin_vel, d_base_ang_vel, d_joint_pos,
                          d_joint_vel, (cudaStream_t)stream, b_cent ? 1 : 0);
}

double* pnc_ws_centroidal_ag(pnc_workspace* ws) { return ws->p.cent_ag; }
double* pnc_ws_centroidal_hg(pnc_workspace* ws) { return ws->p.cent_hg; }
double* pnc_ws_centroidal_ig(pnc_workspace* ws) { return ws->p.cent_ig; }
double* pnc_ws_q(pnc_workspace* ws) { return ws->p.q; }
double* pnc_ws_qdot(pnc_workspace* ws) { return ws->p.qdot; }
double* pnc_ws_mass_matrix(pnc_workspace* ws) { return ws->p.M; }
double* pnc_ws_gravity(pnc_workspace* ws) { return ws->p.grav; }
double* pnc_ws_coriolis(pnc_workspace* ws) { return ws->p.cori; }
double* pnc_ws_com_pos(pnc_workspace* ws) { return ws->p.com; }
double* pnc_ws_com_vel(pnc_workspace* ws) { return ws->p.vcom; }
double* pnc_ws_com_jac(pnc_workspace* ws) { return ws->p.jcom; }
double* pnc_ws_com_jacdot(pnc_workspace* ws) { return ws->p.jcomdot; }
double* pnc_ws_com_jdqd(pnc_workspace* ws) { return ws->p.jcomdot_qdot; }
double* pnc_ws_frame_iso(pnc_workspace* ws) { return ws->p.fr_iso; }
double* pnc_ws_frame_vel(pnc_workspace* ws) { return ws->p.fr_vel; }
double* pnc_ws_frame_jac(pnc_workspace* ws) { return ws->p.fr_jac; }
double* pnc_ws_frame_jdqd(pnc_workspace* ws) { return ws->p.fr_jdqd; }

// ---------------------------------------------------------------------------------------
static int resolve_frames(const pnc_problem* p, const pnc_workspace* ws, FrameMap* fm) {
  for (int i = 0; i < p->h.ntask; i++) {
    fm->task_f[i] = -1;
    int mf = p->task_model_frame[i];
    if (p->h.tasks[i].type == PNC_TASK_LINK_XYZ || p->h.tasks[i].type == PNC_TASK_LINK_ORI) {
      for (int f = 0; f < ws->nfr; f++)
        if (ws->hfr.model_frame[f] == mf) fm->task_f[i] = f;
      if (fm->task_f[i] < 0) return PNC_E_ARG;
    }
  }
  for (int i = 0; i < p->h.ncontact; i++) {
    fm->contact_f[i] = -1;
    for (int f = 0; f < ws->nfr; f++)
      if (ws->hfr.model_frame[f] == p->contact_model_frame[i]) fm->contact_f[i] = f;
    if (fm->contact_f[i] < 0) return PNC_E_ARG;
  }
  return PNC_OK;
}

static int ensure_task_buf(pnc_workspace* ws, const pnc_problem* p) {
  if (ws->task_slab && ws->task_rows >= p->h.ntaskrows) return PNC_OK;
  if (ws->task_slab) { cudaFree(ws->task_slab); ws->task_slab = nullptr; }
  const size_t bytes = (size_t)ws->cap * (p->h.ntaskrows > 0 ? p->h.ntaskrows : 1) * sizeof(double);
  CK(cudaMalloc(&ws->task_slab, bytes));
  ws->task_rows = p->h.ntaskrows;
  ws->p.task_e = ws->task_slab;
  return PNC_OK;
}

static int ensure_setup_buf(pnc_workspace* ws, const pnc_problem* p) {
  const int rec = reg_setup_record_doubles(p->h);
  const int shape = reg_shape_id(p->h);
  if (ws->setup_slab && ws->setup_stride == rec && ws->setup_shape == shape) return PNC_OK;
  if (ws->setup_slab) { cudaFree(ws->setup_slab); ws->setup_slab = nullptr; }
  CK(cudaMalloc(&ws->setup_slab, (size_t)ws->cap * rec * sizeof(double)));
  // the setup kernel only writes the live part of each record (padding columns and unused lanes stay zero)
  CK(cudaMemset(ws->setup_slab, 0, (size_t)ws->cap * rec * sizeof(double)));
  ws->setup_stride = rec;
  ws->setup_shape = shape;
  return PNC_OK;
}

static int ensure_ic(pnc_workspace* ws, const pnc_problem* p) {
  if (p->h.n_ic == 0) return PNC_OK;
  if (ws->ic_slab && ws->ic_nc == p->h.nc && ws->ic_nact == p->h.nact) return PNC_OK;
  if (ws->ic_slab) { cudaFree(ws->ic_slab); ws->ic_slab = nullptr; }
  const size_t B = ws->cap, nv = p->h.nv, nc = p->h.nc, nact = p->h.nact, n = p->h.n;
  size_t sizes[4] = {B * nv, B * nc * nv, B * nact * n, B * nact};
  size_t total = 0, off[5];
  for (int i = 0; i < 4; i++) { off[i] = total; total += align_up(sizes[i] * sizeof(double)); }
  off[4] = total;
  total += align_up(B * sizeof(int));
  CK(cudaMalloc(&ws->ic_slab, total));
  CK(cudaMemset(ws->ic_slab, 0, total));
  ws->ic_bytes = total;
  ws->p.ic_cg = reinterpret_cast<double*>(ws->ic_slab + off[0]);
  ws->p.ic_jcni = reinterpret_cast<double*>(ws->ic_slab + off[1]);
  ws->p.ic_tq = reinterpret_cast<double*>(ws->ic_slab + off[2]);
  ws->p.ic_tq0 = reinterpret_cast<double*>(ws->ic_slab + off[3]);
  ws->p.ic_status = reinterpret_cast<int*>(ws->ic_slab + off[4]);
  ws->ic_nc = p->h.nc; ws->ic_nact = p->h.nact;
  return PNC_OK;
}

static int solve_at(const pnc_problem* p, pnc_workspace* ws, int off, int32_t B, const double* des,
                    const double* w, const double* rfz, double* trq, double* acc, double* rf, double* qddot,
                    int32_t* status, int32_t* iters, uint64_t* active, cudaStream_t st, int* counter,
                    uint64_t* active_ref = nullptr) {
  FrameMap fm;
  int rc = resolve_frames(p, ws, &fm);
  if (rc) return rc;
  rc = ensure_ic(ws, p);
  if (rc) return rc;
  const bool use_reg = reg_kernel_available(p->h);
  if (use_reg) {
    rc = ensure_task_buf(ws, p);
    if (rc) return rc;
  } else {
    ws->p.task_e = nullptr;  // the generic kernel evaluates the tasks itself
  }
  if (use_reg) {
    rc = ensure_setup_buf(ws, p);
    if (rc) return rc;
  }
  ws->p.task_e = use_reg ? ws->task_slab : nullptr;
  ws->p.setup_rec = use_reg ? ws->setup_slab : nullptr;
  ws->p.setup_stride = use_reg ? ws->setup_stride : 0;
  WsPtrs wsp = ws_offset(ws->p, off, ws->m->h.nq, p->h.nv, ws->nfr, p->h.nc, p->h.nact, p->h.n, p->h.ntaskrows);
  if (use_reg) {
    launch_task_prepare(p->d, p->h, wsp, fm, ws->nfr, B, des, st);
    CK(cudaGetLastError());
  }
  if (p->h.n_ic > 0) {
    launch_ic_prepare(p->d, p->h, wsp, fm, ws->nfr, B, ws->m->num_sms, st, counter + 1);
    CK(cudaGetLastError());
  }
  CK(cudaMemsetAsync(counter, 0, 4 * sizeof(int), st));  // work counters of the stage + the overflow flag
  SolveIO io;
  io.des = des; io.w = w; io.rfz = rfz; io.trq = trq; io.acc = acc; io.rf = rf; io.qddot = qddot;
  io.status = status; io.iters = iters; io.active = reinterpret_cast<unsigned long long*>(active); io.counter = counter; io.overflow = counter + 3;
  io.active_ref = reinterpret_cast<unsigned long long*>(active_ref);
  io.psi_refine = g_termination == PNC_TERM_REFERENCE ? 1.0 : 1e-6;
  launch_ihwbc_solve(p->d, p->h, p->lay, wsp, fm, ws->nfr, B, io, ws->m->num_sms, st);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_ihwbc_solve(const pnc_problem* p, pnc_workspace* ws, int32_t B, const double* d_des, const double* d_w,
                    const double* d_rf_z_max, double* d_trq, double* d_acc, double* d_rf, double* d_qddot,
                    int32_t* d_status, int32_t* d_iters, uint64_t* d_active, void* stream) {
  return pnc_ihwbc_solve_ref(p, ws, B, d_des, d_w, d_rf_z_max, d_trq, d_acc, d_rf, d_qddot, d_status, d_iters, d_active,
                             nullptr, stream);
}

int pnc_ihwbc_solve_ref(const pnc_problem* p, pnc_workspace* ws, int32_t B, const double* d_des, const double* d_w,
                        const double* d_rf_z_max, double* d_trq, double* d_acc, double* d_rf, double* d_qddot,
                        int32_t* d_status, int32_t* d_iters, uint64_t* d_active, uint64_t* d_active_ref, void* stream) {
  if (!p || !ws || B < 0 || B > ws->cap || p->m != ws->m) return PNC_E_ARG;
  if (!d_des || !d_w || (p->h.ncontact && !d_rf_z_max) || !d_status) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  return solve_at(p, ws, 0, B, d_des, d_w, d_rf_z_max, d_trq, d_acc, d_rf, d_qddot, d_status, d_iters, d_active,
                  (cudaStream_t)stream, ws->counter, d_active_ref);
}

int pnc_set_precision(int32_t mode) {
  if (mode != PNC_PREC_FP64 && mode != PNC_PREC_FP32_DYNAMICS) return PNC_E_ARG;
  g_precision = mode;
  return PNC_OK;
}

int pnc_get_precision(void) { return g_precision; }

int pnc_set_termination(int32_t mode) {
  if (mode != PNC_TERM_REFINED && mode != PNC_TERM_REFERENCE) return PNC_E_ARG;
  g_termination = mode;
  return PNC_OK;
}

int pnc_get_termination(void) { return g_termination; }

int pnc_task_eval(const pnc_problem* p, pnc_workspace* ws, int32_t B, int32_t itask, const double* d_des,
                  double* d_jac, double* d_jdqd, double* d_op_cmd, double* d_pos_err, void* stream) {
  if (!p || !ws || B < 0 || B > ws->cap || p->m != ws->m || itask < 0 || itask >= p->h.ntask || !d_des)
    return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  FrameMap fm;
  int rc = resolve_frames(p, ws, &fm);
  if (rc) return rc;
  launch_task_eval(p->d, ws->p, fm, ws->nfr, B, itask, p->h.tasks[itask].dim, p->h.nv, d_des, d_jac, d_jdqd,
                   d_op_cmd, d_pos_err, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_contact_eval(const pnc_problem* p, pnc_workspace* ws, int32_t B, int32_t icontact,
                     const double* d_rf_z_max, double* d_jac, double* d_cone_mat, double* d_cone_vec,
                     void* stream) {
  if (!p || !ws || B < 0 || B > ws->cap || p->m != ws->m || icontact < 0 || icontact >= p->h.ncontact ||
      !d_rf_z_max)
    return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  FrameMap fm;
  int rc = resolve_frames(p, ws, &fm);
  if (rc) return rc;
  launch_contact_eval(p->d, ws->p, fm, ws->nfr, B, icontact, p->h.nv, d_rf_z_max, d_jac, d_cone_mat, d_cone_vec,
                      (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_joint_integrate(int32_t B, int32_t na, const double* d_acc, const double* d_joint_pos, double* d_vel_state,
                        double* d_pos_state, const double* d_pos_limit, const double* d_vel_limit,
                        double alpha_vel, double alpha_pos, double dt, void* stream) {
  if (B < 0 || na < 1 || !d_acc || !d_joint_pos || !d_vel_state || !d_pos_state || !d_pos_limit || !d_vel_limit)
    return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  SETDEV(d_acc);
  launch_joint_integrate(B, na, d_acc, d_joint_pos, d_vel_state, d_pos_state, d_pos_limit, d_vel_limit, alpha_vel,
                         alpha_pos, dt, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

// Operational-space control step of a fixed-base arm (manipulator_interface.py:77-114) for B states.
// `frame` is a MODEL frame index that the workspace materialises; d_kp / d_kd are 3 device doubles.
int pnc_osc_step(pnc_workspace* ws, int32_t B, int32_t frame, const double* d_pos_des, const double* d_vel_des,
                 const double* d_kp, const double* d_kd, double smooth, double* d_trq, double* d_qddot, void* stream) {
  if (!ws || B < 0 || B > ws->cap || !d_pos_des || !d_vel_des || !d_kp || !d_kd || !d_trq) return PNC_E_ARG;
  int slot = -1;
  for (int f = 0; f < ws->nfr; f++)
    if (ws->hfr.model_frame[f] == frame) slot = f;
  if (slot < 0) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  launch_osc_step(ws->p, ws->nfr, slot, ws->m->h.nv, B, d_pos_des, d_vel_des, d_kp, d_kd, smooth, 1e-3, d_trq, d_qddot,
                  (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

// ---------------------------------------------------------------------------------------
// Batched dense QP (ihwbc.cu: k_quadprog_dense)
int pnc_quadprog_batch(int32_t device, int32_t B, int32_t n, int32_t p, int32_t m, const double* d_G, const double* d_g0,
                       const double* d_CE, const double* d_ce0, const double* d_CI, const double* d_ci0, double* d_x,
                       double* d_f, int32_t* d_status, int32_t* d_iters, uint64_t* d_active, void* stream) {
  if (B < 0 || n < 1 || p < 0 || m < 0 || !d_G || !d_g0 || !d_x || (p > 0 && (!d_CE || !d_ce0)) ||
      (m > 0 && (!d_CI || !d_ci0)))
    return PNC_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return PNC_E_NOGPU; }
  if (device < 0 || device >= ndev) return PNC_E_ARG;
  if (n > 64 || p > n) return PNC_E_SIZE;
  if (B == 0) return PNC_OK;
  ON_DEVICE(device);
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (!launch_quadprog_dense(n, p, m, B, d_G, d_g0, d_CE, d_ce0, d_CI, d_ci0, d_x, d_f, d_status, d_iters,
                             reinterpret_cast<unsigned long long*>(d_active), sms, (cudaStream_t)stream))
    return PNC_E_SIZE;
  CK(cudaGetLastError());
  return PNC_OK;
}

// ---------------------------------------------------------------------------------------
// The reference generators that feed the tasks (managers.cu).
static int frame_slot(const pnc_workspace* ws, int32_t frame) {
  for (int f = 0; f < ws->nfr; f++)
    if (ws->hfr.model_frame[f] == frame) return f;
  return -1;
}

int pnc_ramp_update(int32_t B, const double* d_start_value, double target, const double* d_start_time,
                    const double* d_duration, double current_time, double* d_out, void* stream) {
  if (B < 0 || !d_start_value || !d_start_time || !d_duration || !d_out) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  SETDEV(d_out);
  launch_ramp_update(B, d_start_value, target, d_start_time, d_duration, current_time, d_out, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_foot_use_current(pnc_workspace* ws, int32_t B, int32_t frame, double* d_pos, double* d_vel, double* d_acc,
                         double* d_quat, double* d_angvel, double* d_angacc, void* stream) {
  if (!ws || B < 0 || B > ws->cap || !d_pos || !d_vel || !d_acc || !d_quat || !d_angvel || !d_angacc) return PNC_E_ARG;
  const int slot = frame_slot(ws, frame);
  if (slot < 0) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  launch_foot_use_current(ws->p, ws->nfr, slot, B, d_pos, d_vel, d_acc, d_quat, d_angvel, d_angacc, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_hand_init(pnc_workspace* ws, int32_t B, int32_t frame, const double* d_target_pos, const double* d_target_quat,
                  const double* d_keypoint, double* d_rec, void* stream) {
  if (!ws || B < 0 || B > ws->cap || !d_target_pos || !d_target_quat || !d_rec) return PNC_E_ARG;
  const int slot = frame_slot(ws, frame);
  if (slot < 0) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  launch_hand_init(ws->p, ws->nfr, slot, B, d_target_pos, d_target_quat, d_keypoint, d_rec, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_hand_eval(int32_t B, const double* d_rec, const double* d_start_time, const double* d_duration, double current_time,
                  int32_t b_keypoint, double* d_pos, double* d_vel, double* d_acc, double* d_quat, double* d_angvel,
                  double* d_angacc, void* stream) {
  if (B < 0 || !d_rec || !d_start_time || !d_duration || !d_pos || !d_vel || !d_acc || !d_quat || !d_angvel || !d_angacc)
    return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  SETDEV(d_rec);
  launch_hand_eval(B, d_rec, d_start_time, d_duration, current_time, b_keypoint != 0, d_pos, d_vel, d_acc, d_quat, d_angvel,
                   d_angacc, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_swing_init(pnc_workspace* ws, int32_t B, int32_t frame, const double* d_land_pos, const double* d_land_quat,
                   const double* d_swing_duration, double swing_height, double* d_rec, void* stream) {
  if (!ws || B < 0 || B > ws->cap || !d_land_pos || !d_land_quat || !d_swing_duration || !d_rec) return PNC_E_ARG;
  const int slot = frame_slot(ws, frame);
  if (slot < 0) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  launch_swing_init(ws->p, ws->nfr, slot, B, d_land_pos, d_land_quat, d_swing_duration, swing_height, d_rec,
                    (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_swing_eval(int32_t B, const double* d_rec, const double* d_start_time, const double* d_swing_duration,
                   double current_time, double* d_pos, double* d_vel, double* d_acc, double* d_quat, double* d_angvel,
                   double* d_angacc, void* stream) {
  if (B < 0 || !d_rec || !d_start_time || !d_swing_duration || !d_pos || !d_vel || !d_acc || !d_quat || !d_angvel ||
      !d_angacc)
    return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  SETDEV(d_rec);
  launch_swing_eval(B, d_rec, d_start_time, d_swing_duration, current_time, d_pos, d_vel, d_acc, d_quat, d_angvel, d_angacc,
                    (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_floating_base_init(pnc_workspace* ws, int32_t B, int32_t base_frame, const double* d_target_com,
                           const double* d_target_quat, double* d_rec, void* stream) {
  if (!ws || B < 0 || B > ws->cap || !d_target_com || !d_target_quat || !d_rec) return PNC_E_ARG;
  const int slot = frame_slot(ws, base_frame);
  if (slot < 0) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  launch_floating_base_init(ws->p, ws->nfr, slot, B, d_target_com, d_target_quat, d_rec, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_floating_base_eval(int32_t B, const double* d_rec, const double* d_start_time, const double* d_duration,
                           double current_time, double* d_pos, double* d_vel, double* d_acc, double* d_quat,
                           double* d_angvel, double* d_angacc, void* stream) {
  if (B < 0 || !d_rec || !d_start_time || !d_duration || !d_pos || !d_vel || !d_acc || !d_quat || !d_angvel || !d_angacc)
    return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  SETDEV(d_rec);
  launch_floating_base_eval(B, d_rec, d_start_time, d_duration, current_time, d_pos, d_vel, d_acc, d_quat, d_angvel,
                            d_angacc, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_sinusoid_eval(int32_t B, const double* d_mid, const double* d_start_time, const double* amp3, const double* freq3,
                      double current_time, double* d_pos, double* d_vel, double* d_acc, void* stream) {
  if (B < 0 || !d_mid || !d_start_time || !amp3 || !freq3 || !d_pos || !d_vel || !d_acc) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  SETDEV(d_mid);
  launch_sinusoid_eval(B, d_mid, d_start_time, amp3, freq3, current_time, d_pos, d_vel, d_acc, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

int pnc_dcm_update(pnc_workspace* ws, int32_t B, double dt, double alpha, double* d_dcm, double* d_prev_dcm,
                   double* d_dcm_vel, void* stream) {
  if (!ws || B < 0 || B > ws->cap || !(dt > 0.0) || !d_dcm || !d_prev_dcm || !d_dcm_vel) return PNC_E_ARG;
  if (B == 0) return PNC_OK;
  ON_DEVICE(ws->m->device);
  launch_dcm_update(ws->p, B, dt, alpha, d_dcm, d_prev_dcm, d_dcm_vel, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return PNC_OK;
}

// ---------------------------------------------------------------------------------------
// Whole tick with host buffers.  The batch is cut into chunks; one stream copies the inputs of
// every chunk in, the kernels of the chunks alternate over two streams, one stream copies the
// outputs out, and per-chunk events order them (see below).  Chunks own disjoint slices of the
// workspace and of the staging slab, so they never alias.  Host buffers should be pinned.
struct TickArgs {
  const double *h_bp, *h_bq, *h_blv, *h_bav, *h_jp, *h_jv;  // robot state, host
  const double *des, *w, *rfz;                               // host, or device when des_on_device
  bool des_on_device;
  cudaStream_t ready;                                        // producer stream of device des / w / rfz (may be 0)
  const pnc_integrator* integ;                               // optional command epilogue
  double *h_trq, *h_acc, *h_rf;
  int32_t *h_status, *h_iters;
  uint64_t* h_active;
  double *h_jtrq, *h_jvel, *h_jpos;                          // all-joint commands (with integ)
};

static int tick_wait(pnc_workspace* ws) {
  if (!ws->tick_init) return PNC_OK;
  DevGuard dev_guard_(ws->device);
  cudaError_t e0 = cudaStreamSynchronize(ws->tick_out);
  cudaError_t e1 = cudaStreamSynchronize(ws->tick_stream[0]);
  cudaError_t e2 = cudaStreamSynchronize(ws->tick_stream[1]);
  cudaError_t e3 = cudaStreamSynchronize(ws->tick_in);
  const bool ok = e0 == cudaSuccess && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess;
  if (ok && ws->tick_last > 0) {
    // How busy was the input link?  With t_in = time the input copies of all chunks took back to back and
    // t_all = time until the last chunk's kernels finished, a chunk may be about t_all / t_in times larger than
    // its predecessor and still have its inputs arrive under the predecessor's kernels.  One B200 on its own
    // x16 link: ~4; eight ranks sharing the host's memory and PCIe switches: ~1.5.
    float t_in = 0.f, t_all = 0.f;
    if (cudaEventElapsedTime(&t_in, ws->tick_ev_start, ws->tick_ev_in[ws->tick_last]) == cudaSuccess &&
        cudaEventElapsedTime(&t_all, ws->tick_ev_start, ws->tick_ev_done[ws->tick_last]) == cudaSuccess && t_in > 0.f && t_all > t_in)
      ws->tick_growth = 0.7 * (double)t_all / (double)t_in;  // 0.7: t_all contains the waits this is meant to remove
  }
  ws->tick_last = -1;
  return ok ? PNC_OK : PNC_E_CUDA;
}

// enqueues everything of one tick on the workspace's own streams; tick_wait() completes it
static int tick_enqueue(const pnc_problem* p, pnc_workspace* ws, int32_t B, const TickArgs& A) {
  ON_DEVICE(ws->m->device);
  const DevProblem& P = p->h;
  const size_t na = ws->m->h.na, nv = P.nv;
  const int mw = (P.m + 63) / 64 > 0 ? (P.m + 63) / 64 : 1;
  // per-state doubles in / out
  const size_t in_d = 3 + 4 + 3 + 3 + 2 * na + P.ndes + P.ntask + P.ncontact;
  const size_t out_d = 2 * (size_t)P.nact + P.nc + nv + 3 * na;
  const size_t per_state = (in_d + out_d) * 8 + 8 /*status, iters*/ + 8 * (size_t)mw;
  const size_t need = per_state * (size_t)ws->cap + 64 * 256;
  if (ws->stage_bytes < need) {
    if (ws->stage) cudaFree(ws->stage);
    ws->stage = nullptr; ws->stage_bytes = 0;
    CK(cudaMalloc(&ws->stage, need));
    ws->stage_bytes = need;
  }
  if (!ws->tick_init) {
    for (int i = 0; i < 2; i++) CK(cudaStreamCreateWithFlags(&ws->tick_stream[i], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ws->tick_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ws->tick_out, cudaStreamNonBlocking));
    for (int i = 0; i < 16; i++) {
      CK(cudaEventCreate(&ws->tick_ev_in[i]));  // timed: the input link rate of one tick sets the next tick's chunks
      CK(cudaEventCreate(&ws->tick_ev_done[i]));
    }
    CK(cudaEventCreateWithFlags(&ws->tick_ev_ready, cudaEventDisableTiming));
    CK(cudaEventCreate(&ws->tick_ev_start));
    ws->tick_init = true;
  }
  // carve the staging slab (capacity-sized arrays)
  char* cur = ws->stage;
  auto take = [&](size_t bytes) { char* r = cur; cur += align_up(bytes); return r; };
  const size_t C = ws->cap;
  double* d_bp = (double*)take(C * 3 * 8);
  double* d_bq = (double*)take(C * 4 * 8);
  double* d_blv = (double*)take(C * 3 * 8);
  double* d_bav = (double*)take(C * 3 * 8);
  double* d_jp = (double*)take(C * na * 8);
  double* d_jv = (double*)take(C * na * 8);
  double* d_des = (double*)take(C * P.ndes * 8);
  double* d_w = (double*)take(C * P.ntask * 8);
  double* d_rfz = (double*)take(C * (P.ncontact ? P.ncontact : 1) * 8);
  double* d_trq = (double*)take(C * P.nact * 8);
  double* d_acc = (double*)take(C * P.nact * 8);
  double* d_rf = (double*)take(C * (P.nc ? P.nc : 1) * 8);
  double* d_jtrq = (double*)take(C * na * 8);
  double* d_jvel = (double*)take(C * na * 8);
  double* d_jpos = (double*)take(C * na * 8);
  int32_t* d_status = (int32_t*)take(C * 4);
  int32_t* d_iters = (int32_t*)take(C * 4);
  uint64_t* d_active = (uint64_t*)take(C * mw * 8);
  int* counters = (int*)take(1024);  // 8 ints per chunk, up to 16 chunks
  if (A.des_on_device) {  // desireds, weights and rf_z_max already live on the device (on-device managers)
    d_des = const_cast<double*>(A.des); d_w = const_cast<double*>(A.w); d_rfz = const_cast<double*>(A.rfz);
    if (A.ready) {
      CK(cudaEventRecord(ws->tick_ev_ready, A.ready));
      for (int i = 0; i < 2; i++) CK(cudaStreamWaitEvent(ws->tick_stream[i], ws->tick_ev_ready, 0));
    }
  }

  // Chunk schedule.  The first chunk's input copy and the last chunk's output copy are the transfers nothing
  // overlaps, so both ends are small; in between every chunk is `growth` times its predecessor, so that its
  // inputs arrive while the predecessor computes.  `growth` comes from the previous tick's own timing (see
  // tick_wait): ~4 with the link to oneself (5 chunks), ~1.4 when eight ranks share the host (9 chunks).
  // Measured on Atlas 64 K (scripts/e2e_split_sweep*.sh): 1 GPU 10.23 -> 10.06 ms per step against the fixed
  // B/16, 0.4 B, 0.4 B, 0.14 B split; 8 GPUs 12.06 -> 11.2-11.3 ms.  PNC_TICK_CHUNKS / PNC_TICK_FIRST force n
  // chunks with a first chunk of the given fraction, PNC_TICK_SPLIT explicit fractions (diagnostics).
  static const int env_chunks = [] { const char* e = getenv("PNC_TICK_CHUNKS"); return e ? atoi(e) : 0; }();
  static const double env_first = [] { const char* e = getenv("PNC_TICK_FIRST"); return e ? atof(e) : 0.0; }();
  int nchunk = B >= 8192 ? (env_chunks > 0 && env_chunks <= 16 ? env_chunks : 0) : 1;
  int bounds[17];
  {
    bounds[0] = 0;
    if (nchunk == 1) {
      bounds[1] = B;
    } else if (env_chunks > 0 || (env_first > 0.0 && env_first < 1.0)) {
      if (nchunk == 0) nchunk = 4;
      int first = env_chunks > 0 ? (B + nchunk - 1) / nchunk : B / 16;
      if (nchunk > 1 && env_first > 0.0 && env_first < 1.0) first = (int)(env_first * B);
      if (first < 1) first = 1;
      const int rest = B - first;
      for (int c = 1; c <= nchunk; c++)
        bounds[c] = c == nchunk ? B : first + (int)((long long)rest * (c - 1) / (nchunk - 1 > 0 ? nchunk - 1 : 1));
    } else {
      double g = ws->tick_growth > 0.0 ? ws->tick_growth : 2.5;
      g = g < 1.3 ? 1.3 : (g > 4.0 ? 4.0 : g);
      const double tail = 0.10, cap = 0.34;
      double f[16], acc = 0.0, sz = 1.0 / 32.0;
      if (sz * B < 2048.0) sz = 2048.0 / B;
      int nf = 0;
      while (nf < 13 && acc + sz < 1.0 - tail) { f[nf++] = sz; acc += sz; sz = sz * g < cap ? sz * g : cap; }
      const double rem = 1.0 - acc;
      if (rem > 1.4 * tail) { f[nf++] = rem - tail; f[nf++] = tail; } else f[nf++] = rem;
      acc = 0.0;
      for (int c = 0; c < nf; c++) { acc += f[c]; bounds[c + 1] = c == nf - 1 ? B : (int)(acc * B); }
      nchunk = nf;
    }
    // PNC_TICK_SPLIT="f0,f1,..." (fractions of the batch, diagnostic) overrides everything above
    static const char* env_split = getenv("PNC_TICK_SPLIT");
    if (env_split && B >= 8192) {
      double f[16];
      int nf = 0;
      for (const char* q = env_split; *q && nf < 16;) {
        f[nf++] = atof(q);
        while (*q && *q != ',') q++;
        if (*q == ',') q++;
      }
      double acc = 0.0;
      for (int c = 0; c < nf; c++) { acc += f[c]; bounds[c + 1] = c == nf - 1 ? B : (int)(acc * B); }
      if (nf > 0) nchunk = nf;
    }
  }
  // Three roles: one stream feeds the inputs of every chunk as fast as the link takes them, the
  // kernels of the chunks alternate over two streams (the tail of one chunk overlaps the start of
  // the next), one stream drains the outputs.  Events tie a chunk's kernels to its inputs and its
  // output copies to its kernels.
  static const bool verbose = getenv("PNC_VERBOSE") != nullptr;
  if (verbose) {
    fprintf(stderr, "pnc_b200: tick of %d states, growth %.2f, %d chunks:", B, ws->tick_growth, nchunk);
    for (int c = 0; c < nchunk; c++) fprintf(stderr, " %d", bounds[c + 1] - bounds[c]);
    fprintf(stderr, "\n");
  }
  CK(cudaEventRecord(ws->tick_ev_start, ws->tick_in));
  ws->tick_last = -1;
  for (int c = 0; c < nchunk; c++) {
    int o = bounds[c], nb = bounds[c + 1] - bounds[c];
    if (nb <= 0) continue;
    ws->tick_last = c;
    cudaStream_t st = ws->tick_in;
#define H2D(dst, src, per) CK(cudaMemcpyAsync((dst) + (size_t)o * (per), (src) + (size_t)o * (per), (size_t)nb * (per) * 8, cudaMemcpyHostToDevice, st))
    H2D(d_bp, A.h_bp, 3); H2D(d_bq, A.h_bq, 4); H2D(d_blv, A.h_blv, 3); H2D(d_bav, A.h_bav, 3);
    H2D(d_jp, A.h_jp, na); H2D(d_jv, A.h_jv, na);
    if (!A.des_on_device) {
      H2D(d_des, A.des, P.ndes); H2D(d_w, A.w, P.ntask);
      if (P.ncontact) H2D(d_rfz, A.rfz, P.ncontact);
    }
#undef H2D
    CK(cudaEventRecord(ws->tick_ev_in[c], st));
    // the chunk's kernels are enqueued right behind its copies (not after the copies of every chunk): the first
    // launch then reaches the GPU ~50 API calls earlier
    st = ws->tick_stream[c & 1];
    CK(cudaStreamWaitEvent(st, ws->tick_ev_in[c], 0));
    int rc = update_system_at(ws, o, nb, d_bp + (size_t)o * 3, d_bq + (size_t)o * 4, d_blv + (size_t)o * 3,
                              d_bav + (size_t)o * 3, d_jp + (size_t)o * na, d_jv + (size_t)o * na, st);
    if (rc) return rc;
    rc = solve_at(p, ws, o, nb, d_des + (size_t)o * P.ndes, d_w + (size_t)o * P.ntask,
                  d_rfz + (size_t)o * P.ncontact, d_trq + (size_t)o * P.nact, d_acc + (size_t)o * P.nact,
                  d_rf + (size_t)o * P.nc, nullptr, d_status + o, d_iters + o, d_active + (size_t)o * mw, st,
                  counters + 8 * c);
    if (rc) return rc;
    if (A.integ) {  // controller epilogue: Sa^T scatter + JointIntegrator.integrate (atlas_controller.py:79-85)
      const pnc_integrator& J = *A.integ;
      launch_command_epilogue(p->d, nb, (int)na, d_trq + (size_t)o * P.nact, d_acc + (size_t)o * P.nact, d_jp + (size_t)o * na,
                              J.d_vel_state + (size_t)o * na, J.d_pos_state + (size_t)o * na, J.d_pos_limit, J.d_vel_limit,
                              J.alpha_vel, J.alpha_pos, J.dt, d_jtrq + (size_t)o * na, d_jvel + (size_t)o * na,
                              d_jpos + (size_t)o * na, st);
      CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ws->tick_ev_done[c], st));
    cudaStream_t so = ws->tick_out;
    CK(cudaStreamWaitEvent(so, ws->tick_ev_done[c], 0));
#define D2H(dst, src, per, esz) if (dst) CK(cudaMemcpyAsync((char*)(dst) + (size_t)o * (per) * (esz), (const char*)(src) + (size_t)o * (per) * (esz), (size_t)nb * (per) * (esz), cudaMemcpyDeviceToHost, so))
    D2H(A.h_trq, d_trq, P.nact, 8); D2H(A.h_acc, d_acc, P.nact, 8); D2H(A.h_rf, d_rf, P.nc, 8);
    D2H(A.h_status, d_status, 1, 4); D2H(A.h_iters, d_iters, 1, 4); D2H(A.h_active, d_active, mw, 8);
    if (A.integ) { D2H(A.h_jtrq, d_jtrq, na, 8); D2H(A.h_jvel, d_jvel, na, 8); D2H(A.h_jpos, d_jpos, na, 8); }
#undef D2H
  }
  return PNC_OK;
}

static int tick_args_ok(const pnc_problem* p, const pnc_workspace* ws, int32_t B, const TickArgs& A) {
  if (!p || !ws || B < 0 || B > ws->cap || p->m != ws->m) return PNC_E_ARG;
  if (!A.h_bp || !A.h_bq || !A.h_blv || !A.h_bav || !A.h_jp || !A.h_jv || !A.des || !A.w || (p->h.ncontact && !A.rfz))
    return PNC_E_ARG;
  if (A.integ && (!A.integ->d_vel_state || !A.integ->d_pos_state || !A.integ->d_pos_limit || !A.integ->d_vel_limit))
    return PNC_E_ARG;
  return PNC_OK;
}

int pnc_wbc_tick_host(const pnc_problem* p, pnc_workspace* ws, int32_t B, const double* h_base_pos,
                      const double* h_base_quat, const double* h_base_lin_vel, const double* h_base_ang_vel,
                      const double* h_joint_pos, const double* h_joint_vel, const double* h_des, const double* h_w,
                      const double* h_rf_z_max, double* h_trq, double* h_acc, double* h_rf, int32_t* h_status,
                      int32_t* h_iters, uint64_t* h_active) {
  TickArgs A = {h_base_pos, h_base_quat, h_base_lin_vel, h_base_ang_vel, h_joint_pos, h_joint_vel, h_des, h_w, h_rf_z_max,
                false, nullptr, nullptr, h_trq, h_acc, h_rf, h_status, h_iters, h_active, nullptr, nullptr, nullptr};
  int rc = tick_args_ok(p, ws, B, A);
  if (rc || B == 0) return rc;
  rc = tick_enqueue(p, ws, B, A);
  const int rw = tick_wait(ws);  // also on an error: nothing of this call stays in flight
  return rc ? rc : rw;
}

int pnc_wbc_tick_host_ex(const pnc_problem* p, pnc_workspace* ws, int32_t B, const double* h_base_pos,
                         const double* h_base_quat, const double* h_base_lin_vel, const double* h_base_ang_vel,
                         const double* h_joint_pos, const double* h_joint_vel, const double* des, const double* w,
                         const double* rf_z_max, int32_t flags, void* ready_stream, const pnc_integrator* integ,
                         double* h_trq, double* h_acc, double* h_rf, int32_t* h_status, int32_t* h_iters, uint64_t* h_active,
                         double* h_joint_trq_cmd, double* h_joint_vel_cmd, double* h_joint_pos_cmd) {
  TickArgs A = {h_base_pos, h_base_quat, h_base_lin_vel, h_base_ang_vel, h_joint_pos, h_joint_vel, des, w, rf_z_max,
                (flags & PNC_TICK_DES_ON_DEVICE) != 0, (cudaStream_t)ready_stream, integ, h_trq, h_acc, h_rf, h_status,
                h_iters, h_active, h_joint_trq_cmd, h_joint_vel_cmd, h_joint_pos_cmd};
  int rc = tick_args_ok(p, ws, B, A);
  if (rc || B == 0) return rc;
  rc = tick_enqueue(p, ws, B, A);
  const int rw = tick_wait(ws);
  return rc ? rc : rw;
}

// One host batch over n devices from ONE process: contiguous split of the state dimension,
// device i takes states [i B / n, (i + 1) B / n); every device's copies and kernels are enqueued before
// any is waited for, so the devices run concurrently.
int pnc_wbc_tick_host_multi(int32_t n_dev, const pnc_problem* const* ps, pnc_workspace* const* wss, int32_t B,
                            const double* h_base_pos, const double* h_base_quat, const double* h_base_lin_vel,
                            const double* h_base_ang_vel, const double* h_joint_pos, const double* h_joint_vel,
                            const double* h_des, const double* h_w, const double* h_rf_z_max, double* h_trq,
                            double* h_acc, double* h_rf, int32_t* h_status, int32_t* h_iters, uint64_t* h_active) {
  if (n_dev < 1 || n_dev > 64 || !ps || !wss || B < 0) return PNC_E_ARG;
  int rc = PNC_OK, launched = 0;
  for (int i = 0; i < n_dev && rc == PNC_OK; i++) {
    const pnc_problem* p = ps[i];
    pnc_workspace* ws = wss[i];
    if (!p || !ws) { rc = PNC_E_ARG; break; }
    const long long lo = (long long)B * i / n_dev, hi = (long long)B * (i + 1) / n_dev;
    const int32_t nb = (int32_t)(hi - lo);
    const DevProblem& P = p->h;
    const size_t na = ws->m->h.na;
    const int mw = (P.m + 63) / 64 > 0 ? (P.m + 63) / 64 : 1;
#define OFF(ptr, per) ((ptr) ? (ptr) + (size_t)lo * (per) : nullptr)
    TickArgs A = {OFF(h_base_pos, 3), OFF(h_base_quat, 4), OFF(h_base_lin_vel, 3), OFF(h_base_ang_vel, 3),
                  OFF(h_joint_pos, na), OFF(h_joint_vel, na), OFF(h_des, P.ndes), OFF(h_w, P.ntask),
                  OFF(h_rf_z_max, P.ncontact), false, nullptr, nullptr, OFF(h_trq, P.nact), OFF(h_acc, P.nact),
                  OFF(h_rf, P.nc), OFF(h_status, 1), OFF(h_iters, 1), OFF(h_active, mw), nullptr, nullptr, nullptr};
#undef OFF
    rc = tick_args_ok(p, ws, nb, A);
    if (rc == PNC_OK && nb > 0) rc = tick_enqueue(p, ws, nb, A);
    launched = i + 1;
  }
  for (int i = 0; i < launched; i++) {
    const int rw = tick_wait(wss[i]);
    if (rc == PNC_OK) rc = rw;
  }
  return rc;
}

int pnc_debug_setup_record(const pnc_problem* p, pnc_workspace* ws, double** d_rec, int32_t* stride, int32_t* layout) {
  if (!p || !ws || !d_rec || !stride || !layout) return PNC_E_ARG;
  *d_rec = ws->setup_slab;
  *stride = ws->setup_stride;
  return reg_record_layout(p->h, layout) ? PNC_OK : PNC_E_SIZE;
}

int pnc_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return PNC_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PNC_E_NOGPU;
  ON_DEVICE(device);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  double best = 0;
  int rc = measure_fp64_peak(prop.multiProcessorCount, &best);
  if (rc) return PNC_E_CUDA;
  *tflops = best;
  return PNC_OK;
}

}  // extern "C"
