// ihwbc_reg.cu -- the IHWBC solve with the active-set operator J = L^-T Q held in REGISTERS.
//
// Same algorithm and reference mapping as ihwbc.cu (IHWBC.solve pnc/wbc/ihwbc/ihwbc.py:90-379,
// Goldfarb-Idnani as in external_source/Goldfarb/QuadProg++.cc:115-502, Householder instead of
// the Givens sweep when a constraint is added); what changes is where the data lives and how
// much code the loop is:
//
//   * one robot state per warp; lane l owns row l of the n x n matrix J (n = nv + nc <= 48) in
//     registers, rows 32.. are split over lane pairs (half a row each) -- 72 doubles per lane
//     for Atlas.  z = J2 d2 and the rank-1 Householder update are pure register FMAs against
//     vectors broadcast from shared memory (LDS.128, one wavefront per two FMAs);
//   * d = J^T n+ goes through a 6-row scratch tile in shared memory: one pass for a wrench-cone
//     normal (3 or 6 non-zeros), n/6 passes for a dense normal (equalities, torque limits);
//   * the Hessian, its Cholesky factor and L^-T are built once per state in shared memory with
//     compact loops, then the rows move into registers;
//   * the kernel is instruction-fetch sensitive (few warps per SM, each walking its own path),
//     so every routine of the active-set loop exists at exactly one call site -- equalities run
//     through the same step code as inequalities -- and the only unrolled code is what register
//     indexing forces.  Dynamic column indices (iq, the Givens pairs of a dropped constraint) are
//     warp-uniform and go through switch statements (jump tables), never local memory.
#include <stdlib.h>

#include "ihwbc_common.cuh"

namespace pnc {

template <int NV_, int NC_>
struct RCfg {
  static constexpr int NV = NV_, NC = NC_, N = NV_ + NC_;
  static constexpr bool EXT = N > 32;                         // rows 32.. live in lane pairs
  static constexpr int E = EXT ? N - 32 : 0;
  static constexpr int HW = EXT ? ((N + 3) / 4) * 2 : 0;      // half-row width (even)
  static constexpr int NP = EXT ? 2 * HW : ((N + 1) & ~1);    // padded vector length (even)
  static constexpr int NB = HW > 0 ? HW : 2;
  static constexpr int ldv_() { int l = (NV_ + 1) & ~1; while ((l & 15) != 6) l += 2; return l; }
  static constexpr int LDV = ldv_();                          // pitch of the nv x nv block: even, 6 (mod 16) ->
                                                              // rows 16-byte aligned, LDS.128 by row conflict-free
  static constexpr int EV = NV_ > 32 ? NV_ - 32 : 0;          // Hessian rows beyond 32
  static constexpr int EVP = EV <= 1 ? 1 : (EV <= 2 ? 2 : (EV <= 4 ? 4 : 8));
  static constexpr int LPR = 32 / EVP;                        // lanes cooperating on one such row in the Cholesky
  static constexpr int LDT = NP + 2;                          // pitch of the d scratch tile
  static constexpr int PMAX = 10;                             // equality rows the record holds
  static_assert(E <= 16, "rows beyond 32 must fit one lane pair each");
  static_assert(N <= 64, "n too large for the register kernel");
};

// Compile-time shared-memory layout of the active-set kernel (doubles).  NU = wrench-cone rows,
// NT = torque-limit rows (0 without torque limits).  Every array sits at an immediate offset from
// the shared-memory base: no address registers, nothing to rematerialise under the register
// pressure of the resident J.  S = R^-1 is capped at RC = 32 columns (active constraints incl.
// equalities); a state that would need more is handed to the generic kernel (status OVERFLOW).
template <int NV_, int NC_, int NU_, int NT_>
struct MCfg : RCfg<NV_, NC_> {
  using R = RCfg<NV_, NC_>;
  static constexpr int NU = NU_, NT = NT_, M = NU_ + 2 * NT_;
  static constexpr int RC = R::N < 32 ? R::N : 32;
  static constexpr int RSZ = (RC * (RC + 3) / 2 + 1) & ~1;
  static constexpr int lda_() { int l = R::NP; while ((l & 15) != 2) l += 2; return l; }
  static constexpr int LDA = lda_();  // torque rows: LDS.128 by row is conflict free at 2 (mod 16)
  static constexpr int ORT = 0, OT = ORT + RSZ, OAR = OT + 6 * R::LDT, OA0 = OAR + NT * LDA, OUF = OA0 + ((3 * NT + 1) & ~1),
                       OUFV = OUF + (NU > 0 ? NU : 1) * 6, OX = OUFV + (((NU > 0 ? NU : 1) + 1) & ~1), OZ = OX + R::NP,
                       ODV = OZ + R::NP, ODM = ODV + R::NP, OVM = ODM + R::NP, ONP = OVM + R::NP, ORR = ONP + R::NP + 6 /* zero pad: the d tile reads np six at a time */,
                       OU = ORR + R::NP + 2, OXOLD = OU + R::NP + 2, OUOLD = OXOLD + R::NP, OXACC = OUOLD + R::NP + 2,
                       OSL = OXACC + R::NP, OINT = OSL + (((M > 0 ? M : 1) + 1) & ~1),
                       SMEM_DOUBLES = (OINT + (((3 * (R::N + 2) + 1) & ~1) + 2 * NU) / 2 + 2) & ~1;  // A, Aold, Aacc, cone-row table
  static_assert(64 + R::PMAX * R::NP <= RSZ || R::N < 20, "equality block must fit behind the first S columns");
  // ---- per-state HBM record (doubles): written by k_ihwbc_setup, consumed by k_ihwbc_reg, whose
  //      result goes back into it for k_ihwbc_finish.  Everything the active-set kernel needs is
  //      here in the layout it is used in, so that kernel holds no staging code at all ----
  static constexpr int IMG = OX - OAR;                      // shared-memory image: Ar, ci0 of the torque rows, Uf, uf
  static constexpr int RJ = 0;                              // J rows as register chunks: [NP/2][32 lanes][2]
  static constexpr int RJB = RJ + 32 * R::NP;               // half rows 32..: [NB/2][32 lanes][2]
  static constexpr int RX = RJB + (R::EXT ? 32 * R::NB : 0);  // x0
  static constexpr int RD = RX + R::NP;                     // D = J^T CE^T, one row per equality
  static constexpr int RE = RD + R::PMAX * R::NP;           // equality normals
  static constexpr int RE0 = RE + R::PMAX * R::NP;          // equality offsets
  static constexpr int RS = RE0 + ((R::PMAX + 1) & ~1);     // c1, c2, status, jrf
  static constexpr int RI = RS + 4;                         // the image
  static constexpr int RO = RI + ((IMG + 1) & ~1);          // result: x[NP], then ints iq, iter, status, A[N+2], iq_ref, Aref[N+2]
  static constexpr int ROI = 2 * (R::N + 2) + 4;            // ints of the result
  static constexpr int RT = (RO + R::NP + (ROI + 1) / 2 + 1) & ~1;  // joint torques (robots with torque-limit rows)
  static constexpr int REC = RT + (NT > 0 ? 32 : 0);
  static_assert((OAR & 1) == 0 && (IMG & 1) == 0, "image is copied in 16-byte pieces");
};

// ---- register-array access by a warp-uniform dynamic index (jump table, no local memory) ----
#define PNC_R8(M, B) M(B + 0) M(B + 1) M(B + 2) M(B + 3) M(B + 4) M(B + 5) M(B + 6) M(B + 7)
#define PNC_R64(M) PNC_R8(M, 0) PNC_R8(M, 8) PNC_R8(M, 16) PNC_R8(M, 24) PNC_R8(M, 32) PNC_R8(M, 40) PNC_R8(M, 48) PNC_R8(M, 56)

template <int LEN>
__device__ __forceinline__ double reg_pick(const double (&A)[LEN], int idx) {
  double v = 0.0;
#define PNC_CASE(K) case (K): if constexpr ((K) < LEN) v = A[(K) < LEN ? (K) : 0]; break;
  switch (idx) { PNC_R64(PNC_CASE) default: break; }
#undef PNC_CASE
  return v;
}

// Givens rotations of the column pairs (j, j+1), j = j0 .. j1-1, of one register row, with the
// coefficients (cc, ss, xny) of pair j at rot[3 j ..] (QuadProg++.cc:660-676).  Enters the
// unrolled chain at j0 (jump table) and leaves at j1.
template <int LEN>
__device__ __forceinline__ void reg_rot_range(double (&A)[LEN], int j0, int j1, const double* rot) {
#define PNC_CASE(K)                                                                         \
  case (K):                                                                                 \
    if constexpr ((K) + 1 < LEN) {                                                          \
      if ((K) >= j1) break;                                                                 \
      const double cc = rot[3 * (K)], ss = rot[3 * (K) + 1], xny = rot[3 * (K) + 2];        \
      if (cc <= 1.5) { /* cc = 2 marks a pair QuadProg++ leaves untouched (|h| < eps) */    \
        const double t1 = A[(K) + 1 < LEN ? (K) : 0], t2 = A[(K) + 1 < LEN ? (K) + 1 : 0];  \
        const double nj = t1 * cc + t2 * ss;                                                \
        A[(K) + 1 < LEN ? (K) : 0] = nj;                                                    \
        A[(K) + 1 < LEN ? (K) + 1 : 0] = xny * (nj + t1) - t2;                              \
      }                                                                                     \
    }
  switch (j0) { PNC_R64(PNC_CASE) default: break; }
#undef PNC_CASE
}

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// CTA barrier over the `nthr` threads still iterating; returns how many of them stay (every
// arriving thread votes `stay`).  A warp that runs out of states votes false once and leaves.
__device__ __forceinline__ int slot_barrier(int nthr, bool stay) {
  int n;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tbarrier.cta.red.popc.u32 %0, 0, %1, p;\n\t}"
               : "=r"(n) : "r"(nthr), "r"((int)stay) : "memory");
  return n;
}

// Torque of actuated row a at x = [qddot; rf] (ihwbc.py:344-354):  S (M qddot + Ni^T (c+g) - (Jc Ni)^T rf).
// Called with lanes = rows: M is symmetric, so lane a reads column act_idx[a] of M -- consecutive
// lanes touch consecutive addresses -- and the constraint rows never have to be copied anywhere.
// Inlined (kernel parameters must not have their address taken: an out-of-line copy goes through
// local memory and generic loads).
template <int NV>
__device__ __forceinline__ double torque_row(const DevProblem& P, const DevProblem& G, const WsPtrs& ws, const FrameMap& fm,
                                             int nfr, size_t s, int a, const double* __restrict__ x) {
  const int nact = P.nact;
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  if (P.n_ic > 0) {
    const int n = P.n;
    const double* col = ws.ic_tq + s * n * nact + a;
    int k = 0;
#pragma unroll 2
    for (; k + 3 < n; k += 4, col += 4 * nact) {
      const double g0 = col[0], g1 = col[nact], g2 = col[2 * nact], g3 = col[3 * nact];
      const double2 xa = *reinterpret_cast<const double2*>(x + k), xb = *reinterpret_cast<const double2*>(x + k + 2);
      t0 = fma(g0, xa.x, t0); t1 = fma(g1, xa.y, t1); t2 = fma(g2, xb.x, t2); t3 = fma(g3, xb.y, t3);
    }
    for (; k < n; k++, col += nact) t0 = fma(col[0], x[k], t0);
    return (t0 + t1) + (t2 + t3) + ws.ic_tq0[s * nact + a];
  }
  const int v = G.act_idx[a];
  const double* col = ws.M + s * NV * NV + v;  // column v of M == row v
  const double cg = ws.cori[s * NV + v] + ws.grav[s * NV + v];
#pragma unroll
  for (int k = 0; k + 3 < NV; k += 4) {
    const double g0 = col[k * NV], g1 = col[(k + 1) * NV], g2 = col[(k + 2) * NV], g3 = col[(k + 3) * NV];
    const double2 xa = *reinterpret_cast<const double2*>(x + k), xb = *reinterpret_cast<const double2*>(x + k + 2);
    t0 = fma(g0, xa.x, t0); t1 = fma(g1, xa.y, t1); t2 = fma(g2, xb.x, t2); t3 = fma(g3, xb.y, t3);
  }
#pragma unroll
  for (int k = NV & ~3; k < NV; k++) t0 = fma(col[k * NV], x[k], t0);
  const double* xr = x + NV;
#pragma unroll 1
  for (int ci = 0; ci < P.ncontact; ci++) {
    const int dim = P.contacts[ci].dim;
    const double* jr = ws.fr_jac + ((s * nfr + fm.contact_f[ci]) * 6 + (6 - dim)) * NV + v;  // point contacts: rows 3..5
    if (dim == 6) {
      const double j0 = jr[0], j1 = jr[NV], j2 = jr[2 * NV], j3 = jr[3 * NV], j4 = jr[4 * NV], j5 = jr[5 * NV];
      t0 = fma(-j0, xr[0], t0); t1 = fma(-j1, xr[1], t1); t2 = fma(-j2, xr[2], t2);
      t3 = fma(-j3, xr[3], t3); t0 = fma(-j4, xr[4], t0); t1 = fma(-j5, xr[5], t1);
    } else {
      const double j0 = jr[0], j1 = jr[NV], j2 = jr[2 * NV];
      t0 = fma(-j0, xr[0], t0); t1 = fma(-j1, xr[1], t1); t2 = fma(-j2, xr[2], t2);
    }
    xr += dim;
  }
  return (t0 + t1) + (t2 + t3) + cg;
}

// =======================================================================================
// Per-state setup, one state per warp, shared-memory resident and small enough to run at high
// occupancy: Hessian task block, its Cholesky factor, J = L^-T, the unconstrained minimiser and
// the equality normals in the J basis.  Results go to a per-state record in HBM that the
// active-set kernel loads (one coalesced-by-sector read of ~14 KB per Atlas state).
struct SetupLayout {
  int oG, oJd, oev, owt, og0, odm, ozv, odl, oD, om6, ojc, per_state;  // shared memory (doubles)
};

template <class C>
__host__ __device__ inline SetupLayout setup_layout(const DevProblem& P, int min_doubles) {
  SetupLayout L;
  int o = 0;
  // the staged task Jacobians share the factor block: the Hessian tiles are accumulated in
  // registers and only written once every lane is done reading the Jacobians
  int blk = (C::NV * C::LDV + 1) & ~1;
  const int jd = (((P.ndense > 0 ? P.ndense : 1) * C::LDV) + 1) & ~1;
  if (blk < jd) blk = jd;
  L.oG = o; L.oJd = o; o += blk;
  const int tr = ((P.ntaskrows > 0 ? P.ntaskrows : 1) + 1) & ~1;
  // task errors, weights, g0, dm and the reciprocal diagonal are all dead when the D block is
  // built: the six floating-base rows of M it needs are staged over them
  L.om6 = o;
  L.oev = o; o += tr;
  L.owt = o; o += tr;
  L.og0 = o; o += C::NP;
  L.odm = o; o += C::NP;
  L.odl = o; o += C::NP;
  if (o - L.om6 < ((6 * C::NV + 1) & ~1)) o = L.om6 + ((6 * C::NV + 1) & ~1);
  L.ozv = o; o += C::NP;  // x0 stays live until the record is written
  L.oD = 0;  // D goes straight to the record in HBM (coalesced by construction): no shared-memory copy
  if (o < min_doubles) o = min_doubles;  // the image of the active-set kernel's arrays is assembled here at the end
  L.ojc = o; o += C::NC > 0 ? C::NC : 1;  // row pointers of Jc Ni, behind the image: they are used while it is built
  L.per_state = (o + 1) & ~1;
  return L;
}

template <int NV, int NC, int NU, int NT>
__global__ void __launch_bounds__(32)
k_ihwbc_setup(const __grid_constant__ DevProblem P, const DevProblem* __restrict__ gp, SetupLayout SL, WsPtrs ws, FrameMap fm,
              int nfr, int B, SolveIO io, double* __restrict__ rec_all, int* __restrict__ counter) {
  using C = MCfg<NV, NC, NU, NT>;
  constexpr int N = C::N, NP = C::NP, LDV = C::LDV, HW = C::HW, NB = C::NB, LDA = C::LDA, E = C::E;
  constexpr bool EXT = C::EXT;
  extern __shared__ __align__(16) double smem[];
  const DevProblem& G = *gp;
  const int lane = threadIdx.x;
  const int p = P.p;
  const bool has_ic = P.n_ic > 0;
  double* Gm = smem + SL.oG;
  double* Jd = smem + SL.oJd;
  double* ev = smem + SL.oev;
  double* wt = smem + SL.owt;
  double* g0 = smem + SL.og0;
  double* dm = smem + SL.odm;
  double* zv = smem + SL.ozv;
  double* vm = smem + SL.odl;
  double* dv = g0;
  double* xs = zv;  // x0
#pragma unroll 1
  for (;;) {
    int si = 0;
    if (lane == 0) si = atomicAdd(counter, 1);
    si = __shfl_sync(FULL, si, 0);
    if (si >= B) break;
    const size_t s = (size_t)si;
    double* rec = rec_all + s * C::REC;
    double* Dt = rec + C::RD;  // equality normals in the J basis, written in place
    int status = PNC_QP_OK;
    if (has_ic && ws.ic_status[s] != 0) status = PNC_QP_EQ_DEPENDENT;
    const double* Mg = ws.M + s * NV * NV;
    // rows of Jc Ni: the pointer of each row is looked up once per state (contact table, frame
    // slot), so the element reads below are single loads instead of dependent chains
    const double** jcrow = reinterpret_cast<const double**>(smem + SL.ojc);
    if (lane < NC) {
      const int c = lane;
      const double* ptr;
      if (has_ic) {
        ptr = ws.ic_jcni + (s * NC + c) * NV;
      } else {
        int ci = 0;
        while (ci + 1 < P.ncontact && c >= G.contacts[ci + 1].rf_off) ci++;
        ptr = contact_jac_row(G.contacts[ci], fm.contact_f[ci], ws, s, nfr, NV, c - G.contacts[ci].rf_off);
      }
      jcrow[c] = ptr;
    }
    __syncwarp();
    auto row_jc = [&](int c, int v) -> double { return jcrow[c][v]; };  // (Jc Ni)[c][v]
    // ---- tasks: e = Jdot qdot - xddot_cmd, weights (ihwbc.py:159-172) ----
    // task error terms e = Jdot qdot - xddot_cmd were evaluated by k_task_prepare (ihwbc.cu)
    {
      const double* te = ws.task_e + s * P.ntaskrows;
#pragma unroll 1
      for (int k = lane; k < P.ntaskrows; k += 32) ev[k] = te[k];
#pragma unroll 1
      for (int t = 0; t < P.ntask; t++) {
        const DevTask& T = P.tasks[t];
        const double wt_t = io.w[s * P.ntask + t];
        for (int i = lane; i < T.dim; i += 32) wt[T.row_off + i] = wt_t;
      }
    }
    // stage the dense task Jacobian rows (flattened copy: several independent loads in flight per
    // lane instead of one row at a time); clear the factor block
    {
      // row -> source pointer table borrows g0 (written later, by the Hessian phase)
      // (shape_ok checks ndense <= NP)
      const double** rptr = reinterpret_cast<const double**>(g0);
#pragma unroll 1
      for (int t = lane; t < P.ntask; t += 32) {
        const DevTask& T = P.tasks[t];
        if (T.dense_off >= 0)
          for (int i = 0; i < 3; i++) rptr[T.dense_off + i] = dense_row_ptr(T, fm.task_f[t], ws, s, nfr, NV, i);
      }
      __syncwarp();
      const int total = P.ndense * NV;
#pragma unroll 1
      for (int e0 = lane; e0 < total; e0 += 32 * 9) {  // nine independent loads in flight per lane
        double v[9];
#pragma unroll
        for (int u = 0; u < 9; u++) {
          const int e = e0 + 32 * u;
          const int r = e < total ? e / NV : 0, c = e < total ? e - r * NV : 0;
          v[u] = rptr[r][c];
        }
#pragma unroll
        for (int u = 0; u < 9; u++) {
          const int e = e0 + 32 * u;
          if (e < total) { const int r = e / NV, c = e - r * NV; Jd[r * LDV + c] = v[u]; }
        }
      }
      __syncwarp();
    }

    // ---- Hessian task block (lower triangle) = sum_t w_t J_t^T J_t + lambda_q_ddot M ----
    // 2x2 tiles, accumulated in registers: the factor block shares its shared memory with the
    // staged Jacobians and is only written once every lane is done reading them
    {
      const double* Mg = ws.M + s * NV * NV;
      constexpr int nt = (NV + 1) >> 1, ntile = nt * (nt + 1) / 2, TPL = (ntile + 31) / 32;
      double h[TPL][4];
#pragma unroll
      for (int u = 0; u < TPL; u++) {
        const int idx = lane + 32 * u;
        h[u][0] = h[u][1] = h[u][2] = h[u][3] = 0.0;
        if (idx >= ntile) continue;
        int ti = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= idx) ti++;
        while (ti * (ti + 1) / 2 > idx) ti--;
        const int tj = idx - ti * (ti + 1) / 2;
        const int r0 = 2 * ti, c0 = 2 * tj;
        const bool r1ok = r0 + 1 < NV, c1ok = c0 + 1 < NV;
        const double m00 = Mg[r0 * NV + c0], m01 = c1ok ? Mg[r0 * NV + c0 + 1] : 0.0;
        const double m10 = r1ok ? Mg[(r0 + 1) * NV + c0] : 0.0, m11 = (r1ok && c1ok) ? Mg[(r0 + 1) * NV + c0 + 1] : 0.0;
        double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
#pragma unroll 1
        for (int t = 0; t < P.ntask; t++) {
          const DevTask& T = P.tasks[t];
          if (T.dense_off < 0) continue;
          const double wk = wt[T.row_off];
          if (wk == 0.0) continue;
          double t00 = 0, t01 = 0, t10 = 0, t11 = 0;
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const double* row = Jd + (T.dense_off + i) * LDV;
            const double ra = row[r0], rb = r1ok ? row[r0 + 1] : 0.0;
            const double ca = row[c0], cb = c1ok ? row[c0 + 1] : 0.0;
            t00 += ra * ca; t01 += ra * cb; t10 += rb * ca; t11 += rb * cb;
          }
          a00 += wk * t00; a01 += wk * t01; a10 += wk * t10; a11 += wk * t11;
        }
        const double lam = P.lambda_q_ddot;
        h[u][0] = a00 + lam * m00; h[u][1] = a01 + lam * m01; h[u][2] = a10 + lam * m10; h[u][3] = a11 + lam * m11;
      }
#pragma unroll 1
      for (int c = lane; c < NV; c += 32) {
        double acc = 0.0;
#pragma unroll 1
        for (int t = 0; t < P.ntask; t++) {
          const DevTask& T = P.tasks[t];
          if (T.dense_off < 0) continue;
          const double wk = wt[T.row_off];
          if (wk == 0.0) continue;
          double tt = 0.0;
          for (int i = 0; i < 3; i++) tt += ev[T.row_off + i] * Jd[(T.dense_off + i) * LDV + c];
          acc += wk * tt;
        }
        g0[c] = acc;
      }
      __syncwarp();
      // the Jacobians are dead: clear the factor block and store the tiles
#pragma unroll 1
      for (int k = lane; k < NV * LDV; k += 32) Gm[k] = 0.0;
      __syncwarp();
#pragma unroll
      for (int u = 0; u < TPL; u++) {
        const int idx = lane + 32 * u;
        if (idx >= ntile) continue;
        int ti = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= idx) ti++;
        while (ti * (ti + 1) / 2 > idx) ti--;
        const int tj = idx - ti * (ti + 1) / 2;
        const int r0 = 2 * ti, c0 = 2 * tj;
        const bool r1ok = r0 + 1 < NV, c1ok = c0 + 1 < NV;
        Gm[r0 * LDV + c0] = h[u][0];
        if (c1ok && c0 + 1 <= r0) Gm[r0 * LDV + c0 + 1] = h[u][1];
        if (r1ok) {
          Gm[(r0 + 1) * LDV + c0] = h[u][2];
          if (c1ok) Gm[(r0 + 1) * LDV + c0 + 1] = h[u][3];
        }
      }
      __syncwarp();
#pragma unroll 1
      for (int t = 0; t < P.ntask; t++) {  // one-hot rows of the joint tasks
        const DevTask& T = P.tasks[t];
        if (T.dense_off >= 0) continue;
#pragma unroll 1
        for (int i = lane; i < T.dim; i += 32) {
          const int vi = T.type == PNC_TASK_JOINT ? 6 + i : G.joint_sel[T.joint_off + i];
          const double wk = wt[T.row_off + i];
          Gm[vi * LDV + vi] += wk;
          g0[vi] += wk * ev[T.row_off + i];
        }
        __syncwarp();
      }
    }

    // ---- Cholesky of the task block (QuadProg++.cc:705-730), left-looking, lanes = rows ----
    // rows 32.. (Atlas has 4) are each shared by LPR lanes that split the dot product
    double c1 = 0.0, c2 = 0.0;
    {
      double tr = 0.0;
      for (int i = lane; i < NV; i += 32) tr += Gm[i * LDV + i];
      c1 = wsum(tr) + NC * P.lambda_rf;
    }
    double* dl = vm;  // reciprocal diagonal of L
    bool pd = true;
    constexpr int EV = C::EV, LPR = C::LPR;
    const int xrow = 32 + lane / LPR, xsl = lane % LPR;
#pragma unroll 1
    for (int j = 0; j < NV; j++) {
      double t0 = 0.0, t1 = 0.0;
      const double* rj = Gm + j * LDV;
      if (lane >= j && lane < NV) {
        const double* ri = Gm + lane * LDV;
        double s0 = ri[j], s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = 0;
#pragma unroll 1
        for (; k + 3 < j; k += 4) {
          const double2 a0 = lds2(ri + k), a1 = lds2(ri + k + 2), b0 = lds2(rj + k), b1 = lds2(rj + k + 2);
          s0 = fma(-a0.x, b0.x, s0); s1 = fma(-a0.y, b0.y, s1); s2 = fma(-a1.x, b1.x, s2); s3 = fma(-a1.y, b1.y, s3);
        }
#pragma unroll 1
        for (; k < j; k++) s0 = fma(-ri[k], rj[k], s0);
        t0 = (s0 + s1) + (s2 + s3);
      }
      if constexpr (EV > 0) {
        const bool xact = xrow < NV && xrow >= j;
        const double* ri = Gm + (xact ? xrow : 0) * LDV;
        double part = 0.0;
        if (xact) {
#pragma unroll 1
          for (int k = xsl; k < j; k += LPR) part = fma(ri[k], rj[k], part);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
        if (xact) t1 = ri[j] - part;
      }
      const double piv = __shfl_sync(FULL, j < 32 ? t0 : t1, j < 32 ? j : (j - 32) * LPR);
      if (!(piv > 0.0)) { pd = false; break; }
      const double dj = sqrt(piv), idj = 1.0 / dj;
      __syncwarp();
      if (j < 32 ? lane == j : lane == (j - 32) * LPR) { Gm[j * LDV + j] = dj; dl[j] = idj; }
      if (lane > j && lane < NV) Gm[lane * LDV + j] = t0 * idj;
      if (EV > 0 && xsl == 0 && xrow > j && xrow < NV) Gm[xrow * LDV + j] = t1 * idj;
      __syncwarp();
    }
    if (!pd || !(P.lambda_rf > 0.0 || NC == 0)) {
      status = PNC_QP_NOT_PD;
    } else if (status == PNC_QP_OK) {
      // ---- J = L^-T in place (QuadProg++.cc:203-210): lane i forward-solves L z = e_i ----
#pragma unroll 1
      for (int j = 0; j < NV; j++) {
        const double* Lj = Gm + j * LDV;
        const double idjj = dl[j];
#pragma unroll
        for (int rr = 0; rr < (NV > 32 ? 2 : 1); rr++) {
          const int i = lane + 32 * rr;
          if ((rr == 0 || j >= 32) && i <= j && i < NV) {
            double* zi = Gm + i * LDV;
            double s0 = i == j ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = i;
            if ((k & 1) && k < j) { s0 = fma(-Lj[k], zi[k], s0); k++; }
#pragma unroll 1
            for (; k + 3 < j; k += 4) {
              const double2 a0 = lds2(zi + k), a1 = lds2(zi + k + 2), b0 = lds2(Lj + k), b1 = lds2(Lj + k + 2);
              s0 = fma(-a0.x, b0.x, s0); s1 = fma(-a0.y, b0.y, s1); s2 = fma(-a1.x, b1.x, s2); s3 = fma(-a1.y, b1.y, s3);
            }
#pragma unroll 1
            for (; k < j; k++) s0 = fma(-Lj[k], zi[k], s0);
            zi[j] = ((s0 + s1) + (s2 + s3)) * idjj;
          }
        }
        __syncwarp();
      }
      const double jrf = 1.0 / sqrt(P.lambda_rf > 0.0 ? P.lambda_rf : 1.0);
      {
        double tr = 0.0;
        for (int i = lane; i < NV; i += 32) tr += Gm[i * LDV + i];
        c2 = wsum(tr) + NC * jrf;
      }
      // ---- unconstrained minimiser x = -J (J^T g0)  (:222-224); J is upper triangular here ----
#pragma unroll 1
      for (int j = lane; j < NV; j += 32) {
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k <= j; k++) acc += Gm[k * LDV + j] * dv[k];
        dm[j] = acc;
      }
      __syncwarp();
#pragma unroll 1
      for (int k = lane; k < NV; k += 32) {
        double acc = 0.0;
#pragma unroll 1
        for (int j = k; j < NV; j++) acc += Gm[k * LDV + j] * dm[j];
        xs[k] = -acc;
      }
      // ---- equality normals in the J basis, all at once while J = L^-T is still triangular in
      //      shared memory: D[i] = J^T ce_i.  The active-set loop then never needs the dense
      //      d = J^T n+ product for an equality: it reads D[i] and keeps the later rows current
      //      with the Householder reflection of each added equality ----
      {
        // floating-base rows i < 6: D[i][j] = sum_{k<=j} J[k][j] M[i][k]; the six rows are staged in
        // shared memory first (Jd is dead), one column j per lane, six accumulators
        double* m6 = smem + SL.om6;
        __syncwarp();  // x0 is done with the vectors m6 overlays
#pragma unroll 1
        for (int e0 = lane; e0 < 6 * NV; e0 += 32 * 7) {
          double v[7];
#pragma unroll
          for (int u = 0; u < 7; u++) v[u] = e0 + 32 * u < 6 * NV ? Mg[e0 + 32 * u] : 0.0;
#pragma unroll
          for (int u = 0; u < 7; u++)
            if (e0 + 32 * u < 6 * NV) m6[e0 + 32 * u] = v[u];
        }
        __syncwarp();
#pragma unroll
        for (int pass = 0; pass < (NV > 32 ? 2 : 1); pass++) {
          const int j = lane + 32 * pass;
          double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0, e4 = 0.0, e5 = 0.0;
          const int kmax = NV < 32 * (pass + 1) ? NV : 32 * (pass + 1);
          const double* gp_ = Gm + (j < NV ? j : 0);
#pragma unroll 2
          for (int k = 0; k < kmax; k++, gp_ += LDV) {
            const double g = (j < NV && k <= j) ? gp_[0] : 0.0;
            e0 = fma(g, m6[k], e0); e1 = fma(g, m6[NV + k], e1); e2 = fma(g, m6[2 * NV + k], e2);
            e3 = fma(g, m6[3 * NV + k], e3); e4 = fma(g, m6[4 * NV + k], e4); e5 = fma(g, m6[5 * NV + k], e5);
          }
          if (j < NV) {
            Dt[j] = e0; Dt[NP + j] = e1; Dt[2 * NP + j] = e2; Dt[3 * NP + j] = e3; Dt[4 * NP + j] = e4; Dt[5 * NP + j] = e5;
          }
        }
#pragma unroll 1
        for (int c = lane; c < NP - NV; c += 32)
#pragma unroll 1
          for (int i = 0; i < 6; i++) Dt[i * NP + NV + c] = c < NC ? -jrf * row_jc(c, i) : 0.0;
      }
#pragma unroll 1
      for (int i = 6; i < p; i++) {  // internal-constraint rows [Ji | 0]
#pragma unroll 1
        for (int j = lane; j < NP; j += 32) {
          double acc = 0.0;
          if (j < NV) {
#pragma unroll 1
            for (int k = 0; k <= j; k++) acc = fma(Gm[k * LDV + j], P.ic_jac[i - 6][k], acc);
          }
          Dt[i * NP + j] = acc;
        }
      }
      __syncwarp();
      // ---- record, part 1: J = L^-T as the register chunks of the active-set kernel (lane = row,
      //      two columns per 16-byte piece, rows 32.. as half rows of lane pairs), x0, D, scalars ----
      {
        const int eh = lane & 1, erow = 32 + (lane >> 1);
        const bool evalid = EXT && (lane >> 1) < E;
        double2* rj = reinterpret_cast<double2*>(rec + C::RJ);
#pragma unroll 2
        for (int k = 0; k < NP; k += 2) {
          double v0 = 0.0, v1 = 0.0;
          if (k < NV) {
            if (lane < NV) {
              const double2 g = lds2(Gm + lane * LDV + k);
              if (k >= lane) v0 = g.x;
              if (k + 1 >= lane && k + 1 < NV) v1 = g.y;
            }
            if (k + 1 >= NV && k + 1 < N && lane == k + 1) v1 = jrf;
          } else {
            if (k < N && lane == k) v0 = jrf;
            if (k + 1 < N && lane == k + 1) v1 = jrf;
          }
          rj[(k >> 1) * 32 + lane] = make_double2(v0, v1);
        }
        if constexpr (EXT) {
          double2* rb = reinterpret_cast<double2*>(rec + C::RJB);
#pragma unroll 2
          for (int c = 0; c < NB; c += 2) {
            double v[2] = {0.0, 0.0};
#pragma unroll
            for (int u = 0; u < 2; u++) {
              const int col = eh * HW + c + u;
              if (evalid && col < N) {
                if (erow < NV) { if (col < NV && col >= erow) v[u] = Gm[erow * LDV + col]; }
                else if (col == erow) v[u] = jrf;
              }
            }
            rb[(c >> 1) * 32 + lane] = make_double2(v[0], v[1]);
          }
        }
      }
#pragma unroll 1
      for (int k = lane; k < NP; k += 32) rec[C::RX + k] = k < NV ? xs[k] : 0.0;
      if (lane == 0) { rec[C::RS] = c1; rec[C::RS + 1] = c2; rec[C::RS + 3] = jrf; }
      // equality normals [M | -(Jc Ni)^T] rows 0..5, then the internal-constraint rows, and offsets
#pragma unroll 1
      for (int i = 0; i < p; i++) {
#pragma unroll 1
        for (int k = lane; k < NP; k += 32) {
          double val = 0.0;
          if (k < N) {
            if (i >= 6) val = k < NV ? G.ic_jac[i - 6][k] : 0.0;
            else if (k < NV) val = Mg[i * NV + k];
            else val = -row_jc(k - NV, i);
          }
          rec[C::RE + i * NP + k] = val;
        }
      }
      for (int i = lane; i < p; i += 32)
        rec[C::RE0 + i] = i >= 6 ? 0.0 : (has_ic ? ws.ic_cg[s * NV + i] : ws.cori[s * NV + i] + ws.grav[s * NV + i]);
      __syncwarp();
      // ---- record, part 2: the image of the active-set kernel's constant arrays (torque-limit rows
      //      S [M | -(Jc Ni)^T] with their offsets, wrench cones), assembled in shared memory over the
      //      dead factor block and written out in one coalesced copy ----
      {
        double* img = smem;
        double* Ar = img + (C::OAR - C::OAR);
        double* a0 = img + (C::OA0 - C::OAR);
        double* Uf = img + (C::OUF - C::OAR);
        double* uf = img + (C::OUFV - C::OAR);
        const int nact = P.nact;
#pragma unroll 1
        for (int k = lane; k < C::IMG; k += 32) img[k] = 0.0;
        __syncwarp();
#pragma unroll 1
        for (int r = lane; r < NU; r += 32) {
          const int ci = G.cone_contact[r];
          const DevContact& Cc = G.contacts[ci];
          const int lr = r - Cc.cone_off;
          double u6[6];
          cone_row_world(Cc, lr, ws.fr_iso + (s * nfr + fm.contact_f[ci]) * 12, u6);
          for (int k = 0; k < 6; k++) Uf[r * 6 + k] = u6[k];
          double rz = io.rfz[s * P.ncontact + ci];
          if (rz <= 0.0) rz = 1e-3;  // contact.py:37-41
          uf[r] = lr == Cc.rows - 1 ? -rz : 0.0;
        }
        if constexpr (NT > 0) {
          // flattened copy with several independent loads in flight; M is symmetric, so row
          // act_idx[a] is read as column act_idx[a] with consecutive lanes on consecutive addresses
          // element (row a, column k) with a fastest: a == lane throughout, so its joint index is loaded once
          const int va = lane < nact ? G.act_idx[lane] : 0;
#pragma unroll 1
          for (int e0 = lane; e0 < N * 32; e0 += 32 * 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const int e = e0 + 32 * u, k = e >> 5, a = e & 31;
              double val = 0.0;
              if (k < N && a < nact) {
                if (has_ic) val = ws.ic_tq[(s * N + k) * nact + a];
                else if (k < NV) val = Mg[k * NV + va];
                else val = -row_jc(k - NV, va);
              }
              v[u] = val;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const int e = e0 + 32 * u, k = e >> 5, a = e & 31;
              if (k < N && a < nact) Ar[a * LDA + k] = v[u];
            }
          }
          for (int a = lane; a < nact; a += 32) {
            const int v = G.act_idx[a];
            // ci0 of the two torque-limit rows of joint a (ihwbc.py:266-296): S Ni^T (c+g) - trq_lo, trq_hi - S Ni^T (c+g)
            const double a0v = has_ic ? ws.ic_tq0[s * nact + a] : ws.cori[s * NV + v] + ws.grav[s * NV + v];
            a0[a] = a0v - G.trq_lo[a];
            a0[NT + a] = G.trq_hi[a] - a0v;
            a0[2 * NT + a] = a0v;  // for the torque recovery at the end of the active-set kernel
          }
        }
        __syncwarp();
        double2* ri = reinterpret_cast<double2*>(rec + C::RI);
#pragma unroll 4
        for (int k = lane; k < C::IMG / 2; k += 32) ri[k] = lds2(img + 2 * k);
      }
    }
    if (lane == 0) rec[C::RS + 2] = (double)status;
    __syncwarp();
  }
}

// =======================================================================================
template <int NV, int NC, int NU, int NT>
#ifndef PNC_REG_LB
#define PNC_REG_LB __launch_bounds__(256, 1)
#endif
__global__ void PNC_REG_LB
k_ihwbc_reg(const __grid_constant__ DevProblem P, const DevProblem* __restrict__ gp, int B, int* __restrict__ counter,
            int* __restrict__ overflow, double* __restrict__ rec_all, int rc_cap, double refine) {
  // P (constant bank) serves warp-uniform reads; tables indexed per lane go through the global copy G
  const DevProblem& G = *gp;
  using C = MCfg<NV, NC, NU, NT>;
  constexpr int N = C::N, NP = C::NP, HW = C::HW, E = C::E, LDT = C::LDT, NB = C::NB, LDA = C::LDA;
  constexpr bool EXT = C::EXT;
  // One warp per state, each with its own slice of shared memory.  The launcher uses one-warp CTAs;
  // with PNC_REG_WARPS=w the w warps of a CTA meet at a barrier at the top of every step-direction
  // pass, which makes them walk the ~31 KB hot loop together (the SM's 32 KB instruction cache
  // thrashes when eight warps wander through it independently, scripts/icache_probe.cu).  Measured
  // on Atlas 64K the wait for the slowest warp costs more than the shared fetch saves (w = 1/2/4/8:
  // 15.3 / 16.2 / 17.3 / 21.4 ms), so this stays a diagnostic knob.
  extern __shared__ __align__(16) double smem_all[];
  const int lane = threadIdx.x & 31;
  double* const smem = smem_all + (threadIdx.x >> 5) * C::SMEM_DOUBLES;
  int nthr = blockDim.x;  // threads still taking part in the slot barrier
  const int p = P.p, m = C::M, nu = NU;
  constexpr int nact = NT;  // torque-limit rows = actuated joints (shape_ok); only used when NT > 0
  constexpr bool trq = NT > 0;
  constexpr bool S2 = C::RC > 32;  // S = R^-1 never has rows beyond 32 (RC caps the active set): second-half code is dead
  double* const Rp = smem + C::ORT;
  double* const Ts = smem + C::OT;
  double* const Ar = smem + C::OAR;   // torque-limit rows S [M | -(Jc Ni)^T], pitch LDA
  double* const a0 = smem + C::OA0;   // ci0 of the torque-limit rows: [0, NT) lower limits, [NT, 2 NT) upper limits; [2 NT, 3 NT) S Ni^T (c+g)
  double* const Uf = smem + C::OUF;
  double* const uf = smem + C::OUFV;
  double* const x = smem + C::OX;
  double* const zv = smem + C::OZ;
  double* const dv = smem + C::ODV;
  double* const dm = smem + C::ODM;
  double* const np = smem + C::ONP;
  double* const rv = smem + C::ORR;
  double* const uv = smem + C::OU;
  double* const xold = smem + C::OXOLD;
  double* const uold = smem + C::OUOLD;
  double* const xacc = smem + C::OXACC;
  double* const sl = smem + C::OSL;
  int* const Aact = reinterpret_cast<int*>(smem + C::OINT);
  int* Aold = Aact + (N + 2);
  int* Aacc = Aold + (N + 2);
  int* const Cm = Aact + ((3 * (N + 2) + 1) & ~1);  // per cone row: first variable, number of variables (8-byte aligned pairs)
#define Rcol(c) (Rp + ((c) * ((c) + 3)) / 2)
  double* rot = Ts;  // the d scratch tile doubles as the Givens coefficient table of a drop
  static_assert(6 * LDT >= 3 * NP, "rotation table must fit the scratch tile");

  // rows beyond 32: lane pair (2e, 2e+1) owns row 32+e, half eh of the columns
  const int eh = lane & 1, epair = lane >> 1;
  const int erow = 32 + epair;
  const bool evalid = EXT && epair < E;

  double Ja[NP];
  double Jb[NB];
#pragma unroll 1
  for (int r = lane; r < NU; r += 32) {
    const DevContact& Cc = G.contacts[G.cone_contact[r]];
    Cm[2 * r] = NV + Cc.rf_off;
    Cm[2 * r + 1] = Cc.dim;
  }
  if (lane < 6) np[NP + lane] = 0.0;  // never written again
  __syncwarp();

#pragma unroll 1
  for (;;) {
    int si = 0;
    if (lane == 0) si = atomicAdd(counter, 1);
    si = __shfl_sync(FULL, si, 0);
    if (si >= B) {  // no work left: leave the barrier group
      slot_barrier(nthr, false);
      break;
    }
    const size_t s = (size_t)si;
    int status = PNC_QP_OK, iter = 0, iq = 0;
    int iq_ref = -1;  // size of the active set at the reference's own termination point (-1: same as the final one)
    double R_norm = 1.0;
    double* rec = rec_all + s * C::REC;  // written by k_ihwbc_setup
    {
      // torque-limit rows, their offsets and the wrench cones arrive in their final layout
      const double2* src = reinterpret_cast<const double2*>(rec + C::RI);
      double2* dst = reinterpret_cast<double2*>(smem + C::OAR);
#pragma unroll 4
      for (int k = lane; k < C::IMG / 2; k += 32) dst[k] = __ldg(src + k);
#pragma unroll 1
      for (int k = lane; k < NP; k += 32) {
        x[k] = rec[C::RX + k];
        zv[k] = 0.0; dm[k] = 0.0; np[k] = 0.0; dv[k] = 0.0;
      }
#pragma unroll 1
      for (int k = lane; k < 6 * LDT; k += 32) Ts[k] = 0.0;
    }
    const double c1 = rec[C::RS], c2 = rec[C::RS + 1], jrf = rec[C::RS + 3];
    status = (int)rec[C::RS + 2];
    if (status == PNC_QP_OK) {
      // ---- move J into registers: row lane, and half rows 32.. in lane pairs ----
      {
        const double2* rj = reinterpret_cast<const double2*>(rec + C::RJ) + lane;
#pragma unroll
        for (int k = 0; k < NP; k += 2) {
          const double2 g = __ldg(rj + (k >> 1) * 32);
          Ja[k] = g.x;
          Ja[k + 1] = g.y;
        }
        if constexpr (EXT) {
          const double2* rb = reinterpret_cast<const double2*>(rec + C::RJB) + lane;
#pragma unroll
          for (int c = 0; c < NB; c += 2) {
            const double2 g = __ldg(rb + (c >> 1) * 32);
            Jb[c] = g.x;
            Jb[c + 1] = g.y;
          }
        } else {
#pragma unroll
          for (int c = 0; c < NB; c++) Jb[c] = 0.0;
        }
      }
      double* Deq = Rp + 64;  // R only uses its first p columns (< 64 doubles) while Deq is live
#pragma unroll 2
      for (int k = lane; k < p * NP; k += 32) Deq[k] = rec[C::RD + k];
#pragma unroll 1
      for (int k = lane; k < N + 2; k += 32) { uv[k] = 0.0; Aact[k] = 0; }
      __syncwarp();

      // ================= dual active-set iterations, equalities first (:232-500) =================
      // Per-lane bit masks: lane l owns the inequalities l, l + 32, l + 64, l + 96 (the ones it scans in
      // l2), bit t of its mask speaks for inequality l + 32 t.
      static_assert(C::M <= 128, "four inequalities per lane");
      unsigned actm = 0u;  // in the active set
      unsigned excm = 0u;  // excluded in this outer iteration
      auto set_bit = [&](unsigned& msk, int i) {
        if (lane == (i & 31)) msk |= 1u << (i >> 5);
      };
      auto clr_bit = [&](unsigned& msk, int i) {
        if (lane == (i & 31)) msk &= ~(1u << (i >> 5));
      };
      const double psi_tol = (double)m * QP_EPS * c1 * c2 * 100.0;  // QuadProg++.cc:306-312
      const double psi_refine = psi_tol * refine;                   // refine = 1e-6 by default, see ihwbc.cu / DESIGN.md 3.2
      bool done = false, accepted = false, need_l1 = true;
      int iq_acc = 0, ieq = 0;

#pragma unroll 1
      while (!done) {
        const bool is_eq = ieq < p;
        int ip = 0, k0 = 0, k1 = N;
        double s_ip = 0.0;
        if (is_eq) {
          ip = ieq;
        } else {
          if (m <= 0) break;
          if (need_l1) {
            // l1: slacks of every inequality, termination test (:282-321)
            need_l1 = false;
            iter++;
            excm = 0u;
            double psi = 0.0;
            if (trq) {
              static_assert(NT <= 32, "one torque-limit row per lane");
              if (lane < nact) {
                const int a = lane;
                const double* row = Ar + a * LDA;
                double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
                for (int k = 0; k < NP; k += 4) {
                  const double2 r0 = lds2(row + k), x0 = lds2(x + k);
                  t0 = fma(r0.x, x0.x, t0); t1 = fma(r0.y, x0.y, t1);
                  if (k + 2 < NP) {
                    const double2 r1 = lds2(row + k + 2), x1 = lds2(x + k + 2);
                    t2 = fma(r1.x, x1.x, t2); t3 = fma(r1.y, x1.y, t3);
                  }
                }
                const double dot = (t0 + t1) + (t2 + t3);
                const double slo = dot + a0[a], shi = a0[NT + a] - dot;  // CI^T x + ci0 (QuadProg++.cc:293-301)
                sl[nu + a] = slo;
                sl[nu + nact + a] = shi;
                psi += fmin(0.0, slo) + fmin(0.0, shi);
              }
            }
#pragma unroll 1
            for (int r = lane; r < nu; r += 32) {
              const int2 cm = *reinterpret_cast<const int2*>(Cm + 2 * r);
              const double* xr = x + cm.x;
              const double* ur = Uf + r * 6;
              double acc = -uf[r];
#pragma unroll
              for (int k = 0; k < 6; k++) acc = fma(ur[k], xr[k], acc);  // ur is zero beyond the contact's dimension
              sl[r] = acc;
              psi += fmin(0.0, acc);
            }
            psi = wsum(psi);
            __syncwarp();
            if (fabs(psi) <= psi_tol) {
              if (fabs(psi) <= psi_refine) break;
              if (!accepted) {
                accepted = true;
                iq_acc = iq;
                for (int i = lane; i < NP; i += 32) xacc[i] = x[i];
                if (lane < iq) Aacc[lane] = Aact[lane];
                __syncwarp();
              }
            }
            if (iter > 40 * (N + m)) { status = PNC_QP_MAX_ITER; break; }
            if (lane < iq) { uold[lane] = uv[lane]; Aold[lane] = Aact[lane]; }  // iq <= RC = 32
            for (int i = lane; i < NP; i += 32) xold[i] = x[i];
            __syncwarp();
          }
          // l2: most violated constraint, first index on ties (:323-334)
          double ss = 0.0;
          ip = 0x7fffffff;
          {
            unsigned blocked = actm | excm;
#pragma unroll
            for (int t4 = 0; t4 < (C::M + 31) / 32; t4++) {  // at most four inequalities per lane
              const int i = lane + 32 * t4;
              if (i < m) {
                const double v = sl[i];
                if (v < ss && !((blocked >> t4) & 1u)) { ss = v; ip = i; }
              }
            }
          }
          wargmin(ss, ip);
          if (ss >= 0.0) break;
          s_ip = ss;
        }
        // ---- the constraint normal np and the rows [k0, k1) where it is non-zero ----
        {
          if (is_eq || ip >= nu) {
            int a = 0;
            double sg = 1.0;
            const bool tq = !is_eq;
            if (tq) { const int tr = ip - nu; a = tr >= nact ? tr - nact : tr; sg = tr >= nact ? -1.0 : 1.0; }
            const double* row = tq ? Ar + a * LDA : rec + C::RE + ip * NP;  // shared or global
            double pnx = 0.0;
#pragma unroll 1
            for (int k = lane; k < NP; k += 32) {
              const double val = k < N ? sg * row[k] : 0.0;
              np[k] = val;
              pnx = fma(val, x[k], pnx);
            }
            if (is_eq) s_ip = wsum(pnx) + rec[C::RE0 + ip];
          } else {
            k0 = Cm[2 * ip]; k1 = k0 + Cm[2 * ip + 1];
#pragma unroll 1
            for (int k = lane; k < NP; k += 32) np[k] = (k >= k0 && k < k1) ? Uf[ip * 6 + (k - k0)] : 0.0;
          }
          if (lane == 0) { uv[iq] = 0.0; Aact[iq] = is_eq ? -ip - 1 : ip; }
          __syncwarp();
        }

        // ---- l2a: step directions, step lengths, step; repeat while constraints are dropped ----
#pragma unroll 1
        for (;;) {
          nthr = slot_barrier(nthr, true);
          // d = J^T np through the scratch tile, six rows of J at a time (:504-516)
          double dA = 0.0, dB = 0.0;
          if (is_eq) {
            if (lane < NP) dA = Deq[ip * NP + lane];
            if (NP > 32 && lane + 32 < NP) dB = Deq[ip * NP + lane + 32];
          } else
#pragma unroll 1
          for (int c0 = k0; c0 < k1; c0 += 6) {
            const int dim = min(6, k1 - c0);
            if (c0 < 32) {
              const int r = lane - c0;
              if (r >= 0 && r < dim) {
                double* dst = Ts + r * LDT;
#pragma unroll
                for (int j = 0; j < NP; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(Ja[j], Ja[j + 1]);
              }
            }
            if constexpr (EXT) {
              if (c0 + dim > 32) {
                const int r = erow - c0;
                if (evalid && r >= 0 && r < dim) {
                  double* dst = Ts + r * LDT + eh * HW;
#pragma unroll
                  for (int c = 0; c < HW; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(Jb[c], Jb[c + 1]);
                }
              }
            }
            __syncwarp();
            {
              const double* npc = np + c0;
              const double* tl = Ts + lane;
#pragma unroll
              for (int r = 0; r < 6; r++) {
                const double nr = npc[r];  // rows beyond the normal's support are zero (np is zero-padded)
                dA = fma(nr, tl[r * LDT], dA);
                if (NP > 32 && lane + 32 < NP) dB = fma(nr, tl[r * LDT + 32], dB);
              }
            }
            __syncwarp();
          }
          if (lane < NP) { dv[lane] = dA; dm[lane] = lane >= iq ? dA : 0.0; }
          if (NP > 32 && lane + 32 < NP) { dv[lane + 32] = dB; dm[lane + 32] = lane + 32 >= iq ? dB : 0.0; }
          __syncwarp();

          // z = J[:, iq:] d[iq:] (:518-528), z.z and z.np
          double za, zb = 0.0, zz, znp, sig2;
          {
            double a0_ = 0.0, a1_ = 0.0, a2_ = 0.0, a3_ = 0.0;
#pragma unroll
            for (int j = 0; j < NP; j += 4) {
              const double2 v = lds2(dm + j);
              a0_ = fma(Ja[j], v.x, a0_);
              a1_ = fma(Ja[j + 1], v.y, a1_);
              if (j + 2 < NP) {
                const double2 w = lds2(dm + j + 2);
                a2_ = fma(Ja[j + 2], w.x, a2_);
                a3_ = fma(Ja[j + 3], w.y, a3_);
              }
            }
            za = (a0_ + a1_) + (a2_ + a3_);
            if constexpr (EXT) {
              double b0_ = 0.0, b1_ = 0.0;
              const double* dh = dm + eh * HW;
#pragma unroll
              for (int c = 0; c < HW; c += 2) {
                const double2 v = lds2(dh + c);
                b0_ = fma(Jb[c], v.x, b0_);
                b1_ = fma(Jb[c + 1], v.y, b1_);
              }
              zb = b0_ + b1_;
              zb += __shfl_xor_sync(FULL, zb, 1);
            }
            double pzz = 0.0, pzn = 0.0;
            double psg = (lane >= iq && lane < N) ? dA * dA : 0.0;   // ||d[iq:]||^2 for the Householder step
            if (lane + 32 >= iq && lane + 32 < N) psg = fma(dB, dB, psg);
            if (lane < N) { zv[lane] = za; pzz = za * za; pzn = za * np[lane]; }
            if (evalid && eh == 0) { zv[erow] = zb; pzz = fma(zb, zb, pzz); pzn = fma(zb, np[erow], pzn); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              pzz += __shfl_xor_sync(FULL, pzz, o);
              pzn += __shfl_xor_sync(FULL, pzn, o);
              psg += __shfl_xor_sync(FULL, psg, o);
            }
            zz = pzz; znp = pzn; sig2 = psg;
          }
          // r = R^-1 d[0:iq] (:530-542).  The kernel keeps S = R^-1 (upper triangular, packed by
          // columns) instead of R: r = S d is a matrix-vector product -- lane i owns row i -- not a
          // back substitution with its dependent chain of iq steps
          {
            static_assert(!S2, "rows of S beyond 32 are not implemented");
            double acc0 = 0.0, acc1 = 0.0;  // even / odd columns: two independent FMA chains
            const double* col = Rp;
            int j = 0;
#pragma unroll 1
            for (; j + 1 < iq; j += 2) {
              const double2 dj = lds2(dv + j);
              const double* cn = col + j + 2;
              if (lane <= j) acc0 = fma(col[lane], dj.x, acc0);
              if (lane <= j + 1) acc1 = fma(cn[lane], dj.y, acc1);
              col = cn + j + 3;
            }
            if (j < iq && lane <= j) acc0 = fma(col[lane], dv[j], acc0);
            if (lane < iq) rv[lane] = acc0 + acc1;
          }
          __syncwarp();
          // step lengths (:366-393); equalities always take the full step (:246-258)
          double t1 = INFINITY;
          int kk = 0x7fffffff;
          if (!is_eq) {
            if (lane >= p && lane < iq) {  // iq <= RC = 32: one candidate per lane
              const double rk = rv[lane];
              if (rk > 0.0) { t1 = uv[lane] / rk; kk = lane; }
            }
            wargmin(t1, kk);
          }
          const int l = kk != 0x7fffffff ? Aact[kk] : 0;
          double t2;
          if (fabs(zz) > QP_EPS) {
            t2 = -s_ip / znp;
            if (!is_eq && t2 < 0.0) t2 = INFINITY;
          } else {
            t2 = is_eq ? 0.0 : INFINITY;
          }
          const double t = fmin(t1, t2);
          if (t >= INFINITY) { status = PNC_QP_INFEASIBLE; done = true; break; }  // :402-407
          const bool dual_only = t2 >= INFINITY;
          if (!dual_only)
            for (int k = lane; k < N; k += 32) x[k] += t * zv[k];
          if (lane < iq) uv[lane] -= t * rv[lane];
          if (lane == 0) uv[iq] += t;
          __syncwarp();

          if (!dual_only && (is_eq || fabs(t - t2) < QP_EPS)) {
            // ---- full step: add the constraint (:441-474), one Householder reflection of J2 ----
            const double sigma = sqrt(sig2);
            if (sigma <= QP_EPS * R_norm) {  // linearly dependent on the active set
              if (is_eq) { status = PNC_QP_EQ_DEPENDENT; done = true; break; }
              set_bit(excm, ip);
              if (lane >= p && lane < iq) { Aact[lane] = Aold[lane]; uv[lane] = uold[lane]; }
              for (int i = lane; i < NP; i += 32) x[i] = xold[i];
              __syncwarp();
              actm = 0u;
#pragma unroll 1
              for (int i = p; i < iq; i++) set_bit(actm, Aact[i]);
              break;  // back to l2
            }
            if (iq >= rc_cap) {  // S is capped at RC columns: hand the state to the generic kernel
              status = PNC_QP_OVERFLOW;
              if (lane == 0) *overflow = 1;
              done = true;
              break;
            }
            const double d0 = dv[iq];
            const double alpha = d0 > 0.0 ? -sigma : sigma;
            const double v0 = d0 - alpha;
            const double beta = 1.0 / (sigma * (sigma + fabs(d0)));
            // R gains the column [d1; alpha]  <=>  S = R^-1 gains the column [-S d1 / alpha; 1 / alpha]
            double* col = Rcol(iq);
            const double ialpha = 1.0 / alpha;
            if (lane < iq) col[lane] = -rv[lane] * ialpha;
            if (lane == 0) col[iq] = ialpha;
            // the Householder vector v = [0 .. 0, v0, d[iq+1:]] is the masked d with one entry patched
            if (lane == 0) dm[iq] = v0;
            const double ca = reg_pick(Ja, iq);  // old column iq of J
            double hwb = 0.0;
            if constexpr (EXT) {
              const int hq = iq >= HW ? 1 : 0;
              double cb = reg_pick(Jb, iq - hq * HW);
              cb = __shfl_sync(FULL, cb, (lane & ~1) | hq);
              hwb = beta * (zb - alpha * cb);
            }
            const double hwa = beta * (za - alpha * ca);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NP; j += 2) {
              const double2 v = lds2(dm + j);
              Ja[j] = fma(-hwa, v.x, Ja[j]);
              Ja[j + 1] = fma(-hwa, v.y, Ja[j + 1]);
            }
            if constexpr (EXT) {
              const double* vh = dm + eh * HW;
#pragma unroll
              for (int c = 0; c < HW; c += 2) {
                const double2 v = lds2(vh + c);
                Jb[c] = fma(-hwb, v.x, Jb[c]);
                Jb[c + 1] = fma(-hwb, v.y, Jb[c + 1]);
              }
            }
            if (is_eq) {  // D[i'] <- (I - beta v v^T) D[i'] for the equalities still to come
#pragma unroll 1
              for (int i2 = ieq + 1; i2 < p; i2++) {
                double* Di = Deq + i2 * NP;
                double e0 = 0.0, e1 = 0.0, w0 = 0.0, w1 = 0.0;
                if (lane < NP) { e0 = Di[lane]; w0 = dm[lane]; }
                if (NP > 32 && lane + 32 < NP) { e1 = Di[lane + 32]; w1 = dm[lane + 32]; }
                const double g = beta * wsum(fma(e0, w0, e1 * w1));
                if (lane < NP) Di[lane] = fma(-g, w0, e0);
                if (NP > 32 && lane + 32 < NP) Di[lane + 32] = fma(-g, w1, e1);
              }
            }
            iq++;
            R_norm = fmax(R_norm, sigma);
            if (is_eq) ieq++;
            else { set_bit(actm, ip); need_l1 = true; }
            __syncwarp();
            break;
          }

          // ---- drop constraint l at position kk (:409-423, :476-499; delete_constraint :610-679) ----
          // With S = R^-1 the Givens sweep is driven by row kk of S: rotating the column pairs
          // (j, j+1), j = kk .. iq-2, so that this row ends up in the last column makes
          // S Q^T (minus row kk and the last column) the inverse of the re-triangularised R, with the
          // same Q that QuadProg++ applies to the columns of J.
          clr_bit(actm, l);
          {
            double carry = Rcol(kk)[kk];
#pragma unroll 1
            for (int j = kk; j < iq - 1; j++) {
              double* cj = Rcol(j);
              double* cn = Rcol(j + 1);
              const double a = carry, b = cn[kk];
              const double hh = qp_distance(a, b);
              const bool skip = fabs(hh) < QP_EPS;
              double cc = 1.0, ss = 0.0, xny = 0.0;
              if (!skip) {
                const double sgn = b >= 0.0 ? 1.0 : -1.0;
                cc = sgn * b / hh;
                ss = -sgn * a / hh;
                xny = ss / (1.0 + cc);
                carry = -sgn * hh;
              } else {
                carry = b;
              }
              const double t1a = lane <= j ? cj[lane] : 0.0, t2a = lane <= j + 1 ? cn[lane] : 0.0;
              double t1b = 0.0, t2b = 0.0;
              if (S2) { t1b = lane + 32 <= j ? cj[lane + 32] : 0.0; t2b = lane + 32 <= j + 1 ? cn[lane + 32] : 0.0; }
              __syncwarp();
              const double nja = skip ? t1a : t1a * cc + t2a * ss, nna = skip ? t2a : xny * (t1a + nja) - t2a;
              // column j is final: store it without row kk; column j+1 keeps its rows until its turn
              if (lane <= j + 1 && lane != kk) cj[lane - (lane > kk ? 1 : 0)] = nja;
              if (lane <= j + 1) cn[lane] = nna;
              if (S2) {
                const double njb = skip ? t1b : t1b * cc + t2b * ss, nnb = skip ? t2b : xny * (t1b + njb) - t2b;
                const int i = lane + 32;
                if (i <= j + 1 && i != kk) cj[i - (i > kk ? 1 : 0)] = njb;
                if (i <= j + 1) cn[i] = nnb;
              }
              if (lane == 0) { rot[3 * j] = skip ? 2.0 : cc; rot[3 * j + 1] = ss; rot[3 * j + 2] = xny; }
              __syncwarp();
            }
          }
          {  // A[i], u[i] <- A[i+1], u[i+1] for i = kk .. iq-1 (the candidate at iq moves down too)
            int a0_ = 0, a1_ = 0;
            double u0_ = 0.0, u1_ = 0.0;
            const bool m0 = lane >= kk && lane < iq, m1 = S2 && lane + 32 >= kk && lane + 32 < iq;
            if (m0) { a0_ = Aact[lane + 1]; u0_ = uv[lane + 1]; }
            if (m1) { a1_ = Aact[lane + 33]; u1_ = uv[lane + 33]; }
            __syncwarp();
            if (m0) { Aact[lane] = a0_; uv[lane] = u0_; }
            if (m1) { Aact[lane + 32] = a1_; uv[lane + 32] = u1_; }
            if (lane == 0) { Aact[iq] = 0; uv[iq] = 0.0; }
          }
          iq--;
          __syncwarp();
          if (iq > kk) {
            // the same rotations on the columns of J (registers)
            reg_rot_range(Ja, kk, iq, rot);
            if constexpr (EXT) {
              // half 0 holds columns [0, HW), half 1 columns [HW, 2 HW); pair (HW-1, HW) straddles
              if (eh == 0) reg_rot_range(Jb, kk, min(iq, HW - 1), rot);
              if (kk <= HW - 1 && iq > HW - 1) {
                const double cc = rot[3 * (HW - 1)], ss = rot[3 * (HW - 1) + 1], xny = rot[3 * (HW - 1) + 2];
                const double mine = eh == 0 ? Jb[HW - 1] : Jb[0];
                const double other = __shfl_xor_sync(FULL, mine, 1);
                const double t1r = eh == 0 ? mine : other, t2r = eh == 0 ? other : mine;
                const double nj = t1r * cc + t2r * ss;
                if (cc <= 1.5) {
                  if (eh == 0) Jb[HW - 1] = nj;
                  else Jb[0] = xny * (nj + t1r) - t2r;
                }
              }
              if (eh == 1 && iq > HW) reg_rot_range(Jb, max(kk, HW) - HW, iq - HW, rot + 3 * HW);
            }
            __syncwarp();
          }
          if (!dual_only) {  // partial step: refresh the slack of ip (:494-499)
            double acc = 0.0;
            for (int k = lane; k < N; k += 32) acc = fma(np[k], x[k], acc);
            acc = wsum(acc);
            double ci0;
            if (ip < nu) ci0 = -uf[ip];
            else ci0 = a0[ip - nu];  // lower-limit rows first, then upper-limit rows, as in sl[]
            s_ip = acc + ci0;
            if (lane == 0) sl[ip] = s_ip;
            __syncwarp();
          }
        }
      }
      if (accepted) iq_ref = iq_acc;
      if (accepted && status != PNC_QP_OK) {  // refinement failed: fall back to the accepted iterate
        status = PNC_QP_OK;
        iq = iq_acc;
        for (int i = lane; i < NP; i += 32) x[i] = xacc[i];
        if (lane < iq_acc) Aact[lane] = Aacc[lane];
        __syncwarp();
      }
    }

    // ---- result back into the record; k_ihwbc_finish turns it into the outputs ----
    {
      int* ro = reinterpret_cast<int*>(rec + C::RO + NP);
#pragma unroll 1
      for (int k = lane; k < NP; k += 32) rec[C::RO + k] = x[k];
      if (lane < iq) ro[3 + lane] = Aact[lane];
      if (lane == 0) { ro[0] = iq; ro[1] = iter; ro[2] = status; ro[3 + N + 2] = iq_ref; }
      if (lane < iq_ref) ro[4 + N + 2 + lane] = Aacc[lane];
      if constexpr (NT > 0) {
        // torque recovery tau = S (M qddot + Ni^T (c+g) - (Jc Ni)^T rf) (ihwbc.py:344-354): the rows
        // S [M | -(Jc Ni)^T] are the torque-limit rows already in shared memory
        if (lane < nact) {
          const double* row = Ar + lane * LDA;
          double t0 = 0.0, t1 = 0.0;
#pragma unroll 1
          for (int k = 0; k < NP; k += 2) {
            const double2 r0 = lds2(row + k), x0 = lds2(x + k);
            t0 = fma(r0.x, x0.x, t0);
            t1 = fma(r0.y, x0.y, t1);
          }
          rec[C::RT + lane] = (t0 + t1) + a0[2 * NT + lane];
        }
      }
    }
    __syncwarp();
  }
}

// Outputs of one state from the result the active-set kernel left in its record: x = [qddot; rf],
// the actuated accelerations, the torque recovery tau = S (M qddot + Ni^T (c + g) - (Jc Ni)^T rf)
// (ihwbc.py:339-356), status, iteration count and the active-set bit mask.  One warp per state,
// lanes = actuated rows: M is symmetric, so lane a reads column act_idx[a] of M and consecutive
// lanes touch consecutive addresses.  States the register kernel handed over (OVERFLOW) keep that
// status here; the generic kernel's second pass overwrites them.
__global__ void __launch_bounds__(256)
k_ihwbc_finish(const DevProblem* __restrict__ gp, WsPtrs ws, FrameMap fm, int nfr, int B, SolveIO io,
               const double* __restrict__ rec_all, int rec_stride, int ro_off, int np_pad, int rt_off, int ref_off) {
  const DevProblem& G = *gp;
  const int lane = threadIdx.x & 31;
  const size_t s = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= (size_t)B) return;
  const int nv = G.nv, nc = G.nc, n = nv + nc, nact = G.nact, p = G.p, m = G.m;
  const double* x = rec_all + s * rec_stride + ro_off;
  const int* ro = reinterpret_cast<const int*>(x + np_pad);
  const int iq = ro[0], iter = ro[1], status = ro[2];
  const bool bad = status != PNC_QP_OK && status != PNC_QP_MAX_ITER;
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  if (io.qddot)
    for (int k = lane; k < nv; k += 32) io.qddot[s * nv + k] = bad ? nanv : x[k];
  if (io.rf)
    for (int k = lane; k < nc; k += 32) io.rf[s * nc + k] = bad ? nanv : x[nv + k];
  if (io.acc)
    for (int a = lane; a < nact; a += 32) io.acc[s * nact + a] = bad ? nanv : x[G.act_idx[a]];
  if (io.trq) {
    for (int a = lane; a < nact; a += 32) {
      double tau = nanv;
      if (!bad && rt_off >= 0) {
        tau = rec_all[s * rec_stride + rt_off + a];  // recovered by the active-set kernel from its torque-limit rows
      } else if (!bad) {
        double t0 = 0.0, t1 = 0.0;
        if (G.n_ic > 0) {
          const double* col = ws.ic_tq + s * n * nact + a;
          int k = 0;
          for (; k + 1 < n; k += 2) { t0 = fma(col[k * nact], x[k], t0); t1 = fma(col[(k + 1) * nact], x[k + 1], t1); }
          if (k < n) t0 = fma(col[k * nact], x[k], t0);
          tau = t0 + t1 + ws.ic_tq0[s * nact + a];
        } else {
          const int v = G.act_idx[a];
          const double* col = ws.M + s * nv * nv + v;  // column v of M == row v
          int k = 0;
          for (; k + 1 < nv; k += 2) { t0 = fma(col[k * nv], x[k], t0); t1 = fma(col[(k + 1) * nv], x[k + 1], t1); }
          if (k < nv) t0 = fma(col[k * nv], x[k], t0);
          const double* xr = x + nv;
          for (int ci = 0; ci < G.ncontact; ci++) {
            const int dim = G.contacts[ci].dim;
            const double* jr = ws.fr_jac + ((s * nfr + fm.contact_f[ci]) * 6 + (6 - dim)) * nv + v;  // point contacts: rows 3..5
            for (int r = 0; r < dim; r++) t1 = fma(-jr[r * nv], xr[r], t1);
            xr += dim;
          }
          tau = t0 + t1 + ws.cori[s * nv + v] + ws.grav[s * nv + v];
        }
      }
      io.trq[s * nact + a] = tau;
    }
  }
  if (io.active) {
    const int mw = (m + 63) / 64 > 0 ? (m + 63) / 64 : 1;
    for (int wdx = lane; wdx < mw; wdx += 32) {
      unsigned long long bits = 0ull;
      for (int i = p; i < iq; i++) {
        const int c = ro[3 + i];
        if (c >= 0 && (c >> 6) == wdx) bits |= 1ull << (c & 63);
      }
      io.active[s * mw + wdx] = bits;
    }
  }
  if (io.active_ref) {
    // the set QuadProg++ itself would have returned: the first iterate that met its bound (:306-312)
    const int mw = (m + 63) / 64 > 0 ? (m + 63) / 64 : 1;
    const int iqr = ro[ref_off];
    const int* ar = iqr >= 0 ? ro + ref_off + 1 : ro + 3;
    const int nr = iqr >= 0 ? iqr : iq;
    for (int wdx = lane; wdx < mw; wdx += 32) {
      unsigned long long bits = 0ull;
      for (int i = p; i < nr; i++) {
        const int c = ar[i];
        if (c >= 0 && (c >> 6) == wdx) bits |= 1ull << (c & 63);
      }
      io.active_ref[s * mw + wdx] = bits;
    }
  }
  if (lane == 0) {
    io.status[s] = status;
    if (io.iters) io.iters[s] = iter;
  }
}

template <int NV, int NC, int NU, int NT>
static bool shape_ok(const DevProblem& hp) {
  using C = MCfg<NV, NC, NU, NT>;
  if (hp.nv != NV || hp.nc != NC || hp.nu != NU || hp.m != C::M) return false;
  if ((hp.b_trq_limit ? hp.nact : 0) != NT || hp.nact > 32 || hp.m > 128 || hp.p > C::PMAX) return false;
  // the equality block D sits behind the first S columns while the equalities are added
  if (64 + hp.p * C::NP > C::RSZ || hp.p * (hp.p + 3) / 2 > 64 || hp.p >= C::RC) return false;
  if (hp.ndense > C::NP) return false;  // the setup kernel keeps one row pointer per dense task row in an NP-sized vector
  return true;
}

template <int NV, int NC, int NU, int NT>
static bool try_launch(const DevProblem* dp, const DevProblem& hp, const WsPtrs& ws, const FrameMap& fm, int nfr, int B,
                       const SolveIO& io, int num_sms, cudaStream_t stream) {
  if (!shape_ok<NV, NC, NU, NT>(hp)) return false;
  using C = MCfg<NV, NC, NU, NT>;
  static const int warps_env = [] { const char* e = getenv("PNC_REG_WARPS"); return e ? atoi(e) : 0; }();
  int warps = warps_env >= 1 && warps_env <= 8 ? warps_env : 1;  // warps (= states in flight) per CTA
  while (warps > 1 && (size_t)warps * C::SMEM_DOUBLES * sizeof(double) > 227 * 1024) warps--;
  const size_t smem = (size_t)warps * C::SMEM_DOUBLES * sizeof(double);
  static SmemOptIn optin;
  if (!optin.ensure(k_ihwbc_reg<NV, NC, NU, NT>, smem)) return false;
  const SetupLayout SL = setup_layout<C>(hp, C::IMG);
  if (!ws.setup_rec || ws.setup_stride != C::REC) return false;
  const size_t smem_s = (size_t)SL.per_state * sizeof(double);
  static SmemOptIn optin_s;
  if (!optin_s.ensure(k_ihwbc_setup<NV, NC, NU, NT>, smem_s)) return false;
  int per_sm_s = 0, per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, k_ihwbc_setup<NV, NC, NU, NT>, 32, smem_s);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ihwbc_reg<NV, NC, NU, NT>, 32 * warps, smem);
  if (per_sm < 1 || per_sm_s < 1) return false;
  static const int cap = [] { const char* e = getenv("PNC_REG_CTAS"); return e ? atoi(e) : 0; }();
  if (cap > 0 && per_sm > cap) per_sm = cap;
  int grid_s = num_sms * per_sm_s, grid = num_sms * per_sm;
  if (grid_s > B) grid_s = B;
  if (grid * warps > B) grid = (B + warps - 1) / warps;
  cudaMemsetAsync(io.counter + 2, 0, sizeof(int), stream);
  k_ihwbc_setup<NV, NC, NU, NT><<<grid_s, 32, smem_s, stream>>>(hp, dp, SL, ws, fm, nfr, B, io, ws.setup_rec, io.counter + 2);
  g_launch_count++;
  // PNC_DEBUG_RC lowers the cap on active constraints so the tests can exercise the hand-over
  static const int rc_dbg = [] { const char* e = getenv("PNC_DEBUG_RC"); return e ? atoi(e) : 0; }();
  const int rc_cap = (rc_dbg > hp.p && rc_dbg < C::RC) ? rc_dbg : C::RC;
  k_ihwbc_reg<NV, NC, NU, NT><<<grid, 32 * warps, smem, stream>>>(hp, dp, B, io.counter, io.overflow, ws.setup_rec, rc_cap,
                                                                  io.psi_refine);
  g_launch_count++;
  k_ihwbc_finish<<<(B + 7) / 8, 256, 0, stream>>>(dp, ws, fm, nfr, B, io, ws.setup_rec, C::REC, C::RO, C::NP, NT > 0 ? C::RT : -1,
                                                  3 + C::N + 2);
  g_launch_count++;
  return true;
}

// register-resident kernels exist for the QP shapes of the shipped robots; anything else takes
// the generic shared-memory kernel of ihwbc.cu
bool launch_ihwbc_solve_reg(const DevProblem* dp, const DevProblem& hp, const WsPtrs& ws, const FrameMap& fm, int nfr,
                            int B, const SolveIO& io, int num_sms, cudaStream_t stream) {
  // <nv, nc, cone rows, torque-limit rows>
  return try_launch<36, 12, 36, 30>(dp, hp, ws, fm, nfr, B, io, num_sms, stream) ||  // Atlas
         try_launch<18, 12, 24, 0>(dp, hp, ws, fm, nfr, B, io, num_sms, stream) ||   // Laikago
         try_launch<33, 12, 36, 0>(dp, hp, ws, fm, nfr, B, io, num_sms, stream) ||   // Draco3
         try_launch<20, 12, 36, 12>(dp, hp, ws, fm, nfr, B, io, num_sms, stream) ||  // Draco3 lower body
         try_launch<32, 12, 36, 26>(dp, hp, ws, fm, nfr, B, io, num_sms, stream);    // Valkyrie
}

template <int NV, int NC, int NU, int NT>
static int rec_doubles(const DevProblem& hp) { return shape_ok<NV, NC, NU, NT>(hp) ? MCfg<NV, NC, NU, NT>::REC : 0; }

// doubles per state of the setup record (0 when no register kernel matches the problem)
int reg_setup_record_doubles(const DevProblem& hp) {
  int r = 0;
  if (!r) r = rec_doubles<36, 12, 36, 30>(hp);
  if (!r) r = rec_doubles<18, 12, 24, 0>(hp);
  if (!r) r = rec_doubles<33, 12, 36, 0>(hp);
  if (!r) r = rec_doubles<20, 12, 36, 12>(hp);
  if (!r) r = rec_doubles<32, 12, 36, 26>(hp);
  return r;
}

bool reg_kernel_available(const DevProblem& hp) {
  static const bool force_smem = [] { const char* e = getenv("PNC_SOLVER"); return e && e[0] == 's'; }();
  if (force_smem) return false;
  return reg_setup_record_doubles(hp) > 0;
}

}  // namespace pnc
