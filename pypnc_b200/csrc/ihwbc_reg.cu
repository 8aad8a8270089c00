// ihwbc_reg.cu -- the IHWBC solve with the active-set operator held in REGISTERS.
//
// Same algorithm and reference mapping as ihwbc.cu (IHWBC.solve pnc/wbc/ihwbc/ihwbc.py:90-379,
// Goldfarb-Idnani as in external_source/Goldfarb/QuadProg++.cc:115-502, Householder instead of
// the Givens sweep when a constraint is added); what changes is where the data lives and how
// much code the loop is:
//
//   * k_ihwbc_setup (one state per warp, shared memory, high occupancy) builds the Hessian, its
//     Cholesky factor, J = L^-T and the unconstrained minimiser, and then adds ALL the equality
//     constraints at once while J is still triangular (QuadProg++.cc:232-272 adds them one by one):
//     a Householder QR of D = J^T CE (n x p, in registers, lanes = rows) gives the reflectors, the
//     rows of J Q are formed streaming (lane = row, reflectors broadcast from shared memory) and
//     written to the per-state record, and x moves to the equality-constrained minimiser
//     x0 - J1 R^-T (CE^T x0 + ce0).  Only the null-space block Z = (J Q)[:, p:] (n x (n - p)) is
//     kept: the first p columns of J and the first p rows of R only ever feed the multipliers of
//     the equalities, which nothing reads;
//   * k_ihwbc_reg runs the inequality iterations with Z in registers: lane l owns row l of Z
//     (n = nv + nc <= 48 rows, n - p columns), rows 32.. are split over lane pairs (half a row
//     each) -- 66 doubles per lane for Atlas.  z = Z2 d2 and the rank-1 Householder update are
//     pure register FMAs against vectors broadcast from shared memory (LDS.128, one wavefront per
//     two FMAs);
//   * d = Z^T n+ goes through a 6-row scratch tile in shared memory: one pass for a wrench-cone
//     normal (3 or 6 non-zeros), n/6 passes for a dense normal (torque limits);
//   * the kernel is instruction-fetch sensitive (few warps per SM, each walking its own path),
//     so every routine of the active-set loop exists at exactly one call site and the only
//     unrolled code is what register indexing forces.  Dynamic column indices (iq, the Givens
//     pairs of a dropped constraint) are warp-uniform and go through switch statements (jump
//     tables), never local memory.
#include <stdio.h>
#include <stdlib.h>

#include "ihwbc_common.cuh"
#include "tri_reg.cuh"

namespace pnc {

template <int NV_, int NC_, int PE_>
struct RCfg {
  static constexpr int NV = NV_, NC = NC_, PE = PE_, N = NV_ + NC_;
  static constexpr int NCOL = N - PE_;                        // columns of the null-space block Z
  static constexpr bool EXT = N > 32;                         // rows 32.. live in lane pairs
  static constexpr int E = EXT ? N - 32 : 0;
  static constexpr int HW = EXT ? ((NCOL + 3) / 4) * 2 : 0;   // half-row width in columns (even)
  static constexpr int NPC = EXT ? 2 * HW : ((NCOL + 1) & ~1);  // padded column count (even)
  static constexpr int NP = (N + 1) & ~1;                     // padded length of a row-space vector (x, z, n+)
  static constexpr int NB = HW > 0 ? HW : 2;
  static constexpr int ldv_() { int l = (NV_ + 1) & ~1; while ((l & 15) != 6) l += 2; return l; }
  static constexpr int LDV = ldv_();                          // pitch of the nv x nv block: even, 6 (mod 16) ->
                                                              // rows 16-byte aligned, LDS.128 by row conflict-free
  static constexpr int EV = NV_ > 32 ? NV_ - 32 : 0;          // Hessian rows beyond 32
  static constexpr int EVP = EV <= 1 ? 1 : (EV <= 2 ? 2 : (EV <= 4 ? 4 : 8));
  static constexpr int LPR = 32 / EVP;                        // lanes cooperating on one such row in the Cholesky
  static constexpr int LDT = NPC + 2;                         // pitch of the d scratch tile
  static_assert(E <= 16, "rows beyond 32 must fit one lane pair each");
  static_assert(N <= 64 && (PE_ & 1) == 0 && PE_ <= 8, "shape not supported by the register kernel");
};

// Compile-time shared-memory layout of the active-set kernel (doubles).  NU = wrench-cone rows,
// NT = torque-limit rows (0 without torque limits).  Every array sits at an immediate offset from
// the shared-memory base: no address registers, nothing to rematerialise under the register
// pressure of the resident Z.  R (upper triangular, packed by columns, one spare slot per column
// for the sub-diagonal entry a dropped constraint leaves behind) is capped at RC columns (active
// inequalities); a state that would need more is handed to the generic kernel (status OVERFLOW).
template <int NV_, int NC_, int NU_, int NT_, int PE_>
struct MCfg : RCfg<NV_, NC_, PE_> {
  using R = RCfg<NV_, NC_, PE_>;
  static constexpr int NU = NU_, NT = NT_, M = NU_ + 2 * NT_;
  static constexpr int RC = R::NCOL < 26 ? R::NCOL : 26;
  static constexpr int RSZ = (RC * (RC + 3) / 2 + 1) & ~1;
  static constexpr int RV = (RC + 3) & ~1;  // length of r, u, uold
  static constexpr int lda_() { int l = R::NP; while ((l & 15) != 2) l += 2; return l; }
  static constexpr int LDA = lda_();  // torque rows: LDS.128 by row is conflict free at 2 (mod 16)
  static constexpr int ORT = 0, OT = ORT + RSZ, OAR = OT + 6 * R::LDT, OA0 = OAR + NT * LDA, OUF = OA0 + ((3 * NT + 1) & ~1),
                       OUFV = OUF + (NU > 0 ? NU : 1) * 6, OX = OUFV + (((NU > 0 ? NU : 1) + 1) & ~1), OZ = OX + R::NP,
                       ODV = OZ + R::NP, ODM = ODV + R::NPC, ONP = ODM + R::NPC, ORR = ONP + R::NP + 6 /* zero pad: the d tile reads np six at a time */,
                       OU = ORR + RV, ORI = OU + RV /* 1 / diagonal of R */, OXOLD = ORI + RV, OUOLD = OXOLD + R::NP, OXACC = OUOLD + RV,
                       OSL = OXACC + R::NP, OINT = OSL + (((M > 0 ? M : 1) + 1) & ~1),
                       SMEM_DOUBLES = (OINT + (((3 * (RC + 2) + 1) & ~1) + 2 * NU) / 2 + 2) & ~1;  // A, Aold, Aacc, cone-row table
  static_assert(6 * R::LDT >= 3 * RC, "rotation table must fit the scratch tile");
  // ---- per-state HBM record (doubles): written by k_ihwbc_setup, consumed by k_ihwbc_reg, whose
  //      result goes back into it for k_ihwbc_finish.  Everything the active-set kernel needs is
  //      here in the layout it is used in, so that kernel holds no staging code at all ----
  static constexpr int IMG = OX - OAR;                      // shared-memory image: Ar, ci0 of the torque rows, Uf, uf
  static constexpr int RJ = 0;                              // rows of Z as register chunks: [NPC/2][32 lanes][2]
  static constexpr int RJB = RJ + 32 * R::NPC;              // half rows 32..: [NB/2][32 lanes][2]
  static constexpr int RX = RJB + (R::EXT ? 32 * R::NB : 0);  // x at the equality-constrained minimiser
  static constexpr int RS = RX + R::NP;                     // c1, c2, status, R_norm
  static constexpr int RI = RS + 4;                         // the image
  static constexpr int RO = RI + ((IMG + 1) & ~1);          // result: x[NP], then ints iq, iter, status, A[RC+2], iq_ref, Aref[RC+2]
  static constexpr int ROI = 2 * (RC + 2) + 4;              // ints of the result
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
  if (o - L.om6 < ((C::PE * C::NV + 1) & ~1)) o = L.om6 + ((C::PE * C::NV + 1) & ~1);
  // ... and once D sits in registers, the Householder vectors of its QR ([row][reflector]) and their
  // Gram matrix take the same place
  if (o - L.om6 < ((C::N * C::PE + C::PE * C::PE + 4 * C::PE + 1) & ~1)) o = L.om6 + ((C::N * C::PE + C::PE * C::PE + 4 * C::PE + 1) & ~1);
  L.ozv = o; o += C::NP;  // x0 stays live until the record is written
  L.oD = 0;
  if (o < min_doubles) o = min_doubles;  // the image of the active-set kernel's arrays is assembled here at the end
  L.ojc = o; o += C::NC > 0 ? C::NC : 1;  // row pointers of Jc Ni, behind the image: they are used while it is built
  L.per_state = (o + 1) & ~1;
  return L;
}

// WS warps per CTA, one state each, claimed together and meeting at CTA-wide barriers before the long
// straight-line sections (Hessian, the unrolled Cholesky / inverse of tri_reg.cuh: 60 KB of code run once per
// state), so that an instruction line fetched once serves all the warps of the CTA.  Measured in the bench
// (solve stage, 64 K states, WS = 7 against 1): Atlas 7.16 against 7.19 ms, but Valkyrie 6.85 against 6.30,
// Laikago 2.26 against 2.10, Draco3 9.83 against 9.51 -- the 224-thread CTAs pack worse and the register cap
// that goes with them costs the 8-equality shapes spills.  Kept as a build knob, default one warp.
#ifndef PNC_SETUP_WARPS
#define PNC_SETUP_WARPS 1
#endif
template <int NV, int NC, int NU, int NT, int PE, int WS>
__global__ void __launch_bounds__(32 * WS, 16 / WS)
k_ihwbc_setup(const __grid_constant__ DevProblem P, const DevProblem* __restrict__ gp, SetupLayout SL, WsPtrs ws, FrameMap fm,
              int nfr, int B, SolveIO io, double* __restrict__ rec_all, int* __restrict__ counter) {
  using C = MCfg<NV, NC, NU, NT, PE>;
  constexpr int N = C::N, NP = C::NP, LDV = C::LDV, HW = C::HW, LDA = C::LDA;
  constexpr bool EXT = C::EXT;
  extern __shared__ __align__(16) double smem_all[];
  __shared__ int s_base;
  const DevProblem& G = *gp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const smem = smem_all + (size_t)warp * SL.per_state;
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
  // 2x2 tiles of the Hessian's lower triangle, tile idx -> (first row, first column), packed r0 | c0 << 8
  constexpr int nt = (NV + 1) >> 1, ntile = nt * (nt + 1) / 2, TPL = (ntile + 31) / 32;
  __shared__ unsigned short tile_tab[TPL * 32];
  for (int idx = threadIdx.x; idx < TPL * 32; idx += blockDim.x) {
    int ti = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= idx) ti++;
    while (ti * (ti + 1) / 2 > idx) ti--;
    const int tj = idx - ti * (ti + 1) / 2;
    tile_tab[idx] = idx < ntile ? (unsigned short)((2 * ti) | ((2 * tj) << 8)) : (unsigned short)0;
  }
  __syncthreads();
#pragma unroll 1
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_base = atomicAdd(counter, WS);
    __syncthreads();
    const int base = s_base;
    if (base >= B) break;
    // (the warps of a last, partial group redo its first state: identical values to identical places)
    const size_t s = (size_t)(base + warp < B ? base + warp : base);
    double* rec = rec_all + s * C::REC;
    int status = PNC_QP_OK;
    double R_norm = 1.0;
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
      // the image phase at the end of the state reads these rows in full (the phases before it only their first
      // six entries): ask L2 for them now, a DRAM round trip ahead
#pragma unroll
      for (int o = 0; o < NV * 8; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(ptr) + o));
    }
    __syncwarp();
    auto row_jc = [&](int c, int v) -> double { return jcrow[c][v]; };  // (Jc Ni)[c][v]
    // The Hessian accumulators start as the lambda_q_ddot M term: its global loads are issued here, a memory round
    // trip before their first use, and travel while the task data is staged (added after the task loop they were
    // 4 % of the kernel's stall samples, all of it waiting on these loads).
    // lane i < 6 also fetches the bias term of floating-base row i, (c + g)_i or (Ni^T (c + g))_i, for the equality
    // residuals far below (read one row at a time there, it was six exposed round trips)
    // (the two addends stay apart until then where that pays: adding them here waits for the loads on the spot)
    // measured per shape (solve stage, 64 K states): apart, Valkyrie 6.02 -> 5.94 ms, but Atlas 6.55 -> 6.59 (the
    // second live register moves a spill into a hot loop of that instantiation), so Atlas adds on the spot
    constexpr bool kBiasApart = !(NV == 36 && NT == 30);
    double cg6a, cg6b = 0.0;
    if (kBiasApart) {
      cg6a = lane < 6 ? (has_ic ? ws.ic_cg[s * NV + lane] : ws.cori[s * NV + lane]) : 0.0;
      cg6b = (lane < 6 && !has_ic) ? ws.grav[s * NV + lane] : 0.0;
    } else {
      cg6a = lane < 6 ? (has_ic ? ws.ic_cg[s * NV + lane] : ws.cori[s * NV + lane] + ws.grav[s * NV + lane]) : 0.0;
    }
    double h[TPL][4];
    int ro[TPL], co[TPL];
#pragma unroll
    for (int u = 0; u < TPL; u++) {
      const int tab = tile_tab[lane + 32 * u];  // r0 | c0 << 8 (0 | 0 past the last tile: computed, never stored)
      const int r0 = tab & 0xff, c0 = tab >> 8;
      ro[u] = r0;
      co[u] = c0;
      const bool r1ok = r0 + 1 < NV, c1ok = c0 + 1 < NV;
      h[u][0] = Mg[r0 * NV + c0];
      h[u][1] = c1ok ? Mg[r0 * NV + c0 + 1] : 0.0;
      h[u][2] = r1ok ? Mg[(r0 + 1) * NV + c0] : 0.0;
      h[u][3] = (r1ok && c1ok) ? Mg[(r0 + 1) * NV + c0 + 1] : 0.0;
    }
    // ---- tasks: e = Jdot qdot - xddot_cmd, weights (ihwbc.py:159-172) ----
    // task error terms e = Jdot qdot - xddot_cmd were evaluated by k_task_prepare (ihwbc.cu).  They and the dense
    // task Jacobian rows go from global to shared memory as asynchronous copies (cp.async, 8 bytes each, no
    // registers in between): every load of the lane is in flight at once, one memory round trip for the phase
    // (batches of nine register loads paid three), and the weight loads below overlap it.
#ifndef PNC_STAGE_SYNC
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
      const unsigned jd0 = (unsigned)__cvta_generic_to_shared(Jd), ev0 = (unsigned)__cvta_generic_to_shared(ev);
#pragma unroll 4
      for (int e = lane; e < total; e += 32) {
        const int r = e / NV, c = e - r * NV;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(jd0 + 8u * (r * LDV + c)), "l"(rptr[r] + c) : "memory");
      }
      const double* te = ws.task_e + s * P.ntaskrows;
#pragma unroll 1
      for (int k = lane; k < P.ntaskrows; k += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ev0 + 8u * k), "l"(te + k) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      // the task weights in one coalesced load (lane t holds w_t; ntask <= PNC_MAX_TASKS = 16), broadcast per task
      const double w_lane = lane < P.ntask ? io.w[s * P.ntask + lane] : 0.0;
#pragma unroll 1
      for (int t = 0; t < P.ntask; t++) {
        const DevTask& T = P.tasks[t];
        const double wt_t = __shfl_sync(FULL, w_lane, t);
        for (int i = lane; i < T.dim; i += 32) wt[T.row_off + i] = wt_t;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
#else
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
#endif

    __syncthreads();
    // ---- Hessian task block (lower triangle) = sum_t w_t J_t^T J_t + lambda_q_ddot M ----
    // 2x2 tiles, accumulated in registers: the factor block shares its shared memory with the
    // staged Jacobians and is only written once every lane is done reading them.  Loop order: task, row, tile --
    // a lane's TPL tiles are TPL independent accumulator sets fed by two LDS.128 per row (row[r0..r0+1],
    // row[c0..c0+1]); the tile coordinates come from a table built once per CTA.
    {
      {
        const double lam = P.lambda_q_ddot;
#pragma unroll
        for (int u = 0; u < TPL; u++) { h[u][0] *= lam; h[u][1] *= lam; h[u][2] *= lam; h[u][3] *= lam; }
      }
#pragma unroll 1
      for (int t = 0; t < P.ntask; t++) {
        const DevTask& T = P.tasks[t];
        if (T.dense_off < 0) continue;
        const double wk = wt[T.row_off];
        if (wk == 0.0) continue;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const double* row = Jd + (T.dense_off + i) * LDV;
#pragma unroll
          for (int u = 0; u < TPL; u++) {
            double2 rr = lds2(row + ro[u]), cc = lds2(row + co[u]);
            if (NV & 1) {  // the pad column of a staged row is not data
              if (ro[u] + 1 >= NV) rr.y = 0.0;
              if (co[u] + 1 >= NV) cc.y = 0.0;
            }
            const double wa = wk * rr.x, wb = wk * rr.y;
            h[u][0] = fma(wa, cc.x, h[u][0]); h[u][1] = fma(wa, cc.y, h[u][1]);
            h[u][2] = fma(wb, cc.x, h[u][2]); h[u][3] = fma(wb, cc.y, h[u][3]);
          }
        }
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
        if (lane + 32 * u >= ntile) continue;
        const int r0 = ro[u], c0 = co[u];
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

    // ---- Cholesky of the task block (QuadProg++.cc:705-730): right-looking, rows in registers (tri_reg.cuh) ----
    double c1 = 0.0, c2 = 0.0;
    {
      double tr = 0.0;
      for (int i = lane; i < NV; i += 32) tr += Gm[i * LDV + i];
      c1 = wsum(tr) + NC * P.lambda_rf;
    }
    double* dl = vm;  // reciprocal diagonal of L
    __syncthreads();
    const bool pd = chol_rows_reg<NV, LDV>(Gm, dl, dm, lane);
    if (!pd || !(P.lambda_rf > 0.0 || NC == 0)) {
      status = PNC_QP_NOT_PD;
    } else if (status == PNC_QP_OK) {
      // ---- J = L^-T (QuadProg++.cc:203-210): lane c forward-solves L x = e_c with x in registers (fully
      //      unrolled: no shuffles, no barriers, every lane the same instruction stream, tri_reg.cuh) ----
      linv_cols_reg<NV, LDV>(Gm, dl, lane);
      const double jrf = 1.0 / sqrt(P.lambda_rf > 0.0 ? P.lambda_rf : 1.0);
      {
        double tr = 0.0;
        for (int i = lane; i < NV; i += 32) tr += Gm[i * LDV + i];
        c2 = wsum(tr) + NC * jrf;
      }
      // ---- unconstrained minimiser x = -J (J^T g0)  (:222-224); J is upper triangular with its zeros written out,
      //      so both products run over full rows / columns with uniform trip counts (columns 32.. of a robot with
      //      nv > 32 ride along in lanes 0 .. nv-33) ----
      {
        constexpr int NX = NV > 32 ? NV - 32 : 0;
        const double* cj = Gm + (lane < NV ? lane : 0);
        const double* cx = Gm + (NX > 0 && lane < NX ? 32 + lane : 0);
        double a0 = 0.0, a1 = 0.0, b0 = 0.0;
#pragma unroll 2
        for (int k = 0; k + 1 < NV; k += 2) {
          const double2 d2 = lds2(dv + k);
          a0 = fma(cj[k * LDV], d2.x, a0);
          a1 = fma(cj[(k + 1) * LDV], d2.y, a1);
          if (NX > 0) { b0 = fma(cx[k * LDV], d2.x, b0); b0 = fma(cx[(k + 1) * LDV], d2.y, b0); }
        }
        if (NV & 1) {
          a0 = fma(cj[(NV - 1) * LDV], dv[NV - 1], a0);
          if (NX > 0) b0 = fma(cx[(NV - 1) * LDV], dv[NV - 1], b0);
        }
        if (lane < NV) dm[lane] = a0 + a1;
        if (NX > 0 && lane < NX) dm[32 + lane] = b0;
        if (lane == 0 && (NV & 1)) dm[NV] = 0.0;  // pad of the pair loads below
      }
      __syncwarp();
      {
        constexpr int NX = NV > 32 ? NV - 32 : 0;
        const double* rk = Gm + (lane < NV ? lane : 0) * LDV;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 2
        for (int j = 0; j < NV; j += 2) {  // the pad column of J (nv odd) is zero
          const double2 g = lds2(rk + j), d2 = lds2(dm + j);
          a0 = fma(g.x, d2.x, a0);
          a1 = fma(g.y, d2.y, a1);
        }
        if (lane < NV) xs[lane] = -(a0 + a1);
        if (NX > 0) {
          const double* rx = Gm + (32 + (lane < NX ? lane : 0)) * LDV;
          double b0 = 0.0;
#pragma unroll
          for (int j = 32; j < NV; j++) b0 = fma(rx[j], dm[j], b0);
          if (lane < NX) xs[32 + lane] = -b0;
        }
      }
      // ---- all p equality constraints at once (QuadProg++.cc:232-272 adds them one by one), while
      //      J = L^-T is still triangular in shared memory.  D = J^T CE (n x p) is built in registers,
      //      lanes = rows (row j of D in lane j & 31, slot j >> 5): D[j][i] = sum_{k<=j} J[k][j] ce_i[k]
      //      for j < nv, -jrf (Jc Ni)[j-nv][i] for the reaction-force rows of the floating-base
      //      equalities, 0 for those of the internal-constraint rows ----
      double dq[2][PE];
      double* Vs = smem + SL.om6;            // Householder vectors of the QR, [N][PE] (over m6 once it is dead)
      double* Gs = Vs + N * PE;              // their Gram matrix, [PE][PE]
      double* Bs = Gs + PE * PE;             // beta_k
      double* Ys = Bs + PE;                  // y = R^-T (CE^T x0 + ce0)
      {
        double* m6 = smem + SL.om6;  // the nv-part of the PE equality normals, [PE][NV]
        const bool has_tr = P.n_trans > 0;
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
        if constexpr (PE > 6) {
          // rows 6..: internal-constraint rows Ji (ihwbc.py:225-249), or IHWBC2's transmission rows
          // (Sd - Sv) M (ihwbc2.py:249-256)
#pragma unroll 1
          for (int e = lane; e < (PE - 6) * NV; e += 32) {
            const int i = e / NV, k = e - i * NV;
            m6[6 * NV + e] = has_tr ? Mg[P.trans_d[i] * NV + k] - Mg[P.trans_p[i] * NV + k] : G.ic_jac[i][k];
          }
        }
        __syncwarp();
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
          const int j = lane + 32 * pass;
#pragma unroll
          for (int i = 0; i < PE; i++) dq[pass][i] = 0.0;
          if (pass == 1 && N <= 32) continue;
          if (pass == 0 || NV > 32) {
            double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0, e4 = 0.0, e5 = 0.0, e6 = 0.0, e7 = 0.0;
            const int kmax = NV < 32 * (pass + 1) ? NV : 32 * (pass + 1);
            const double* gp_ = Gm + (j < NV ? j : 0);
#pragma unroll 2
            for (int k = 0; k < kmax; k++, gp_ += LDV) {
              const double g = gp_[0];  // J[k][j]: zero below the diagonal (k > j), written out by linv_cols_reg
              e0 = fma(g, m6[k], e0); e1 = fma(g, m6[NV + k], e1); e2 = fma(g, m6[2 * NV + k], e2);
              e3 = fma(g, m6[3 * NV + k], e3); e4 = fma(g, m6[4 * NV + k], e4); e5 = fma(g, m6[5 * NV + k], e5);
              if constexpr (PE > 6) { e6 = fma(g, m6[6 * NV + k], e6); e7 = fma(g, m6[7 * NV + k], e7); }
            }
            if (j >= NV) e0 = e1 = e2 = e3 = e4 = e5 = e6 = e7 = 0.0;  // (those lanes walked column 0)
            dq[pass][0] = e0; dq[pass][1] = e1; dq[pass][2] = e2; dq[pass][3] = e3; dq[pass][4] = e4; dq[pass][5] = e5;
            if constexpr (PE > 6) { dq[pass][6] = e6; dq[pass][7] = e7; }
          }
          if (j >= NV && j < N) {
#pragma unroll
            for (int i = 0; i < 6; i++) dq[pass][i] = -jrf * row_jc(j - NV, i);
            if constexpr (PE > 6) {
              if (has_tr) {  // -(Sd - Sv) (Jc)^T
#pragma unroll
                for (int i = 6; i < PE; i++) dq[pass][i] = -jrf * (row_jc(j - NV, P.trans_d[i - 6]) - row_jc(j - NV, P.trans_p[i - 6]));
              }
            }
          }
        }
        // equality residuals at x0 (its reaction-force part is zero): floating-base rows M[i,:] x0 + (c+g)_i
        // (Ni^T (c+g) with internal constraints), internal-constraint rows Ji x0.  Kept in lane i.
        double cmine = 0.0;
#pragma unroll 1
        for (int i = 0; i < PE; i++) {
          double part = 0.0;
#pragma unroll 1
          for (int k = lane; k < NV; k += 32) part = fma(m6[i * NV + k], xs[k], part);
          part = wsum(part);
          if (i < 6) part += __shfl_sync(FULL, cg6a + cg6b, i);
          else if (has_tr) {
            const int vd = P.trans_d[i - 6], vp = P.trans_p[i - 6];
            part += (ws.cori[s * NV + vd] + ws.grav[s * NV + vd]) - (ws.cori[s * NV + vp] + ws.grav[s * NV + vp]);
          }
          if (lane == i) cmine = part;
        }
        __syncwarp();  // m6 is dead: the reflectors go there
        if (lane < PE) Ys[lane] = cmine;  // parked until the triangular solve below
      }
      // ---- Householder QR of D in registers: reflector k zeroes column k below row k.  v_k goes to
      //      shared memory as Vs[row][k] (one broadcast load serves a row of all reflectors), R stays
      //      in the lanes that own its rows (R[k][c] = row k of column c after reflector k) ----
      double rmine = 1.0;  // R[lane][lane]
#pragma unroll
      for (int k = 0; k < PE; k++) {
        double part = lane >= k ? dq[0][k] * dq[0][k] : 0.0;
        if (N > 32) part = fma(dq[1][k], dq[1][k], part);  // rows beyond n hold zeros
        const double sigma = sqrt(wsum(part));
        if (sigma <= QP_EPS * R_norm && status == PNC_QP_OK) status = PNC_QP_EQ_DEPENDENT;  // :268-271
        R_norm = fmax(R_norm, sigma);
        const double dkk = __shfl_sync(FULL, dq[0][k], k);
        const double alpha = dkk > 0.0 ? -sigma : sigma;
        const double bk = sigma > 0.0 ? fast_rcp(sigma * (sigma + fabs(dkk))) : 0.0;
        if (lane == k) { rmine = alpha; Bs[k] = bk; }
        const double v0 = lane > k ? dq[0][k] : (lane == k ? dkk - alpha : 0.0);
        const double v1 = N > 32 ? dq[1][k] : 0.0;
        if (lane < N) Vs[lane * PE + k] = v0;
        if (N > 32 && lane + 32 < N) Vs[(lane + 32) * PE + k] = v1;
#pragma unroll
        for (int c = k + 1; c < PE; c++) {
          const double sc = bk * wsum(fma(v0, dq[0][c], v1 * dq[1][c]));
          dq[0][c] = fma(-sc, v0, dq[0][c]);
          dq[1][c] = fma(-sc, v1, dq[1][c]);
        }
      }
      __syncwarp();
      // Gram matrix of the reflectors: one (l, k) pair per lane
      if (lane < PE * (PE - 1) / 2) {
        int k = 1, idx = lane;
        while (idx >= k) { idx -= k; k++; }  // pairs in the order (0,1) (0,2) (1,2) (0,3) ...
        const int l = idx;
        double acc = 0.0;
#pragma unroll 4
        for (int j = 0; j < N; j++) acc = fma(Vs[j * PE + l], Vs[j * PE + k], acc);
        Gs[l * PE + k] = acc;
      }
      // y = R^-T (CE^T x0 + ce0): x at the equality-constrained minimiser is x0 - J1 y (J1 = first p
      // columns of J Q).  Forward substitution down the lanes: lane k holds c_k, R[k][k] and R[k][c] (c > k).
      {
        double ymine = 0.0;
        double acc = lane < PE ? Ys[lane] : 0.0;
        const double rimine = rmine != 0.0 ? fast_rcp(rmine) : 0.0;  // one reciprocal per lane, off the chain below
#pragma unroll
        for (int k = 0; k < PE; k++) {
          const double yk = __shfl_sync(FULL, acc * rimine, k);  // y_k, final once rows < k are folded in
          if (lane == k) ymine = yk;
          // fold y_k into the rows below: c_c -= R[k][c] y_k, R[k][c] = dq[0][c] of lane k
#pragma unroll
          for (int c = k + 1; c < PE; c++) {
            const double rkc = __shfl_sync(FULL, dq[0][c], k);
            if (lane == c) acc = fma(-rkc, yk, acc);
          }
        }
        __syncwarp();
        if (lane < PE) Ys[lane] = ymine;
      }
      __syncwarp();
      // ---- rows of J Q, streaming: lane = row i (slot 0: i = lane, slot 1: i = lane + 32).  With
      //      g_k = r . v_k for the ORIGINAL row r, the dot products of the successively reflected row are
      //      y_k = g_k - sum_{l<k} beta_l y_l (v_l . v_k), and r Q = r - sum_k beta_k y_k v_k^T.  Columns
      //      0..p-1 only feed x; columns p.. are Z, written as the register chunks of the active-set kernel
      //      (lane = row, two columns per 16-byte piece; rows 32.. as half rows of lane pairs) ----
#pragma unroll 1
      for (int slot = 0; slot < (N > 32 ? 2 : 1); slot++) {
        const int i = lane + 32 * slot;
        const bool rowok = i < N, rowj = i < NV;
        const double* ri = Gm + (rowj ? i : 0) * LDV;
        // J[i][j], J[i][j + 1] for even j: rows of the nv block carry their zeros explicitly (left of the diagonal,
        // pad column), jrf on the reaction-force diagonal
#define PNC_ROWVAL(j, rx, ry)                                               \
        do {                                                                \
          rx = 0.0; ry = 0.0;                                               \
          if (rowj) {                                                       \
            if ((j) < NV) {                                                 \
              const double2 g_ = lds2(ri + (j));                            \
              rx = g_.x; ry = g_.y;                                         \
            }                                                               \
          } else if (rowok) {                                               \
            if ((j) == i) rx = jrf;                                         \
            if ((j) + 1 == i) ry = jrf;                                     \
          }                                                                 \
        } while (0)
        double gk[PE];
#pragma unroll
        for (int k = 0; k < PE; k++) gk[k] = 0.0;
        // row i of J is zero left of its diagonal and, for i < nv, right of column nv: rows 0..31 of a robot
        // with nv >= 32 only need columns [0, nv), rows 32.. only columns [32, n)
        const int jlo = slot == 0 ? 0 : 32, jhi = (slot == 0 && NV >= 32) ? ((NV + 1) & ~1) : NP;
#pragma unroll 1
        for (int j = jlo; j < jhi; j += 2) {
          double rx, ry;
          PNC_ROWVAL(j, rx, ry);
#pragma unroll
          for (int k = 0; k < PE; k += 2) {
            const double2 va = lds2(Vs + j * PE + k);
            gk[k] = fma(rx, va.x, gk[k]);
            gk[k + 1] = fma(rx, va.y, gk[k + 1]);
            if (j + 1 < N) {
              const double2 vb = lds2(Vs + (j + 1) * PE + k);
              gk[k] = fma(ry, vb.x, gk[k]);
              gk[k + 1] = fma(ry, vb.y, gk[k + 1]);
            }
          }
        }
        // gk[k] becomes beta_k y_k
#pragma unroll
        for (int k = 0; k < PE; k++) {
          double yk = gk[k];
#pragma unroll
          for (int l = 0; l < k; l++) yk = fma(-gk[l], Gs[l * PE + k], yk);
          gk[k] = Bs[k] * yk;
        }
        // x_i = x0_i - sum_{k<p} (J Q)[i][k] y_k
        {
          double acc = 0.0;
#pragma unroll 1
          for (int k = 0; k < PE; k += 2) {
            double o0, o1;
            PNC_ROWVAL(k, o0, o1);
#pragma unroll
            for (int l = 0; l < PE; l++) {
              o0 = fma(-gk[l], Vs[k * PE + l], o0);
              o1 = fma(-gk[l], Vs[(k + 1) * PE + l], o1);
            }
            acc = fma(o0, Ys[k], acc);
            acc = fma(o1, Ys[k + 1], acc);
          }
          if (rowok) rec[C::RX + i] = (i < NV ? xs[i] : 0.0) - acc;
        }
        double2* rj = reinterpret_cast<double2*>(rec + C::RJ);
        double2* rb = reinterpret_cast<double2*>(rec + C::RJB);
#pragma unroll 1
        for (int j = PE; j < NP; j += 2) {
          double o0, o1;
          PNC_ROWVAL(j, o0, o1);
#pragma unroll
          for (int k = 0; k < PE; k += 2) {
            const double2 va = lds2(Vs + j * PE + k);
            o0 = fma(-gk[k], va.x, o0);
            o0 = fma(-gk[k + 1], va.y, o0);
            if (j + 1 < N) {
              const double2 vb = lds2(Vs + (j + 1) * PE + k);
              o1 = fma(-gk[k], vb.x, o1);
              o1 = fma(-gk[k + 1], vb.y, o1);
            }
          }
          if (j + 1 >= N) o1 = 0.0;
          const int cz = j - PE;  // column of Z (even)
          if (slot == 0) {
            if (rowok) rj[(cz >> 1) * 32 + lane] = make_double2(o0, o1);
          } else if (EXT && rowok) {
            const int h = cz >= HW ? 1 : 0, c = cz - h * HW;  // HW is even: a column pair never straddles the halves
            rb[(c >> 1) * 32 + 2 * lane + h] = make_double2(o0, o1);
          }
        }
#undef PNC_ROWVAL
      }
      if (NP > N && lane == 0) rec[C::RX + N] = 0.0;
      if (lane == 0) { rec[C::RS] = c1; rec[C::RS + 1] = c2; rec[C::RS + 3] = R_norm; }
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
#ifndef PNC_IMG_SYNC
        // Torque-limit rows, element (row a, column k) with a == lane.  M is symmetric, so row act_idx[a] is read as
        // column act_idx[a] with consecutive lanes on consecutive addresses.  The rows go from global to shared memory
        // by asynchronous copies (cp.async, 8 bytes each): no registers in between, so all N loads of a lane are in
        // flight at once and the phase pays ONE memory round trip (batches of eight register loads paid six: 10.7 %
        // of the kernel's stall samples).  Issued first: the rows travel while the cones and offsets below are built;
        // -(Jc Ni)^T takes its sign once the data has landed.
        if constexpr (NT > 0) {
          if (lane < nact) {
            const int va = G.act_idx[lane];
            const unsigned dst0 = (unsigned)__cvta_generic_to_shared(Ar + lane * LDA);
#pragma unroll 4
            for (int k = 0; k < N; k++) {
              const double* src = has_ic ? ws.ic_tq + (s * N + k) * nact + lane : (k < NV ? Mg + k * NV + va : jcrow[k - NV] + va);
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst0 + 8u * k), "l"(src) : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
#endif
#pragma unroll 1
        for (int r = lane; r < NU; r += 32) {
          const int ci = G.cone_contact[r];
          const DevContact& Cc = G.contacts[ci];
          const int lr = r - Cc.cone_off;
          double u6[6];
          cone_row_world(Cc, lr, ws.fr_iso + (s * nfr + fm.contact_f[ci]) * 12, u6);
#pragma unroll
          for (int k = 0; k < 6; k++) Uf[r * 6 + k] = u6[k];
          double rz = io.rfz[s * P.ncontact + ci];
          if (rz <= 0.0) rz = 1e-3;  // contact.py:37-41
          uf[r] = lr == Cc.rows - 1 ? -rz : 0.0;
        }
        if constexpr (NT > 0) {
#ifdef PNC_IMG_SYNC  // the register formulation: batches of eight loads, six memory round trips
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
#endif
          for (int a = lane; a < nact; a += 32) {
            const int v = G.act_idx[a];
            // ci0 of the two torque-limit rows of joint a (ihwbc.py:266-296): S Ni^T (c+g) - trq_lo, trq_hi - S Ni^T (c+g)
            const double a0v = has_ic ? ws.ic_tq0[s * nact + a] : ws.cori[s * NV + v] + ws.grav[s * NV + v];
            a0[a] = a0v - G.trq_lo[a];
            a0[NT + a] = G.trq_hi[a] - a0v;
            a0[2 * NT + a] = a0v;  // for the torque recovery at the end of the active-set kernel
          }
#ifndef PNC_IMG_SYNC
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          if (!has_ic && lane < nact) {
#pragma unroll
            for (int k = NV; k < N; k++) Ar[lane * LDA + k] = -Ar[lane * LDA + k];
          }
#endif
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
template <int NV, int NC, int NU, int NT, int PE>
#ifndef PNC_REG_MINB
#define PNC_REG_MINB 8  // resident one-warp CTAs per SM the register allocation aims for
#endif
__global__ void __launch_bounds__(32, PNC_REG_MINB)
k_ihwbc_reg(const __grid_constant__ DevProblem P, const DevProblem* __restrict__ gp, int B, int* __restrict__ counter,
            int* __restrict__ overflow, double* __restrict__ rec_all, int rc_cap, double refine) {
  // P (constant bank) serves warp-uniform reads; tables indexed per lane go through the global copy G
  const DevProblem& G = *gp;
  using C = MCfg<NV, NC, NU, NT, PE>;
  constexpr int N = C::N, NP = C::NP, NPC = C::NPC, NCOL = C::NCOL, HW = C::HW, E = C::E, LDT = C::LDT, NB = C::NB, LDA = C::LDA,
                RC = C::RC;
  constexpr bool EXT = C::EXT;
  // One warp per state, one-warp CTAs.  (Eight warps per CTA meeting at a barrier at the top of every
  // step-direction pass, so that they walk the hot loop together, was measured in round 1: the wait for
  // the slowest warp costs more than the shared instruction fetch saves.)
  extern __shared__ __align__(16) double smem_all[];
  const int lane = threadIdx.x;
  double* const smem = smem_all;
  const int m = C::M, nu = NU;
  constexpr int nact = NT;  // torque-limit rows = actuated joints (shape_ok); only used when NT > 0
  constexpr bool trq = NT > 0;
  constexpr bool S2 = RC > 32;  // S = R^-1 never has rows beyond 32 (RC caps the active set): second-half code is dead
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
  double* const rinv = smem + C::ORI;
  double* const xold = smem + C::OXOLD;
  double* const uold = smem + C::OUOLD;
  double* const xacc = smem + C::OXACC;
  double* const sl = smem + C::OSL;
  int* const Aact = reinterpret_cast<int*>(smem + C::OINT);
  int* Aold = Aact + (RC + 2);
  int* Aacc = Aold + (RC + 2);
  int* const Cm = Aact + ((3 * (RC + 2) + 1) & ~1);  // per cone row: first variable, number of variables (8-byte aligned pairs)
#define Rcol(c) (Rp + ((c) * ((c) + 3)) / 2)
  double* rot = Ts;  // the d scratch tile doubles as the Givens coefficient table of a drop

  // rows beyond 32: lane pair (2e, 2e+1) owns row 32+e, half eh of the columns
  const int eh = lane & 1, epair = lane >> 1;
  const int erow = 32 + epair;
  const bool evalid = EXT && epair < E;

  double Ja[NPC];
  double Jb[NB];
#pragma unroll 1
  for (int r = lane; r < NU; r += 32) {
    const DevContact& Cc = G.contacts[G.cone_contact[r]];
    Cm[2 * r] = NV + Cc.rf_off;
    Cm[2 * r + 1] = Cc.dim;
  }
  if (lane < 6) np[NP + lane] = 0.0;  // never written again
#ifndef PNC_REG_NO_BULK
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned mbar_addr = (unsigned)__cvta_generic_to_shared(&mbar);
  const unsigned img_addr = (unsigned)__cvta_generic_to_shared(smem + C::OAR);
  unsigned mbar_phase = 0u;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#endif
  __syncwarp();

#pragma unroll 1
  for (;;) {
    int si = 0;
    if (lane == 0) si = atomicAdd(counter, 1);
    si = __shfl_sync(FULL, si, 0);
    if (si >= B) break;
    const size_t s = (size_t)si;
    int status = PNC_QP_OK, iter = 0, iq = 0;
    int iq_ref = -1;  // size of the active set at the reference's own termination point (-1: same as the final one)
    double R_norm = 1.0;
    double* rec = rec_all + s * C::REC;  // written by k_ihwbc_setup
#ifndef PNC_REG_NO_BULK
    // The shared-memory image (torque-limit rows, their offsets, the wrench cones: 14 KB for Atlas, already in
    // its final layout) comes in as ONE bulk asynchronous copy (cp.async.bulk, the TMA engine) signalled on an
    // mbarrier: no load / store pairs through registers, and it is in flight while Z is loaded below.
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the previous state's reads of the image are done
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr), "r"((unsigned)(C::IMG * 8)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(img_addr),
                   "l"(rec + C::RI), "r"((unsigned)(C::IMG * 8)), "r"(mbar_addr)
                   : "memory");
    }
#endif
    // Every global load of the state is issued before anything waits on one: Z (row lane, and half rows 32..
    // in lane pairs) straight into its registers -- also for a state the setup kernel gave up on, whose
    // record holds stale numbers nobody reads --, the scalars, x, then the shared-memory image.
    {
      const double2* rj = reinterpret_cast<const double2*>(rec + C::RJ) + lane;
#pragma unroll
      for (int k = 0; k < NPC; k += 2) {
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
    const double c1 = rec[C::RS], c2 = rec[C::RS + 1];
    R_norm = rec[C::RS + 3];  // max(1, |diagonal of R|) after the equalities
    status = (int)rec[C::RS + 2];
    {
#pragma unroll 1
      for (int k = lane; k < NP; k += 32) {
        x[k] = rec[C::RX + k];  // the equality-constrained minimiser (k_ihwbc_setup)
        zv[k] = 0.0; np[k] = 0.0;
      }
#ifdef PNC_REG_NO_BULK
      // torque-limit rows, their offsets and the wrench cones arrive in their final layout
      const double2* src = reinterpret_cast<const double2*>(rec + C::RI);
      double2* dst = reinterpret_cast<double2*>(smem + C::OAR);
#pragma unroll 8
      for (int k = lane; k < C::IMG / 2; k += 32) dst[k] = __ldg(src + k);
#endif
#pragma unroll 1
      for (int k = lane; k < NPC; k += 32) { dm[k] = 0.0; dv[k] = 0.0; }
#pragma unroll 1
      for (int k = lane; k < 6 * LDT; k += 32) Ts[k] = 0.0;
    }
#ifndef PNC_REG_NO_BULK
    {
      unsigned landed;
      do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(landed)
                     : "r"(mbar_addr), "r"(mbar_phase)
                     : "memory");
      } while (!landed);
      mbar_phase ^= 1u;
    }
#endif
    if (status == PNC_QP_OK) {
#pragma unroll 1
      for (int k = lane; k < RC + 2; k += 32) { uv[k] = 0.0; Aact[k] = 0; }
      __syncwarp();

      // ================= dual active-set iterations over the inequalities (:274-500) =================
      // (the equalities of :232-272 were all added by k_ihwbc_setup; iq counts active inequalities)
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
      int iq_acc = 0;

#pragma unroll 1
      while (!done) {
        int ip = 0, k0 = 0, k1 = N;
        double s_ip = 0.0;
        {
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
          if (ip >= nu) {  // torque-limit row: + / - row a of S [M | -(Jc Ni)^T]
            const int tr = ip - nu;
            const int a = tr >= nact ? tr - nact : tr;
            const double sg = tr >= nact ? -1.0 : 1.0;
            const double* row = Ar + a * LDA;
#pragma unroll 1
            for (int k = lane; k < NP; k += 32) np[k] = k < N ? sg * row[k] : 0.0;
          } else {
            k0 = Cm[2 * ip]; k1 = k0 + Cm[2 * ip + 1];
#pragma unroll 1
            for (int k = lane; k < NP; k += 32) np[k] = (k >= k0 && k < k1) ? Uf[ip * 6 + (k - k0)] : 0.0;
          }
          if (lane == 0) { uv[iq] = 0.0; Aact[iq] = ip; }
          __syncwarp();
        }

        // ---- l2a: step directions, step lengths, step; repeat while constraints are dropped ----
#pragma unroll 1
        for (;;) {
          // d = Z^T np through the scratch tile, six rows of Z at a time (:504-516)
          double dA = 0.0, dB = 0.0;
#pragma unroll 1
          for (int c0 = k0; c0 < k1; c0 += 6) {
            const int dim = min(6, k1 - c0);
            if (c0 < 32) {
              const int r = lane - c0;
              if (r >= 0 && r < dim) {
                double* dst = Ts + r * LDT;
#pragma unroll
                for (int j = 0; j < NPC; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(Ja[j], Ja[j + 1]);
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
                if (NPC > 32 && lane + 32 < NPC) dB = fma(nr, tl[r * LDT + 32], dB);
              }
            }
            __syncwarp();
          }
          if (lane < NPC) { dv[lane] = dA; dm[lane] = lane >= iq ? dA : 0.0; }
          if (NPC > 32 && lane + 32 < NPC) { dv[lane + 32] = dB; dm[lane + 32] = lane + 32 >= iq ? dB : 0.0; }
          __syncwarp();

          // z = Z[:, iq:] d[iq:] (:518-528), z.z and z.np
          double za, zb = 0.0, zz, znp, sig2;
          {
            double a0_ = 0.0, a1_ = 0.0, a2_ = 0.0, a3_ = 0.0;
#pragma unroll
            for (int j = 0; j < NPC; j += 4) {
              const double2 v = lds2(dm + j);
              a0_ = fma(Ja[j], v.x, a0_);
              a1_ = fma(Ja[j + 1], v.y, a1_);
              if (j + 2 < NPC) {
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
            double psg = (lane >= iq && lane < NCOL) ? dA * dA : 0.0;   // ||d[iq:]||^2 for the Householder step
            if (lane + 32 >= iq && lane + 32 < NCOL) psg = fma(dB, dB, psg);
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
          // r = R^-1 d[0:iq] (:530-542): column-oriented back substitution, lane i owns row i, two rows
          // per step.  Every lane carries val = (d_i - sum_{j>i} R[i][j] r_j) / R[i][i] for its row, with
          // the row scaled by the reciprocal diagonal off the dependent chain, so that chain is one
          // shuffle and three fmas per pair of rows.  (An explicit S = R^-1 would make this a plain
          // matrix-vector product, but it loses cond(R) eps -- and with it the Givens pairs of a drop --
          // at degenerate vertices, where the active normals are nearly dependent: measured 4e-9 from
          // the exact solution on 6 of 262144 Valkyrie states.)
          {
            static_assert(!S2, "rows of R beyond 32 are not implemented");
            const double rin = lane < iq ? rinv[lane] : 0.0;
            double val = (lane < iq ? dv[lane] : 0.0) * rin;
            int i = iq - 1;
#pragma unroll 1
            for (; i >= 1; i -= 2) {
              const double e1 = lane < i ? Rcol(i)[lane] * rin : 0.0;          // R[lane][i] / R[lane][lane]
              const double e0 = lane < i - 1 ? Rcol(i - 1)[lane] * rin : 0.0;  // R[lane][i-1] / R[lane][lane]
              const double r1 = __shfl_sync(FULL, val, i);
              const double s0 = __shfl_sync(FULL, val, i - 1);
              const double h = __shfl_sync(FULL, e1, i - 1);
              const double r0 = fma(-h, r1, s0);
              val = fma(-e1, r1, val);
              val = fma(-e0, r0, val);
            }
            if (lane < iq) rv[lane] = val;
          }
          __syncwarp();
          // step lengths (:366-393)
          double t1 = INFINITY;
          int kk = 0x7fffffff;
          if (lane < iq) {  // iq <= RC <= 32: one candidate per lane
            const double rk = rv[lane];
            if (rk > 1e-300) { t1 = uv[lane] * fast_rcp(rk); kk = lane; }  // (a denormal r_k would flush to 1 / 0)
          }
          wargmin(t1, kk);
          const int l = kk != 0x7fffffff ? Aact[kk] : 0;
          double t2 = INFINITY;
          if (fabs(zz) > QP_EPS) {
            t2 = -s_ip * fast_rcp(znp);
            if (t2 < 0.0) t2 = INFINITY;
          }
          const double t = fmin(t1, t2);
          if (t >= INFINITY) { status = PNC_QP_INFEASIBLE; done = true; break; }  // :402-407
          const bool dual_only = t2 >= INFINITY;
          if (!dual_only)
            for (int k = lane; k < N; k += 32) x[k] += t * zv[k];
          if (lane < iq) uv[lane] -= t * rv[lane];
          if (lane == 0) uv[iq] += t;
          __syncwarp();

          if (!dual_only && fabs(t - t2) < QP_EPS) {
            // ---- full step: add the constraint (:441-474), one Householder reflection of Z2 ----
            const double isig = sig2 > 0.0 ? rsqrt(sig2) : 0.0;
            const double sigma = sig2 * isig;  // sqrt(sig2), with its reciprocal for free
            if (sigma <= QP_EPS * R_norm) {  // linearly dependent on the active set
              set_bit(excm, ip);
              if (lane < iq) { Aact[lane] = Aold[lane]; uv[lane] = uold[lane]; }
              for (int i = lane; i < NP; i += 32) x[i] = xold[i];
              __syncwarp();
              actm = 0u;
#pragma unroll 1
              for (int i = 0; i < iq; i++) set_bit(actm, Aact[i]);
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
            const double beta = fast_rcp(sigma * (sigma + fabs(d0)));
            // R gains the column [d1; alpha]
            double* col = Rcol(iq);
            if (lane < iq) col[lane] = dv[lane];
            if (lane == 0) { col[iq] = alpha; rinv[iq] = d0 > 0.0 ? -isig : isig; }
            // the Householder vector v = [0 .. 0, v0, d[iq+1:]] is the masked d with one entry patched
            if (lane == 0) dm[iq] = v0;
            const double ca = reg_pick(Ja, iq);  // old column iq of Z
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
            for (int j = 0; j < NPC; j += 2) {
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
            iq++;
            R_norm = fmax(R_norm, sigma);
            set_bit(actm, ip);
            need_l1 = true;
            __syncwarp();
            break;
          }

          // ---- drop constraint l at position kk (:409-423, :476-499; delete_constraint :610-679) ----
          // Column kk of R goes, the columns behind it move up (each keeps one sub-diagonal entry), and the
          // Givens rotations of the row pairs (j, j+1), j = kk .. iq-2, make R triangular again; the same
          // pairs rotate the columns of Z afterwards (coefficient table `rot`).
          clr_bit(actm, l);
          // (lane l moves entry l of every column and nothing else touches those slots: no barrier inside)
#pragma unroll 1
          for (int i = kk; i < iq - 1; i++) {
            if (lane <= i + 1) Rcol(i)[lane] = Rcol(i + 1)[lane];
          }
          __syncwarp();
#pragma unroll 1
          for (int j = kk; j < iq - 1; j++) {
            double* cj = Rcol(j);
            double cc = cj[j], ss = cj[j + 1];
            // h = distance(cc, ss) (QuadProg++.cc:681-692) and 1 / h from one reciprocal square root: R's entries
            // are O(1e-3 .. 1e3) here, so the overflow guard of the reference's hypot is not needed, and the
            // one reciprocal serves cc / h, ss / h and 1 / diag (an ulp or two from :636-638)
            const double d2 = fma(cc, cc, ss * ss);
            const bool skip = !(d2 >= QP_EPS * QP_EPS);  // |h| < eps: QuadProg++ leaves such a pair untouched
            double xny = 0.0;
            if (!skip) {
              double ih = rsqrt(d2);
              const double hh = d2 * ih;
              cc = cc * ih;
              ss = ss * ih;
              double diag = hh;
              if (cc < 0.0) { diag = -hh; ih = -ih; cc = -cc; ss = -ss; }
              xny = ss * fast_rcp(1.0 + cc);
              __syncwarp();
              if (lane == 0) { cj[j] = diag; cj[j + 1] = 0.0; rinv[j] = ih; }
              const int k = j + 1 + lane;  // iq <= RC <= 32: one column per lane
              if (k < iq - 1) {
                double* ck = Rcol(k);
                const double t1r = ck[j], t2r = ck[j + 1];
                const double nj = t1r * cc + t2r * ss;
                ck[j] = nj;
                ck[j + 1] = xny * (t1r + nj) - t2r;
              }
            } else if (lane == 0) {
              rinv[j] = 1.0 / cc;
            }
            if (lane == 0) { rot[3 * j] = skip ? 2.0 : cc; rot[3 * j + 1] = ss; rot[3 * j + 2] = xny; }
            __syncwarp();
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
      if (lane == 0) { ro[0] = iq; ro[1] = iter; ro[2] = status; ro[3 + RC + 2] = iq_ref; }
      if (lane < iq_ref) ro[4 + RC + 2 + lane] = Aacc[lane];
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
  const int nv = G.nv, nc = G.nc, n = nv + nc, nact = G.nact, m = G.m;
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
          // IHWBC2 transmission (ihwbc2.py:386-403): the internal force Sv (M qddot + c + g - Jc^T rf) at the
          // passive joint comes back through Ji^T onto its distal joint
          for (int kt = 0; kt < G.n_trans; kt++) {
            if (G.trans_d[kt] != v) continue;
            const int vp = G.trans_p[kt];
            const double* cp = ws.M + s * nv * nv + vp;
            double tp = 0.0;
            for (int k2 = 0; k2 < nv; k2++) tp = fma(cp[k2 * nv], x[k2], tp);
            const double* xr2 = x + nv;
            for (int ci = 0; ci < G.ncontact; ci++) {
              const int dim = G.contacts[ci].dim;
              const double* jr = ws.fr_jac + ((s * nfr + fm.contact_f[ci]) * 6 + (6 - dim)) * nv + vp;
              for (int r = 0; r < dim; r++) tp = fma(-jr[r * nv], xr2[r], tp);
              xr2 += dim;
            }
            tau += tp + ws.cori[s * nv + vp] + ws.grav[s * nv + vp];
          }
        }
      }
      io.trq[s * nact + a] = tau;
    }
  }
  if (io.active) {
    const int mw = (m + 63) / 64 > 0 ? (m + 63) / 64 : 1;
    for (int wdx = lane; wdx < mw; wdx += 32) {
      unsigned long long bits = 0ull;
      for (int i = 0; i < iq; i++) {  // the recorded sets hold inequalities only
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
      for (int i = 0; i < nr; i++) {
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

template <int NV, int NC, int NU, int NT, int PE>
static bool shape_ok(const DevProblem& hp) {
  using C = MCfg<NV, NC, NU, NT, PE>;
  if (hp.nv != NV || hp.nc != NC || hp.nu != NU || hp.m != C::M || hp.p != PE) return false;
  if ((hp.b_trq_limit ? hp.nact : 0) != NT || hp.nact > 32 || hp.m > 128) return false;
  if (hp.ndense > C::NP) return false;  // the setup kernel keeps one row pointer per dense task row in an NP-sized vector
  return true;
}

template <int NV, int NC, int NU, int NT, int PE>
static bool try_launch(const DevProblem* dp, const DevProblem& hp, const WsPtrs& ws, const FrameMap& fm, int nfr, int B,
                       const SolveIO& io, int num_sms, cudaStream_t stream) {
  if (!shape_ok<NV, NC, NU, NT, PE>(hp)) return false;
  using C = MCfg<NV, NC, NU, NT, PE>;
  constexpr int warps = 1;
  const size_t smem = (size_t)C::SMEM_DOUBLES * sizeof(double);
  static SmemOptIn optin;
  if (!optin.ensure(k_ihwbc_reg<NV, NC, NU, NT, PE>, smem)) return false;
  const SetupLayout SL = setup_layout<C>(hp, C::IMG);
  if (!ws.setup_rec || ws.setup_stride != C::REC) return false;
  constexpr int WS = PNC_SETUP_WARPS;  // warps per CTA of the setup kernel (see k_ihwbc_setup)
  const size_t smem_s = (size_t)SL.per_state * sizeof(double) * WS;
  static SmemOptIn optin_s;
  if (!optin_s.ensure(k_ihwbc_setup<NV, NC, NU, NT, PE, WS>, smem_s)) return false;
  int per_sm_s = 0, per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, k_ihwbc_setup<NV, NC, NU, NT, PE, WS>, 32 * WS, smem_s);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ihwbc_reg<NV, NC, NU, NT, PE>, 32 * warps, smem);
  if (per_sm < 1 || per_sm_s < 1) return false;
  static const int cap = [] { const char* e = getenv("PNC_REG_CTAS"); return e ? atoi(e) : 0; }();
  if (cap > 0 && per_sm > cap) per_sm = cap;
  static const bool verbose = getenv("PNC_VERBOSE") != nullptr;
  if (verbose)
    fprintf(stderr, "pnc_b200: <%d,%d,%d,%d,%d> setup %d CTAs of %d warps per SM (%zu B smem), active-set %d warps/SM (%zu B smem), record %d B\n",
            NV, NC, NU, NT, PE, per_sm_s, WS, smem_s, per_sm * warps, smem / warps, (int)(C::REC * sizeof(double)));
  int grid_s = num_sms * per_sm_s, grid = num_sms * per_sm;
  if (grid_s * WS > B) grid_s = (B + WS - 1) / WS;
  if (grid * warps > B) grid = (B + warps - 1) / warps;
  cudaMemsetAsync(io.counter + 2, 0, sizeof(int), stream);
  k_ihwbc_setup<NV, NC, NU, NT, PE, WS><<<grid_s, 32 * WS, smem_s, stream>>>(hp, dp, SL, ws, fm, nfr, B, io, ws.setup_rec, io.counter + 2);
  g_launch_count++;
  // PNC_DEBUG_RC lowers the cap on active constraints so the tests can exercise the hand-over
  static const int rc_dbg = [] { const char* e = getenv("PNC_DEBUG_RC"); return e ? atoi(e) : 0; }();
  const int rc_cap = (rc_dbg > 0 && rc_dbg < C::RC) ? rc_dbg : C::RC;
  k_ihwbc_reg<NV, NC, NU, NT, PE><<<grid, 32 * warps, smem, stream>>>(hp, dp, B, io.counter, io.overflow, ws.setup_rec, rc_cap,
                                                                      io.psi_refine);
  g_launch_count++;
  k_ihwbc_finish<<<(B + 7) / 8, 256, 0, stream>>>(dp, ws, fm, nfr, B, io, ws.setup_rec, C::REC, C::RO, C::NP, NT > 0 ? C::RT : -1,
                                                  3 + C::RC + 2);
  g_launch_count++;
  return true;
}

// register-resident kernels exist for the QP shapes of the shipped robots; anything else takes
// the generic shared-memory kernel of ihwbc.cu
// <nv, nc, cone rows, torque-limit rows, equalities>
#define PNC_REG_SHAPES(X) \
  X(36, 12, 36, 30, 6) /* Atlas */ X(18, 12, 24, 0, 6) /* Laikago */ X(33, 12, 36, 0, 8) /* Draco3 */ \
  X(20, 12, 36, 12, 8) /* Draco3 lower body */ X(32, 12, 36, 26, 6) /* Valkyrie */

bool launch_ihwbc_solve_reg(const DevProblem* dp, const DevProblem& hp, const WsPtrs& ws, const FrameMap& fm, int nfr,
                            int B, const SolveIO& io, int num_sms, cudaStream_t stream) {
#define X(a, b, c, d, e) if (try_launch<a, b, c, d, e>(dp, hp, ws, fm, nfr, B, io, num_sms, stream)) return true;
  PNC_REG_SHAPES(X)
#undef X
  return false;
}

// doubles per state of the setup record (0 when no register kernel matches the problem)
int reg_setup_record_doubles(const DevProblem& hp) {
#define X(a, b, c, d, e) if (shape_ok<a, b, c, d, e>(hp)) return MCfg<a, b, c, d, e>::REC;
  PNC_REG_SHAPES(X)
#undef X
  return 0;
}

// 1-based index of the register-kernel shape that serves the problem (0: none).  A record slab is laid
// out for ONE shape: the setup kernel leaves padding columns and unused lanes untouched (zero).
int reg_shape_id(const DevProblem& hp) {
  int id = 0;
#define X(a, b, c, d, e) id++; if (shape_ok<a, b, c, d, e>(hp)) return id;
  PNC_REG_SHAPES(X)
#undef X
  return 0;
}

bool reg_record_layout(const DevProblem& hp, int* o) {
#define X(a, b, c, d, e)                                                                                      \
  if (shape_ok<a, b, c, d, e>(hp)) {                                                                          \
    using C = MCfg<a, b, c, d, e>;                                                                            \
    const int v[10] = {C::N, C::PE, C::NCOL, C::NPC, C::HW, C::RJ, C::RJB, C::RX, C::RS, C::REC};             \
    for (int i = 0; i < 10; i++) o[i] = v[i];                                                                 \
    return true;                                                                                              \
  }
  PNC_REG_SHAPES(X)
#undef X
  return false;
}

bool reg_kernel_available(const DevProblem& hp) {
  static const bool force_smem = [] { const char* e = getenv("PNC_SOLVER"); return e && e[0] == 's'; }();
  if (force_smem) return false;
  return reg_setup_record_doubles(hp) > 0;
}

}  // namespace pnc
