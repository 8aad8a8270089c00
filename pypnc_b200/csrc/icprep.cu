// icprep.cu -- internal-constraint projection products, one robot state per warp.
//
// Replaces the numpy block at the top of IHWBC.solve (reference pnc/wbc/ihwbc/ihwbc.py:124-146
// with util/util.py:91-95 weighted_pinv) for problems with internal constraints
// (Draco3 rolling knee joints, pnc/draco3_pnc/draco3_rolling_joint_constraint.py):
//   Lambda = pinv(Ji M^-1 Ji^T)            Jibar = M^-1 Ji^T Lambda       Ni = I - Jibar Ji
//   A = (Sa Ni)[:, 6:]   W = M^-1[6:, 6:]  B = (W A^T pinv(A W A^T))^T    S = [0 | B]
// and writes what the QP kernel consumes: Ni^T (c+g), Jc Ni, S [M | -(Jc Ni)^T], S Ni^T (c+g).
//
// No explicit inverse is formed.  With M = L L^T:  Ji M^-1 Ji^T = Y^T Y (Y = L^-1 Ji^T),
// A W A^T = Z^T Z with Z = L^-1 [0 | A]^T, and W A^T = rows 6.. of L^-T Z.  Both pseudo-inverses
// are applied as Cholesky solves: for the shipped robots the two Gram matrices are full rank
// (singular-value ratio > 1e-4, far above the reference's rcond = 1e-15), where pinv == inverse.
// A rank-deficient Gram matrix is reported through ic_status (the QP then returns
// PNC_QP_EQ_DEPENDENT for that state) instead of being silently truncated.
#include <math.h>

#include <stdlib.h>

#include "pnc_launch.h"
#include "tri_reg.cuh"

namespace pnc {

#define FULL 0xffffffffu

// sum_c a[c sa] b[c sb], c < n, with four independent accumulators: the factorisations and
// triangular solves below are chains of such dot products, and a single accumulator makes every
// FMA wait for the previous one
// Inlined where the dimensions (hence the strides) are compile-time constants; out of line for the
// runtime-dimension instantiation, whose seven inlined, unrolled copies pushed no_inst stalls from
// 2 % to 24 % (the kernel's warps wander through the code independently).
static __device__ __forceinline__ double dot4_inl(const double* a, int sa, const double* b, int sb, int n) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int c = 0;
#pragma unroll 1
  for (; c + 3 < n; c += 4) {
    s0 = fma(a[c * sa], b[c * sb], s0);
    s1 = fma(a[(c + 1) * sa], b[(c + 1) * sb], s1);
    s2 = fma(a[(c + 2) * sa], b[(c + 2) * sb], s2);
    s3 = fma(a[(c + 3) * sa], b[(c + 3) * sb], s3);
  }
#pragma unroll 1
  for (; c < n; c++) s0 = fma(a[c * sa], b[c * sb], s0);
  return (s0 + s1) + (s2 + s3);
}
static __device__ __noinline__ double dot4_out(const double* a, int sa, const double* b, int sb, int n) {
  return dot4_inl(a, sa, b, sb, n);
}


// Lambda = (Y^T Y)^-1 (k x k, k <= PNC_MAX_IC) by Gauss-Jordan with partial pivoting, one thread.
// Returns 1 when the Gram matrix is singular to 1e-13 of its largest diagonal entry.
static __device__ __noinline__ int ic_lambda_inverse(const double* __restrict__ Y, double* __restrict__ Lam, int nv, int k) {
  double a[PNC_MAX_IC][2 * PNC_MAX_IC];
  for (int r = 0; r < k; r++)
    for (int c = 0; c < k; c++) {
      double sum = 0.0;
      for (int i = 0; i < nv; i++) sum += Y[i * k + r] * Y[i * k + c];
      a[r][c] = sum;
      a[r][k + c] = r == c ? 1.0 : 0.0;
    }
  double amax = 0.0;
  for (int r = 0; r < k; r++) amax = fmax(amax, fabs(a[r][r]));
  for (int c = 0; c < k; c++) {
    int pr = c;
    for (int r = c + 1; r < k; r++)
      if (fabs(a[r][c]) > fabs(a[pr][c])) pr = r;
    if (!(fabs(a[pr][c]) > 1e-13 * amax)) return 1;
    for (int cc = 0; cc < 2 * k; cc++) { double tmp = a[c][cc]; a[c][cc] = a[pr][cc]; a[pr][cc] = tmp; }
    const double d = a[c][c];
    for (int cc = 0; cc < 2 * k; cc++) a[c][cc] /= d;
    for (int r = 0; r < k; r++) {
      if (r == c) continue;
      const double f = a[r][c];
      for (int cc = 0; cc < 2 * k; cc++) a[r][cc] -= f * a[c][cc];
    }
  }
  for (int r = 0; r < k; r++)
    for (int c = 0; c < k; c++) Lam[r * k + c] = a[r][k + c];
  return 0;
}

#ifndef PNC_IC_WARPS
#define PNC_IC_WARPS 10
#endif

struct IcLayout {
  int ld, ldz, ldb;
  int oL, oZ, oV, oG, oBt, oY, oJb, oLam, odiag, ocg, oncg, oJcJb, oLdi, oGdi;
  int total;
};

__host__ __device__ inline IcLayout ic_layout(int nv, int nact, int nc, int k) {
  IcLayout L;
  L.ld = nv | 1; L.ldz = nact | 1; L.ldb = nact | 1;
  int o = 0;
  L.oL = o; o += nv * L.ld;
  L.oZ = o; o += nv * L.ldz;       // Z = L^-1 Atilde^T, later V = L^-T Z
  L.oG = o; o += nact * L.ldz;     // Z^T Z and its Cholesky factor
  L.oBt = o; o += (nv - 6) * L.ldb;  // B^T  [(nv-6) x nact]
  L.oY = o; o += nv * k;
  L.oJb = o; o += nv * k;
  L.oLam = o; o += k * k + 2 * k;
  L.odiag = o; o += nv;
  L.ocg = o; o += nv;
  L.oncg = o; o += nv;
  L.oJcJb = o; o += (nc > 0 ? nc : 1) * k;
  L.oLdi = o; o += nv;
  L.oGdi = o; o += nact;
  L.oV = L.oZ;
  L.total = (o + 1) & ~1;
  return L;
}

// NV_ / NACT_ / NC_ / K_ > 0 fix the dimensions at compile time (the shipped robots); 0 = read them
// from the problem (any other shape).  Same code either way.
template <int NV_, int NACT_, int NC_, int K_>
__global__ void __launch_bounds__(32)
k_ic_prepare(const DevProblem* __restrict__ gp, IcLayout L, WsPtrs ws, FrameMap fm, int nfr, int B, int* counter) {
  extern __shared__ __align__(16) double sm[];
  const DevProblem& P = *gp;
  const int lane = threadIdx.x;
  const int nv = NV_ ? NV_ : P.nv, k = K_ ? K_ : P.n_ic, nact = NACT_ ? NACT_ : P.nact, nc = NV_ ? NC_ : P.nc;
  const int n = nv + nc, ld = nv | 1, ldz = nact | 1, ldb = nact | 1;
  auto dot4 = [](const double* a, int sa, const double* b, int sb, int cnt) {
    if constexpr (NV_ > 0) return dot4_inl(a, sa, b, sb, cnt);
    else return dot4_out(a, sa, b, sb, cnt);
  };
  const int nj = nv - 6;
  double* Lm = sm + L.oL;
  double* Z = sm + L.oZ;
  double* G = sm + L.oG;
  double* Bt = sm + L.oBt;
  double* Y = sm + L.oY;
  double* Jb = sm + L.oJb;
  double* Lam = sm + L.oLam;
  double* Mdiag = sm + L.odiag;
  double* cg = sm + L.ocg;
  double* ncg = sm + L.oncg;
  double* JcJb = sm + L.oJcJb;
  double* Ldi = sm + L.oLdi;  // 1 / L[i][i]
  double* Gdi = sm + L.oGdi;  // 1 / G-factor[i][i]

  for (;;) {
    int si = 0;
    if (lane == 0) si = atomicAdd(counter, 1);
    si = __shfl_sync(FULL, si, 0);
    if (si >= B) break;
    const size_t s = (size_t)si;
    int bad = 0;
    const double* Mg = ws.M + s * nv * nv;
    for (int e0 = lane; e0 < nv * nv; e0 += 32 * 8) {  // flattened: eight independent loads in flight per lane
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = e0 + 32 * u < nv * nv ? Mg[e0 + 32 * u] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int e = e0 + 32 * u;
        if (e < nv * nv) { const int i = e / nv; Lm[i * ld + (e - i * nv)] = v[u]; }
      }
    }
    for (int c = lane; c < nv; c += 32) { Mdiag[c] = Mg[c * nv + c]; cg[c] = ws.cori[s * nv + c] + ws.grav[s * nv + c]; }
    __syncwarp();
    // ---- M = L L^T (lower triangle of Lm; the strict upper triangle keeps M) ----
    for (int j = 0; j < nv; j++) {
      double t[2] = {0.0, 0.0};
      const double* rj = Lm + j * ld;
      for (int rr = 0; rr < 2; rr++) {
        const int i = lane + 32 * rr;
        if (i >= j && i < nv) {
          const double* ri = Lm + i * ld;
          t[rr] = ri[j] - dot4(ri, 1, rj, 1, j);
        }
      }
      const double piv = __shfl_sync(FULL, (j & 32) ? t[1] : t[0], j & 31);
      if (!(piv > 0.0)) { bad = 1; break; }
      const double dj = sqrt(piv);
      __syncwarp();
      const double idj = 1.0 / dj;
      for (int rr = 0; rr < 2; rr++) {
        const int i = lane + 32 * rr;
        if (i == j) { Lm[j * ld + j] = dj; Ldi[j] = idj; }
        else if (i > j && i < nv) Lm[i * ld + j] = t[rr] * idj;
      }
      __syncwarp();
    }
    if (!bad) {
      // ---- Y = L^-1 Ji^T, Lambda = (Y^T Y)^-1.  Column-oriented forward substitution, lanes = rows
      //      (two per lane), all k right-hand sides at once: row c is final after step c-1, its owner
      //      broadcasts y_c and every later row subtracts L[i][c] y_c ----
      {
        double rh[2][PNC_MAX_IC];
#pragma unroll
        for (int rr = 0; rr < 2; rr++)
#pragma unroll
          for (int t = 0; t < PNC_MAX_IC; t++) {
            const int i = lane + 32 * rr;
            rh[rr][t] = (t < k && i < nv) ? P.ic_jac[t][i] : 0.0;
          }
        for (int c = 0; c < nv; c++) {
          const double idc = Ldi[c];
          double yc[PNC_MAX_IC];
#pragma unroll
          for (int t = 0; t < PNC_MAX_IC; t++) yc[t] = __shfl_sync(FULL, ((c & 32) ? rh[1][t] : rh[0][t]) * idc, c & 31);
#pragma unroll
          for (int rr = 0; rr < 2; rr++) {
            const int i = lane + 32 * rr;
            if (i > c && i < nv) {
              const double l = Lm[i * ld + c];
#pragma unroll
              for (int t = 0; t < PNC_MAX_IC; t++) rh[rr][t] = fma(-l, yc[t], rh[rr][t]);
            }
          }
          if (lane < k) {
            double v = yc[0];
#pragma unroll
            for (int t = 1; t < PNC_MAX_IC; t++) v = lane == t ? yc[t] : v;
            Y[c * k + lane] = v;
          }
        }
      }
      __syncwarp();
      if (lane == 0) bad = ic_lambda_inverse(Y, Lam, nv, k);
      bad = __shfl_sync(FULL, bad, 0);
      __syncwarp();
    }
    if (!bad) {
      // ---- Jibar = L^-T (Y Lambda): column-oriented back substitution, lanes = rows, k columns at once ----
      {
        double rh[2][PNC_MAX_IC];
#pragma unroll
        for (int rr = 0; rr < 2; rr++)
#pragma unroll
          for (int t = 0; t < PNC_MAX_IC; t++) {
            const int i = lane + 32 * rr;
            double sum = 0.0;
            if (t < k && i < nv)
              for (int u = 0; u < k; u++) sum += Y[i * k + u] * Lam[u * k + t];
            rh[rr][t] = sum;
          }
        for (int c = nv - 1; c >= 0; c--) {
          const double idc = Ldi[c];
          double xc[PNC_MAX_IC];
#pragma unroll
          for (int t = 0; t < PNC_MAX_IC; t++) xc[t] = __shfl_sync(FULL, ((c & 32) ? rh[1][t] : rh[0][t]) * idc, c & 31);
#pragma unroll
          for (int rr = 0; rr < 2; rr++) {
            const int i = lane + 32 * rr;
            if (i < c) {
              const double l = Lm[c * ld + i];
#pragma unroll
              for (int t = 0; t < PNC_MAX_IC; t++) rh[rr][t] = fma(-l, xc[t], rh[rr][t]);
            }
          }
          if (lane < k) {
            double v = xc[0];
#pragma unroll
            for (int t = 1; t < PNC_MAX_IC; t++) v = lane == t ? xc[t] : v;
            Jb[c * k + lane] = v;
          }
        }
      }
      __syncwarp();
      // ---- Ni^T (c+g) = cg - Ji^T (Jibar^T cg) ----
      {
        double jt[PNC_MAX_IC];
        for (int t = 0; t < k; t++) {
          double part = 0.0;
          for (int i = lane; i < nv; i += 32) part += Jb[i * k + t] * cg[i];
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
          jt[t] = part;
        }
        for (int i = lane; i < nv; i += 32) {
          double v = cg[i];
          for (int t = 0; t < k; t++) v -= P.ic_jac[t][i] * jt[t];
          ncg[i] = v;
          ws.ic_cg[s * nv + i] = v;
        }
      }
      // ---- Jc Ni = Jc - (Jc Jibar) Ji ----
      // the contact Jacobian rows are staged in shared memory first (Z is not live yet): coalesced
      // loads instead of one strided row walk per lane
      const bool stage_jc = nc * ld <= nv * ldz;
      double* Jcs = Z;
      if (stage_jc) {
        for (int c = 0; c < nc; c++) {
          int ci = 0;
          while (ci + 1 < P.ncontact && c >= P.contacts[ci + 1].rf_off) ci++;
          const DevContact& C = P.contacts[ci];
          const double* Jrow = ws.fr_jac + (s * nfr + fm.contact_f[ci]) * 6 * nv +
                               (C.type == PNC_CONTACT_POINT ? 3 + (c - C.rf_off) : (c - C.rf_off)) * nv;
          for (int i = lane; i < nv; i += 32) Jcs[c * ld + i] = Jrow[i];
        }
        __syncwarp();
      }
      for (int c = lane; c < nc; c += 32) {
        int ci = 0;
        while (ci + 1 < P.ncontact && c >= P.contacts[ci + 1].rf_off) ci++;
        const DevContact& C = P.contacts[ci];
        const double* Jrow = stage_jc ? Jcs + c * ld
                                      : ws.fr_jac + (s * nfr + fm.contact_f[ci]) * 6 * nv +
                                            (C.type == PNC_CONTACT_POINT ? 3 + (c - C.rf_off) : (c - C.rf_off)) * nv;
        double acc[PNC_MAX_IC];
        for (int t = 0; t < k; t++) acc[t] = 0.0;
        for (int i = 0; i < nv; i++) {
          const double v = Jrow[i];
          for (int t = 0; t < k; t++) acc[t] += v * Jb[i * k + t];
        }
        for (int t = 0; t < k; t++) JcJb[c * k + t] = acc[t];
        for (int i = 0; i < nv; i++) {
          double v = Jrow[i];
          for (int t = 0; t < k; t++) v -= acc[t] * P.ic_jac[t][i];
          ws.ic_jcni[(s * nc + c) * nv + i] = v;
        }
      }
      __syncwarp();
      // ---- Z = L^-1 Atilde^T, Atilde[a][j] = Ni[act_idx[a]][j] for j >= 6 (lane a solves column a) ----
      for (int a = lane; a < nact; a += 32) {
        const int va = P.act_idx[a];
        for (int i = 0; i < nv; i++) {
          double sum = 0.0;
          if (i >= 6) {
            sum = i == va ? 1.0 : 0.0;
            for (int t = 0; t < k; t++) sum -= Jb[va * k + t] * P.ic_jac[t][i];
          }
          sum -= dot4(Lm + i * ld, 1, Z + a, ldz, i);
          Z[i * ldz + a] = sum * Ldi[i];
        }
      }
      __syncwarp();
      // ---- G = Z^T Z (lower triangle), then its Cholesky factor ----
      for (int idx = lane; idx < nact * (nact + 1) / 2; idx += 32) {
        int r = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while ((r + 1) * (r + 2) / 2 <= idx) r++;
        while (r * (r + 1) / 2 > idx) r--;
        const int c = idx - r * (r + 1) / 2;
        G[r * ldz + c] = dot4(Z + r, ldz, Z + c, ldz, nv);
      }
      __syncwarp();
      double gmax = 0.0;
      for (int a = lane; a < nact; a += 32) gmax = fmax(gmax, G[a * ldz + a]);
      for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(FULL, gmax, o));
      for (int j = 0; j < nact; j++) {
        double t[2] = {0.0, 0.0};
        const double* rj = G + j * ldz;
        for (int rr = 0; rr < 2; rr++) {
          const int i = lane + 32 * rr;
          if (i >= j && i < nact) {
            const double* ri = G + i * ldz;
            t[rr] = ri[j] - dot4(ri, 1, rj, 1, j);
          }
        }
        const double piv = __shfl_sync(FULL, (j & 32) ? t[1] : t[0], j & 31);
        if (!(piv > 1e-13 * gmax)) { bad = 1; break; }
        const double dj = sqrt(piv);
        __syncwarp();
        const double idj = 1.0 / dj;
        for (int rr = 0; rr < 2; rr++) {
          const int i = lane + 32 * rr;
          if (i == j) { G[j * ldz + j] = dj; Gdi[j] = idj; }
          else if (i > j && i < nact) G[i * ldz + j] = t[rr] * idj;
        }
        __syncwarp();
      }
    }
    if (!bad) {
      // ---- V = L^-T Z in place (lane a back-substitutes column a); W A^T = rows 6.. of V ----
      for (int a = lane; a < nact; a += 32) {
        for (int i = nv - 1; i >= 0; i--) {
          const double sum = Z[i * ldz + a] - dot4(Lm + (i + 1) * ld + i, ld, Z + (i + 1) * ldz + a, ldz, nv - i - 1);
          Z[i * ldz + a] = sum * Ldi[i];
        }
      }
      __syncwarp();
      // ---- B^T[j][:] = (A W A^T)^-1 (W A^T)[j][:]^T : lane j solves with the factor of G ----
      for (int j = lane; j < nj; j += 32) {
        double* b = Bt + j * ldb;
        const double* rhs = Z + (6 + j) * ldz;
        for (int i = 0; i < nact; i++) {  // forward
          b[i] = (rhs[i] - dot4(G + i * ldz, 1, b, 1, i)) * Gdi[i];
        }
        for (int i = nact - 1; i >= 0; i--) {  // backward
          b[i] = (b[i] - dot4(G + (i + 1) * ldz + i, ldz, b + i + 1, 1, nact - i - 1)) * Gdi[i];
        }
      }
      __syncwarp();
      // ---- S [M | -(Jc Ni)^T] and S Ni^T (c+g), S = [0 | B]: lane = actuated row a, so the
      //      transposed store [n][nact] is coalesced.  The right factor X = [M[6:, :] | (Jc Ni)[:, 6:]^T]
      //      (nj x n) is staged over the dead factor L and Z, then read as shared-memory broadcasts ----
      const int ldx = n | 1;
      const bool stage_x = nj * ldx <= nv * ld + nv * ldz;
      double* X = Lm;
      if (stage_x) {
        for (int e0 = lane; e0 < nj * nv; e0 += 32 * 8) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = e0 + 32 * u < nj * nv ? Mg[6 * nv + e0 + 32 * u] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int e = e0 + 32 * u;
            if (e < nj * nv) { const int j = e / nv; X[j * ldx + (e - j * nv)] = v[u]; }
          }
        }
        for (int e = lane; e < nc * nj; e += 32) {
          const int cc = e / nj, j = e - cc * nj;
          X[j * ldx + nv + cc] = ws.ic_jcni[(s * nc + cc) * nv + 6 + j];
        }
        __syncwarp();
      }
      for (int a = lane; a < nact; a += 32) {
        double* out = ws.ic_tq + s * nact * n + a;
        if (stage_x) {
          for (int c = 0; c < n; c += 4) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            const double* xc = X + c;
            const bool ok1 = c + 1 < n, ok2 = c + 2 < n, ok3 = c + 3 < n;
            for (int j = 0; j < nj; j++, xc += ldx) {
              const double b = Bt[j * ldb + a];
              s0 = fma(b, xc[0], s0);
              if (ok1) s1 = fma(b, xc[1], s1);
              if (ok2) s2 = fma(b, xc[2], s2);
              if (ok3) s3 = fma(b, xc[3], s3);
            }
            out[c * nact] = c < nv ? s0 : -s0;
            if (ok1) out[(c + 1) * nact] = c + 1 < nv ? s1 : -s1;
            if (ok2) out[(c + 2) * nact] = c + 2 < nv ? s2 : -s2;
            if (ok3) out[(c + 3) * nact] = c + 3 < nv ? s3 : -s3;
          }
        } else {
          for (int c = 0; c < n; c++) {
            const double* src = c < nv ? Mg + 6 * nv + c : ws.ic_jcni + (s * nc + (c - nv)) * nv + 6;
            const int stride = c < nv ? nv : 1;
            double s0 = 0.0, s1 = 0.0;
            int j = 0;
            for (; j + 1 < nj; j += 2) {
              s0 += Bt[j * ldb + a] * src[j * stride];
              s1 += Bt[(j + 1) * ldb + a] * src[(j + 1) * stride];
            }
            if (j < nj) s0 += Bt[j * ldb + a] * src[j * stride];
            out[c * nact] = c < nv ? (s0 + s1) : -(s0 + s1);
          }
        }
        double part = 0.0;
        for (int j = 0; j < nj; j++) part += Bt[j * ldb + a] * ncg[6 + j];
        ws.ic_tq0[s * nact + a] = part;
      }
    }
    if (lane == 0) ws.ic_status[s] = bad;
    __syncwarp();
  }
}


// =======================================================================================
// Register formulation for the shipped shapes (compile-time dimensions).  Same quantities as
// k_ic_prepare, other data flow: the two Cholesky factors and their triangular inverses run with the
// working rows in registers (tri_reg.cuh), every later product is "lane = one actuated row / column,
// operand rows broadcast from shared memory as LDS.128", fully unrolled:
//   T = L^-T (M = L L^T)                       rows of T = columns of L^-1
//   Y = L^-1 Ji^T, Lambda = (Y^T Y)^-1, Jibar = T (Y Lambda), Ni^T (c+g), Jc Ni        (as k_ic_prepare)
//   lane a:  z = L^-1 Atilde[a]^T              z_i = sum_{6<=c<=i} T[c][i] at_c, at_c formed on the fly
//            g = row a of G = Z^T Z            Z broadcast from shared memory
//            V[6+j][a] = sum_{i>=6+j} T[6+j][i] z_i                                    (W A^T = rows 6.. of L^-T Z)
//   G = Lg Lg^T, Tg = Lg^-T;  lane a': row a' of G^-1 = sum_i Tg[a'][i] Tg[a][i],
//            row a' of B = G^-1 V6^T,  row a' of S [M | -(Jc Ni)^T] = B X  and  S Ni^T (c+g) = B ncg[6:]
// No explicit 33 x 33 / 25 x 25 inverse of an ill-conditioned matrix is applied to data: both triangular
// inverses are of Cholesky factors (cond = sqrt of the matrix's), exactly the solves k_ic_prepare does.
template <int NV, int NACT, int NC, int K>
struct IcRegCfg {
  static constexpr int N = NV + NC, NJ = NV - 6;
  static constexpr int ld_() { int l = (NV + 1) & ~1; while ((l & 15) != 6) l += 2; return l; }
  static constexpr int LD = ld_();                // pitch of T: even, 6 (mod 16): LDS.128 by row conflict-free
  static constexpr int LDZ = (NACT + 1) & ~1;     // pitch of Z / V6 / the G block (rows broadcast)
  static constexpr int LDX = (N + 1) & ~1;        // pitch of X = [M[6:, :] | (Jc Ni)[:, 6:]^T]
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  static constexpr int OT = 0;
  static constexpr int SZT = cmax(NV * LD, NACT * LDZ);                  // T, later the G block
  static constexpr int OZ = OT + SZT;
  static constexpr int SZZ = cmax(cmax(NV * LDZ, NC * LD), NJ * LDZ);     // staged Jc, then Z, then V6
  static constexpr int OSM = OZ + SZZ;                                    // small vectors
  static constexpr int OCG = OSM, ONCG = OCG + NV, OY = ONCG + NV, OYL = OY + NV * K, OJB = OYL + NV * K,
                       OJI = OJB + NV * K, OLAM = OJI + K * NV, OJCJB = OLAM + K * K + 2 * K, OLDI = OJCJB + NC * K,
                       OGDI = OLDI + NV, OPV = OGDI + NACT, OEND = OPV + NV;
  static constexpr int TOTAL = (cmax(OEND, NJ * LDX) + 1) & ~1;           // X is staged over T and Z at the end
  static_assert(NV <= 64 && NACT <= 32 && NJ * LDX <= OSM, "shape not supported by the register formulation");
  static_assert((SZT & 1) == 0 && (SZZ & 1) == 0, "16-byte alignment of the regions");
};

// W warps per CTA, one state each, walking the code TOGETHER: the kernel is ~17 K straight-line instructions
// (270 KB) executed once per state, and ten independent one-warp CTAs per SM each streaming their own copy
// through the instruction cache spent a third of their cycles waiting for instructions (ncu: no_instruction
// 3.2 of 10 cycles per issue).  Nothing here is data dependent (fixed trip counts; a failed factorisation is a
// flag, not a branch), so CTA-wide barriers between the sections cost no waiting and every fetched line
// serves W warps.  States are claimed W at a time; the warps of a last, partial group redo its first state.
template <int NV, int NACT, int NC, int K, int W>
__global__ void __launch_bounds__(32 * W, 1)
k_ic_prepare_reg(const DevProblem* __restrict__ gp, WsPtrs ws, FrameMap fm, int nfr, int B, int* counter) {
  using C = IcRegCfg<NV, NACT, NC, K>;
  constexpr int N = C::N, NJ = C::NJ, LD = C::LD, LDZ = C::LDZ, LDX = C::LDX;
  extern __shared__ __align__(16) double sm_all[];
  __shared__ int s_base;
  const DevProblem& P = *gp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sm = sm_all + (size_t)warp * C::TOTAL;
  double* T = sm + C::OT;
  double* ZR = sm + C::OZ;
  double* cg = sm + C::OCG;
  double* ncg = sm + C::ONCG;
  double* Y = sm + C::OY;
  double* YL = sm + C::OYL;
  double* Jb = sm + C::OJB;
  double* Jis = sm + C::OJI;
  double* Lam = sm + C::OLAM;
  double* JcJb = sm + C::OJCJB;
  double* Ldi = sm + C::OLDI;
  double* Gdi = sm + C::OGDI;
  double* pv = sm + C::OPV;
  for (int e = lane; e < K * NV; e += 32) Jis[e] = P.ic_jac[e / NV][e % NV];
  const int va = lane < NACT ? P.act_idx[lane] : 6;
  __syncwarp();

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_base = atomicAdd(counter, W);
    __syncthreads();
    const int base = s_base;
    if (base >= B) break;
    const size_t s = (size_t)(base + warp < B ? base + warp : base);
    int bad = 0;
    const double* Mg = ws.M + s * NV * NV;
    // M goes global -> shared as asynchronous copies (cp.async, 8 bytes each, no registers in between): all of a
    // lane's loads are in flight at once, one memory round trip instead of one per batch of eight register loads
    {
      const unsigned t0 = (unsigned)__cvta_generic_to_shared(T);
#pragma unroll 4
      for (int e = lane; e < NV * NV; e += 32) {
        const int i = e / NV;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(t0 + 8u * (i * LD + (e - i * NV))), "l"(Mg + e) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int c = lane; c < NV; c += 32) cg[c] = ws.cori[s * NV + c] + ws.grav[s * NV + c];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- M = L L^T, T = L^-T ----
    __syncthreads();
    if (!chol_rows_reg<NV, LD>(T, Ldi, pv, lane)) bad = 1;
    __syncthreads();
    {
      linv_cols_reg<NV, LD>(T, Ldi, lane);
      // ---- Y = L^-1 Ji^T: lanes = rows, Y[i][t] = sum_{c<=i} T[c][i] Ji[t][c] ----
      for (int i = lane; i < NV; i += 32) {
        double y[K];
#pragma unroll
        for (int t = 0; t < K; t++) y[t] = 0.0;
#pragma unroll 4
        for (int c = 0; c < NV; c++) {
          const double tv = T[c * LD + i];  // zero below the diagonal of T (c > i)
#pragma unroll
          for (int t = 0; t < K; t++) y[t] = fma(tv, Jis[t * NV + c], y[t]);
        }
#pragma unroll
        for (int t = 0; t < K; t++) Y[i * K + t] = y[t];
      }
      __syncwarp();
      if (lane == 0) bad |= ic_lambda_inverse(Y, Lam, NV, K);
      bad = __shfl_sync(FULL, bad, 0);
      __syncwarp();
    }
    __syncthreads();
    {
      // ---- Jibar = L^-T (Y Lambda) = T (Y Lambda): lanes = rows of T ----
      for (int i = lane; i < NV; i += 32) {
#pragma unroll
        for (int t = 0; t < K; t++) {
          double sum = 0.0;
#pragma unroll
          for (int u = 0; u < K; u++) sum = fma(Y[i * K + u], Lam[u * K + t], sum);
          YL[i * K + t] = sum;
        }
      }
      __syncwarp();
      for (int c = lane; c < NV; c += 32) {
        double acc[K];
#pragma unroll
        for (int t = 0; t < K; t++) acc[t] = 0.0;
        const double* Tc = T + c * LD;
#pragma unroll 4
        for (int i = 0; i < NV; i++) {
          const double tv = Tc[i];  // zero for i < c
#pragma unroll
          for (int t = 0; t < K; t++) acc[t] = fma(tv, YL[i * K + t], acc[t]);
        }
#pragma unroll
        for (int t = 0; t < K; t++) Jb[c * K + t] = acc[t];
      }
      __syncwarp();
      // ---- Ni^T (c+g) = cg - Ji^T (Jibar^T cg) ----
      {
        double jt[K];
#pragma unroll
        for (int t = 0; t < K; t++) {
          double part = 0.0;
          for (int i = lane; i < NV; i += 32) part += Jb[i * K + t] * cg[i];
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
          jt[t] = part;
        }
        for (int i = lane; i < NV; i += 32) {
          double v = cg[i];
#pragma unroll
          for (int t = 0; t < K; t++) v -= Jis[t * NV + i] * jt[t];
          ncg[i] = v;
          ws.ic_cg[s * NV + i] = v;
        }
      }
      // ---- Jc Ni = Jc - (Jc Jibar) Ji; the contact Jacobian rows are staged in the Z region first ----
      double* Jcs = ZR;
#pragma unroll 1
      for (int c = 0; c < NC; c++) {
        int ci = 0;
        while (ci + 1 < P.ncontact && c >= P.contacts[ci + 1].rf_off) ci++;
        const DevContact& Cc = P.contacts[ci];
        const double* Jrow = ws.fr_jac + (s * nfr + fm.contact_f[ci]) * 6 * NV +
                             (Cc.type == PNC_CONTACT_POINT ? 3 + (c - Cc.rf_off) : (c - Cc.rf_off)) * NV;
        // asynchronous copies: the NC rows travel together (one row at a time through registers paid NC round trips)
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(Jcs + c * LD);
        for (int i = lane; i < NV; i += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + 8u * i), "l"(Jrow + i) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      for (int c = lane; c < NC; c += 32) {
        const double* Jrow = Jcs + c * LD;
        double acc[K];
#pragma unroll
        for (int t = 0; t < K; t++) acc[t] = 0.0;
#pragma unroll 4
        for (int i = 0; i < NV; i++) {
          const double v = Jrow[i];
#pragma unroll
          for (int t = 0; t < K; t++) acc[t] = fma(v, Jb[i * K + t], acc[t]);
        }
#pragma unroll
        for (int t = 0; t < K; t++) JcJb[c * K + t] = acc[t];
      }
      __syncwarp();
      // (Jc Ni)[c][i], lanes along i: coalesced stores
#pragma unroll 1
      for (int c = 0; c < NC; c++) {
        for (int i = lane; i < NV; i += 32) {
          double v = Jcs[c * LD + i];
#pragma unroll
          for (int t = 0; t < K; t++) v = fma(-JcJb[c * K + t], Jis[t * NV + i], v);
          ws.ic_jcni[(s * NC + c) * NV + i] = v;
        }
      }
      __syncwarp();
      // ---- lane a: z = L^-1 Atilde[a]^T, Atilde[a][c] = Ni[act_idx[a]][c] for c >= 6 (0 before) ----
      __syncthreads();
      double z[NV];
#pragma unroll
      for (int i = 0; i < NV; i++) z[i] = 0.0;
      {
        double jb[K];
#pragma unroll
        for (int t = 0; t < K; t++) jb[t] = Jb[va * K + t];
#pragma unroll
        for (int c = 6; c < NV; c++) {
          double at = c == va ? 1.0 : 0.0;
#pragma unroll
          for (int t = 0; t < K; t++) at = fma(-jb[t], Jis[t * NV + c], at);
          const double* Tc = T + c * LD;
#pragma unroll
          for (int i = c & ~1; i + 1 < NV; i += 2) {  // T[c][i] = 0 for i < c
            const double2 tv = tri_lds2(Tc + i);
            z[i] = fma(tv.x, at, z[i]);
            z[i + 1] = fma(tv.y, at, z[i + 1]);
          }
          if (NV & 1) z[NV - 1] = fma(Tc[NV - 1], at, z[NV - 1]);
        }
      }
      if (lane < NACT) {
#pragma unroll
        for (int i = 0; i < NV; i++) ZR[i * LDZ + lane] = z[i];
      } else if (lane < LDZ) {
#pragma unroll
        for (int i = 0; i < NV; i++) ZR[i * LDZ + lane] = 0.0;  // the pad column
      }
      __syncthreads();
      // ---- row a of G = Z^T Z ----
      double g[LDZ];
#pragma unroll
      for (int b = 0; b < LDZ; b++) g[b] = 0.0;
#pragma unroll
      for (int i = 0; i < NV; i++) {
        const double* Zi = ZR + i * LDZ;
#pragma unroll
        for (int b = 0; b < LDZ; b += 2) {
          const double2 q = tri_lds2(Zi + b);
          g[b] = fma(z[i], q.x, g[b]);
          g[b + 1] = fma(z[i], q.y, g[b + 1]);
        }
      }
      __syncthreads();  // Z is dead: V6 takes its place
      // ---- V6[j][a] = (L^-T Z)[6 + j][a] = sum_{i >= 6+j} T[6+j][i] z_i ----
#pragma unroll
      for (int c = 6; c < NV; c++) {
        const double* Tc = T + c * LD;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = c & ~1; i + 1 < NV; i += 2) {
          const double2 tv = tri_lds2(Tc + i);
          a0 = fma(tv.x, z[i], a0);
          a1 = fma(tv.y, z[i + 1], a1);
        }
        if (NV & 1) a0 = fma(Tc[NV - 1], z[NV - 1], a0);
        if (lane < LDZ) ZR[(c - 6) * LDZ + lane] = lane < NACT ? a0 + a1 : 0.0;
      }
      __syncwarp();  // T is dead: the G block takes its place
      // ---- G = Lg Lg^T, Tg = Lg^-T ----
      double* Gs = T;
      if (lane < NACT) {
#pragma unroll
        for (int b = 0; b < LDZ; b += 2) *reinterpret_cast<double2*>(Gs + lane * LDZ + b) = make_double2(g[b], g[b + 1]);
      }
      __syncwarp();
      double gmax = lane < NACT ? Gs[lane * LDZ + lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(FULL, gmax, o));
      __syncthreads();
      if (!chol_rows_reg<NACT, LDZ>(Gs, Gdi, pv, lane, 1e-13 * gmax)) bad = 1;
    }
    __syncthreads();
    {
      double* Gs = T;
      linv_cols_reg<NACT, LDZ>(Gs, Gdi, lane);  // row c of Gs = column c of Lg^-1, zeros left of the diagonal
      // ---- row a' of G^-1 = Tg Tg^T: gi[a] = sum_i Tg[a'][i] Tg[a][i] ----
      double gi[LDZ];
      {
        double xg[LDZ];
        const double* mine = Gs + (lane < NACT ? lane : 0) * LDZ;
#pragma unroll
        for (int i = 0; i < LDZ; i += 2) {
          const double2 q = tri_lds2(mine + i);
          xg[i] = q.x; xg[i + 1] = i + 1 < NACT ? q.y : 0.0;
        }
#pragma unroll
        for (int a = 0; a < NACT; a++) {
          const double* Ga = Gs + a * LDZ;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int i = a & ~1; i < LDZ; i += 2) {
            const double2 q = tri_lds2(Ga + i);
            a0 = fma(q.x, xg[i], a0);
            a1 = fma(q.y, xg[i + 1], a1);  // xg of the pad column is zero
          }
          gi[a] = a0 + a1;
        }
        if (LDZ > NACT) gi[LDZ - 1] = 0.0;
      }
      // ---- row a' of B = G^-1 V6^T: b[j] = sum_a gi[a] V6[j][a] ----
      __syncthreads();
      double bq[NJ];
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const double* Vj = ZR + j * LDZ;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int a = 0; a < LDZ; a += 2) {
          const double2 q = tri_lds2(Vj + a);
          a0 = fma(gi[a], q.x, a0);
          a1 = fma(gi[a + 1], q.y, a1);  // pad column of V6 and of gi are zero
        }
        bq[j] = a0 + a1;
      }
      __syncwarp();  // every lane is done with the G block and V6: X = [M[6:, :] | (Jc Ni)[:, 6:]^T] goes over them
      double* X = sm;
      {
        const unsigned x0 = (unsigned)__cvta_generic_to_shared(X);
#pragma unroll 4
        for (int e = lane; e < NJ * NV; e += 32) {
          const int j = e / NV;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(x0 + 8u * (j * LDX + (e - j * NV))), "l"(Mg + 6 * NV + e) : "memory");
        }
#pragma unroll 2
        for (int e = lane; e < NC * NJ; e += 32) {
          const int cc = e / NJ, j = e - cc * NJ;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(x0 + 8u * (j * LDX + NV + cc)),
                       "l"(ws.ic_jcni + (s * NC + cc) * NV + 6 + j) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (LDX > N)
        for (int j = lane; j < NJ; j += 32) X[j * LDX + N] = 0.0;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      // ---- S [M | -(Jc Ni)^T] (transposed store [n][nact]: coalesced) and S Ni^T (c+g) ----
      __syncthreads();
      double* out = ws.ic_tq + s * NACT * N + lane;
#pragma unroll 1
      for (int c = 0; c < N; c += 4) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const double* xc = X + c;
        const bool pair2 = c + 2 < LDX;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          const double2 q = tri_lds2(xc + j * LDX);
          s0 = fma(bq[j], q.x, s0);
          s1 = fma(bq[j], q.y, s1);
          if (pair2) {
            const double2 q2 = tri_lds2(xc + j * LDX + 2);
            s2 = fma(bq[j], q2.x, s2);
            s3 = fma(bq[j], q2.y, s3);
          }
        }
        if (lane < NACT) {
          out[c * NACT] = c < NV ? s0 : -s0;
          if (c + 1 < N) out[(c + 1) * NACT] = c + 1 < NV ? s1 : -s1;
          if (c + 2 < N) out[(c + 2) * NACT] = c + 2 < NV ? s2 : -s2;
          if (c + 3 < N) out[(c + 3) * NACT] = c + 3 < NV ? s3 : -s3;
        }
      }
      {
        double part = 0.0;
#pragma unroll
        for (int j = 0; j < NJ; j++) part = fma(bq[j], ncg[6 + j], part);
        if (lane < NACT) ws.ic_tq0[s * NACT + lane] = part;
      }
    }
    if (lane == 0) ws.ic_status[s] = bad;
    __syncwarp();
  }
}

void launch_ic_prepare(const DevProblem* dp, const DevProblem& hp, const WsPtrs& ws, const FrameMap& fm, int nfr,
                       int B, int num_sms, cudaStream_t stream, int* counter) {
  IcLayout L = ic_layout(hp.nv, hp.nact, hp.nc, hp.n_ic);
  const size_t smem = (size_t)L.total * sizeof(double);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  int grid = num_sms * per_sm;
  if (grid > B) grid = B;
  cudaMemsetAsync(counter, 0, sizeof(int), stream);
  auto go = [&](auto kernel) {
    static SmemOptIn optin;
    optin.ensure(kernel, smem);
    kernel<<<grid, 32, smem, stream>>>(dp, L, ws, fm, nfr, B, counter);
  };
  // PNC_IC=smem forces the shared-memory formulation (A/B runs, cross-check in the tests)
  static const bool force_smem = [] { const char* e = getenv("PNC_IC"); return e && e[0] == 's'; }();
  auto go_reg = [&](auto kernel, size_t smem_reg, int warps) {
    static SmemOptIn optin;
    optin.ensure(kernel, smem_reg);
    int per = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, 32 * warps, smem_reg);
    if (per < 1) per = 1;
    int g = num_sms * per;
    if (g * warps > B) g = (B + warps - 1) / warps;
    kernel<<<g, 32 * warps, smem_reg, stream>>>(dp, ws, fm, nfr, B, counter);
  };
  constexpr int ICW = PNC_IC_WARPS;  // warps per CTA walking the code together
  if (!force_smem && hp.nv == 33 && hp.nact == 25 && hp.nc == 12 && hp.n_ic == 2)          // Draco3
    go_reg(k_ic_prepare_reg<33, 25, 12, 2, ICW>, sizeof(double) * ICW * IcRegCfg<33, 25, 12, 2>::TOTAL, ICW);
  else if (!force_smem && hp.nv == 20 && hp.nact == 12 && hp.nc == 12 && hp.n_ic == 2)     // Draco3 lower body
    go_reg(k_ic_prepare_reg<20, 12, 12, 2, ICW>, sizeof(double) * ICW * IcRegCfg<20, 12, 12, 2>::TOTAL, ICW);
  else if (hp.nv == 33 && hp.nact == 25 && hp.nc == 12 && hp.n_ic == 2) go(k_ic_prepare<33, 25, 12, 2>);
  else if (hp.nv == 20 && hp.nact == 12 && hp.nc == 12 && hp.n_ic == 2) go(k_ic_prepare<20, 12, 12, 2>);   // Draco3 lower body
  else go(k_ic_prepare<0, 0, 0, 0>);
  g_launch_count++;
}

}  // namespace pnc
