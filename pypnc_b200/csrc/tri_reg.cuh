// tri_reg.cuh -- Cholesky factor and triangular inverse of a small SPD matrix (n <= 64, in practice 12..36)
// by ONE WARP with the working rows in REGISTERS: fully unrolled straight-line code, no shuffles, the only
// cross-lane traffic a warp-uniform broadcast through the shared-memory copy of the matrix itself.
// Shared by k_ihwbc_setup (Hessian task block, QuadProg++.cc:705-730 / :203-210) and k_ic_prepare_reg
// (mass matrix and A W A^T, ihwbc.py:124-146).
#pragma once

namespace pnc {

__device__ __forceinline__ double2 tri_lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// In: lower triangle (incl. diagonal) of the SPD matrix in Gm[i * LD + k], k <= i (LD even, rows 16-byte
// aligned).  Out: L (lower, incl. diagonal) in the same place, dl[j] = 1 / L[j][j].  pv: N doubles of
// scratch.  Returns false when a pivot is not above pmin (the factor is then garbage).
//
// Right-looking: lane l owns row r = l + NX (NX = N - 32 leading rows, if any, are factored redundantly by
// every lane from warp-uniform loads, so that the remaining <= 32 rows map one to one onto the lanes).
// Per pivot the owner publishes the pivot, every lane scales its entry of column j and stores the final
// L[r][j] into the factor block, and the trailing update runs two pivots at a time as a rank-2 update
// whose multipliers (L[k][j], L[k][j+1]) come as one warp-uniform LDS.128 per column k.
template <int N, int LD>
__device__ __forceinline__ bool chol_rows_reg(double* __restrict__ Gm, double* __restrict__ dl, double* __restrict__ pv,
                                              const int lane, const double pmin = 0.0) {
  static_assert(N <= 64 && (LD & 1) == 0, "one row per lane, rows 16-byte aligned");
  constexpr int NX = N > 32 ? N - 32 : 0;
  const int r = lane + NX;
  const bool rok = r < N;
  bool pd = true;
  double a[N];
  {
    const double* ri = Gm + (rok ? r : 0) * LD;
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
      const double2 g = tri_lds2(ri + k);
      a[k] = g.x; a[k + 1] = g.y;
    }
    if (N & 1) a[N - 1] = ri[N - 1];
  }
  if constexpr (NX > 0) {
    double u[NX][NX], uid[NX];
#pragma unroll
    for (int i = 0; i < NX; i++)
#pragma unroll
      for (int k = 0; k <= i; k++) u[i][k] = Gm[i * LD + k];
#pragma unroll
    for (int j = 0; j < NX; j++) {
      double p = u[j][j];
#pragma unroll
      for (int q = 0; q < j; q++) p = fma(-u[j][q], u[j][q], p);
      if (!(p > pmin)) pd = false;
      
#ifdef PNC_IEEE_DIV
      const double d = sqrt(p), id = 1.0 / d;
#else
      const double id = rsqrt(p), d = p * id;  // sqrt(p) and its reciprocal from one reciprocal square root
#endif

      u[j][j] = d; uid[j] = id;
#pragma unroll
      for (int i = j + 1; i < NX; i++) {
        double sacc = u[i][j];
#pragma unroll
        for (int q = 0; q < j; q++) sacc = fma(-u[i][q], u[j][q], sacc);
        u[i][j] = sacc * id;
      }
      double sacc = a[j];
#pragma unroll
      for (int q = 0; q < j; q++) sacc = fma(-a[q], u[j][q], sacc);
      a[j] = sacc * id;
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NX; i++) {
        dl[i] = uid[i];
#pragma unroll
        for (int k = 0; k <= i; k++) Gm[i * LD + k] = u[i][k];
      }
    }
    if (rok) {
#pragma unroll
      for (int k = 0; k < NX; k++) Gm[r * LD + k] = a[k];
    }
    __syncwarp();
#pragma unroll
    for (int k = NX; k < N; k++) {
      const double* Lk = Gm + k * LD;
#pragma unroll
      for (int j = 0; j < NX; j++) a[k] = fma(-a[j], Lk[j], a[k]);
    }
  }
  double lprev = 0.0;  // L[r][j-1] of the even pivot a pair starts with
#pragma unroll
  for (int j = NX; j < N; j++) {
    if (r == j) pv[j] = a[j];  // the pivot
    __syncwarp();
    const double p = pv[j];
    if (!(p > pmin)) pd = false;
    
#ifdef PNC_IEEE_DIV
      const double d = sqrt(p), id = 1.0 / d;
#else
      const double id = rsqrt(p), d = p * id;  // sqrt(p) and its reciprocal from one reciprocal square root
#endif

    double l = a[j] * id;
    if (r == j) l = d;
    a[j] = l;
    if (rok && r >= j) Gm[r * LD + j] = l;
    if (lane == 0) dl[j] = id;
    __syncwarp();
    if (!(j & 1)) {
      // even pivot: its update of the very next column now, the rest together with pivot j + 1
      if (j + 1 < N) a[j + 1] = fma(-l, Gm[(j + 1) * LD + j], a[j + 1]);
      lprev = l;
    } else if (j == NX) {
      // an odd first pivot has no partner: rank-1 update
#pragma unroll
      for (int k = j + 1; k < N; k++) a[k] = fma(-l, Gm[k * LD + j], a[k]);
    } else {
#pragma unroll
      for (int k = j + 1; k < N; k++) {
        const double2 lk = tri_lds2(Gm + k * LD + j - 1);
        a[k] = fma(-lprev, lk.x, a[k]);
        a[k] = fma(-l, lk.y, a[k]);
      }
    }
  }
  __syncwarp();
  return pd;
}

// In: L (lower, incl. diagonal) in Gm, dl = 1 / diagonal.  Out: J = L^-T (upper triangular, zeros below the
// diagonal written out) in the same place: row c of J = column c of L^-1.  Lane c forward-solves L x = e_c
// with x in registers; L is read as warp-uniform LDS.128 pairs.  Columns 32.. (N > 32) only reach rows
// 32..: a second, tiny solve in lanes 0 .. N-33.
template <int N, int LD>
__device__ __forceinline__ void linv_cols_reg(double* __restrict__ Gm, const double* __restrict__ dl, const int lane) {
  double xr[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    const double* Li = Gm + i * LD;
    double s0 = i == lane ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < i; k += 2) {
      const double2 l = tri_lds2(Li + k);
      s0 = fma(-l.x, xr[k], s0);
      s1 = fma(-l.y, xr[k + 1], s1);
    }
    if (i & 1) s0 = fma(-Li[i - 1], xr[i - 1], s0);
    xr[i] = (s0 + s1) * dl[i];
  }
  constexpr int NX = N > 32 ? N - 32 : 0;
  double xe[NX > 0 ? NX : 1];
  if constexpr (NX > 0) {
#pragma unroll
    for (int i = 32; i < N; i++) {
      const double* Li = Gm + i * LD;
      double s0 = i == 32 + lane ? 1.0 : 0.0;
#pragma unroll
      for (int k = 32; k < i; k++) s0 = fma(-Li[k], xe[k - 32], s0);
      xe[i - 32] = s0 * dl[i];
    }
  }
  __syncwarp();  // every lane is done reading L
  if (lane < N) {
    double* zr = Gm + lane * LD;
#pragma unroll
    for (int i = 0; i + 1 < N; i += 2) *reinterpret_cast<double2*>(zr + i) = make_double2(xr[i], xr[i + 1]);
    if (N & 1) zr[N - 1] = xr[N - 1];
  }
  if constexpr (NX > 0) {
    if (lane < NX) {
      double* zr = Gm + (32 + lane) * LD;
#pragma unroll
      for (int i = 0; i < N; i++) zr[i] = i >= 32 ? xe[i >= 32 ? i - 32 : 0] : 0.0;
    }
  }
  __syncwarp();
}

}  // namespace pnc
