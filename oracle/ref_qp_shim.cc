// ORACLE (test infrastructure): C shim over the REFERENCE's own Goldfarb-Idnani solver,
// external_source/Goldfarb/QuadProg++.cc:115 solve_quadprog(GMatr...), compiled from the
// sources where they lie under /root/reference (see oracle/Makefile, target _ref).
// Used to pin oracle/pnc_oracle.c:orc_solve_quadprog and as the qpsolvers stand-in of
// tests/golden/make_golden.py.
#include <cmath>
#include <exception>
#include "QuadProg++.hh"

extern "C" double ref_solve_quadprog(int n, int p, int m, const double* G, const double* g0,
                                     const double* CE, const double* ce0, const double* CI,
                                     const double* ci0, double* x, int* status) {
  GMatr<double> g(n, n), ce(n, p > 0 ? p : 0), ci(n, m > 0 ? m : 0);
  GVect<double> g0v(n), ce0v(p), ci0v(m), xv(n);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) g[i][j] = G[i * n + j];
    for (int j = 0; j < p; j++) ce[i][j] = CE[i * p + j];
    for (int j = 0; j < m; j++) ci[i][j] = CI[i * m + j];
    g0v[i] = g0[i];
  }
  for (int j = 0; j < p; j++) ce0v[j] = ce0[j];
  for (int j = 0; j < m; j++) ci0v[j] = ci0[j];
  *status = 0;
  double f;
  try {
    f = solve_quadprog(g, g0v, ce, ce0v, ci, ci0v, xv);
  } catch (const std::logic_error&) {
    *status = 2;  // Cholesky failure (QuadProg++.cc:715-722)
    return INFINITY;
  } catch (const std::exception&) {
    *status = 3;  // linearly dependent equalities (:268-271)
    return INFINITY;
  }
  if (std::isinf(f)) *status = 1;
  for (int i = 0; i < n; i++) x[i] = xv[i];
  return f;
}
