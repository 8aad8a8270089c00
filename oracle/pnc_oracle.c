/*
 * pnc_oracle.c -- CPU ORACLE (test infrastructure, NOT the product; see pnc_oracle.h).
 *
 * Rigid-body part: body-frame (Featherstone / pinocchio-style) recursions, i.e. a
 * DIFFERENT formulation from the world-frame prefix/subtree sums the CUDA kernels use,
 * so the two implementations cross-check each other.  File:line citations refer to the
 * reference checkout (/root/reference).
 */
#include "pnc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <time.h>
#include <pthread.h>

#define MAXB 64
#define MAXV 64

struct orc_model {
  int nb, nq, nv, na, nfl, nf;
  double total_mass, mass_dyn;
  int *parent, *jtype, *idx_v, *subtree, *frame_parent;
  double *place_R, *place_p, *axis, *mass, *com, *inertia, *frame_R, *frame_p;
};

static void* dupmem(const void* src, size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (bytes) memcpy(p, src, bytes);
  return p;
}

orc_model* orc_model_create(const pnc_model_desc* d) {
  orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
  m->nb = d->nb; m->nq = d->nq; m->nv = d->nv; m->na = d->na; m->nfl = d->n_floating; m->nf = d->nf;
  m->total_mass = d->total_mass; m->mass_dyn = d->mass_dyn;
  m->parent = (int*)dupmem(d->parent, sizeof(int) * d->nb);
  m->jtype = (int*)dupmem(d->jtype, sizeof(int) * d->nb);
  m->idx_v = (int*)dupmem(d->idx_v, sizeof(int) * d->nb);
  m->subtree = (int*)dupmem(d->subtree, sizeof(int) * d->nb);
  m->place_R = (double*)dupmem(d->place_R, sizeof(double) * 9 * d->nb);
  m->place_p = (double*)dupmem(d->place_p, sizeof(double) * 3 * d->nb);
  m->axis = (double*)dupmem(d->axis, sizeof(double) * 3 * d->nb);
  m->mass = (double*)dupmem(d->mass, sizeof(double) * d->nb);
  m->com = (double*)dupmem(d->com, sizeof(double) * 3 * d->nb);
  m->inertia = (double*)dupmem(d->inertia, sizeof(double) * 9 * d->nb);
  m->frame_parent = (int*)dupmem(d->frame_parent, sizeof(int) * d->nf);
  m->frame_R = (double*)dupmem(d->frame_R, sizeof(double) * 9 * d->nf);
  m->frame_p = (double*)dupmem(d->frame_p, sizeof(double) * 3 * d->nf);
  return m;
}

void orc_model_destroy(orc_model* m) {
  if (!m) return;
  free(m->parent); free(m->jtype); free(m->idx_v); free(m->subtree); free(m->place_R); free(m->place_p);
  free(m->axis); free(m->mass); free(m->com); free(m->inertia); free(m->frame_parent); free(m->frame_R);
  free(m->frame_p); free(m);
}

/* ---------------------------------------------------------------- 3-vector algebra */
static inline void cross3(const double* a, const double* b, double* c) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  c[0] = x; c[1] = y; c[2] = z;
}
static inline void mv3(const double* R, const double* a, double* b) { /* b = R a */
  double x = R[0] * a[0] + R[1] * a[1] + R[2] * a[2];
  double y = R[3] * a[0] + R[4] * a[1] + R[5] * a[2];
  double z = R[6] * a[0] + R[7] * a[1] + R[8] * a[2];
  b[0] = x; b[1] = y; b[2] = z;
}
static inline void mtv3(const double* R, const double* a, double* b) { /* b = R^T a */
  double x = R[0] * a[0] + R[3] * a[1] + R[6] * a[2];
  double y = R[1] * a[0] + R[4] * a[1] + R[7] * a[2];
  double z = R[2] * a[0] + R[5] * a[1] + R[8] * a[2];
  b[0] = x; b[1] = y; b[2] = z;
}
static inline void mm3(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, T, sizeof(T));
}
static inline void mmt3(const double* A, const double* B, double* C) { /* C = A B^T */
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      T[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
  memcpy(C, T, sizeof(T));
}
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* scalar-last unit quaternion -> rotation (scipy Rotation.from_quat normalises) */
void orc_quat_to_rot(const double* qq, double* R) {
  double nrm = sqrt(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
  double x = qq[0] / nrm, y = qq[1] / nrm, z = qq[2] / nrm, w = qq[3] / nrm;
  double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);     R[2] = 2 * (xz + yw);
  R[3] = 2 * (xy + zw);     R[4] = -x2 + y2 - z2 + w2; R[5] = 2 * (yz - xw);
  R[6] = 2 * (xz - yw);     R[7] = 2 * (yz + xw);      R[8] = -x2 - y2 + z2 + w2;
}

/* scipy Rotation.from_matrix(...).as_quat(): largest-of-(diag, trace) branch choice */
void orc_rot_to_quat(const double* m, double* q) {
  double dec[4] = {m[0], m[4], m[8], m[0] + m[4] + m[8]};
  int choice = 0;
  for (int i = 1; i < 4; i++)
    if (dec[i] > dec[choice]) choice = i;
  if (choice != 3) {
    int i = choice, j = (i + 1) % 3, k = (j + 1) % 3;
    q[i] = 1 - dec[3] + 2 * m[3 * i + i];
    q[j] = m[3 * j + i] + m[3 * i + j];
    q[k] = m[3 * k + i] + m[3 * i + k];
    q[3] = m[3 * k + j] - m[3 * j + k];
  } else {
    q[0] = m[7] - m[5]; q[1] = m[2] - m[6]; q[2] = m[3] - m[1]; q[3] = 1 + dec[3];
  }
  double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= nrm;
}

/* util/util.py:61-72 */
void orc_quat_to_exp(const double* q, double* e) {
  double theta = 2.0 * asin(sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]));
  if (fabs(theta) < 1e-4) { e[0] = e[1] = e[2] = 0; return; }
  double s = sin(theta / 2.0);
  for (int i = 0; i < 3; i++) e[i] = q[i] / s * theta;
}

/* ---------------------------------------------------------------- kinematics data */
typedef struct {
  double liR[MAXB][9], lip[MAXB][3]; /* child joint frame in parent joint frame */
  double oR[MAXB][9], op[MAXB][3];   /* joint frame in world */
  double v[MAXB][6];                 /* body-frame spatial velocity [lin; ang] */
  double a[MAXB][6];                 /* body-frame spatial acceleration (qddot = 0, no gravity) */
} kin_t;

static void joint_rot(const double* ax, double qj, double* R) { /* R = c I + s [a]x + (1-c) a a^T */
  double c = cos(qj), s = sin(qj), t = 1 - c;
  R[0] = c + t * ax[0] * ax[0];         R[1] = t * ax[0] * ax[1] - s * ax[2]; R[2] = t * ax[0] * ax[2] + s * ax[1];
  R[3] = t * ax[1] * ax[0] + s * ax[2]; R[4] = c + t * ax[1] * ax[1];         R[5] = t * ax[1] * ax[2] - s * ax[0];
  R[6] = t * ax[2] * ax[0] - s * ax[1]; R[7] = t * ax[2] * ax[1] + s * ax[0]; R[8] = c + t * ax[2] * ax[2];
}

/* motion: X^{-1} act, M=(R,p) child in parent: v_c = [R^T (v - p x w); R^T w] */
static void mot_actinv(const double* R, const double* p, const double* vp, double* vc) {
  double t[3], pw[3];
  cross3(p, vp + 3, pw);
  t[0] = vp[0] - pw[0]; t[1] = vp[1] - pw[1]; t[2] = vp[2] - pw[2];
  mtv3(R, t, vc);
  mtv3(R, vp + 3, vc + 3);
}
/* motion cross: (v,w) x (v',w') = (w x v' + v x w', w x w') */
static void mot_cross(const double* a, const double* b, double* c) {
  double t1[3], t2[3], t3[3];
  cross3(a + 3, b, t1); cross3(a, b + 3, t2); cross3(a + 3, b + 3, t3);
  c[0] = t1[0] + t2[0]; c[1] = t1[1] + t2[1]; c[2] = t1[2] + t2[2];
  c[3] = t3[0]; c[4] = t3[1]; c[5] = t3[2];
}
/* force act child->parent: f_p = [R f; R n + p x (R f)] */
static void frc_act(const double* R, const double* p, const double* fc, double* fp) {
  double f[3], n[3], pf[3];
  mv3(R, fc, f); mv3(R, fc + 3, n); cross3(p, f, pf);
  fp[0] = f[0]; fp[1] = f[1]; fp[2] = f[2];
  fp[3] = n[0] + pf[0]; fp[4] = n[1] + pf[1]; fp[5] = n[2] + pf[2];
}
/* force cross: v x* f = [w x f; w x n + v x f] */
static void frc_cross(const double* v, const double* f, double* o) {
  double t1[3], t2[3], t3[3];
  cross3(v + 3, f, t1); cross3(v + 3, f + 3, t2); cross3(v, f, t3);
  o[0] = t1[0]; o[1] = t1[1]; o[2] = t1[2];
  o[3] = t2[0] + t3[0]; o[4] = t2[1] + t3[1]; o[5] = t2[2] + t3[2];
}
/* Y v with Y = (m, c, Ic): lin = m (v - c x w); ang = Ic w + c x lin */
static void inertia_mul(double m, const double* c, const double* Ic, const double* v, double* f) {
  double cw[3], Iw[3], cf[3];
  cross3(c, v + 3, cw);
  f[0] = m * (v[0] - cw[0]); f[1] = m * (v[1] - cw[1]); f[2] = m * (v[2] - cw[2]);
  mv3(Ic, v + 3, Iw); cross3(c, f, cf);
  f[3] = Iw[0] + cf[0]; f[4] = Iw[1] + cf[1]; f[5] = Iw[2] + cf[2];
}

/* forwardKinematics(q, v [, a=0]) */
static void fk(const orc_model* m, const double* q, const double* v, kin_t* k, int with_v) {
  for (int i = 0; i < m->nb; i++) {
    double Rj[9], pj[3] = {0, 0, 0};
    int iv = m->idx_v[i];
    if (m->jtype[i] == PNC_JOINT_FREEFLYER) {
      orc_quat_to_rot(q + 3, Rj);
      pj[0] = q[0]; pj[1] = q[1]; pj[2] = q[2];
    } else {
      int iq = m->nfl ? iv + 1 : iv;
      joint_rot(m->axis + 3 * i, q[iq], Rj);
    }
    /* liMi = placement * Mj */
    mm3(m->place_R + 9 * i, Rj, k->liR[i]);
    double t[3];
    mv3(m->place_R + 9 * i, pj, t);
    for (int c = 0; c < 3; c++) k->lip[i][c] = m->place_p[3 * i + c] + t[c];
    int par = m->parent[i];
    if (par < 0) {
      memcpy(k->oR[i], k->liR[i], sizeof(double) * 9);
      memcpy(k->op[i], k->lip[i], sizeof(double) * 3);
    } else {
      mm3(k->oR[par], k->liR[i], k->oR[i]);
      mv3(k->oR[par], k->lip[i], t);
      for (int c = 0; c < 3; c++) k->op[i][c] = k->op[par][c] + t[c];
    }
    if (!with_v) continue;
    double vj[6] = {0, 0, 0, 0, 0, 0};
    if (m->jtype[i] == PNC_JOINT_FREEFLYER) {
      for (int c = 0; c < 6; c++) vj[c] = v[c];
    } else {
      for (int c = 0; c < 3; c++) vj[3 + c] = m->axis[3 * i + c] * v[iv];
    }
    if (par < 0) {
      for (int c = 0; c < 6; c++) { k->v[i][c] = vj[c]; k->a[i][c] = 0; }
    } else {
      double vp[6], ap[6], cr[6];
      mot_actinv(k->liR[i], k->lip[i], k->v[par], vp);
      for (int c = 0; c < 6; c++) k->v[i][c] = vp[c] + vj[c];
      mot_actinv(k->liR[i], k->lip[i], k->a[par], ap);
      mot_cross(k->v[i], vj, cr);
      for (int c = 0; c < 6; c++) k->a[i][c] = ap[c] + cr[c];
    }
  }
}

void orc_pack_state(const orc_model* m, const double* bp, const double* bq, const double* blv,
                    const double* bav, const double* jp, const double* jv, double* q, double* v) {
  if (m->nfl) {
    double R[9];
    for (int c = 0; c < 3; c++) q[c] = bp[c];
    for (int c = 0; c < 4; c++) q[3 + c] = bq[c];
    orc_quat_to_rot(bq, R); /* util.quat_to_rot, :152 */
    mtv3(R, blv, v);
    mtv3(R, bav, v + 3);
    for (int i = 0; i < m->na; i++) { q[7 + i] = jp[i]; v[6 + i] = jv[i]; }
  } else {
    for (int i = 0; i < m->na; i++) { q[i] = jp[i]; v[i] = jv[i]; }
  }
}

/* RNEA, pinocchio rnea(): a_gf[universe] = -gravity */
void orc_rnea(const orc_model* m, const double* q, const double* v, const double* a, int with_gravity,
              double* tau) {
  kin_t k;
  double zero[MAXV] = {0};
  if (!v) v = zero;
  fk(m, q, v, &k, 1);
  static const double g0[3] = {0, 0, 9.81};
  double agf[MAXB][6], f[MAXB][6];
  for (int i = 0; i < m->nb; i++) {
    int par = m->parent[i], iv = m->idx_v[i];
    double ap[6], vj[6] = {0}, aj[6] = {0}, cr[6];
    if (m->jtype[i] == PNC_JOINT_FREEFLYER) {
      for (int c = 0; c < 6; c++) { vj[c] = v[c]; aj[c] = a ? a[c] : 0; }
    } else {
      for (int c = 0; c < 3; c++) { vj[3 + c] = m->axis[3 * i + c] * v[iv]; aj[3 + c] = m->axis[3 * i + c] * (a ? a[iv] : 0); }
    }
    if (par < 0) {
      double au[6] = {0, 0, 0, 0, 0, 0};
      if (with_gravity) { au[0] = g0[0]; au[1] = g0[1]; au[2] = g0[2]; }
      mot_actinv(k.liR[i], k.lip[i], au, ap);
    } else {
      mot_actinv(k.liR[i], k.lip[i], agf[par], ap);
    }
    mot_cross(k.v[i], vj, cr);
    for (int c = 0; c < 6; c++) agf[i][c] = ap[c] + aj[c] + cr[c];
    double h[6], fa[6], fv[6];
    inertia_mul(m->mass[i], m->com + 3 * i, m->inertia + 9 * i, k.v[i], h);
    inertia_mul(m->mass[i], m->com + 3 * i, m->inertia + 9 * i, agf[i], fa);
    frc_cross(k.v[i], h, fv);
    for (int c = 0; c < 6; c++) f[i][c] = fa[c] + fv[c];
  }
  for (int i = m->nb - 1; i >= 0; i--) {
    int par = m->parent[i], iv = m->idx_v[i];
    if (m->jtype[i] == PNC_JOINT_FREEFLYER) {
      for (int c = 0; c < 6; c++) tau[c] = f[i][c];
    } else {
      tau[iv] = dot3(m->axis + 3 * i, f[i] + 3);
    }
    if (par >= 0) {
      double fp[6];
      frc_act(k.liR[i], k.lip[i], f[i], fp);
      for (int c = 0; c < 6; c++) f[par][c] += fp[c];
    }
  }
}

void orc_gravity(const orc_model* m, const double* q, double* g) { orc_rnea(m, q, NULL, NULL, 1, g); }

void orc_coriolis(const orc_model* m, const double* q, const double* v, double* c) {
  /* nonLinearEffects(q,v) - computeGeneralizedGravity(q), :209-212 */
  double nle[MAXV], g[MAXV];
  orc_rnea(m, q, v, NULL, 1, nle);
  orc_rnea(m, q, NULL, NULL, 1, g);
  for (int i = 0; i < m->nv; i++) c[i] = nle[i] - g[i];
}

/* composite inertia kept as (m, h = m c, Io about the frame origin) */
typedef struct { double m, h[3], Io[9]; } cinertia;

static void cin_from_body(double m, const double* c, const double* Ic, cinertia* Y) {
  Y->m = m;
  double cc = dot3(c, c);
  for (int i = 0; i < 3; i++) {
    Y->h[i] = m * c[i];
    for (int j = 0; j < 3; j++) Y->Io[3 * i + j] = Ic[3 * i + j] + m * ((i == j ? cc : 0) - c[i] * c[j]);
  }
}
/* express Y (given in child frame) in the parent frame, (R,p) = child in parent, and add */
static void cin_add_transformed(cinertia* P, const cinertia* C, const double* R, const double* p) {
  if (C->m <= 0) return;
  double c[3] = {C->h[0] / C->m, C->h[1] / C->m, C->h[2] / C->m};
  double cc = dot3(c, c), Ic[9], t[9], RIc[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Ic[3 * i + j] = C->Io[3 * i + j] - C->m * ((i == j ? cc : 0) - c[i] * c[j]);
  mm3(R, Ic, t); mmt3(t, R, RIc);
  double cp[3];
  mv3(R, c, cp);
  for (int i = 0; i < 3; i++) cp[i] += p[i];
  double cpcp = dot3(cp, cp);
  P->m += C->m;
  for (int i = 0; i < 3; i++) {
    P->h[i] += C->m * cp[i];
    for (int j = 0; j < 3; j++) P->Io[3 * i + j] += RIc[3 * i + j] + C->m * ((i == j ? cpcp : 0) - cp[i] * cp[j]);
  }
}
static void cin_mul(const cinertia* Y, const double* v, double* f) { /* lin = m v - h x w; ang = h x v + Io w */
  double hw[3], hv[3], Iw[3];
  cross3(Y->h, v + 3, hw); cross3(Y->h, v, hv); mv3(Y->Io, v + 3, Iw);
  for (int i = 0; i < 3; i++) { f[i] = Y->m * v[i] - hw[i]; f[3 + i] = hv[i] + Iw[i]; }
}

/* CRBA (composite rigid body algorithm), full symmetric result as the pinocchio binding returns */
void orc_mass_matrix(const orc_model* m, const double* q, double* M) {
  kin_t k;
  fk(m, q, NULL, &k, 0);
  int nv = m->nv;
  memset(M, 0, sizeof(double) * nv * nv);
  cinertia Y[MAXB];
  for (int i = 0; i < m->nb; i++) cin_from_body(m->mass[i], m->com + 3 * i, m->inertia + 9 * i, &Y[i]);
  for (int i = m->nb - 1; i >= 0; i--) {
    int nd = m->jtype[i] == PNC_JOINT_FREEFLYER ? 6 : 1, iv = m->idx_v[i];
    for (int d = 0; d < nd; d++) {
      double S[6] = {0}, F[6];
      if (nd == 6) S[d] = 1; else for (int c = 0; c < 3; c++) S[3 + c] = m->axis[3 * i + c];
      cin_mul(&Y[i], S, F);
      /* own block */
      for (int e = 0; e < nd; e++) {
        double Se[6] = {0};
        if (nd == 6) Se[e] = 1; else for (int c = 0; c < 3; c++) Se[3 + c] = m->axis[3 * i + c];
        double val = 0;
        for (int c = 0; c < 6; c++) val += Se[c] * F[c];
        M[(iv + e) * nv + iv + d] = val;
      }
      /* ancestors */
      int j = i;
      while (m->parent[j] >= 0) {
        double Fp[6];
        frc_act(k.liR[j], k.lip[j], F, Fp);
        memcpy(F, Fp, sizeof(F));
        j = m->parent[j];
        int ndj = m->jtype[j] == PNC_JOINT_FREEFLYER ? 6 : 1, jv = m->idx_v[j];
        for (int e = 0; e < ndj; e++) {
          double val = ndj == 6 ? F[e] : dot3(m->axis + 3 * j, F + 3);
          M[(jv + e) * nv + iv + d] = val;
          M[(iv + d) * nv + jv + e] = val;
        }
      }
    }
    if (m->parent[i] >= 0) cin_add_transformed(&Y[m->parent[i]], &Y[i], k.liR[i], k.lip[i]);
  }
}

void orc_com(const orc_model* m, const double* q, const double* v, double* com, double* vcom) {
  kin_t k;
  fk(m, q, v, &k, v != NULL);
  double sc[3] = {0, 0, 0}, sv[3] = {0, 0, 0};
  for (int i = 0; i < m->nb; i++) {
    double cw[3];
    mv3(k.oR[i], m->com + 3 * i, cw);
    for (int c = 0; c < 3; c++) sc[c] += m->mass[i] * (cw[c] + k.op[i][c]);
    if (v) {
      double wc[3], vl[3], vw[3];
      cross3(k.v[i] + 3, m->com + 3 * i, wc);
      for (int c = 0; c < 3; c++) vl[c] = k.v[i][c] + wc[c];
      mv3(k.oR[i], vl, vw);
      for (int c = 0; c < 3; c++) sv[c] += m->mass[i] * vw[c];
    }
  }
  for (int c = 0; c < 3; c++) { com[c] = sc[c] / m->mass_dyn; if (vcom) vcom[c] = v ? sv[c] / m->mass_dyn : 0; }
}

/* world joint-Jacobian columns (pinocchio data.J): oMi.act(S), motion at the world origin */
static void world_jac(const orc_model* m, const kin_t* k, double* Jw /* 6 x nv */) {
  int nv = m->nv;
  memset(Jw, 0, sizeof(double) * 6 * nv);
  for (int i = 0; i < m->nb; i++) {
    int iv = m->idx_v[i];
    if (m->jtype[i] == PNC_JOINT_FREEFLYER) {
      for (int d = 0; d < 6; d++) {
        double lin[3] = {0, 0, 0}, ang[3] = {0, 0, 0}, col[3], t[3];
        for (int c = 0; c < 3; c++) col[c] = k->oR[i][3 * c + (d % 3)];
        if (d < 3) { memcpy(lin, col, sizeof(col)); }
        else { memcpy(ang, col, sizeof(col)); cross3(k->op[i], col, t); memcpy(lin, t, sizeof(t)); }
        for (int c = 0; c < 3; c++) { Jw[c * nv + iv + d] = lin[c]; Jw[(3 + c) * nv + iv + d] = ang[c]; }
      }
    } else {
      double aw[3], lin[3];
      mv3(k->oR[i], m->axis + 3 * i, aw);
      cross3(k->op[i], aw, lin);
      for (int c = 0; c < 3; c++) { Jw[c * nv + iv] = lin[c]; Jw[(3 + c) * nv + iv] = aw[c]; }
    }
  }
}

static int is_ancestor_or_self(const orc_model* m, int j, int i) { return j <= i && i < j + m->subtree[j]; }

void orc_com_jacobian(const orc_model* m, const double* q, double* J) {
  kin_t k;
  fk(m, q, NULL, &k, 0);
  int nv = m->nv;
  double Jw[6 * MAXV];
  world_jac(m, &k, Jw);
  memset(J, 0, sizeof(double) * 3 * nv);
  /* J_com = (1/M) sum_i m_i [ J_lin + J_ang x c_i ] over the dofs supporting body i */
  for (int i = 0; i < m->nb; i++) {
    double cw[3];
    mv3(k.oR[i], m->com + 3 * i, cw);
    for (int c = 0; c < 3; c++) cw[c] += k.op[i][c];
    for (int j = 0; j <= i; j++) {
      if (!is_ancestor_or_self(m, j, i)) continue;
      int nd = m->jtype[j] == PNC_JOINT_FREEFLYER ? 6 : 1;
      for (int d = 0; d < nd; d++) {
        int col = m->idx_v[j] + d;
        double lin[3] = {Jw[col], Jw[nv + col], Jw[2 * nv + col]};
        double ang[3] = {Jw[3 * nv + col], Jw[4 * nv + col], Jw[5 * nv + col]}, wc[3];
        cross3(ang, cw, wc);
        for (int c = 0; c < 3; c++) J[c * nv + col] += m->mass[i] * (lin[c] + wc[c]) / m->mass_dyn;
      }
    }
  }
}

/* linear rows of dAg/dt (pinocchio dccrba), divided by the wrapper's total mass, :226-229 */
void orc_com_jacobian_dot(const orc_model* m, const double* q, const double* v, double* dJ) {
  kin_t k;
  fk(m, q, v, &k, 1);
  int nv = m->nv;
  double Jw[6 * MAXV];
  world_jac(m, &k, Jw);
  /* world composite (mass, h, hdot) per subtree and world body velocities */
  double ms[MAXB], h[MAXB][3], hd[MAXB][3], ov[MAXB][6];
  for (int i = 0; i < m->nb; i++) {
    double cw[3], vl[3], t[3];
    mv3(k.oR[i], m->com + 3 * i, cw);
    for (int c = 0; c < 3; c++) cw[c] += k.op[i][c];
    /* world-origin spatial velocity of body i */
    mv3(k.oR[i], k.v[i] + 3, ov[i] + 3);
    mv3(k.oR[i], k.v[i], vl);
    cross3(k.op[i], ov[i] + 3, t);
    for (int c = 0; c < 3; c++) ov[i][c] = vl[c] + t[c];
    cross3(ov[i] + 3, cw, t);
    ms[i] = m->mass[i];
    for (int c = 0; c < 3; c++) { h[i][c] = m->mass[i] * cw[c]; hd[i][c] = m->mass[i] * (ov[i][c] + t[c]); }
  }
  for (int i = m->nb - 1; i > 0; i--) {
    int par = m->parent[i];
    if (par < 0) continue;
    ms[par] += ms[i];
    for (int c = 0; c < 3; c++) { h[par][c] += h[i][c]; hd[par][c] += hd[i][c]; }
  }
  memset(dJ, 0, sizeof(double) * 3 * nv);
  for (int i = 0; i < m->nb; i++) {
    int nd = m->jtype[i] == PNC_JOINT_FREEFLYER ? 6 : 1;
    for (int d = 0; d < nd; d++) {
      int col = m->idx_v[i] + d;
      double Jc[6] = {Jw[col], Jw[nv + col], Jw[2 * nv + col], Jw[3 * nv + col], Jw[4 * nv + col], Jw[5 * nv + col]};
      double dJc[6], t1[3], t2[3];
      mot_cross(ov[i], Jc, dJc);
      /* lin( dY J + Y dJ ) = -hdot x w_J + m dJ_lin - h x dJ_ang */
      cross3(hd[i], Jc + 3, t1);
      cross3(h[i], dJc + 3, t2);
      for (int c = 0; c < 3; c++) dJ[c * nv + col] = (-t1[c] + ms[i] * dJc[c] - t2[c]) / m->total_mass;
    }
  }
}

static void frame_world(const orc_model* m, const kin_t* k, int frame, double* R, double* p, int* body) {
  int b = m->frame_parent[frame];
  *body = b;
  if (b < 0) {
    memcpy(R, m->frame_R + 9 * frame, sizeof(double) * 9);
    memcpy(p, m->frame_p + 3 * frame, sizeof(double) * 3);
  } else {
    double t[3];
    mm3(k->oR[b], m->frame_R + 9 * frame, R);
    mv3(k->oR[b], m->frame_p + 3 * frame, t);
    for (int c = 0; c < 3; c++) p[c] = k->op[b][c] + t[c];
  }
}

void orc_link_iso(const orc_model* m, const double* q, int frame, double* T) {
  kin_t k; double R[9], p[3]; int b;
  fk(m, q, NULL, &k, 0);
  frame_world(m, &k, frame, R, p, &b);
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T[4 * i + j] = R[3 * i + j]; T[4 * i + 3] = p[i]; }
  T[12] = T[13] = T[14] = 0; T[15] = 1;
}

/* getFrameVelocity(LOCAL_WORLD_ALIGNED) reordered [ang; lin], :239-250 */
void orc_link_vel(const orc_model* m, const double* q, const double* v, int frame, double* out) {
  kin_t k; double R[9], p[3]; int b;
  fk(m, q, v, &k, 1);
  frame_world(m, &k, frame, R, p, &b);
  if (b < 0) { memset(out, 0, sizeof(double) * 6); return; }
  double vf[6];
  mot_actinv(m->frame_R + 9 * frame, m->frame_p + 3 * frame, k.v[b], vf);
  mv3(R, vf + 3, out);
  mv3(R, vf, out + 3);
}

/* getFrameJacobian(LOCAL_WORLD_ALIGNED) with rows swapped to [ang; lin], :252-263 */
void orc_link_jacobian(const orc_model* m, const double* q, int frame, double* J) {
  kin_t k; double R[9], p[3]; int b;
  fk(m, q, NULL, &k, 0);
  frame_world(m, &k, frame, R, p, &b);
  int nv = m->nv;
  memset(J, 0, sizeof(double) * 6 * nv);
  if (b < 0) return;
  double Jw[6 * MAXV];
  world_jac(m, &k, Jw);
  for (int j = 0; j <= b; j++) {
    if (!is_ancestor_or_self(m, j, b)) continue;
    int nd = m->jtype[j] == PNC_JOINT_FREEFLYER ? 6 : 1;
    for (int d = 0; d < nd; d++) {
      int col = m->idx_v[j] + d;
      double lin[3] = {Jw[col], Jw[nv + col], Jw[2 * nv + col]};
      double ang[3] = {Jw[3 * nv + col], Jw[4 * nv + col], Jw[5 * nv + col]}, wp[3];
      cross3(ang, p, wp);
      for (int c = 0; c < 3; c++) { J[c * nv + col] = ang[c]; J[(3 + c) * nv + col] = lin[c] + wp[c]; }
    }
  }
}

/* forwardKinematics(q,v,0) + getFrameClassicalAcceleration(LWA) -> [ang; lin], :265-278 */
void orc_link_jdqd(const orc_model* m, const double* q, const double* v, int frame, double* out) {
  kin_t k; double R[9], p[3]; int b;
  fk(m, q, v, &k, 1);
  frame_world(m, &k, frame, R, p, &b);
  if (b < 0) { memset(out, 0, sizeof(double) * 6); return; }
  double vf[6], af[6], wv[3];
  mot_actinv(m->frame_R + 9 * frame, m->frame_p + 3 * frame, k.v[b], vf);
  mot_actinv(m->frame_R + 9 * frame, m->frame_p + 3 * frame, k.a[b], af);
  cross3(vf + 3, vf, wv);
  for (int c = 0; c < 3; c++) af[c] += wv[c];
  mv3(R, af + 3, out);
  mv3(R, af, out + 3);
}

/* ccrba: Ag, hg, Ig re-ordered as the wrapper does (:181-194): angular rows first,
 * Ig keeps only the two diagonal 3x3 blocks. */
void orc_centroidal(const orc_model* m, const double* q, const double* v, double* Ag, double* hg,
                    double* Ig) {
  kin_t k;
  fk(m, q, v, &k, 1);
  int nv = m->nv;
  double Jw[6 * MAXV];
  world_jac(m, &k, Jw);
  /* world composite inertia per subtree about the world origin */
  cinertia Y[MAXB];
  double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z3[3] = {0, 0, 0};
  (void)I3; (void)z3;
  for (int i = 0; i < m->nb; i++) {
    cinertia B;
    cin_from_body(m->mass[i], m->com + 3 * i, m->inertia + 9 * i, &B);
    memset(&Y[i], 0, sizeof(cinertia));
    cin_add_transformed(&Y[i], &B, k.oR[i], k.op[i]);
  }
  for (int i = m->nb - 1; i > 0; i--) {
    int par = m->parent[i];
    if (par < 0) continue;
    Y[par].m += Y[i].m;
    for (int c = 0; c < 3; c++) Y[par].h[c] += Y[i].h[c];
    for (int c = 0; c < 9; c++) Y[par].Io[c] += Y[i].Io[c];
  }
  /* total (sum over roots) */
  cinertia T;
  memset(&T, 0, sizeof(T));
  for (int i = 0; i < m->nb; i++)
    if (m->parent[i] < 0) { T.m += Y[i].m; for (int c = 0; c < 3; c++) T.h[c] += Y[i].h[c]; for (int c = 0; c < 9; c++) T.Io[c] += Y[i].Io[c]; }
  double com[3] = {T.h[0] / T.m, T.h[1] / T.m, T.h[2] / T.m};
  double A[6 * MAXV];
  for (int i = 0; i < m->nb; i++) {
    int nd = m->jtype[i] == PNC_JOINT_FREEFLYER ? 6 : 1;
    for (int d = 0; d < nd; d++) {
      int col = m->idx_v[i] + d;
      double Jc[6] = {Jw[col], Jw[nv + col], Jw[2 * nv + col], Jw[3 * nv + col], Jw[4 * nv + col], Jw[5 * nv + col]};
      double f[6], t[3];
      cin_mul(&Y[i], Jc, f);
      cross3(com, f, t); /* move the moment from the origin to the CoM: n_c = n_o - com x f */
      for (int c = 0; c < 3; c++) { A[c * nv + col] = f[c]; A[(3 + c) * nv + col] = f[3 + c] - t[c]; }
    }
  }
  for (int c = 0; c < 3; c++)
    for (int j = 0; j < nv; j++) { Ag[c * nv + j] = A[(3 + c) * nv + j]; Ag[(3 + c) * nv + j] = A[c * nv + j]; }
  for (int r = 0; r < 6; r++) { double s = 0; for (int j = 0; j < nv; j++) s += Ag[r * nv + j] * v[j]; hg[r] = s; }
  memset(Ig, 0, sizeof(double) * 36);
  double cc = dot3(com, com);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Ig[6 * i + j] = T.Io[3 * i + j] - T.m * ((i == j ? cc : 0) - com[i] * com[j]);
      Ig[6 * (3 + i) + 3 + j] = (i == j) ? T.m : 0;
    }
}

/* ================================================================= dense helpers */
static void mat_mul(const double* A, const double* B, double* C, int m, int k, int n) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double s = 0;
      for (int l = 0; l < k; l++) s += A[i * k + l] * B[l * n + j];
      C[i * n + j] = s;
    }
}
static void mat_tr(const double* A, double* At, int m, int n) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) At[j * m + i] = A[i * n + j];
}

/* one-sided Jacobi SVD of A (m x n, m >= n): A = U diag(s) V^T */
static void svd_jacobi(const double* A, int m, int n, double* U, double* s, double* V) {
  memcpy(U, A, sizeof(double) * m * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < m; i++) { alpha += U[i * n + p] * U[i * n + p]; beta += U[i * n + q] * U[i * n + q]; gamma += U[i * n + p] * U[i * n + q]; }
        if (gamma == 0 || fabs(gamma) <= 1e-300) continue;
        double lim = sqrt(alpha * beta);
        if (fabs(gamma) <= DBL_EPSILON * 1e-2 * lim) continue;
        off = fmax(off, fabs(gamma) / (lim > 0 ? lim : 1));
        double zeta = (beta - alpha) / (2 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
        double c = 1 / sqrt(1 + t * t), sn = c * t;
        for (int i = 0; i < m; i++) { double up = U[i * n + p], uq = U[i * n + q]; U[i * n + p] = c * up - sn * uq; U[i * n + q] = sn * up + c * uq; }
        for (int i = 0; i < n; i++) { double vp = V[i * n + p], vq = V[i * n + q]; V[i * n + p] = c * vp - sn * vq; V[i * n + q] = sn * vp + c * vq; }
      }
    if (off < 1e-15) break;
  }
  for (int j = 0; j < n; j++) {
    double nr = 0;
    for (int i = 0; i < m; i++) nr += U[i * n + j] * U[i * n + j];
    s[j] = sqrt(nr);
    if (s[j] > 0) for (int i = 0; i < m; i++) U[i * n + j] /= s[j];
  }
}

/* np.linalg.pinv(A, rcond): singular values <= rcond * max are dropped */
void orc_pinv(const double* A, int m, int n, double rcond, double* P) {
  if (m < n) { /* pinv(A) = pinv(A^T)^T */
    double* At = (double*)malloc(sizeof(double) * m * n);
    double* Pt = (double*)malloc(sizeof(double) * m * n);
    mat_tr(A, At, m, n);
    orc_pinv(At, n, m, rcond, Pt);
    mat_tr(Pt, P, m, n);
    free(At); free(Pt);
    return;
  }
  double* U = (double*)malloc(sizeof(double) * m * n);
  double* V = (double*)malloc(sizeof(double) * n * n);
  double* s = (double*)malloc(sizeof(double) * n);
  svd_jacobi(A, m, n, U, s, V);
  double smax = 0;
  for (int j = 0; j < n; j++) smax = fmax(smax, s[j]);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) {
      double acc = 0;
      for (int l = 0; l < n; l++)
        if (s[l] > rcond * smax) acc += V[i * n + l] * U[j * n + l] / s[l];
      P[i * m + j] = acc;
    }
  free(U); free(V); free(s);
}

/* ================================================================= QuadProg++ restated */
static double qp_distance(double a, double b) { /* QuadProg++.cc:681-692 */
  double a1 = fabs(a), b1 = fabs(b), t;
  if (a1 > b1) { t = b1 / a1; return a1 * sqrt(1.0 + t * t); }
  else if (b1 > a1) { t = a1 / b1; return b1 * sqrt(1.0 + t * t); }
  return a1 * sqrt(2.0);
}
static double qp_dot(const double* x, const double* y, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}
static int qp_cholesky(double* A, int n) { /* :705-730 */
  for (int i = 0; i < n; i++) {
    for (int j = i; j < n; j++) {
      double sum = A[i * n + j];
      for (int k = i - 1; k >= 0; k--) sum -= A[i * n + k] * A[j * n + k];
      if (i == j) {
        if (sum <= 0.0) return -1;
        A[i * n + i] = sqrt(sum);
      } else
        A[j * n + i] = sum / A[i * n + i];
    }
    for (int k = i + 1; k < n; k++) A[i * n + k] = A[k * n + i];
  }
  return 0;
}
static void qp_forward(const double* L, double* y, const double* b, int n) { /* :743-754 */
  y[0] = b[0] / L[0];
  for (int i = 1; i < n; i++) {
    y[i] = b[i];
    for (int j = 0; j < i; j++) y[i] -= L[i * n + j] * y[j];
    y[i] = y[i] / L[i * n + i];
  }
}
static void qp_backward(const double* U, double* x, const double* y, int n) { /* :756-767 */
  x[n - 1] = y[n - 1] / U[(n - 1) * n + n - 1];
  for (int i = n - 2; i >= 0; i--) {
    x[i] = y[i];
    for (int j = i + 1; j < n; j++) x[i] -= U[i * n + j] * x[j];
    x[i] = x[i] / U[i * n + i];
  }
}
static void qp_compute_d(double* d, const double* J, const double* np, int n) { /* :504-516 */
  for (int i = 0; i < n; i++) {
    double sum = 0.0;
    for (int j = 0; j < n; j++) sum += J[j * n + i] * np[j];
    d[i] = sum;
  }
}
static void qp_update_z(double* z, const double* J, const double* d, int iq, int n) { /* :518-528 */
  for (int i = 0; i < n; i++) {
    z[i] = 0.0;
    for (int j = iq; j < n; j++) z[i] += J[i * n + j] * d[j];
  }
}
static void qp_update_r(const double* R, double* r, const double* d, int iq, int n) { /* :530-542 */
  for (int i = iq - 1; i >= 0; i--) {
    double sum = 0.0;
    for (int j = i + 1; j < iq; j++) sum += R[i * n + j] * r[j];
    r[i] = (d[i] - sum) / R[i * n + i];
  }
}
static int qp_add_constraint(double* R, double* J, double* d, int* iq, double* R_norm, int n) { /* :544-608 */
  double cc, ss, h, t1, t2, xny;
  for (int j = n - 1; j >= *iq + 1; j--) {
    cc = d[j - 1];
    ss = d[j];
    h = qp_distance(cc, ss);
    if (fabs(h) < DBL_EPSILON) continue;
    d[j] = 0.0;
    ss = ss / h;
    cc = cc / h;
    if (cc < 0.0) { cc = -cc; ss = -ss; d[j - 1] = -h; }
    else d[j - 1] = h;
    xny = ss / (1.0 + cc);
    for (int k = 0; k < n; k++) {
      t1 = J[k * n + j - 1];
      t2 = J[k * n + j];
      J[k * n + j - 1] = t1 * cc + t2 * ss;
      J[k * n + j] = xny * (t1 + J[k * n + j - 1]) - t2;
    }
  }
  (*iq)++;
  for (int i = 0; i < *iq; i++) R[i * n + *iq - 1] = d[i];
  if (fabs(d[*iq - 1]) <= DBL_EPSILON * *R_norm) return 0;
  *R_norm = fmax(*R_norm, fabs(d[*iq - 1]));
  return 1;
}
static void qp_delete_constraint(double* R, double* J, int* A, double* u, int n, int p, int* iq, int l) { /* :610-679 */
  int qq = -1;
  double cc, ss, h, xny, t1, t2;
  for (int i = p; i < *iq; i++)
    if (A[i] == l) { qq = i; break; }
  for (int i = qq; i < *iq - 1; i++) {
    A[i] = A[i + 1];
    u[i] = u[i + 1];
    for (int j = 0; j < n; j++) R[j * n + i] = R[j * n + i + 1];
  }
  A[*iq - 1] = A[*iq];
  u[*iq - 1] = u[*iq];
  A[*iq] = 0;
  u[*iq] = 0.0;
  for (int j = 0; j < *iq; j++) R[j * n + *iq - 1] = 0.0;
  (*iq)--;
  if (*iq == 0) return;
  for (int j = qq; j < *iq; j++) {
    cc = R[j * n + j];
    ss = R[(j + 1) * n + j];
    h = qp_distance(cc, ss);
    if (fabs(h) < DBL_EPSILON) continue;
    cc = cc / h;
    ss = ss / h;
    R[(j + 1) * n + j] = 0.0;
    if (cc < 0.0) { R[j * n + j] = -h; cc = -cc; ss = -ss; }
    else R[j * n + j] = h;
    xny = ss / (1.0 + cc);
    for (int k = j + 1; k < *iq; k++) {
      t1 = R[j * n + k];
      t2 = R[(j + 1) * n + k];
      R[j * n + k] = t1 * cc + t2 * ss;
      R[(j + 1) * n + k] = xny * (t1 + R[j * n + k]) - t2;
    }
    for (int k = 0; k < n; k++) {
      t1 = J[k * n + j];
      t2 = J[k * n + j + 1];
      J[k * n + j] = t1 * cc + t2 * ss;
      J[k * n + j + 1] = xny * (J[k * n + j] + t1) - t2;
    }
  }
}

/* Decision margin of the last orc_solve_quadprog call on this thread (test instrumentation, does
 * not touch the arithmetic): the smallest relative distance, over every data-dependent branch the
 * solver took, between the two sides of that branch.  A state whose margin is far above rounding
 * has one well-defined iteration path and final active set; a margin at rounding level means the
 * reference itself picks among exact ties by noise (redundant wrench-cone rows, degenerate vertices).
 * kind: 1 termination test (:306-312), 2 most-violated-constraint choice (:323-334), 3 blocking
 * constraint choice (:366-378), 4 partial vs full step (:390-393, :441), 5 near-dependent normal
 * (|d[iq:]| / |d|: decides :380 and the degeneracy test of add_constraint :597-603). */
static __thread double g_qp_margin;
static __thread int g_qp_margin_kind;
static __thread double g_qp_psi_rel; /* |psi| / max|s| at the iterate the solver returned: 0 = exactly feasible;
                                        > 0 = it stopped inside its loose bound (:306-312) with a residual violation */
#define QP_MARGIN(val, kind)                                    \
  do {                                                          \
    const double v_ = (val);                                    \
    if (v_ < g_qp_margin) { g_qp_margin = v_; g_qp_margin_kind = (kind); } \
  } while (0)
void orc_last_qp_margin(double* margin, int* kind) {
  if (margin) *margin = g_qp_margin;
  if (kind) *kind = g_qp_margin_kind;
}
double orc_last_qp_psi_rel(void) { return g_qp_psi_rel; }

double orc_solve_quadprog(int n, int p, int m, double* G, const double* g0, const double* CE,
                          const double* ce0, const double* CI, const double* ci0, double* x, int* A,
                          double* u, int* iq_out, int* iters, int* status, double* flops) {
  int i, j, k, l = 0, ip, iq, iter = 0;
  g_qp_margin = 1.0;
  g_qp_margin_kind = 0;
  g_qp_psi_rel = 0.0;
  double s_scale = 1.0;
  int mp = m + p;
  double* R = (double*)calloc((size_t)n * n, sizeof(double));
  double* J = (double*)calloc((size_t)n * n, sizeof(double));
  double* w = (double*)calloc((size_t)6 * n + 4 * (mp + 1), sizeof(double));
  double *s = w, *z = s + mp + 1, *r = z + n, *d = r + mp + 1, *np = d + n, *x_old = np + n,
         *u_old = x_old + n, *yy = u_old + mp + 1;
  int* iw = (int*)calloc((size_t)3 * (mp + 1), sizeof(int));
  int *A_old = iw, *iai = A_old + mp + 1, *iaexcl = iai + mp + 1;
  double f_value, psi, c1, c2, sum, ss, R_norm, t, t1, t2;
  const double inf = INFINITY;
  double F = 0;
  *status = PNC_QP_OK;
  for (i = 0; i < mp; i++) { A[i] = 0; u[i] = 0; }

  c1 = 0.0;
  for (i = 0; i < n; i++) c1 += G[i * n + i];
  if (qp_cholesky(G, n) != 0) { *status = PNC_QP_NOT_PD; f_value = inf; iq = 0; goto done; }
  F += (double)n * n * n / 3.0;
  for (i = 0; i < n; i++) d[i] = 0.0;
  R_norm = 1.0;
  c2 = 0.0;
  for (i = 0; i < n; i++) {
    d[i] = 1.0;
    qp_forward(G, z, d, n);
    for (j = 0; j < n; j++) J[i * n + j] = z[j];
    c2 += z[i];
    d[i] = 0.0;
  }
  F += (double)n * n * n / 3.0;
  /* x = -G^-1 g0 */
  qp_forward(G, yy, g0, n);
  qp_backward(G, x, yy, n);
  for (i = 0; i < n; i++) x[i] = -x[i];
  F += 2.0 * n * n;
  f_value = 0.5 * qp_dot(g0, x, n);

  iq = 0;
  for (i = 0; i < p; i++) {
    for (j = 0; j < n; j++) np[j] = CE[j * p + i];
    qp_compute_d(d, J, np, n);
    qp_update_z(z, J, d, iq, n);
    qp_update_r(R, r, d, iq, n);
    t2 = 0.0;
    if (fabs(qp_dot(z, z, n)) > DBL_EPSILON) t2 = (-qp_dot(np, x, n) - ce0[i]) / qp_dot(z, np, n);
    for (k = 0; k < n; k++) x[k] += t2 * z[k];
    u[iq] = t2;
    for (k = 0; k < iq; k++) u[k] -= t2 * r[k];
    f_value += 0.5 * (t2 * t2) * qp_dot(z, np, n);
    A[i] = -i - 1;
    F += 10.0 * n * n;
    if (!qp_add_constraint(R, J, d, &iq, &R_norm, n)) { *status = PNC_QP_EQ_DEPENDENT; f_value = inf; goto done; }
  }
  for (i = 0; i < m; i++) iai[i] = i;

l1:
  iter++;
  for (i = p; i < iq; i++) { ip = A[i]; iai[ip] = -1; }
  ss = 0.0;
  psi = 0.0;
  ip = 0;
  for (i = 0; i < m; i++) {
    iaexcl[i] = 1;
    sum = 0.0;
    for (j = 0; j < n; j++) sum += CI[j * m + i] * x[j];
    sum += ci0[i];
    s[i] = sum;
    psi += fmin(0.0, sum);
  }
  F += 2.0 * m * n;
  {
    const double tol = m * DBL_EPSILON * c1 * c2 * 100.0;
    s_scale = 0.0;
    for (i = 0; i < m; i++) s_scale = fmax(s_scale, fabs(s[i]));
    if (!(s_scale > 0.0)) s_scale = 1.0;
    if (tol > 0.0) QP_MARGIN(fabs(fabs(psi) - tol) / tol, 1);
    g_qp_psi_rel = fabs(psi) / s_scale;
  }
  if (fabs(psi) <= m * DBL_EPSILON * c1 * c2 * 100.0) goto done;
  if (iter > 40 * (n + m)) { *status = PNC_QP_MAX_ITER; goto done; }
  for (i = 0; i < iq; i++) { u_old[i] = u[i]; A_old[i] = A[i]; }
  for (i = 0; i < n; i++) x_old[i] = x[i];

l2:
  for (i = 0; i < m; i++) {
    if (s[i] < ss && iai[i] != -1 && iaexcl[i]) { ss = s[i]; ip = i; }
  }
  {
    /* runner-up of the choice just made; when nothing is violated, the distance of the smallest
     * eligible slack from zero */
    double s2 = INFINITY;
    for (i = 0; i < m; i++)
      if (!(ss < 0.0 && i == ip) && iai[i] != -1 && iaexcl[i] && s[i] < s2) s2 = s[i];
    if (ss < 0.0) QP_MARGIN((fmin(s2, 0.0) - ss) / s_scale, 2);
    else if (s2 < INFINITY) QP_MARGIN(fabs(s2) / s_scale, 2);
  }
  if (ss >= 0.0) goto done;
  for (i = 0; i < n; i++) np[i] = CI[i * m + ip];
  u[iq] = 0.0;
  A[iq] = ip;

l2a:
  qp_compute_d(d, J, np, n);
  qp_update_z(z, J, d, iq, n);
  qp_update_r(R, r, d, iq, n);
  F += 2.0 * n * n + 2.0 * n * (n - iq) + (double)iq * iq;
  l = 0;
  t1 = inf;
  for (k = p; k < iq; k++) {
    if (r[k] > 0.0) {
      if (u[k] / r[k] < t1) { t1 = u[k] / r[k]; l = A[k]; }
    }
  }
  if (fabs(qp_dot(z, z, n)) > DBL_EPSILON) {
    t2 = -s[ip] / qp_dot(z, np, n);
    if (t2 < 0) t2 = inf;
  } else
    t2 = inf;
  t = fmin(t1, t2);
  {
    double t1b = inf, dd = 0.0, d2 = 0.0;
    for (k = p; k < iq; k++)
      if (r[k] > 0.0 && A[k] != l && u[k] / r[k] < t1b) t1b = u[k] / r[k];
    if (t1 < inf && t1b < inf) QP_MARGIN(fabs(t1b - t1) / fmax(fabs(t1b), 1e-300), 3);
    if (t1 < inf && t2 < inf) QP_MARGIN(fabs(t1 - t2) / fmax(fmax(fabs(t1), fabs(t2)), 1e-300), 4);
    for (k = 0; k < n; k++) { dd += d[k] * d[k]; if (k >= iq) d2 += d[k] * d[k]; }
    if (dd > 0.0) QP_MARGIN(sqrt(d2 / dd), 5);
  }
  if (t >= inf) { *status = PNC_QP_INFEASIBLE; f_value = inf; goto done; }
  if (t2 >= inf) {
    for (k = 0; k < iq; k++) u[k] -= t * r[k];
    u[iq] += t;
    iai[l] = l;
    F += 6.0 * n * iq;
    qp_delete_constraint(R, J, A, u, n, p, &iq, l);
    goto l2a;
  }
  for (k = 0; k < n; k++) x[k] += t * z[k];
  f_value += t * qp_dot(z, np, n) * (0.5 * t + u[iq]);
  for (k = 0; k < iq; k++) u[k] -= t * r[k];
  u[iq] += t;
  if (fabs(t - t2) < DBL_EPSILON) {
    F += 6.0 * n * (n - iq);
    if (!qp_add_constraint(R, J, d, &iq, &R_norm, n)) {
      iaexcl[ip] = 0;
      qp_delete_constraint(R, J, A, u, n, p, &iq, ip);
      for (i = 0; i < m; i++) iai[i] = i;
      for (i = p; i < iq; i++) { A[i] = A_old[i]; u[i] = u_old[i]; iai[A[i]] = -1; }
      for (i = 0; i < n; i++) x[i] = x_old[i];
      goto l2;
    } else
      iai[ip] = -1;
    goto l1;
  }
  iai[l] = l;
  F += 6.0 * n * iq;
  qp_delete_constraint(R, J, A, u, n, p, &iq, l);
  sum = 0.0;
  for (k = 0; k < n; k++) sum += CI[k * m + ip] * x[k];
  s[ip] = sum + ci0[ip];
  goto l2a;

done:
  *iq_out = iq;
  *iters = iter;
  if (flops) *flops = F;
  free(R); free(J); free(w); free(iw);
  return f_value;
}

/* ================================================================= WBC problem */
struct orc_problem {
  const orc_model* m;
  int ntask, ncontact, n_joint_sel, n_gain, n_ic, nact, b_trq_limit, ndes;
  int n_trans, trans_d[8], trans_p[8]; /* IHWBC2 transmission rows (ihwbc2.py), v indices of the distal / passive joints */
  pnc_task_desc* tasks;
  pnc_contact_desc* contacts;
  int *joint_sel, *act_idx;
  double *kp, *kd, *ic_jac, *trq_limit;
  double lambda_q_ddot, lambda_rf;
  int nc, ncone, T;
};

orc_problem* orc_problem_create(const orc_model* m, const pnc_problem_desc* d) {
  orc_problem* p = (orc_problem*)calloc(1, sizeof(orc_problem));
  p->m = m;
  p->ntask = d->ntask; p->ncontact = d->ncontact; p->n_joint_sel = d->n_joint_sel; p->n_gain = d->n_gain;
  p->n_ic = d->n_ic; p->nact = d->nact; p->b_trq_limit = d->b_trq_limit; p->ndes = d->ndes;
  p->n_trans = d->n_trans;
  for (int i = 0; i < d->n_trans && i < 8; i++) { p->trans_d[i] = d->trans_distal[i]; p->trans_p[i] = d->trans_passive[i]; }
  p->lambda_q_ddot = d->lambda_q_ddot; p->lambda_rf = d->lambda_rf;
  p->tasks = (pnc_task_desc*)dupmem(d->tasks, sizeof(pnc_task_desc) * d->ntask);
  p->contacts = (pnc_contact_desc*)dupmem(d->contacts, sizeof(pnc_contact_desc) * d->ncontact);
  p->joint_sel = (int*)dupmem(d->joint_sel, sizeof(int) * d->n_joint_sel);
  p->act_idx = (int*)dupmem(d->act_idx, sizeof(int) * d->nact);
  p->kp = (double*)dupmem(d->kp, sizeof(double) * d->n_gain);
  p->kd = (double*)dupmem(d->kd, sizeof(double) * d->n_gain);
  p->ic_jac = (double*)dupmem(d->ic_jac, sizeof(double) * d->n_ic * m->nv);
  p->trq_limit = (double*)dupmem(d->trq_limit, sizeof(double) * 2 * d->nact * (d->b_trq_limit ? 1 : 0));
  for (int i = 0; i < p->ncontact; i++) {
    p->nc += p->contacts[i].type == PNC_CONTACT_SURFACE ? 6 : 3;
    p->ncone += p->contacts[i].type == PNC_CONTACT_SURFACE ? 18 : 6;
  }
  for (int i = 0; i < p->ntask; i++) p->T += p->tasks[i].dim;
  return p;
}
void orc_problem_destroy(orc_problem* p) {
  if (!p) return;
  free(p->tasks); free(p->contacts); free(p->joint_sel); free(p->act_idx); free(p->kp); free(p->kd);
  free(p->ic_jac); free(p->trq_limit); free(p);
}
void orc_problem_dims(const orc_problem* p, int* nc, int* ncone, int* n, int* peq, int* mineq) {
  *nc = p->nc; *ncone = p->ncone; *n = p->m->nv + p->nc; *peq = 6 + p->n_ic + p->n_trans;
  *mineq = p->ncone + (p->b_trq_limit ? 2 * p->nact : 0);
}

/* basic_task.py:25-137 */
void orc_task_eval(const orc_problem* p, const double* q, const double* v, const double* des, int it,
                   double* jac, double* jdqd, double* op_cmd, double* pos_err) {
  const orc_model* m = p->m;
  const pnc_task_desc* t = &p->tasks[it];
  int nv = m->nv, dim = t->dim;
  double J6[6 * MAXV], a6[6], v6[6], T[16], perr[MAXV], vact[MAXV], jd[MAXV];
  double* Jt = (double*)calloc((size_t)dim * nv, sizeof(double));
  const double* pos_des = des + t->des_off;
  int npos = t->type == PNC_TASK_LINK_ORI ? 4 : dim;
  const double* vel_des = pos_des + npos;
  const double* acc_des = vel_des + dim;
  const double* jpos = q + (m->nfl ? 7 : 0);
  const double* jvel = v + m->nfl;
  switch (t->type) {
    case PNC_TASK_JOINT:
      for (int i = 0; i < dim; i++) { Jt[i * nv + m->nfl + i] = 1; jd[i] = 0; perr[i] = pos_des[i] - jpos[i]; vact[i] = jvel[i]; }
      break;
    case PNC_TASK_SELECTED_JOINT:
      for (int i = 0; i < dim; i++) {
        int vi = p->joint_sel[t->joint_off + i];
        Jt[i * nv + vi] = 1; jd[i] = 0;
        perr[i] = pos_des[i] - jpos[vi - m->nfl];
        vact[i] = jvel[vi - m->nfl];
      }
      break;
    case PNC_TASK_LINK_XYZ:
      orc_link_jacobian(m, q, t->frame, J6);
      orc_link_jdqd(m, q, v, t->frame, a6);
      orc_link_iso(m, q, t->frame, T);
      orc_link_vel(m, q, v, t->frame, v6);
      memcpy(Jt, J6 + 3 * nv, sizeof(double) * 3 * nv);
      for (int i = 0; i < 3; i++) { jd[i] = a6[3 + i]; perr[i] = pos_des[i] - T[4 * i + 3]; vact[i] = v6[3 + i]; }
      break;
    case PNC_TASK_LINK_ORI: {
      orc_link_jacobian(m, q, t->frame, J6);
      orc_link_jdqd(m, q, v, t->frame, a6);
      orc_link_iso(m, q, t->frame, T);
      orc_link_vel(m, q, v, t->frame, v6);
      memcpy(Jt, J6, sizeof(double) * 3 * nv);
      /* :62-78.  prevent_quat_jump only flips the sign of a quaternion that is turned
       * back into a rotation, so it does not change R_act. */
      double Rdes[9], Ract[9], Rerr[9], qerr[4], qa[4];
      orc_quat_to_rot(pos_des, Rdes);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ract[3 * i + j] = T[4 * i + j];
      orc_rot_to_quat(Ract, qa); /* R.from_matrix(R_link) -> as_quat -> from_quat -> as_matrix */
      orc_quat_to_rot(qa, Ract);
      mmt3(Rdes, Ract, Rerr);
      orc_rot_to_quat(Rerr, qerr);
      orc_quat_to_exp(qerr, perr);
      for (int i = 0; i < 3; i++) { jd[i] = a6[i]; vact[i] = v6[i]; }
    } break;
    case PNC_TASK_COM: {
      double com[3], vcom[3], dJ[3 * MAXV];
      orc_com_jacobian(m, q, Jt);
      orc_com_jacobian_dot(m, q, v, dJ);
      orc_com(m, q, v, com, vcom);
      for (int i = 0; i < 3; i++) {
        double s = 0;
        for (int j = 0; j < nv; j++) s += dJ[i * nv + j] * v[j];
        jd[i] = s; perr[i] = pos_des[i] - com[i]; vact[i] = vcom[i];
      }
    } break;
  }
  for (int i = 0; i < dim; i++) {
    if (op_cmd) op_cmd[i] = acc_des[i] + p->kp[t->gain_off + i] * perr[i] + p->kd[t->gain_off + i] * (vel_des[i] - vact[i]);
    if (pos_err) pos_err[i] = perr[i];
    if (jdqd) jdqd[i] = jd[i];
  }
  if (jac) memcpy(jac, Jt, sizeof(double) * dim * nv);
  free(Jt);
}

/* basic_contact.py:24-51 (point), :68-173 (surface) */
void orc_contact_eval(const orc_problem* p, const double* q, const double* v, double rf_z_max, int ic,
                      double* jac, double* jdqd, double* cone_mat, double* cone_vec) {
  const orc_model* m = p->m;
  const pnc_contact_desc* c = &p->contacts[ic];
  int nv = m->nv;
  double J6[6 * MAXV], a6[6] = {0}, T[16];
  if (rf_z_max <= 0.) rf_z_max = 1e-3; /* contact.py:37-41 */
  orc_link_jacobian(m, q, c->frame, J6);
  if (v) orc_link_jdqd(m, q, v, c->frame, a6);
  if (c->type == PNC_CONTACT_POINT) {
    if (jac) memcpy(jac, J6 + 3 * nv, sizeof(double) * 3 * nv);
    if (jdqd) for (int i = 0; i < 3; i++) jdqd[i] = a6[3 + i];
    double mu = c->mu;
    double U[18] = {0, 0, 1, 1, 0, mu, -1, 0, mu, 0, 1, mu, 0, -1, mu, 0, 0, -1};
    if (cone_mat) memcpy(cone_mat, U, sizeof(U));
    if (cone_vec) { memset(cone_vec, 0, sizeof(double) * 6); cone_vec[5] = -rf_z_max; }
  } else {
    if (jac) memcpy(jac, J6, sizeof(double) * 6 * nv);
    if (jdqd) memcpy(jdqd, a6, sizeof(a6));
    double mu = c->mu, x = c->x, y = c->y, U[18 * 6];
    memset(U, 0, sizeof(U));
#define UU(r, cc) U[(r) * 6 + (cc)]
    UU(0, 5) = 1.;
    UU(1, 3) = 1.; UU(1, 5) = mu; UU(2, 3) = -1.; UU(2, 5) = mu;
    UU(3, 4) = 1.; UU(3, 5) = mu; UU(4, 4) = -1.; UU(4, 5) = mu;
    UU(5, 0) = 1.; UU(5, 5) = y; UU(6, 0) = -1.; UU(6, 5) = y;
    UU(7, 1) = 1.; UU(7, 5) = x; UU(8, 1) = -1.; UU(8, 5) = x;
    static const int sg[8][5] = {/* signs of (mu, mu, 1, y, x) columns 0..4, rows 9..16 */
        {-1, -1, 1, 1, 1}, {-1, 1, 1, 1, -1}, {1, -1, 1, -1, 1}, {1, 1, 1, -1, -1},
        {-1, -1, -1, -1, -1}, {-1, 1, -1, -1, 1}, {1, -1, -1, 1, -1}, {1, 1, -1, 1, 1}};
    for (int r = 0; r < 8; r++) {
      UU(9 + r, 0) = sg[r][0] * mu; UU(9 + r, 1) = sg[r][1] * mu; UU(9 + r, 2) = sg[r][2] * 1.;
      UU(9 + r, 3) = sg[r][3] * y; UU(9 + r, 4) = sg[r][4] * x; UU(9 + r, 5) = (x + y) * mu;
    }
    UU(17, 5) = -1.;
#undef UU
    orc_link_iso(m, q, c->frame, T);
    double RF[36];
    memset(RF, 0, sizeof(RF));
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) { RF[6 * i + j] = T[4 * j + i]; RF[6 * (3 + i) + 3 + j] = T[4 * j + i]; }
    if (cone_mat) mat_mul(U, RF, cone_mat, 18, 6, 6);
    if (cone_vec) { memset(cone_vec, 0, sizeof(double) * 18); cone_vec[17] = -rf_z_max; }
  }
}

/* work buffers of one assembled IHWBC problem */
typedef struct {
  int nv, nc, ncone, n, peq, mineq, nact, npass;
  double *M, *Minv, *cg, *ni, *jit, *B, *Jc, *JcNi, *Uf, *uf, *P, *qv, *Gm, *h, *Am, *b, *S, *SNtg;
} asm_t;

static void asm_alloc(const orc_problem* p, asm_t* a) {
  const orc_model* m = p->m;
  int nv = m->nv;
  a->nv = nv; a->nc = p->nc; a->ncone = p->ncone; a->n = nv + p->nc; a->peq = 6 + p->n_ic + p->n_trans;
  a->mineq = p->ncone + (p->b_trq_limit ? 2 * p->nact : 0); a->nact = p->nact; a->npass = nv - 6 - p->nact;
  size_t n = a->n, mi = a->mineq > 0 ? a->mineq : 1;
  a->M = (double*)calloc((size_t)nv * nv, 8); a->Minv = (double*)calloc((size_t)nv * nv, 8);
  a->cg = (double*)calloc(nv, 8); a->ni = (double*)calloc((size_t)nv * nv, 8); a->jit = (double*)calloc(nv, 8);
  a->B = (double*)calloc((size_t)p->nact * (nv - 6), 8);
  a->Jc = (double*)calloc((size_t)(p->nc ? p->nc : 1) * nv, 8); a->JcNi = (double*)calloc((size_t)(p->nc ? p->nc : 1) * nv, 8);
  a->Uf = (double*)calloc((size_t)(p->ncone ? p->ncone : 1) * (p->nc ? p->nc : 1), 8); a->uf = (double*)calloc(p->ncone ? p->ncone : 1, 8);
  a->P = (double*)calloc(n * n, 8); a->qv = (double*)calloc(n, 8); a->Gm = (double*)calloc(mi * n, 8);
  a->h = (double*)calloc(mi, 8); a->Am = (double*)calloc((size_t)a->peq * n, 8); a->b = (double*)calloc(a->peq, 8);
  a->S = (double*)calloc((size_t)p->nact * nv, 8); a->SNtg = (double*)calloc(p->nact, 8);
}
static void asm_free(asm_t* a) {
  free(a->M); free(a->Minv); free(a->cg); free(a->ni); free(a->jit); free(a->B); free(a->Jc); free(a->JcNi);
  free(a->Uf); free(a->uf); free(a->P); free(a->qv); free(a->Gm); free(a->h); free(a->Am); free(a->b);
  free(a->S); free(a->SNtg);
}

static void mat_inv_spd_or_general(const double* A, double* Ainv, int n) { /* np.linalg.inv: LU w/ partial pivoting */
  double* a = (double*)malloc(sizeof(double) * n * n);
  memcpy(a, A, sizeof(double) * n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ainv[i * n + j] = (i == j);
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (fabs(a[r * n + c]) > fabs(a[piv * n + c])) piv = r;
    if (piv != c) for (int j = 0; j < n; j++) {
      double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t;
      t = Ainv[c * n + j]; Ainv[c * n + j] = Ainv[piv * n + j]; Ainv[piv * n + j] = t;
    }
    double d = a[c * n + c];
    for (int j = 0; j < n; j++) { a[c * n + j] /= d; Ainv[c * n + j] /= d; }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      double f = a[r * n + c];
      if (f == 0) continue;
      for (int j = 0; j < n; j++) { a[r * n + j] -= f * a[c * n + j]; Ainv[r * n + j] -= f * Ainv[c * n + j]; }
    }
  }
  free(a);
}

/* ihwbc.py:124-328 line by line (dense, as the reference builds it) */
static void assemble(const orc_problem* p, const double* q, const double* v, const double* des,
                     const double* w, const double* rf_z_max, asm_t* a, double* flops) {
  const orc_model* m = p->m;
  int nv = a->nv, nc = a->nc, n = a->n, nact = a->nact, nu = a->ncone;
  double F = 0;
  /* atlas_controller.py:57-62 */
  orc_mass_matrix(m, q, a->M);
  double cor[MAXV], grav[MAXV];
  orc_coriolis(m, q, v, cor);
  orc_gravity(m, q, grav);
  for (int i = 0; i < nv; i++) a->cg[i] = cor[i] + grav[i];

  /* ---- internal constraints, :124-146 */
  int k = p->n_ic;
  int na_all = nv - 6; /* n_active + n_passive: columns of snf */
  if (k > 0) {
    mat_inv_spd_or_general(a->M, a->Minv, nv); /* np.linalg.inv, atlas_controller.py:58 */
    double* JiMinv = (double*)calloc((size_t)k * nv, 8);
    double* JiT = (double*)calloc((size_t)k * nv, 8);
    double* L0 = (double*)calloc((size_t)k * k, 8);
    double* lmd = (double*)calloc((size_t)k * k, 8);
    double* MinvJiT = (double*)calloc((size_t)nv * k, 8);
    double* jibar = (double*)calloc((size_t)nv * k, 8);
    mat_mul(p->ic_jac, a->Minv, JiMinv, k, nv, nv);
    mat_tr(p->ic_jac, JiT, k, nv);
    mat_mul(JiMinv, JiT, L0, k, nv, k);
    orc_pinv(L0, k, k, 1e-15, lmd);
    mat_mul(a->Minv, JiT, MinvJiT, nv, nv, k);
    mat_mul(MinvJiT, lmd, jibar, nv, k, k);
    mat_mul(jibar, p->ic_jac, a->ni, nv, k, nv);
    for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) a->ni[i * nv + j] = (i == j) - a->ni[i * nv + j];
    /* jit_lmd_jidot_qdot = Ji^T lmd Jidot_qdot with Jidot_qdot = 0 (rolling joint, :16) */
    for (int i = 0; i < nv; i++) a->jit[i] = 0;
    /* sa_ni_trc = (Sa Ni)[:, 6:], weighted_pinv(., Minv[6:,6:]) */
    double* A = (double*)calloc((size_t)nact * na_all, 8);
    double* W = (double*)calloc((size_t)na_all * na_all, 8);
    for (int i = 0; i < nact; i++) for (int j = 0; j < na_all; j++) A[i * na_all + j] = a->ni[p->act_idx[i] * nv + 6 + j];
    for (int i = 0; i < na_all; i++) for (int j = 0; j < na_all; j++) W[i * na_all + j] = a->Minv[(6 + i) * nv + 6 + j];
    double* AW = (double*)calloc((size_t)nact * na_all, 8);
    double* At = (double*)calloc((size_t)nact * na_all, 8);
    double* AWAt = (double*)calloc((size_t)nact * nact, 8);
    double* pin = (double*)calloc((size_t)nact * nact, 8);
    double* WAt = (double*)calloc((size_t)na_all * nact, 8);
    double* Abar = (double*)calloc((size_t)na_all * nact, 8);
    mat_mul(A, W, AW, nact, na_all, na_all);
    mat_tr(A, At, nact, na_all);
    mat_mul(AW, At, AWAt, nact, na_all, nact);
    orc_pinv(AWAt, nact, nact, 1e-15, pin);
    mat_mul(W, At, WAt, na_all, na_all, nact);
    mat_mul(WAt, pin, Abar, na_all, nact, nact);
    mat_tr(Abar, a->B, na_all, nact); /* sa_ni_trc_bar_tr: nact x na_all */
    F += 2.0 * k * nv * nv + 2.0 * nv * nv * k + 2.0 * nv * nv * nv;
    free(JiMinv); free(JiT); free(L0); free(lmd); free(MinvJiT); free(jibar); free(A); free(W); free(AW);
    free(At); free(AWAt); free(pin); free(WAt); free(Abar);
  } else {
    for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) a->ni[i * nv + j] = (i == j);
    for (int i = 0; i < nv; i++) a->jit[i] = 0;
    /* sa_ni_trc_bar_tr = eye(n_active): requires n_active == nv - 6 in the reference */
    memset(a->B, 0, sizeof(double) * nact * na_all);
    for (int i = 0; i < nact; i++) a->B[i * na_all + (p->act_idx[i] - 6)] = 1;
  }
  /* S = sa_ni_trc_bar_tr . snf  (nact x nv) */
  memset(a->S, 0, sizeof(double) * nact * nv);
  for (int i = 0; i < nact; i++) for (int j = 0; j < na_all; j++) a->S[i * nv + 6 + j] = a->B[i * na_all + j];

  /* ---- cost, :159-175 */
  double* Pt = (double*)calloc((size_t)nv * nv, 8);
  double* qt = (double*)calloc(nv, 8);
  for (int it = 0; it < p->ntask; it++) {
    int d = p->tasks[it].dim;
    double* J = (double*)calloc((size_t)d * nv, 8);
    double jd[MAXV], xdd[MAXV];
    orc_task_eval(p, q, v, des, it, J, jd, xdd, NULL);
    for (int r = 0; r < nv; r++)
      for (int c = 0; c < nv; c++) {
        double s = 0;
        for (int l = 0; l < d; l++) s += J[l * nv + r] * J[l * nv + c];
        Pt[r * nv + c] += w[it] * s;
      }
    for (int c = 0; c < nv; c++) {
      double s = 0;
      for (int l = 0; l < d; l++) s += (jd[l] - xdd[l]) * J[l * nv + c];
      qt[c] += w[it] * s;
    }
    F += 2.0 * d * nv * nv + 2.0 * d * nv;
    free(J);
  }
  for (int i = 0; i < nv * nv; i++) Pt[i] += p->lambda_q_ddot * a->M[i];
  F += (double)nv * nv;

  /* ---- contacts, :177-198 */
  memset(a->Uf, 0, sizeof(double) * (nu ? nu : 1) * (nc ? nc : 1));
  int ro = 0, co = 0;
  for (int ic = 0; ic < p->ncontact; ic++) {
    int d = p->contacts[ic].type == PNC_CONTACT_SURFACE ? 6 : 3, rows = d == 6 ? 18 : 6;
    double cm[18 * 6], cv[18];
    orc_contact_eval(p, q, v, rf_z_max[ic], ic, a->Jc + (size_t)co * nv, NULL, cm, cv);
    for (int r = 0; r < rows; r++) {
      for (int c = 0; c < d; c++) a->Uf[(ro + r) * nc + co + c] = cm[r * d + c];
      a->uf[ro + r] = cv[r];
    }
    ro += rows; co += d;
  }
  memset(a->P, 0, sizeof(double) * n * n);
  for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) a->P[i * n + j] = Pt[i * nv + j];
  for (int i = 0; i < nc; i++) a->P[(nv + i) * n + nv + i] = p->lambda_rf; /* w_rf = 0 in every shipped controller */
  for (int i = 0; i < nv; i++) a->qv[i] = qt[i];
  for (int i = 0; i < nc; i++) a->qv[nv + i] = 0;
  free(Pt); free(qt);

  /* (Jc Ni) and Ni^T (c+g) */
  mat_mul(a->Jc, a->ni, a->JcNi, nc, nv, nv);
  double Ntg[MAXV];
  for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += a->ni[j * nv + i] * a->cg[j]; Ntg[i] = s; }

  /* ---- equalities, :225-249:  [Sf M | -Sf (Jc Ni)^T] x = -Sf Ni^T (c+g);  [Ji | 0] x = 0 */
  memset(a->Am, 0, sizeof(double) * a->peq * n);
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < nv; j++) a->Am[i * n + j] = a->M[i * nv + j];
    for (int j = 0; j < nc; j++) a->Am[i * n + nv + j] = -a->JcNi[j * nv + i];
    a->b[i] = -Ntg[i];
  }
  for (int i = 0; i < k; i++) {
    for (int j = 0; j < nv; j++) a->Am[(6 + i) * n + j] = p->ic_jac[i * nv + j];
    a->b[6 + i] = 0;
  }
  /* IHWBC2, b_transmission_constraint (ihwbc2.py:249-256): [(Sd - Sv) M | (-Sd + Sv) Jc^T] x = (-Sd + Sv)(c+g) */
  for (int i = 0; i < p->n_trans; i++) {
    int vd = p->trans_d[i], vp = p->trans_p[i], r = 6 + k + i;
    for (int j = 0; j < nv; j++) a->Am[r * n + j] = a->M[vd * nv + j] - a->M[vp * nv + j];
    for (int j = 0; j < nc; j++) a->Am[r * n + nv + j] = -a->Jc[j * nv + vd] + a->Jc[j * nv + vp];
    a->b[r] = -a->cg[vd] + a->cg[vp];
  }

  /* ---- inequalities, :255-317 */
  memset(a->Gm, 0, sizeof(double) * (a->mineq > 0 ? a->mineq : 1) * n);
  for (int r = 0; r < nu; r++) {
    for (int c = 0; c < nc; c++) a->Gm[r * n + nv + c] = -a->Uf[r * nc + c];
    a->h[r] = -a->uf[r];
  }
  if (p->b_trq_limit) {
    double* SM = (double*)calloc((size_t)nact * nv, 8);
    double* SJt = (double*)calloc((size_t)nact * (nc ? nc : 1), 8);
    mat_mul(a->S, a->M, SM, nact, nv, nv);
    for (int i = 0; i < nact; i++)
      for (int j = 0; j < nc; j++) { double s = 0; for (int l = 0; l < nv; l++) s += a->S[i * nv + l] * a->JcNi[j * nv + l]; SJt[i * nc + j] = s; }
    for (int i = 0; i < nact; i++) {
      double sg = 0, sj = 0;
      for (int l = 0; l < nv; l++) { sg += a->S[i * nv + l] * Ntg[l]; sj += a->S[i * nv + l] * a->jit[l]; }
      a->SNtg[i] = sg;
      for (int j = 0; j < nv; j++) { a->Gm[(nu + i) * n + j] = -SM[i * nv + j]; a->Gm[(nu + nact + i) * n + j] = SM[i * nv + j]; }
      for (int j = 0; j < nc; j++) { a->Gm[(nu + i) * n + nv + j] = SJt[i * nc + j]; a->Gm[(nu + nact + i) * n + nv + j] = -SJt[i * nc + j]; }
      a->h[nu + i] = sg + sj - p->trq_limit[2 * i];
      a->h[nu + nact + i] = -sg - sj + p->trq_limit[2 * i + 1];
    }
    free(SM); free(SJt);
  } else {
    for (int i = 0; i < nact; i++) { double sg = 0; for (int l = 0; l < nv; l++) sg += a->S[i * nv + l] * Ntg[l]; a->SNtg[i] = sg; }
  }
  if (flops) *flops = F;
}

void orc_ihwbc_assemble(const orc_problem* p, const double* q, const double* v, const double* des,
                        const double* w, const double* rf_z_max, double* P, double* qv, double* Gm,
                        double* h, double* Am, double* b) {
  asm_t a;
  asm_alloc(p, &a);
  assemble(p, q, v, des, w, rf_z_max, &a, NULL);
  memcpy(P, a.P, sizeof(double) * a.n * a.n);
  memcpy(qv, a.qv, sizeof(double) * a.n);
  if (a.mineq) { memcpy(Gm, a.Gm, sizeof(double) * a.mineq * a.n); memcpy(h, a.h, sizeof(double) * a.mineq); }
  memcpy(Am, a.Am, sizeof(double) * a.peq * a.n);
  memcpy(b, a.b, sizeof(double) * a.peq);
  asm_free(&a);
}

void orc_wbc_tick(const orc_problem* p, const double* bp, const double* bq, const double* blv,
                  const double* bav, const double* jp, const double* jv, const double* des,
                  const double* w, const double* rf_z_max, double* trq, double* acc, double* rf,
                  double* qddot, unsigned char* active, orc_tick_info* info) {
  const orc_model* m = p->m;
  double q[MAXV + 1], v[MAXV];
  orc_pack_state(m, bp, bq, blv, bav, jp, jv, q, v);
  asm_t a;
  asm_alloc(p, &a);
  double Fasm = 0, Fqp = 0;
  assemble(p, q, v, des, w, rf_z_max, &a, &Fasm);
  int n = a.n, pe = a.peq, mi = a.mineq, nv = a.nv, nc = a.nc;
  /* qpsolvers quadprog shim: CE = A^T, ce0 = -b, CI = -G^T, ci0 = h */
  double* CE = (double*)calloc((size_t)n * pe, 8);
  double* ce0 = (double*)calloc(pe, 8);
  double* CI = (double*)calloc((size_t)n * (mi ? mi : 1), 8);
  double* ci0 = (double*)calloc(mi ? mi : 1, 8);
  for (int i = 0; i < pe; i++) { for (int j = 0; j < n; j++) CE[j * pe + i] = a.Am[i * n + j]; ce0[i] = -a.b[i]; }
  for (int i = 0; i < mi; i++) { for (int j = 0; j < n; j++) CI[j * mi + i] = -a.Gm[i * n + j]; ci0[i] = a.h[i]; }
  double* x = (double*)calloc(n, 8);
  int* A = (int*)calloc(mi + pe + 1, sizeof(int));
  double* u = (double*)calloc(mi + pe + 1, 8);
  int iq = 0, iters = 0, status = 0;
  orc_solve_quadprog(n, pe, mi, a.P, a.qv, CE, ce0, CI, ci0, x, A, u, &iq, &iters, &status, &Fqp);
  /* torque recovery, :344-356:  trq = S (M qdd + Ni^T (c+g) - (Jc Ni)^T rf) */
  double tmp[MAXV];
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) s += a.M[i * nv + j] * x[j];
    double ntg = 0;
    for (int j = 0; j < nv; j++) ntg += a.ni[j * nv + i] * a.cg[j];
    double jr = 0;
    for (int j = 0; j < nc; j++) jr += a.JcNi[j * nv + i] * x[nv + j];
    tmp[i] = s + ntg - jr;
  }
  double* Str = NULL; /* S of the transmission variant */
  if (p->n_trans > 0) {
    /* ihwbc2.py:131-139,386-403: sa_trc_bar = weighted_pinv(Sa[:, 6:], Minv[6:, 6:]); f_int = Sv (...);
     * trq = sa_trc_bar^T Snf (M qdd + c + g - Jc^T rf - Ji^T f_int), Ji row k = +1 at the passive (proximal),
     * -1 at the distal joint (draco3_rolling_joint_constraint.py:9-16) */
    int na_all = nv - 6, nact = a.nact;
    mat_inv_spd_or_general(a.M, a.Minv, nv);
    double* A = (double*)calloc((size_t)nact * na_all, 8);
    double* W = (double*)calloc((size_t)na_all * na_all, 8);
    double* AW = (double*)calloc((size_t)nact * na_all, 8);
    double* At = (double*)calloc((size_t)nact * na_all, 8);
    double* AWAt = (double*)calloc((size_t)nact * nact, 8);
    double* pin = (double*)calloc((size_t)nact * nact, 8);
    double* WAt = (double*)calloc((size_t)na_all * nact, 8);
    double* Abar = (double*)calloc((size_t)na_all * nact, 8);
    for (int i = 0; i < nact; i++) A[i * na_all + (p->act_idx[i] - 6)] = 1.0;
    for (int i = 0; i < na_all; i++) for (int j = 0; j < na_all; j++) W[i * na_all + j] = a.Minv[(6 + i) * nv + 6 + j];
    mat_mul(A, W, AW, nact, na_all, na_all);
    mat_tr(A, At, nact, na_all);
    mat_mul(AW, At, AWAt, nact, na_all, nact);
    orc_pinv(AWAt, nact, nact, 1e-15, pin);
    mat_mul(W, At, WAt, na_all, na_all, nact);
    mat_mul(WAt, pin, Abar, na_all, nact, nact);
    Str = (double*)calloc((size_t)nact * nv, 8);
    for (int i = 0; i < nact; i++) for (int j = 0; j < na_all; j++) Str[i * nv + 6 + j] = Abar[j * nact + i];
    for (int k = 0; k < p->n_trans; k++) {
      double f_int = tmp[p->trans_p[k]];
      tmp[p->trans_p[k]] -= f_int;
      tmp[p->trans_d[k]] += f_int;
    }
    free(A); free(W); free(AW); free(At); free(AWAt); free(pin); free(WAt); free(Abar);
  }
  for (int i = 0; i < a.nact; i++) {
    double s = 0;
    const double* Srow = Str ? Str + (size_t)i * nv : a.S + (size_t)i * nv;
    for (int j = 0; j < nv; j++) s += Srow[j] * tmp[j];
    if (trq) trq[i] = s;
    if (acc) acc[i] = x[p->act_idx[i]];
  }
  free(Str);
  if (rf) for (int j = 0; j < nc; j++) rf[j] = x[nv + j];
  if (qddot) for (int j = 0; j < nv; j++) qddot[j] = x[j];
  if (active) {
    memset(active, 0, mi);
    for (int i = pe; i < iq; i++) active[A[i]] = 1;
  }
  if (info) {
    info->status = status; info->iters = iters; info->iq = iq; info->qp_flops = Fqp; info->asm_flops = Fasm;
    orc_last_qp_margin(&info->margin, &info->margin_kind);
    info->psi_rel = orc_last_qp_psi_rel();
  }
  free(CE); free(ce0); free(CI); free(ci0); free(x); free(A); free(u);
  asm_free(&a);
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

typedef struct {
  const orc_problem* p;
  int lo, hi;
  const double *bp, *bq, *blv, *bav, *jp, *jv, *des, *w, *rfz;
  double *trq, *acc, *rf, *qddot, *flops;
  unsigned char* active;
  int *status, *iters;
  double* margin;
  int* margin_kind;
  double* psi_rel;
} shard_t;

static void* shard_run(void* arg) {
  shard_t* s = (shard_t*)arg;
  const orc_problem* p = s->p;
  const orc_model* m = p->m;
  int na = m->na, nv = m->nv, nc = p->nc, nact = p->nact, ndes = p->ndes, nt = p->ntask, ncn = p->ncontact;
  int mi = p->ncone + (p->b_trq_limit ? 2 * p->nact : 0);
  int fl = m->nfl != 0;
  for (int i = s->lo; i < s->hi; i++) {
    orc_tick_info info;
    double ltrq[MAXV], lacc[MAXV], lrf[MAXV], lqdd[MAXV];
    unsigned char lact[512];
    orc_wbc_tick(p, fl ? s->bp + 3 * (size_t)i : NULL, fl ? s->bq + 4 * (size_t)i : NULL,
                 fl ? s->blv + 3 * (size_t)i : NULL, fl ? s->bav + 3 * (size_t)i : NULL, s->jp + (size_t)na * i,
                 s->jv + (size_t)na * i, s->des + (size_t)ndes * i, s->w + (size_t)nt * i, s->rfz + (size_t)ncn * i,
                 ltrq, lacc, lrf, lqdd, lact, &info);
    if (s->trq) memcpy(s->trq + (size_t)nact * i, ltrq, 8 * nact);
    if (s->acc) memcpy(s->acc + (size_t)nact * i, lacc, 8 * nact);
    if (s->rf) memcpy(s->rf + (size_t)nc * i, lrf, 8 * nc);
    if (s->qddot) memcpy(s->qddot + (size_t)nv * i, lqdd, 8 * nv);
    if (s->active) memcpy(s->active + (size_t)mi * i, lact, mi);
    if (s->status) s->status[i] = info.status;
    if (s->iters) s->iters[i] = info.iters;
    if (s->flops) s->flops[i] = info.qp_flops + info.asm_flops;
    if (s->margin) s->margin[i] = info.margin;
    if (s->margin_kind) s->margin_kind[i] = info.margin_kind;
    if (s->psi_rel) s->psi_rel[i] = info.psi_rel;
  }
  return NULL;
}

double orc_wbc_tick_batch(const orc_problem* p, int B, int nthreads, const double* bp, const double* bq,
                          const double* blv, const double* bav, const double* jp, const double* jv,
                          const double* des, const double* w, const double* rfz, double* trq, double* acc,
                          double* rf, double* qddot, unsigned char* active, int* status, int* iters,
                          double* flops) {
  return orc_wbc_tick_batch_margin(p, B, nthreads, bp, bq, blv, bav, jp, jv, des, w, rfz, trq, acc, rf, qddot, active,
                                   status, iters, flops, NULL, NULL, NULL);
}

double orc_wbc_tick_batch_margin(const orc_problem* p, int B, int nthreads, const double* bp, const double* bq,
                                 const double* blv, const double* bav, const double* jp, const double* jv,
                                 const double* des, const double* w, const double* rfz, double* trq, double* acc,
                                 double* rf, double* qddot, unsigned char* active, int* status, int* iters,
                                 double* flops, double* margin, int* margin_kind, double* psi_rel) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  shard_t sh[256];
  pthread_t th[256];
  double t0 = now_s();
  for (int t = 0; t < nthreads; t++) {
    shard_t s = {p, (int)((long long)B * t / nthreads), (int)((long long)B * (t + 1) / nthreads),
                 bp, bq, blv, bav, jp, jv, des, w, rfz, trq, acc, rf, qddot, flops, active, status, iters,
                 margin, margin_kind, psi_rel};
    sh[t] = s;
    if (nthreads > 1) pthread_create(&th[t], NULL, shard_run, &sh[t]);
  }
  if (nthreads == 1) shard_run(&sh[0]);
  else for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return now_s() - t0;
}

/* manipulator_interface.py:77-114 (without the time-dependent blend, which is an input) */
void orc_osc_step(const orc_model* m, int ee, const double* q, const double* v, const double* pos_des,
                  const double* vel_des, const double* acc_des, const double* kp, const double* kd,
                  double* trq, double* qddot) {
  int nv = m->nv;
  double J6[6 * MAXV], T[16], v6[6], xdd[3], Jp[3 * MAXV], M[MAXV * MAXV], c[MAXV], g[MAXV], qdd[MAXV];
  (void)acc_des; /* the reference drops acc_des: xddot_des = KP err + KD err_d (:98-99) */
  orc_link_jacobian(m, q, ee, J6);
  orc_link_iso(m, q, ee, T);
  orc_link_vel(m, q, v, ee, v6);
  for (int i = 0; i < 3; i++) xdd[i] = kp[i] * (pos_des[i] - T[4 * i + 3]) + kd[i] * (vel_des[i] - v6[3 + i]);
  orc_pinv(J6 + 3 * nv, 3, nv, 1e-3, Jp); /* nv x 3 */
  for (int i = 0; i < nv; i++) qdd[i] = Jp[i * 3] * xdd[0] + Jp[i * 3 + 1] * xdd[1] + Jp[i * 3 + 2] * xdd[2];
  orc_mass_matrix(m, q, M);
  orc_coriolis(m, q, v, c);
  orc_gravity(m, q, g);
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) s += M[i * nv + j] * qdd[j];
    trq[i] = s + c[i] + g[i];
    if (qddot) qddot[i] = qdd[i];
  }
}

/* B OSC steps in a plain loop (the CPU baseline of BASELINE config 1); returns wall seconds */
double orc_osc_step_batch(const orc_model* m, int ee, int B, const double* q, const double* v, const double* pos_des,
                          const double* vel_des, const double* kp, const double* kd, double* trq, double* qddot) {
  const int nv = m->nv;
  const double zero3[3] = {0, 0, 0};
  double t0 = now_s();
  for (int b = 0; b < B; b++)
    orc_osc_step(m, ee, q + (size_t)b * m->nq, v + (size_t)b * nv, pos_des + 3 * (size_t)b, vel_des + 3 * (size_t)b, zero3, kp,
                 kd, trq + (size_t)b * nv, qddot ? qddot + (size_t)b * nv : NULL);
  return now_s() - t0;
}
