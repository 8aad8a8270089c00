/*
 * pnc_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C fp64 restatement of the reference's hot path, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs to
 * check the CUDA kernels.  Nothing under pypnc_b200/ may link or call this.
 *
 * Parity status of this oracle:
 *   - QP: the port of QuadProg++ (orc_solve_quadprog) is pinned bit-for-bit against the
 *     reference's own external_source/Goldfarb/QuadProg++.cc compiled into oracle/_ref
 *     and against the known-answer QPs (tests/test_oracle_qp.py).
 *   - IHWBC assembly, BasicTask, contacts: pinned against the reference's own Python
 *     classes (imported from /root/reference with qpsolvers stubbed onto oracle/_ref;
 *     fixtures in tests/golden/, generator tests/golden/make_golden.py).
 *   - Rigid-body algorithms: the reference delegates these to pinocchio 2.5.5, which is
 *     not vendored and not installable here => "parity unpinned" at that boundary.
 *     They are restated from the definitions and checked against
 *     the reference's one closed-form fixture (test/two_link_manipulator_test.py:19-48)
 *     and independent identities (energy-derived M, potential-gradient g, finite
 *     differences of J, Christoffel Coriolis).
 */
#ifndef PNC_ORACLE_H
#define PNC_ORACLE_H
#include "../include/pnc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;
typedef struct orc_problem orc_problem;

orc_model* orc_model_create(const pnc_model_desc* d);
void orc_model_destroy(orc_model* m);
orc_problem* orc_problem_create(const orc_model* m, const pnc_problem_desc* d);
void orc_problem_destroy(orc_problem* p);

/* pinocchio_robot_system.py:128-176: pack (q, qdot) from sensor-style inputs */
void orc_pack_state(const orc_model* m, const double* base_pos, const double* base_quat,
                    const double* base_lin_vel, const double* base_ang_vel, const double* joint_pos,
                    const double* joint_vel, double* q, double* v);

/* the getters of PinocchioRobotSystem, on one (q, v) */
void orc_mass_matrix(const orc_model* m, const double* q, double* M);                   /* :202 */
void orc_gravity(const orc_model* m, const double* q, double* g);                       /* :205 */
void orc_coriolis(const orc_model* m, const double* q, const double* v, double* c);     /* :209 */
void orc_com(const orc_model* m, const double* q, const double* v, double* com, double* vcom); /* :214-220 */
void orc_com_jacobian(const orc_model* m, const double* q, double* J);                  /* :222 */
void orc_com_jacobian_dot(const orc_model* m, const double* q, const double* v, double* dJ); /* :226 */
void orc_link_iso(const orc_model* m, const double* q, int frame, double* T16);         /* :231 */
void orc_link_vel(const orc_model* m, const double* q, const double* v, int frame, double* v6); /* :239 */
void orc_link_jacobian(const orc_model* m, const double* q, int frame, double* J);      /* :252 */
void orc_link_jdqd(const orc_model* m, const double* q, const double* v, int frame, double* a6); /* :265 */
void orc_centroidal(const orc_model* m, const double* q, const double* v, double* Ag, double* hg,
                    double* Ig);                                                       /* :181-194 */
/* RNEA(q, v, a) with gravity on/off (building block; tau[nv]) */
void orc_rnea(const orc_model* m, const double* q, const double* v, const double* a, int with_gravity,
              double* tau);

/* QuadProg++.cc:115-502 restated.  G[n*n] (overwritten by its Cholesky factor), g0[n],
 * CE[n*p], ce0[p], CI[n*m], ci0[m] row-major with one COLUMN per constraint.
 * Returns f (inf when infeasible); *status = PNC_QP_*.  A[0..*iq) = active set
 * (equalities as -i-1, inequalities as index), u[0..*iq) multipliers. */
double orc_solve_quadprog(int n, int p, int m, double* G, const double* g0, const double* CE,
                          const double* ce0, const double* CI, const double* ci0, double* x, int* A,
                          double* u, int* iq, int* iters, int* status, double* flops);

/* util.py helpers */
void orc_quat_to_rot(const double* quat_xyzw, double* R9);            /* util.py:33-44 */
void orc_rot_to_quat(const double* R9, double* quat_xyzw);            /* util.py:47-58 (scipy from_matrix) */
void orc_quat_to_exp(const double* quat_xyzw, double* exp3);          /* util.py:61-72 */
/* np.linalg.pinv(A[m*n], rcond) -> P[n*m] */
void orc_pinv(const double* A, int m, int n, double rcond, double* P);

typedef struct {
  int status, iters, iq;
  double qp_flops, asm_flops;
  double margin;    /* decision margin of the QP solve, see orc_last_qp_margin */
  int margin_kind;
  double psi_rel;   /* summed violation / max |slack| left at the returned iterate (0: exactly feasible) */
} orc_tick_info;

/* Smallest relative distance between the two sides of any data-dependent branch the last
 * orc_solve_quadprog call on this thread took (1 = never close), and which branch kind it was. */
void orc_last_qp_margin(double* margin, int* kind);
double orc_last_qp_psi_rel(void);

/* BasicTask.update_jacobian + update_cmd of task `it` (basic_task.py:25-137) */
void orc_task_eval(const orc_problem* p, const double* q, const double* v, const double* des, int it,
                   double* jac, double* jdqd, double* op_cmd, double* pos_err);
/* Contact.update_contact (basic_contact.py) */
void orc_contact_eval(const orc_problem* p, const double* q, const double* v, double rf_z_max, int ic,
                      double* jac, double* jdqd, double* cone_mat, double* cone_vec);

/* The dense QP the reference hands to qpsolvers (ihwbc.py:124-328), for inspection:
 * P[n*n], qv[n], Gm[m*n], h[m], Am[p*n], b[p] */
void orc_ihwbc_assemble(const orc_problem* p, const double* q, const double* v, const double* des,
                        const double* w, const double* rf_z_max, double* P, double* qv, double* Gm,
                        double* h, double* Am, double* b);

/* One controller tick (atlas_controller.py:51-77): update + IHWBC.solve.
 * active[m] : 1 when inequality i is in the final active set */
void orc_wbc_tick(const orc_problem* p, const double* base_pos, const double* base_quat,
                  const double* base_lin_vel, const double* base_ang_vel, const double* joint_pos,
                  const double* joint_vel, const double* des, const double* w, const double* rf_z_max,
                  double* trq, double* acc, double* rf, double* qddot, unsigned char* active,
                  orc_tick_info* info);

/* B ticks on `nthreads` host threads (contiguous shards) -- the CPU baseline loop.
 * Any output may be NULL.  Returns wall seconds. */
double orc_wbc_tick_batch(const orc_problem* p, int B, int nthreads, const double* base_pos,
                          const double* base_quat, const double* base_lin_vel, const double* base_ang_vel,
                          const double* joint_pos, const double* joint_vel, const double* des,
                          const double* w, const double* rf_z_max, double* trq, double* acc, double* rf,
                          double* qddot, unsigned char* active, int* status, int* iters,
                          double* flops);

/* same, plus the per-state decision margin of the QP solve (margin[B], margin_kind[B]; may be NULL) */
double orc_wbc_tick_batch_margin(const orc_problem* p, int B, int nthreads, const double* base_pos,
                                 const double* base_quat, const double* base_lin_vel, const double* base_ang_vel,
                                 const double* joint_pos, const double* joint_vel, const double* des,
                                 const double* w, const double* rf_z_max, double* trq, double* acc, double* rf,
                                 double* qddot, unsigned char* active, int* status, int* iters, double* flops,
                                 double* margin, int* margin_kind, double* psi_rel);

/* manipulator_interface.py:77-114 OSC step on a fixed-base arm: returns trq[na] */
void orc_osc_step(const orc_model* m, int ee_frame, const double* q, const double* v, const double* pos_des,
                  const double* vel_des, const double* acc_des, const double* kp, const double* kd,
                  double* trq, double* qddot);

/* B OSC steps in a loop: q [B,nq], v [B,nv], pos_des / vel_des [B,3]; returns wall seconds */
double orc_osc_step_batch(const orc_model* m, int ee_frame, int B, const double* q, const double* v,
                          const double* pos_des, const double* vel_des, const double* kp, const double* kd,
                          double* trq, double* qddot);

void orc_problem_dims(const orc_problem* p, int* nc, int* ncone, int* n, int* peq, int* mineq);

#ifdef __cplusplus
}
#endif
#endif
