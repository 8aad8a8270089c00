/*
 * pnc_b200.h -- C ABI of the batched whole-body-control engine (B200, sm_100a).
 *
 * The reference (junhyeokahn/PyPnC) has no FFI on this path: the boundary is two
 * Python classes.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference checkout):
 *
 *   pnc_model_create         PinocchioRobotSystem._config_robot
 *                            pnc/robot_system/pinocchio_robot_system.py:33-93
 *   pnc_update_system        PinocchioRobotSystem.update_system        :128-179
 *                            + every getter the controller calls per tick:
 *                            get_mass_matrix :202, get_gravity :205, get_coriolis :209,
 *                            get_com_pos/lin_vel :214-220, get_com_lin_jacobian :222,
 *                            get_com_lin_jacobian_dot :226, get_link_iso :231,
 *                            get_link_vel :239, get_link_jacobian :252,
 *                            get_link_jacobian_dot_times_qdot :265
 *   pnc_ws_*                 the getters above, as device pointers into the workspace
 *   pnc_problem_create       <Robot>TCIContainer.__init__ + <Robot>Controller.__init__
 *                            pnc/atlas_pnc/atlas_tci_container.py:10-99,
 *                            pnc/atlas_pnc/atlas_controller.py:10-49
 *   pnc_ihwbc_solve          BasicTask.update_jacobian/update_cmd  pnc/wbc/basic_task.py:25-137
 *                            Contact.update_contact                pnc/wbc/basic_contact.py:24-173
 *                            IHWBC.update_setting + IHWBC.solve    pnc/wbc/ihwbc/ihwbc.py:84-379
 *                            (QP: qpsolvers->quadprog at ihwbc.py:330; algorithm of
 *                             external_source/Goldfarb/QuadProg++.cc:115-502)
 *   pnc_task_eval            Task.jacobian / jacobian_dot_q_dot / op_cmd (debug + parity)
 *   pnc_joint_integrate      JointIntegrator.integrate  pnc/wbc/ihwbc/joint_integrator.py:73-82
 *   pnc_wbc_tick_host        <Robot>Controller.get_command body  atlas_controller.py:51-86,
 *                            host buffers in, host buffers out (the e2e call)
 *
 * Conventions: fp64 everywhere; quaternions scalar-last (x,y,z,w);
 * spatial quantities at the API are [angular; linear]; generalized velocity =
 * [base lin (base frame); base ang (base frame); joints]; reaction force of a surface
 * contact = [torque(3); force(3)] in world axes; all batched arrays are row-major with
 * the batch index first.  Pointers named d_* are DEVICE pointers, h_* HOST pointers.
 * Every call is asynchronous on `stream` (a cudaStream_t passed as void*) unless it
 * says otherwise; calls on one handle are stream-ordered, not re-entrant.  Different handles may be driven
 * from different host threads (e.g. one thread per GPU); the process-wide modes (pnc_set_termination,
 * pnc_set_precision) are read at launch time and are meant to be set before the threads start.
 *
 * Return value: 0 on success, a negative PNC_E_* code on API misuse / CUDA failure.
 * Per-state solver outcomes are reported in `status[B]` (PNC_QP_*), never by abort.
 */
#ifndef PNC_B200_H
#define PNC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNC_OK 0
#define PNC_E_ARG (-1)
#define PNC_E_CUDA (-2)
#define PNC_E_SIZE (-3)    /* model/problem exceeds a compiled-in limit */
#define PNC_E_NOGPU (-4)

/* joint types (model tables) */
#define PNC_JOINT_FREEFLYER 0
#define PNC_JOINT_REVOLUTE 1

/* task types: pnc/wbc/basic_task.py:26-105 */
#define PNC_TASK_JOINT 0
#define PNC_TASK_SELECTED_JOINT 1
#define PNC_TASK_LINK_XYZ 2
#define PNC_TASK_LINK_ORI 3
#define PNC_TASK_COM 4

/* contact types: pnc/wbc/basic_contact.py:13,54 */
#define PNC_CONTACT_POINT 0   /* dim 3, 6 cone rows  */
#define PNC_CONTACT_SURFACE 1 /* dim 6, 18 cone rows */

/* per-state QP status */
#define PNC_QP_OK 0
#define PNC_QP_INFEASIBLE 1     /* QuadProg++.cc:402-407 returns inf            */
#define PNC_QP_NOT_PD 2         /* Cholesky pivot <= 0, QuadProg++.cc:715-722   */
#define PNC_QP_EQ_DEPENDENT 3   /* "Constraints are linearly dependent" :268-271 */
#define PNC_QP_MAX_ITER 4

/* Termination of the active-set loop (pnc_set_termination).  QuadProg++.cc:306-312 stops when the
 * summed infeasibility |psi| <= m eps c1 c2 100, which is 1e-4 .. 1e-3 for these problems.
 * REFINED (default) remembers the first iterate that meets that bound and keeps iterating to
 * 1e-6 x the bound (falling back to the remembered iterate if that cannot complete);
 * REFERENCE stops exactly where QuadProg++ stops. */
#define PNC_TERM_REFINED 0
#define PNC_TERM_REFERENCE 1

/* Arithmetic of the robot_system update (pnc_set_precision).  FP64 (default) is the path every 1e-9 parity claim
 * refers to.  FP32_DYNAMICS is the north star's optional fp32 mode: k_update_system computes forward kinematics,
 * Jacobians, CRBA and RNEA in fp32 and stores fp64; the Cholesky factorisations and the active-set iterations of
 * the solve stage stay fp64 (the Hessian's lambda_q = 1e-8 regularisation, config/atlas_config.py:65-66, is below
 * fp32 resolution).  Tolerance: 1e-4 relative on states whose active set is the fp64 one. */
#define PNC_PREC_FP64 0
#define PNC_PREC_FP32_DYNAMICS 1

/* limits compiled into the kernels */
#define PNC_MAX_BODIES 32   /* one warp lane per movable body */
#define PNC_MAX_NV 40
#define PNC_MAX_FRAMES 12   /* distinct task/contact frames of one problem */
#define PNC_MAX_TASKS 16
#define PNC_MAX_CONTACTS 4
#define PNC_MAX_IC 4        /* internal-constraint rows */
#define PNC_SWING_REC 26    /* doubles per state of a swing-foot trajectory record (pnc_swing_init) */
#define PNC_FB_REC 14       /* doubles per state of a floating-base trajectory record (pnc_floating_base_init) */
#define PNC_HAND_REC 24     /* doubles per state of a hand trajectory record (pnc_hand_init) */

/* ---- model ------------------------------------------------------------- */
typedef struct {
  int32_t nb;         /* movable bodies; body 0 is the free-flyer on a floating base */
  int32_t nq, nv, na; /* as RobotSystem.n_q / n_q_dot / n_a */
  int32_t n_floating; /* 0 or 6 */
  int32_t nf;         /* frames */
  double total_mass;  /* sum of all link masses (wrapper's total_mass, :73-74) */
  double mass_dyn;    /* mass seen by the centre-of-mass algorithms (moving bodies) */
  const int32_t* parent;   /* [nb] -1 = universe; parent[i] < i, DFS numbering */
  const int32_t* jtype;    /* [nb] */
  const int32_t* idx_v;    /* [nb] */
  const int32_t* subtree;  /* [nb] bodies i..i+subtree[i]-1 form the subtree of i */
  const double* place_R;   /* [nb*9] joint placement in parent joint frame, row-major */
  const double* place_p;   /* [nb*3] */
  const double* axis;      /* [nb*3] unit axis (revolute) */
  const double* mass;      /* [nb] */
  const double* com;       /* [nb*3] joint frame */
  const double* inertia;   /* [nb*9] about com, joint-frame axes */
  const int32_t* frame_parent; /* [nf] body index, -1 = universe */
  const double* frame_R;   /* [nf*9] placement in the parent joint frame */
  const double* frame_p;   /* [nf*3] */
} pnc_model_desc;

typedef struct pnc_model pnc_model;

int pnc_model_create(const pnc_model_desc* desc, int device, pnc_model** out);
void pnc_model_destroy(pnc_model* m);

/* ---- WBC problem (tasks, contacts, internal constraints, selection) ---- */
typedef struct {
  int32_t type;      /* PNC_TASK_* */
  int32_t dim;
  int32_t frame;     /* model frame index (LINK_*), else -1 */
  int32_t joint_off; /* SELECTED_JOINT: offset into joint_sel[] (v indices) */
  int32_t des_off;   /* offset of this task in the per-state `des` row:
                        pos_des (dim; 4 = quat xyzw for LINK_ORI), vel_des (dim), acc_des (dim) */
  int32_t gain_off;  /* offset into kp[] / kd[] */
} pnc_task_desc;

typedef struct {
  int32_t type;  /* PNC_CONTACT_* */
  int32_t frame; /* model frame index */
  double mu, x, y; /* friction, half sizes (surface only) */
} pnc_contact_desc;

typedef struct {
  int32_t ntask;
  const pnc_task_desc* tasks;
  int32_t ncontact;
  const pnc_contact_desc* contacts;
  int32_t n_joint_sel;
  const int32_t* joint_sel; /* v indices referenced by SELECTED_JOINT tasks */
  int32_t n_gain;
  const double* kp;
  const double* kd;
  int32_t n_ic;            /* rows of the internal-constraint Jacobian (0 = none) */
  const double* ic_jac;    /* [n_ic*nv] constant rows (draco3_rolling_joint_constraint.py:9-16) */
  int32_t nact;            /* rows of Sa */
  const int32_t* act_idx;  /* [nact] v index of each actuated joint (Sa one-hot columns) */
  int32_t b_trq_limit;     /* WBCConfig.B_TRQ_LIMIT */
  const double* trq_limit; /* [nact*2] lower, upper (ihwbc.trq_limit) */
  double lambda_q_ddot;
  double lambda_rf;
  int32_t ndes;            /* length of one per-state `des` row */
  /* IHWBC2 transmission constraint (pnc/wbc/ihwbc/ihwbc2.py:131-139,249-256,386-403; b_transmission_constraint):
   * n_trans equality rows (Sd - Sv)(M qddot + c + g - Jc^T rf) = 0 instead of the internal-constraint rows, no
   * projection (Ni = I), and the internal force Sv(...) fed back through Ji^T in the torque recovery.  Row k pairs
   * the distal joint trans_distal[k] (one-hot column of Sd) with the passive joint trans_passive[k] (column of
   * Sv, the joint the internal-constraint row k carries with +1); v indices.  Requires n_ic = 0, no torque limits. */
  int32_t n_trans;
  const int32_t* trans_distal;
  const int32_t* trans_passive;
} pnc_problem_desc;

typedef struct pnc_problem pnc_problem;

int pnc_problem_create(const pnc_model* m, const pnc_problem_desc* desc, pnc_problem** out);
void pnc_problem_destroy(pnc_problem* p);
/* sizes derived from the descriptor: QP n = nv + nc, p = 6 + n_ic + n_trans, m = ncone + 2*nact */
int pnc_problem_dims(const pnc_problem* p, int32_t* nc, int32_t* ncone, int32_t* n, int32_t* p_eq,
                     int32_t* m_ineq);

/* ---- workspace: per-state rigid-body quantities kept in HBM ------------- */
typedef struct pnc_workspace pnc_workspace;

/* frames[nframes]: the model frames whose iso/vel/J/Jdot*qdot are materialised */
int pnc_workspace_create(const pnc_model* m, int32_t capacity, const int32_t* frames, int32_t nframes,
                         pnc_workspace** out);
void pnc_workspace_destroy(pnc_workspace* ws);
int64_t pnc_workspace_bytes(const pnc_workspace* ws);

/* robot_system.update_system for B states (fixed base: pass NULL base pointers).
 * base_lin_vel / base_ang_vel are WORLD-frame (base_joint_*), rotated into the base
 * frame inside, as pinocchio_robot_system.py:149-162 does. */
int pnc_update_system(pnc_workspace* ws, int32_t B, const double* d_base_pos, const double* d_base_quat,
                      const double* d_base_lin_vel, const double* d_base_ang_vel,
                      const double* d_joint_pos, const double* d_joint_vel, void* stream);

/* update_system(..., b_cent): additionally fills the centroidal quantities of
 * PinocchioRobotSystem._update_centroidal_quantities (:181-194, pin.ccrba) when b_cent != 0:
 * Ag [B,6,nv] and hg [B,6] with rows [angular; linear], Ig [B,6,6] (diagonal 3x3 blocks only). */
int pnc_update_system_cent(pnc_workspace* ws, int32_t B, const double* d_base_pos, const double* d_base_quat,
                           const double* d_base_lin_vel, const double* d_base_ang_vel,
                           const double* d_joint_pos, const double* d_joint_vel, int32_t b_cent, void* stream);
double* pnc_ws_centroidal_ag(pnc_workspace* ws);
double* pnc_ws_centroidal_hg(pnc_workspace* ws);
double* pnc_ws_centroidal_ig(pnc_workspace* ws);

/* device pointers into the workspace, valid until the workspace is destroyed.
 * Shapes (row-major, batch first): q [B,nq], qdot [B,nv], M [B,nv,nv], grav [B,nv],
 * cori [B,nv], com [B,3], vcom [B,3], jcom [B,3,nv], jcomdot [B,3,nv], jcomdot_qdot [B,3],
 * frame_iso [B,nfr,12] (R row-major 9, p 3), frame_vel [B,nfr,6] ([ang;lin]),
 * frame_jac [B,nfr,6,nv] ([ang;lin] rows), frame_jdqd [B,nfr,6] ([ang;lin]). */
double* pnc_ws_q(pnc_workspace* ws);
double* pnc_ws_qdot(pnc_workspace* ws);
double* pnc_ws_mass_matrix(pnc_workspace* ws);
double* pnc_ws_gravity(pnc_workspace* ws);
double* pnc_ws_coriolis(pnc_workspace* ws);
double* pnc_ws_com_pos(pnc_workspace* ws);
double* pnc_ws_com_vel(pnc_workspace* ws);
double* pnc_ws_com_jac(pnc_workspace* ws);
double* pnc_ws_com_jacdot(pnc_workspace* ws);
double* pnc_ws_com_jdqd(pnc_workspace* ws);
double* pnc_ws_frame_iso(pnc_workspace* ws);
double* pnc_ws_frame_vel(pnc_workspace* ws);
double* pnc_ws_frame_jac(pnc_workspace* ws);
double* pnc_ws_frame_jdqd(pnc_workspace* ws);

/* ---- IHWBC ---------------------------------------------------------------- */
/* One call (stream-ordered kernels: task errors, per-state setup, active-set loop, outputs): task
 * Jacobians/commands, contact cones, cost + constraint assembly, batched Goldfarb-Idnani solve,
 * torque recovery.  Requires pnc_update_system on `ws`.
 *   d_des        [B,ndes]     task desireds (layout: pnc_task_desc.des_off)
 *   d_w          [B,ntask]    w_hierarchy per task
 *   d_rf_z_max   [B,ncontact] (values <= 0 are replaced by 1e-3, contact.py:37-41)
 * outputs
 *   d_trq [B,nact]  d_acc [B,nact]  d_rf [B,nc]  d_qddot [B,nv] (may be NULL)
 *   d_status [B] int32 PNC_QP_*   d_iters [B] int32 (active-set iterations)
 *   d_active [B, ceil(m/64)] uint64 bit i set <=> inequality i active at the solution
 *            (inequality order = rows of ineq_mat, ihwbc.py:255-317: cones, -trq, +trq) */
int pnc_ihwbc_solve(const pnc_problem* p, pnc_workspace* ws, int32_t B, const double* d_des,
                    const double* d_w, const double* d_rf_z_max, double* d_trq, double* d_acc,
                    double* d_rf, double* d_qddot, int32_t* d_status, int32_t* d_iters,
                    uint64_t* d_active, void* stream);

/* pnc_ihwbc_solve with one more output: d_active_ref [B, ceil(m/64)] = the active set at the first
 * iterate that met the termination test of QuadProg++.cc:306-312, i.e. the set solve_quadprog
 * itself returns (equal to d_active when the refinement added nothing; may be NULL). */
int pnc_ihwbc_solve_ref(const pnc_problem* p, pnc_workspace* ws, int32_t B, const double* d_des,
                        const double* d_w, const double* d_rf_z_max, double* d_trq, double* d_acc,
                        double* d_rf, double* d_qddot, int32_t* d_status, int32_t* d_iters,
                        uint64_t* d_active, uint64_t* d_active_ref, void* stream);

/* Process-wide termination mode of the IHWBC active-set loop: PNC_TERM_REFINED (default, or
 * PNC_TERMINATION=reference in the environment for the other) / PNC_TERM_REFERENCE. */
int pnc_set_termination(int32_t mode);
int pnc_get_termination(void);

/* Process-wide arithmetic of pnc_update_system / the dynamics part of the pnc_wbc_tick_host calls: PNC_PREC_FP64
 * (default, or PNC_PRECISION=fp32 in the environment for the other) / PNC_PREC_FP32_DYNAMICS. */
int pnc_set_precision(int32_t mode);
int pnc_get_precision(void);

/* Task.jacobian [B,dim,nv], jacobian_dot_q_dot [B,dim], op_cmd [B,dim], pos_err [B,dim]
 * of task `itask` (any output may be NULL). */
int pnc_task_eval(const pnc_problem* p, pnc_workspace* ws, int32_t B, int32_t itask, const double* d_des,
                  double* d_jac, double* d_jdqd, double* d_op_cmd, double* d_pos_err, void* stream);

/* Contact.jacobian [B,dim,nv], cone_constraint_mat [B,rows,dim], cone_constraint_vec [B,rows] */
int pnc_contact_eval(const pnc_problem* p, pnc_workspace* ws, int32_t B, int32_t icontact,
                     const double* d_rf_z_max, double* d_jac, double* d_cone_mat, double* d_cone_vec,
                     void* stream);

/* JointIntegrator.integrate: state vel/pos [B,na] updated in place, returned as the
 * commands.  limits are [na*2]. */
int pnc_joint_integrate(int32_t B, int32_t na, const double* d_acc, const double* d_joint_pos,
                        double* d_vel_state, double* d_pos_state, const double* d_pos_limit,
                        const double* d_vel_limit, double alpha_vel, double alpha_pos, double dt,
                        void* stream);

/* Operational-space control step of the fixed-base manipulator, ManipulatorInterface._compute_osc_command
 * (pnc/manipulator_pnc/manipulator_interface.py:77-114), for B states after pnc_update_system:
 * xddot = kp (pos_des - pos) + kd (vel_des - vel) on the LINEAR rows of frame `frame` (a model frame index
 * the workspace materialises), qddot = smooth * pinv(J, rcond = 1e-3) xddot, trq = M qddot + c + g.
 * d_pos_des / d_vel_des [B,3], d_kp / d_kd [3] (device), outputs d_trq [B,nv], d_qddot [B,nv] (may be NULL). */
int pnc_osc_step(pnc_workspace* ws, int32_t B, int32_t frame, const double* d_pos_des, const double* d_vel_des,
                 const double* d_kp, const double* d_kd, double smooth, double* d_trq, double* d_qddot, void* stream);

/* Batched dense QP: B independent problems of one shape, each
 *   min 1/2 x^T G x + g0^T x   s.t.  CE^T x + ce0 = 0,  CI^T x + ci0 >= 0
 * -- the reference's own C++ entry point solve_quadprog(G, g0, CE, ce0, CI, ci0, x)
 * (external_source/Goldfarb/QuadProg++.hh:70, QuadProg++.cc:115-502; its in-repo caller is the
 * wrench-distribution QP of towr_plus/src/dcm_planner.cpp:752-855, n = 12, p = 6, m = 36).
 * Layout as QuadProg++'s Matrix<double>: row-major d_G [B,n,n], d_CE [B,n,p], d_CI [B,n,m] (one
 * COLUMN per constraint), d_g0 [B,n], d_ce0 [B,p], d_ci0 [B,m].  Outputs d_x [B,n], d_f [B] (the
 * cost; +inf for an infeasible problem as in :402-407), d_status [B] (PNC_QP_*), d_iters [B],
 * d_active [B, ceil(m/64)] bit mask of the active inequalities; any output but d_x may be NULL.
 * n <= 64; returns PNC_E_SIZE when the shape does not fit shared memory.  `device` is the CUDA
 * device the pointers live on. */
int pnc_quadprog_batch(int32_t device, int32_t B, int32_t n, int32_t p, int32_t m, const double* d_G, const double* d_g0,
                       const double* d_CE, const double* d_ce0, const double* d_CI, const double* d_ci0, double* d_x,
                       double* d_f, int32_t* d_status, int32_t* d_iters, uint64_t* d_active, void* stream);

/* ---- the reference generators that feed the tasks each tick ----
 * All pointers are device pointers, fp64, batch first; `frame` is a model frame index the
 * workspace materialises; quaternions scalar-last.  Outputs go straight into
 * BasicTask.update_desired(pos, vel, acc) / the w and rf_z_max arrays of pnc_ihwbc_solve. */

/* TaskHierarchyManager.update_ramp_to_{min,max} (pnc/wbc/manager/task_hierarchy_manager.py:23-35) and
 * ReactionForceManager.update_ramp_to_{min,max} (reaction_force_manager.py:23-35):
 * out = (target - start_value) / duration * (clip(t, t0, t0 + duration) - t0) + start_value.  Arrays [B]. */
int pnc_ramp_update(int32_t B, const double* d_start_value, double target, const double* d_start_time,
                    const double* d_duration, double current_time, double* d_out, void* stream);

/* FootTrajectoryManager.use_current (pnc/wbc/manager/foot_trajectory_manager.py:41-52): desireds of the
 * LINK_XYZ / LINK_ORI task pair of `frame` = its current pose and velocity, zero acceleration.
 * pos/vel/acc/angvel/angacc [B,3], quat [B,4]. */
int pnc_foot_use_current(pnc_workspace* ws, int32_t B, int32_t frame, double* d_pos, double* d_vel, double* d_acc,
                         double* d_quat, double* d_angvel, double* d_angacc, void* stream);

/* FootTrajectoryManager.initialize_swing_foot_trajectory (:54-84): from the current pose of `frame`
 * and the landing foot (d_land_pos [B,3], d_land_quat [B,4]) build the two Hermite position
 * segments through the raised mid foot (footstep.py:64-69 interpolate, swing_height) and the
 * HermiteCurveQuat (util/interpolation.py:104-134) into d_rec [B, PNC_SWING_REC]. */
int pnc_swing_init(pnc_workspace* ws, int32_t B, int32_t frame, const double* d_land_pos, const double* d_land_quat,
                   const double* d_swing_duration, double swing_height, double* d_rec, void* stream);

/* FootTrajectoryManager.update_swing_foot_desired (:86-110).  Derivatives are with respect to the
 * segment parameter, exactly as the reference passes them on. */
int pnc_swing_eval(int32_t B, const double* d_rec, const double* d_start_time, const double* d_swing_duration,
                   double current_time, double* d_pos, double* d_vel, double* d_acc, double* d_quat, double* d_angvel,
                   double* d_angacc, void* stream);

/* HandTrajectoryManager.initialize_hand_trajectory / initialize_keypoint_hand_trajectory
 * (pnc/wbc/manager/hand_trajectory_manager.py:57-102): from the current placement and angular velocity of
 * `frame` and the target hand pose (d_target_pos [B,3], d_target_quat [B,4] = rot_to_quat of target_hand_iso;
 * d_keypoint [B,3] or NULL) build the record of the cosine position blend and the HermiteCurveQuat
 * (util/interpolation.py:104-134, start angular velocity = the hand's) into d_rec [B, PNC_HAND_REC].
 * use_current (:40-55) is pnc_foot_use_current on the hand frame. */
int pnc_hand_init(pnc_workspace* ws, int32_t B, int32_t frame, const double* d_target_pos, const double* d_target_quat,
                  const double* d_keypoint, double* d_rec, void* stream);

/* update_hand_trajectory (:104-129; b_keypoint = 0) / update_keypoint_hand_trajectory (:131-178; b_keypoint = 1). */
int pnc_hand_eval(int32_t B, const double* d_rec, const double* d_start_time, const double* d_duration, double current_time,
                  int32_t b_keypoint, double* d_pos, double* d_vel, double* d_acc, double* d_quat, double* d_angvel,
                  double* d_angacc, void* stream);

/* FloatingBaseTrajectoryManager.initialize_floating_base_interpolation_trajectory
 * (pnc/wbc/manager/floating_base_trajectory_manager.py:35-54): record of the current CoM, the target,
 * the current orientation of `base_frame` and the exponential-coordinate error, [B, PNC_FB_REC]. */
int pnc_floating_base_init(pnc_workspace* ws, int32_t B, int32_t base_frame, const double* d_target_com,
                           const double* d_target_quat, double* d_rec, void* stream);

/* update_floating_base_desired, interpolation branch (:81-115; util/interpolation.py:8-32). */
int pnc_floating_base_eval(int32_t B, const double* d_rec, const double* d_start_time, const double* d_duration,
                           double current_time, double* d_pos, double* d_vel, double* d_acc, double* d_quat,
                           double* d_angvel, double* d_angacc, void* stream);

/* util.get_sinusoid_trajectory (util/util.py:98-107), the swaying branch (:76-80): d_mid [B,3],
 * amp / freq are 3 HOST doubles. */
int pnc_sinusoid_eval(int32_t B, const double* d_mid, const double* d_start_time, const double* amp3, const double* freq3,
                      double current_time, double* d_pos, double* d_vel, double* d_acc, void* stream);

/* AtlasStateEstimator._update_dcm (pnc/atlas_pnc/atlas_state_estimator.py:35-44) from the workspace CoM
 * and CoM velocity: d_dcm [B,3] is state (in: last dcm, out: new dcm), d_prev_dcm / d_dcm_vel [B,3] out.
 * As executed by the reference: dcm_vel = alpha (dcm - prev) / dt (its low-pass term is a no-op statement). */
int pnc_dcm_update(pnc_workspace* ws, int32_t B, double dt, double alpha, double* d_dcm, double* d_prev_dcm,
                   double* d_dcm_vel, void* stream);

/* Whole controller tick with HOST buffers: H2D of the inputs, update_system, solve,
 * D2H of the outputs, pipelined over chunks of the batch on the library's own streams;
 * synchronous (returns when every output has landed).  Pinned host buffers let the copies
 * overlap the kernels.  Any output may be NULL. */
int pnc_wbc_tick_host(const pnc_problem* p, pnc_workspace* ws, int32_t B, const double* h_base_pos,
                      const double* h_base_quat, const double* h_base_lin_vel,
                      const double* h_base_ang_vel, const double* h_joint_pos, const double* h_joint_vel,
                      const double* h_des, const double* h_w, const double* h_rf_z_max, double* h_trq,
                      double* h_acc, double* h_rf, int32_t* h_status, int32_t* h_iters,
                      uint64_t* h_active);

/* JointIntegrator state and constants (pnc/wbc/ihwbc/joint_integrator.py:6-82) for the command epilogue of
 * pnc_wbc_tick_host_ex: d_vel_state / d_pos_state [capacity, na] device, in/out; limits [na,2] device. */
typedef struct {
  double* d_vel_state;
  double* d_pos_state;
  const double* d_pos_limit;
  const double* d_vel_limit;
  double alpha_vel, alpha_pos, dt;
} pnc_integrator;

#define PNC_TICK_DES_ON_DEVICE 1 /* des / w / rf_z_max are DEVICE pointers (written by the on-device managers) */

/* The whole <Robot>Controller.get_command (atlas_controller.py:51-86, draco3_controller.py:51-103) as one
 * stream-ordered call: update_system, solve and -- when `integ` is given -- the command epilogue: actuated
 * torques / accelerations back to all joints through Sa[:, 6:]^T and JointIntegrator.integrate, returned as
 * h_joint_trq_cmd / h_joint_vel_cmd / h_joint_pos_cmd [B, na] (create_cmd_ordered_dict's three arrays).
 * With PNC_TICK_DES_ON_DEVICE in `flags` the task desireds, weights and rf_z_max are read from device memory
 * (the pnc_*_eval manager entry points wrote them there; `ready_stream` is the stream they were enqueued on, or
 * NULL when they are known to be complete) and only the robot state crosses the bus: 584 of the 1704 input
 * bytes per Atlas state.  Synchronous like pnc_wbc_tick_host. */
int pnc_wbc_tick_host_ex(const pnc_problem* p, pnc_workspace* ws, int32_t B, const double* h_base_pos,
                         const double* h_base_quat, const double* h_base_lin_vel, const double* h_base_ang_vel,
                         const double* h_joint_pos, const double* h_joint_vel, const double* des, const double* w,
                         const double* rf_z_max, int32_t flags, void* ready_stream, const pnc_integrator* integ,
                         double* h_trq, double* h_acc, double* h_rf, int32_t* h_status, int32_t* h_iters,
                         uint64_t* h_active, double* h_joint_trq_cmd, double* h_joint_vel_cmd,
                         double* h_joint_pos_cmd);

/* One host batch sharded over n devices from one process: problems[i] / workspaces[i] live on
 * device i's GPU, device i takes the contiguous states [i B / n, (i + 1) B / n) of every host array; all
 * devices are enqueued before any is waited for.  No collective: the shards are independent. */
int pnc_wbc_tick_host_multi(int32_t n_dev, const pnc_problem* const* problems, pnc_workspace* const* workspaces,
                            int32_t B, const double* h_base_pos, const double* h_base_quat,
                            const double* h_base_lin_vel, const double* h_base_ang_vel, const double* h_joint_pos,
                            const double* h_joint_vel, const double* h_des, const double* h_w,
                            const double* h_rf_z_max, double* h_trq, double* h_acc, double* h_rf,
                            int32_t* h_status, int32_t* h_iters, uint64_t* h_active);

/* number of kernels this library has launched since load (bench `gpu_launches`) */
int64_t pnc_launch_count(void);
/* DFMA-chain microbenchmark: measured fp64 FMA peak of the device in TFLOP/s */
int pnc_measure_fp64_peak(int device, double* tflops);
const char* pnc_version(void);
/* debug: while d_buf != NULL the solve kernel writes, for state `state`, the assembled QP pieces
 * (Hessian lower triangle n*n, g0 n, constraint rows nrow*n, their offsets nrow, cone matrix nu*6,
 * cone vector nu, unconstrained minimiser n) into d_buf.  Pass NULL to switch it off. */
int pnc_debug_set_dump(double* d_buf, int state);

/* debug: device pointer and per-state stride (doubles) of the setup records of the last pnc_ihwbc_solve on `ws`
 * (NULL / 0 when the problem took the generic kernel), and the record layout of the problem's register
 * shape: layout[0..9] = {n, p, ncol, npc, hw, RJ, RJB, RX, RS, REC} (offsets in doubles; see ihwbc_reg.cu). */
int pnc_debug_setup_record(const pnc_problem* p, pnc_workspace* ws, double** d_rec, int32_t* stride, int32_t* layout);

#ifdef __cplusplus
}
#endif
#endif /* PNC_B200_H */
