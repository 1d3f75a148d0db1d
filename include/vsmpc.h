/*
 * vsmpc.h — C-ABI of the B200-native batched variable-sampling ("multi-rate") MPC.
 *
 * Drop-in boundary for ONE path of ami-iit/paper_gorbani_2025_humanoids_multi-rate-mpc-ironcub:
 *   VariableSamplingMPC::{configure, update, solveMPC, get*}
 *   (src/flight-controller/momentum-based-linear-mpc-lib/include/variableSamplingMPC/variableSamplingMPC.h:15-41,
 *    .../include/IMPCProblem/IMPCProblem.h:35-118), as bound to Python in
 *   .../bindings/python/MPCPyBindings.cpp:22-90.
 *
 * A handle owns B independent MPC instances on one GPU.  All arrays crossing this boundary are
 * plain FP64 / int32 buffers owned by the caller; nothing is retained past a call except the
 * handle.  Every function returns 0 on success or a VSMPC_ERR_* code; vsmpc_last_error() gives
 * the message (the reference returns `bool` and logs through yError()).
 *
 * Layout conventions
 *   pack      : structure-of-arrays, double[VSMPC_PACK_DOUBLES][B]  (row = scalar, column = instance)
 *   outputs   : one row per instance, double[B][VSMPC_OUT_DOUBLES]
 *   solution  : one row per instance, double[B][n_var]   (reference variable order)
 */
#ifndef VSMPC_H
#define VSMPC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSMPC_NX 26        /* states:   VSconstant.h:9-16 */
#define VSMPC_NJ 8         /* N_JOINTS: VSconstant.h:6    */
#define VSMPC_NT 4         /* N_THRUSTS: VSconstant.h:7   */

/* ---- per-tick input pack: what update(QPInput&) reads (row offsets) -------- */
#define VSMPC_PK_WRB            0   /* 9  Robot::getBasePose().getRotation(), row-major          */
#define VSMPC_PK_OMEGA_WORLD    9   /* 3  getBaseVel().getAngularVec3()                           */
#define VSMPC_PK_RPY           12   /* 3  getRotation().asRPY()                                   */
#define VSMPC_PK_MASS          15   /* 1  getTotalMass() (already float-rounded, Robot.h:338)     */
#define VSMPC_PK_GRAVITY       16   /* 3  getGravity()                                            */
#define VSMPC_PK_MB            19   /* 36 getMassMatrix().block(0,0,6,6), row-major               */
#define VSMPC_PK_BASE_POS      55   /* 3  getBasePose().getPosition()                             */
#define VSMPC_PK_P_COM         58   /* 3  getPositionCoM()                                        */
#define VSMPC_PK_MOMENTUM_BODY 61   /* 6  getMomentum(true)                                       */
#define VSMPC_PK_AMOM_BODY     67   /* 24 getMatrixAmomJets(true), 6x4 row-major                  */
#define VSMPC_PK_JET_AXES      91   /* 12 getMatrixOfJetAxes()[i], world, 4x3                     */
#define VSMPC_PK_JET_ARMS     103   /* 12 getMatrixOfJetArms()[i], world, 4x3                     */
#define VSMPC_PK_J_REL_ANG    115   /* 96 getRelativeJacobianJetsBodyFrame()[i].bottomRows(3), controlled joints, 4x3x8 */
#define VSMPC_PK_J_JET_LIN    211   /* 96 getJacobian(jet).topRightCorner(3,nJ), controlled joints, 4x3x8 */
#define VSMPC_PK_J_COM        307   /* 24 getJacobianCoM().topRightCorner(3,nJ), controlled joints, 3x8   */
#define VSMPC_PK_THRUST       331   /* 4  getJetThrusts()                                         */
#define VSMPC_PK_THRUST_DOT_EST 335 /* 4  QPInput::getEstimatedThrustDot()                        */
#define VSMPC_PK_THRUST_DES   339   /* 4  QPInput::getThrustDesMPC()                              */
#define VSMPC_PK_THRUST_DOT_DES 343 /* 4  QPInput::getThrustDotDesMPC()                           */
#define VSMPC_PK_THROTTLE_PREV 347  /* 4  QPInput::getThrottleMPC()  [percent]                    */
#define VSMPC_PK_Q_CMD        351   /* 8  QPInput::getOutputQPJointsPosition()[controlled]        */
#define VSMPC_PACK_DOUBLES    359

/* ---- per-instance output row: the getters of variableSamplingMPC.cpp:88-227 ------------------ */
#define VSMPC_OUT_DELTA_Q       0   /* 8  m_deltaJointsPositionReference                          */
#define VSMPC_OUT_THROTTLE      8   /* 4  getThrottleReference()  [percent, clipped 0..100]       */
#define VSMPC_OUT_THRUST       12   /* 4  getThrustReference()                                    */
#define VSMPC_OUT_THRUST_DOT   16   /* 4  getThrustDotReference()                                 */
#define VSMPC_OUT_FINAL_STATE  20   /* 26 m_finalState -> getFinalCoMPosition/LinMom/RPY/AngMom   */
#define VSMPC_OUT_JOINTS_REF   46   /* 8  getJointsReferencePosition()[controlled] (accumulated)  */
#define VSMPC_OUT_DOUBLES      54

/* per-instance solver status; outputs of a non-solved instance are held (variableSamplingMPC.cpp:91) */
#define VSMPC_STATUS_SOLVED     0
#define VSMPC_STATUS_MAX_ITER   1   /* active-set iteration / working-set cap reached             */
#define VSMPC_STATUS_NUMERICAL  2   /* non-finite data or non-positive pivot                      */

#define VSMPC_OK                0
#define VSMPC_ERR_ARG           1
#define VSMPC_ERR_CUDA          2
#define VSMPC_ERR_UNSUPPORTED   3
#define VSMPC_ERR_STATE         4   /* call order violated (e.g. solve before configure)          */

#define VSMPC_MAX_ITER        128   /* max nIter supported by one handle                          */

typedef struct vsmpc_handle vsmpc_handle;

/* Problem configuration = group VS_MPC_CONFIG of src/config/vs_mcp_config.xml:5-44 plus the two
 * trajectory files it names (:34-40) as plain arrays (TrajectoryManager.cpp:67-140) and the jet
 * model constants (UT/src/JetModel.cpp:13-26). */
typedef struct vsmpc_config
{
    int n_iter;               /* nIter            */
    int n_iter_small;         /* nIterSmall       */
    int control_horizon;      /* controlHorizon   */
    double period_mpc;        /* periodMPC        */
    double period_large;      /* periodMPCLargeSteps */
    double period_small;      /* periodMPCSmallSteps */
    int use_jet_dynamic;      /* useJetDynamic    */
    int use_estimated_thrust; /* useEstimatedThrust */
    int joints_lambda_option; /* 0 = "unfiltered" (only option implemented), 1 = "constant" */
    double weight_com_pos[3];
    double weight_com_pos_error[3];
    double weight_lin_mom[3];
    double weight_rpy[3];
    double weight_rpy_error[3];
    double weight_ang_mom[3];
    double weight_delta_joint[VSMPC_NJ];
    double weight_throttle;
    double weight_initial_throttle;
    double weight_regularization_joint_pos;
    double throttle_min;
    double throttle_max;
    double jet_coeff[13];     /* m_u2TCoeff          */
    double jet_norm[4];       /* m_u2Tnormalization  */
    /* TRAJECTORY_MANAGER: alphaGravity[alpha_len] at alpha_fps */
    const double* alpha_gravity;
    int alpha_len;
    int alpha_fps;
    /* POSITION_TRAJECTORY: 3 x traj_len each, sample-major (element [3*s + axis]), at traj_fps */
    const double* position_com;
    const double* velocity_com;
    const double* rpy;
    const double* rpy_dot;
    int traj_len;
    int traj_fps;
    int solver;               /* 0 = default (condensed-throttle Riccati kernel: two warps per instance at the reference
                                 horizon, 1 + G warps for long horizons, generic kernel beyond ~37 throttle blocks);
                                 1 = generic dense variant; 2 = structured one-warp kernel (cross-checks) */
    /* Optional joint-limit rows (BASELINE configs[4] "per-instance constraint sets"): the reference's JointPositionConstraint
     * (constraintsVSMPC.cpp:388-468; present but not registered in the shipped problem, its parameters jointPos_max /
     * jointPos_min [degrees, :421-423] are absent from the XML).  When enabled, 8 * nIter rows are appended after the throttle
     * rows (:391), block i < controlHorizon bounding dq_i — the displacement from the commanded posture acting on knot i — by
     * jointPos_min - q_cmd <= dq_i <= jointPos_max - q_cmd (:450-453), per block.  The reference's m_firstIteriation slip
     * (:440-449, identity rows only for block 0) is fixed, not ported.  Default solver only.  The QP kernels carry the joint boxes in
     * their own working set (clamped increments inside the 8 x 8 eliminations, one more factorisation per change of the
     * working set, vsmpc_get_counts reports them; reference horizon and up to twice its knot count); an instance whose working
     * set does not settle goes to the fallback kernel (vsmpc_set_fallback), whose size limit — 4 x throttle blocks +
     * 8 x controlHorizon <= 256 — is also the limit of the rows: vsmpc_create refuses them beyond. */
    int use_joint_limits;
    double joint_pos_min_deg[VSMPC_NJ];   /* jointPos_min of the controlled joints */
    double joint_pos_max_deg[VSMPC_NJ];   /* jointPos_max */
} vsmpc_config;

/* lifecycle ------------------------------------------------------------------------------------ */
int vsmpc_create(const vsmpc_config* cfg, int n_instances, int device, vsmpc_handle** out);
int vsmpc_destroy(vsmpc_handle* h);
const char* vsmpc_last_error(const vsmpc_handle* h);
/* launch everything on this CUDA stream (a cudaStream_t; NULL = the handle's own stream) */
int vsmpc_set_stream(vsmpc_handle* h, void* cuda_stream);

int vsmpc_n_var(const vsmpc_handle* h);          /* IMPCProblem::getNOptimizationVariables */
int vsmpc_n_constraints(const vsmpc_handle* h);  /* IMPCProblem::getNConstraints           */
int vsmpc_n_instances(const vsmpc_handle* h);

/* IMPCProblem::configure (IMPCProblem.cpp:3-148): initialise every instance's persistent state
 * from the robot state at configure time and run "tick 0" of all counters / trajectory cursors.
 * joint_pos_sel: double[8][B], Robot::getJointPos() of the controlled joints (costsVSMPC.cpp:540-550,
 * variableSamplingMPC.cpp:60).  phase0: optional int[B] number of extra ticks (0..ratio-1) by which
 * the 20-tick phase counters start ahead (NULL = reference behaviour). */
int vsmpc_configure(vsmpc_handle* h, const double* pack_host, const double* joint_pos_sel_host,
                    const int* phase0_host);

/* Per-instance model parameters (BASELINE.json configs[4]: jet time constants / thrust limits sweeps): one SoA
 * buffer double[VSMPC_INSTANCE_PARAM_DOUBLES][B] replacing the handle-wide jet_coeff / jet_norm / throttle_min /
 * throttle_max of vsmpc_config (JetModel.cpp:13-26, vs_mcp_config.xml throttleMin/Max) for every later call;
 * NULL returns to the handle-wide values.  Mass and inertia are per instance already (they are in the pack). */
#define VSMPC_IP_JET_COEFF      0   /* 13 */
#define VSMPC_IP_JET_NORM      13   /* 4  thrust mean, thrust std, throttle mean, throttle std */
#define VSMPC_IP_THROTTLE_MIN  17   /* 1  [percent] */
#define VSMPC_IP_THROTTLE_MAX  18   /* 1  [percent] */
#define VSMPC_INSTANCE_PARAM_DOUBLES 19
int vsmpc_set_instance_params(vsmpc_handle* h, const double* instance_params_host);

/* per-instance joint limits [rad] of the controlled joints, SoA double[8][B] each, replacing the handle-wide jointPos_min /
 * jointPos_max for every later call (NULL, NULL: back to the handle-wide values); needs use_joint_limits at create */
int vsmpc_set_joint_limits(vsmpc_handle* h, const double* q_min_host, const double* q_max_host);

/* Warm start — the counterpart of the reference running OSQP with setWarmStart(true) (IMPCProblem.cpp:140).  Long horizons
 * (more than 6 throttle blocks): the active set of the reduced throttle QP starts from the working set the instance ended its
 * previous solve with (device-resident; all-lower vertex after vsmpc_configure).  Joint-limit rows (any horizon that has them):
 * the working set of the joint boxes starts from the one the previous solve ended with (empty after vsmpc_configure and after
 * vsmpc_set_joint_limits).  Any guess gives the same minimiser; enable = 0 starts every solve cold (tests, A/B timing).
 * Default: on. */
int vsmpc_set_warm_start(vsmpc_handle* h, int enable);
/* tests: overwrite the stored working sets, signed char[B][4 * throttle blocks] (+1 upper bound, -1 lower bound, 0 free) */
int vsmpc_debug_set_working_set(vsmpc_handle* h, const signed char* working_set_host);

/* Page-locked host memory (cudaHostAlloc) for the SoA buffers a caller hands to vsmpc_set_state / vsmpc_get_output_async, so that
 * a host program that does not link the CUDA runtime itself (include/vsmpc_adapter.hpp: vsmpc::PackBatch) gets copies that are
 * asynchronous and run at the full PCIe rate.  Not tied to a handle; free with vsmpc_host_free. */
int vsmpc_host_alloc(size_t bytes, void** out);
int vsmpc_host_free(void* p);

/* IMPCProblem::update (IMPCProblem.cpp:150-194): copy the pack H2D and run the linearise kernel.
 * Asynchronous: the call returns once the copy and the kernel are ENQUEUED (copy on the handle's own copy stream into one of
 * two staging buffers, so that it overlaps the QP kernel of the tick before).  From pinned host memory the copy itself is
 * still in flight when the call returns: pack_host must stay untouched until vsmpc_wait / vsmpc_wait_output / a blocking
 * getter of THIS tick has returned (pageable memory is staged by the driver before the call returns).  A handle is driven by
 * one host thread; vsmpc_set_stream between vsmpc_set_state and vsmpc_solve keeps the order (the new stream waits for the
 * old one).  The linearise kernel of this call runs on a stream of the handle's own, concurrently with the QP kernel of the
 * tick before (vsmpc_solve_async of that tick may still be running): every later call of the handle that reads what it
 * produced waits for it. */
int vsmpc_set_state(vsmpc_handle* h, const double* pack_host);
/* same, pack already resident on the handle's GPU */
int vsmpc_set_state_device(vsmpc_handle* h, const double* pack_dev);

/* the same two calls for a COLUMN WINDOW of a larger SoA batch: row r of this handle's instances starts at
 * pack_host + r * row_stride (joint_pos_sel likewise; row_stride >= n_instances, in doubles).  This is how one host batch is
 * split over several devices without repacking (vsmpc_multi_* below). */
int vsmpc_configure_strided(vsmpc_handle* h, const double* pack_host, const double* joint_pos_sel_host, const int* phase0_host,
                            size_t row_stride);
int vsmpc_set_state_strided(vsmpc_handle* h, const double* pack_host, size_t row_stride);

/* VariableSamplingMPC::solveMPC (variableSamplingMPC.cpp:88-112): structured QP solve + output
 * extraction + joint accumulator.  vsmpc_solve blocks; _async + vsmpc_wait split it. */
int vsmpc_solve(vsmpc_handle* h);
int vsmpc_solve_async(vsmpc_handle* h);
int vsmpc_wait(vsmpc_handle* h);

/* getters: out_rows double[B][VSMPC_OUT_DOUBLES], status int[B] (either may be NULL) */
int vsmpc_get_output(vsmpc_handle* h, double* out_rows_host, int* status_host);
/* pipelined variant: enqueue the device->host copies behind the solve on the handle's stream and return at once;
 * vsmpc_wait_output(h, ticket) blocks until that copy has landed.  Two tickets are in flight at most, so a caller can
 * enqueue tick j+1 (vsmpc_set_state stages its pack on a separate copy stream) before waiting for tick j. */
int vsmpc_get_output_async(vsmpc_handle* h, double* out_rows_host, int* status_host, int* ticket);
int vsmpc_wait_output(vsmpc_handle* h, int ticket);
/* device pointers of the same buffers (valid for the handle's lifetime) */
int vsmpc_get_output_device(vsmpc_handle* h, double** out_rows_dev, int** status_dev);
/* IMPCProblem::getSolution: double[B][n_var].  The full 588-vector is optional output: the controller
 * only consumes the getters above, so the structured kernel skips it unless enabled here (when
 * disabled vsmpc_get_full_solution returns VSMPC_ERR_STATE). */
int vsmpc_set_full_solution(vsmpc_handle* h, int enable);
int vsmpc_get_full_solution(vsmpc_handle* h, double* z_host);

/* ---- inner seams, separately callable for parity tests -------------------------- */
/* K1 alone — seam (ii), DynamicTemplateVariableSampling::updateInitialStates + the four getters
 * (MPC/include/IMPCProblem/systemDynamic.h:39-65) plus a8-a14: copy the pack H2D and run the linearise kernel, nothing
 * else (same work as vsmpc_set_state, whose name follows the outer surface).  Read its product with
 * vsmpc_get_dynamics / vsmpc_get_qp_vectors. */
int vsmpc_linearise(vsmpc_handle* h, const double* pack_host);
/* K2 alone — seam (iii), the OsqpEigen calls of IMPCProblem::solve (IMPCProblem.cpp:225-279,296): solve the QP the last
 * vsmpc_linearise / vsmpc_set_state left on the device and extract the outputs; blocks (same work as vsmpc_solve). */
int vsmpc_solve_qp(vsmpc_handle* h);
/* dense expansions of what the linearise kernel produced for the current tick:
 *   A double[B][26*26], BJ double[B][26*8], BT double[B][26*4], c double[B][26] (row-major),
 *   dt double[n_iter]  — SystemDynamicVS::get{A,BJoints,BThrottle}Matrix/getCVector + dt grid      */
int vsmpc_get_dynamics(vsmpc_handle* h, double* A, double* BJ, double* BT, double* c, double* dt);
/* IMPCProblem::getGradient / getLowerBound / getUpperBound: q double[B][n_var], l,u double[B][n_con] */
int vsmpc_get_qp_vectors(vsmpc_handle* h, double* q, double* l, double* u);
/* executed Riccati factorisations / back-solves per instance in the last solve (int[B] each) */
int vsmpc_get_counts(vsmpc_handle* h, int* n_factor, int* n_solve);
/* exchange pivots of the reduced throttle QP executed per instance in the last solve (inverse of the reduced Hessian +
 * dual active-set iterations; int[B]; the condensed kernels report it, the cross-check kernels leave 0) */
int vsmpc_get_pivot_counts(vsmpc_handle* h, int* n_pivot);

/* ---- the QPInput fields the path WRITES (what a caller of update(QPInput&) reads back afterwards) ---------------
 * ReferenceTrackingCost::computeHessianAndGradient publishes the tracked references on every window shift
 * (costsVSMPC.cpp:155-160: QPInput::setPosCoMReference / setRPYReference / setMomentumReference) and
 * LinearMomentumDynamicVS publishes the gravity-compensation factor every tick (systemDynamicsVSMPC.cpp:310:
 * QPInput::setAlphaGravity).  refs: double[B][VSMPC_REF_DOUBLES], valid after vsmpc_configure / vsmpc_set_state. */
#define VSMPC_REF_ALPHA_GRAVITY  0   /* 1  QPInput::getAlphaGravity()      */
#define VSMPC_REF_POS_COM        1   /* 3  QPInput::getPosCoMReference()   */
#define VSMPC_REF_RPY            4   /* 3  QPInput::getRPYReference()      */
#define VSMPC_REF_MOMENTUM       7   /* 6  QPInput::getMomentumReference() (linear, angular) */
#define VSMPC_REF_DOUBLES       13
int vsmpc_get_references(vsmpc_handle* h, double* refs_host);
/* IMPCProblem::getHessian (IMPCProblem.h:88): dense n_var x n_var, row-major, of one instance (constant after configure,
 * IMPCProblem.cpp:152-175) */
int vsmpc_get_hessian(vsmpc_handle* h, int instance, double* P_host);
/* IMPCProblem::getLinearConstraintMatrix (IMPCProblem.h:100): dense n_con x n_var, row-major, of one instance for the
 * current tick, in the reference's row order (dynamics, initial state, throttle; variableSamplingMPC.cpp:77-84) */
int vsmpc_get_constraint_matrix(vsmpc_handle* h, int instance, double* A_host);
/* Fallback QP kernel of the default solver.  The Riccati recursion of the condensed kernels breaks down (status 2 on finite
 * data) when the open-loop transition of the linearised model expands strongly — |omega_B| >~ 30 rad/s, a lost vehicle —
 * although the QP stays well posed (the reference's OSQP returns its minimiser).  Such instances are solved by a pivoted LU
 * of the KKT system (csrc/vsmpc_qp_fallback.cu) launched behind the QP kernel.  mode 0: off (status 2 holds the outputs),
 * 1: on (default), 2: EVERY instance goes through the fallback kernel (parity tests of the fallback itself). */
int vsmpc_set_fallback(vsmpc_handle* h, int mode);
/* test hook: overwrite the two 20-tick phase counters (ReferenceTrackingCost::m_counter,
 * ThrottleConstraint::m_counter) of every instance; -1 leaves a counter unchanged */
int vsmpc_debug_set_counters(vsmpc_handle* h, int ref_counter, int throttle_counter);

/* ---- device-resident closed loop (the loop body of src/variable_sampling_mpc.py:106-161) --------
 * A SURROGATE plant replaces MuJoCo (not available): it integrates the MPC's own nonlinear model
 * (centroidal momentum driven by the four jets, gravity scaled by the take-off factor alpha_g the MPC publishes — the
 * ground carries the rest —, and the second-order jet model of
 * src/mujoco_lib/jet_kalman_filter.py:30-45); the jet frames follow the (position-controlled) arm joints through
 * first-order kinematics about the posture q0 with frozen relative Jacobians; n_sub steps of dt_sim per
 * controller tick, and rebuilds the pack from the plant state on the device.  Per tick: plant -> pack ->
 * linearise kernel -> QP kernel -> feedback (:124-131); no host round trip. */
#define VSMPC_PS_P_COM            0   /* 3  CoM position (world)                                         */
#define VSMPC_PS_LIN_MOM_WORLD    3   /* 3  linear momentum (world)                                      */
#define VSMPC_PS_RPY              6   /* 3  base roll / pitch / yaw                                      */
#define VSMPC_PS_ANG_MOM_BODY     9   /* 3  angular momentum about the CoM (body)                        */
#define VSMPC_PS_THRUST          12   /* 4  jet thrusts                                                  */
#define VSMPC_PS_THRUST_DOT      16   /* 4  jet thrust rates                                             */
#define VSMPC_PS_THROTTLE        20   /* 4  throttle command in effect [percent]  (QPInput::setThrottleMPC)     */
#define VSMPC_PS_THRUST_DES      24   /* 4  QPInput::setThrustDesMPC                                     */
#define VSMPC_PS_THRUST_DOT_DES  28   /* 4  QPInput::setThrustDotDesMPC                                  */
#define VSMPC_PS_Q_CMD           32   /* 8  QPInput::setOutputQPJointsPosition (controlled joints)        */
#define VSMPC_PS_THRUST_NN       40   /* 4  thrust state of the neural jet plant (float32 values), jet-NN mode only      */
#define VSMPC_PS_EKF_P           44   /* 16 per-jet 2x2 EKF covariance, row-major, jet-NN mode only                      */
#define VSMPC_PLANT_STATE_DOUBLES 60
#define VSMPC_PP_MASS             0   /* 1  total mass (float-rounded like Robot::m_totalMass)           */
#define VSMPC_PP_INERTIA_BODY     1   /* 9  locked inertia about the CoM, body frame, row-major          */
#define VSMPC_PP_THRUST_DISTURBANCE 10 /* 4 constant thrust disturbance added in the plant [N]            */
#define VSMPC_PLANT_PARAM_DOUBLES 14
/* recorded per instance and recorded tick: p_com(3) rpy(3) thrust(4) throttle(4) status(1) pad(1) */
#define VSMPC_ROLLOUT_REC_DOUBLES 16

typedef struct vsmpc_plant_model
{
    double com_from_base_body[3];
    double jet_pos_body[12];      /* 4x3, jet positions relative to the CoM, body frame                  */
    double jet_axes_body[12];     /* 4x3, thrust directions, body frame                                  */
    double J_rel_ang_body[96];    /* 4x3x8 angular relative Jacobians of the jets, controlled joints     */
    double J_jet_lin_body[96];    /* 4x3x8 linear Jacobians of the jet frames (body frame)               */
    double J_com_body[24];        /* 3x8  CoM Jacobian, joint part (body frame)                          */
    double gravity[3];
    double q0[8];                 /* controlled-joint posture the frozen kinematics above refer to                */
    double dt_sim;                /* plant step (reference: MuJoCo timestep 1 ms)                        */
    int n_sub;                    /* plant steps per controller tick (reference: periodMPC / dt = 5)     */
} vsmpc_plant_model;

/* upload the plant, build the first pack on the device and run configure (tick 0) from it.
 * plant_state: double[VSMPC_PLANT_STATE_DOUBLES][B]; plant_param: double[VSMPC_PLANT_PARAM_DOUBLES][B];
 * joint_pos_sel: double[8][B]; phase0: int[B] or NULL (as in vsmpc_configure) */
int vsmpc_rollout_init(vsmpc_handle* h, const vsmpc_plant_model* model, const double* plant_state_host,
                       const double* plant_param_host, const double* joint_pos_sel_host, const int* phase0_host);
/* n_ticks closed-loop controller ticks.  record_every > 0: every record_every-th tick (after the plant step) a
 * record row per instance is written; rec_host receives double[n_ticks / record_every][B][VSMPC_ROLLOUT_REC_DOUBLES].
 * With use_graph != 0 the three kernels of a tick are captured once in a CUDA graph and replayed. */
int vsmpc_rollout_run(vsmpc_handle* h, int n_ticks, int record_every, double* rec_host, int use_graph);
int vsmpc_rollout_get_state(vsmpc_handle* h, double* plant_state_host);
/* Jet plant + estimator of the reference simulator instead of the second-order jet model: per 1 ms plant
 * step and jet, the neural jet model (src/mujoco_lib/nn_jet_model.py:21-30,86-109 — an LSTM cell evaluated from a zero
 * state, float32) advances its thrust state under the commanded throttle, and the per-jet EKF
 * (src/mujoco_lib/jet_kalman_filter.py:56-65) fuses it with the second-order model; the EKF estimate is the thrust applied
 * to the plant and reported to the MPC (ironcub_mujoco_simulator.py:129-133).  w_ih float[320][2], b_ih / b_hh float[320],
 * fc_w float[80], fc_b float[1], norm = {thrust mean, thrust std, throttle mean, throttle std} of the checkpoint,
 * ekf_R / ekf_Q double[4] (2x2 row-major).  Call before vsmpc_rollout_init; w_ih == NULL switches back. */
int vsmpc_rollout_set_jet_nn(vsmpc_handle* h, const float* w_ih, const float* b_ih, const float* b_hh, const float* fc_w,
                             const float* fc_b, const double* norm, const double* ekf_R, const double* ekf_Q);
/* parity seam of the neural jet plant: one NeuralJetModel.get_state step (nn_jet_model.py:21-30) for n_groups x 4
 * (thrust [N], throttle [percent]) pairs on the device; float[n_groups][4] each */
int vsmpc_jet_nn_eval(vsmpc_handle* h, int n_groups, double dt, const float* T_host, const float* throttle_host,
                      float* T_next_host, float* T_dot_host);
/* the pack the plant built for the next tick (double[VSMPC_PACK_DOUBLES][B]) — parity tests */
int vsmpc_rollout_get_pack(vsmpc_handle* h, double* pack_host);

/* ---- batched reduced kinematics: the part of Robot::setState the path consumes ----------------------------
 * UT/src/Robot.cpp:212-278,325-332 computes, through iDynTree, the kinematic rows of the pack (base block of the mass matrix,
 * CoM, centroidal momentum, jet axes / arms, A_mom, relative / free-floating / CoM Jacobians of the controlled joints).  Here
 * the same rows are computed ON THE DEVICE for the whole batch from a kinematic tree given as arrays, so that the host sends
 * the 18 + 2 n_dof + 28 doubles of a robot state per instance (92 for the 23-joint robot) instead of the 359-double pack.
 * Conventions: iDynTree's MIXED velocity representation (oracle/kinematics_oracle.py states them); parity against iDynTree
 * itself is unpinned — neither the library nor the iRonCub URDF is available; the oracle is checked against closed forms and
 * finite differences of its own forward kinematics. */
#define VSMPC_KIN_MAX_LINKS 32
typedef struct vsmpc_kin_model
{
    int n_links;                                /* link 0 = floating base; every other link hangs on one revolute joint   */
    int n_dof;                                  /* length of the joint vector (axesList order, robot.toml:3-27)           */
    int parent[VSMPC_KIN_MAX_LINKS];            /* parent[0] = -1, parent[l] < l                                          */
    int dof[VSMPC_KIN_MAX_LINKS];               /* position of link l's joint in the joint vector, dof[0] = -1            */
    double R0[VSMPC_KIN_MAX_LINKS][9];          /* rotation parent link frame -> joint frame at q = 0, row-major          */
    double p0[VSMPC_KIN_MAX_LINKS][3];          /* joint origin in the parent link frame                                  */
    double axis[VSMPC_KIN_MAX_LINKS][3];        /* unit joint axis in the joint (= child link) frame                      */
    double mass[VSMPC_KIN_MAX_LINKS];
    double com[VSMPC_KIN_MAX_LINKS][3];         /* link CoM in the link frame                                             */
    double inertia[VSMPC_KIN_MAX_LINKS][9];     /* link inertia about its CoM, link axes, row-major                       */
    int jet_link[VSMPC_NT];                     /* link carrying jet i                                                    */
    double jet_pos[VSMPC_NT][3];                /* jet frame origin in that link's frame                                  */
    double jet_axis[VSMPC_NT][3];               /* thrust axis in that link's frame (Robot::m_jetsAxesLocalFrames)        */
    double delta_com[3];                        /* Robot::m_deltaCoM (base axes), Robot.cpp:253-255                       */
    double gravity[3];
    int sel[VSMPC_NJ];                          /* controlled joints as positions in the joint vector                     */
} vsmpc_kin_model;
/* robot state of one tick, SoA double[vsmpc_kin_state_doubles][B]: rows in this order */
#define VSMPC_KS_WRB          0   /* 9  base rotation, row-major                    */
#define VSMPC_KS_BASE_POS     9   /* 3                                              */
#define VSMPC_KS_BASE_LIN_VEL 12  /* 3  velocity of the base origin, world axes     */
#define VSMPC_KS_OMEGA_WORLD  15  /* 3                                              */
#define VSMPC_KS_Q            18  /* n_dof joint positions, then n_dof joint velocities, then the 28 QPInput rows of the pack
                                     (VSMPC_PK_THRUST .. VSMPC_PK_Q_CMD, same order) */
int vsmpc_set_kin_model(vsmpc_handle* h, const vsmpc_kin_model* model);
int vsmpc_kin_state_doubles(const vsmpc_handle* h);          /* 18 + 2 n_dof + 28, -1 before vsmpc_set_kin_model */
/* vsmpc_configure / vsmpc_set_state with the pack built on the device by the kinematics kernel (joint_pos_sel = q[sel]) */
int vsmpc_configure_kinematics(vsmpc_handle* h, const double* kin_state_host, const int* phase0_host);
int vsmpc_set_state_kinematics(vsmpc_handle* h, const double* kin_state_host);
/* the pack of the last vsmpc_*_kinematics call (double[VSMPC_PACK_DOUBLES][B]) — parity tests */
int vsmpc_get_kinematics_pack(vsmpc_handle* h, double* pack_host);

/* ---- one process, several GPUs (SURVEY §8b "vsmpc_create(cfg, n_instances, n_gpus, ...)", §8e) -----------------------------
 * B instances sharded in contiguous ranges [g B / G, (g + 1) B / G) over G devices (devices == NULL: 0 .. G-1; an entry may
 * repeat, e.g. {0, 0} puts two shards on one GPU), one vsmpc_handle with its own streams per shard, driven by ONE host
 * thread: every call below enqueues on all shards before it waits on any, and there is no inter-GPU traffic — the only
 * gather is vsmpc_multi_get_output, where every device copies its own rows into the caller's [B][54] array.  All host
 * arrays have the single-GPU layouts for the WHOLE batch. */
typedef struct vsmpc_multi vsmpc_multi;
int vsmpc_create_multi(const vsmpc_config* cfg, int n_instances, int n_gpus, const int* devices, vsmpc_multi** out);
int vsmpc_multi_destroy(vsmpc_multi* m);
const char* vsmpc_multi_last_error(const vsmpc_multi* m);
int vsmpc_multi_n_shards(const vsmpc_multi* m);
int vsmpc_multi_n_instances(const vsmpc_multi* m);
/* instance range and per-device handle of one shard (handle == NULL for a shard without instances) */
int vsmpc_multi_shard(const vsmpc_multi* m, int shard, int* first, int* count, vsmpc_handle** handle);
int vsmpc_multi_configure(vsmpc_multi* m, const double* pack_host, const double* joint_pos_sel_host, const int* phase0_host);
int vsmpc_multi_set_instance_params(vsmpc_multi* m, const double* instance_params_host);
int vsmpc_multi_set_state(vsmpc_multi* m, const double* pack_host);
int vsmpc_multi_solve(vsmpc_multi* m);
int vsmpc_multi_solve_async(vsmpc_multi* m);
int vsmpc_multi_wait(vsmpc_multi* m);
int vsmpc_multi_get_output(vsmpc_multi* m, double* out_rows_host, int* status_host);
int vsmpc_multi_set_full_solution(vsmpc_multi* m, int enable);
int vsmpc_multi_get_full_solution(vsmpc_multi* m, double* z_host);

/* development hook: per-instance clock64() stamps of the condensed kernel's phases of the last launch (long long
 * [n][8]); returns VSMPC_ERR_UNSUPPORTED unless the library was built with -DVSMPC_PHASE_CLOCKS */
int vsmpc_debug_phase_clocks(long long* clocks_host, int n_instances);

/* measured FP64 throughput of `device` in TFLOP/s: kind 0 = DFMA on the CUDA cores, kind 1 = DMMA
 * (mma.sync.m8n8k4.f64).  Used as the roofline denominator of the QP kernel (bench.py). */
int vsmpc_microbench_fp64(int device, int kind, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* VSMPC_H */
