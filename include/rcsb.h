/*
 * rcsb.h — C ABI of the batched B200 path of rcs_b200.
 *
 * The reference (HRI-EU/Rcs) has no FFI seam on this path: its boundary is the extern "C"
 * API of libRcsCore itself, one state per call on one RcsGraph. The entry points below are
 * the batched equivalents; each evaluates N independent joint states in hand-written
 * sm_100a CUDA kernels and returns, per state, exactly what the cited reference function
 * leaves in the graph / output matrices for that state. Layouts are the reference's
 * (HTr = org[3] + row-major rot[3][3]; matrices row-major), with the state index slowest.
 *
 *   rcsb_fk              RcsGraph_setState / RcsGraph_computeForwardKinematics
 *                        src/RcsCore/Rcs_graph.c:231, :301 (kernel: :61-226)
 *   rcsb_fk_jacobians    + RcsGraph_bodyPointJacobian  src/RcsCore/Rcs_kinematics.c:200
 *                        + RcsGraph_rotationJacobian   src/RcsCore/Rcs_kinematics.c:416
 *   rcsb_fk_jacobians_mass   the same + M(q) (RcsGraph_computeMassMatrix, Rcs_dynamics.c:610) in
 *                        one launch: BASELINE.json's metric "FK+Jacobian+M(q) states/sec"
 *   rcsb_task_jacobians  RcsGraph_3dPosJacobian / 3dOmegaJacobian  Rcs_kinematics.c:566, :789
 *   rcsb_kinetic_terms   RcsGraph_computeKineticTerms  src/RcsCore/Rcs_dynamics.c:439
 *   rcsb_cog, rcsb_kinetic_terms_cog   RcsGraph_COG / COGJacobian  Rcs_kinematics.c:904, :973
 *   rcsb_direct_dynamics Rcs_directDynamics            src/RcsCore/Rcs_dynamics.c:673
 *   rcsb_wholebody       RcsGraph_setState + ControllerBase::computeJ + RcsGraph_computeMassMatrix
 *   rcsb_ik_rmr          ControllerBase::computeDX / computeJ / computeJointlimitGradient
 *                        (src/RcsCore/ControllerBase.cpp:776, :896, :1332) +
 *                        IkSolverRMR::solveRightInverse (src/RcsCore/IkSolverRMR.cpp:403)
 *                        iterated as in examples/ExampleIK.cpp:154-177
 *   rcsb_ik_rmr_ex       the same with activations, lambda vector, dq_ts / dq_ns, left inverse with
 *                        TaskSpaceBlender weights (IkSolverRMR.cpp:205-272, :385-596)
 *   rcsb_ik_tasks_rmr    the same over XYZ / ABC / XYZABC with refBdy / refFrame, COGX/Y/Z, Joint, POLAR
 *   rcsb_ik_prio_rmr     IkSolverPrioRMR::solve        src/RcsCore/IkSolverPrioRMR.cpp:178
 *   rcsb_ik_constraint_rmr   IkSolverConstraintRMR::solveRightInverse  IkSolverConstraintRMR.cpp:156
 *
 * Conventions
 *  - Every call returns 0 on success, a negative RCSB_E* code otherwise; the message is
 *    available from rcsb_last_error() (thread-local). Nothing here calls exit().
 *  - q / q_dot rows have qdim entries, qdim == dof (full state, joint order = jointIndex)
 *    or qdim == nJ (IK state, order = jacobiIndex; constrained joints then take the value
 *    the graph held when the model was created). Kinematically coupled (slave) joints are
 *    overwritten from their master, as RcsGraph_setState does (Rcs_graph.c:2756).
 *  - Array arguments are DEVICE pointers unless RCSB_HOST_PTRS is set; with RCSB_HOST_PTRS
 *    they are host pointers (pinned memory from rcsb_host_alloc() is fastest) and the call
 *    copies inputs in and results out, pipelined with the kernels, and returns when the
 *    results are in host memory. Selector arrays (bodySel, effBody, effPt) are always small
 *    host arrays.
 *  - stream is a cudaStream_t (NULL = default stream). Device-pointer calls are
 *    asynchronous on that stream. A model is immutable and may be shared by threads; the
 *    kinetic-terms, COG, direct-dynamics and whole-body calls keep one per-state record workspace
 *    per model, so device-pointer calls of those four on ONE model must be issued on one stream
 *    at a time (host-pointer calls serialise themselves).
 *  - There is no CPU fallback: without a CUDA device every compute call fails with
 *    RCSB_ENODEVICE.
 */
#ifndef RCSB_H
#define RCSB_H

#include "rcsb_graph.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RCSB_OK 0
#define RCSB_EINVAL (-1)     /* bad argument / shape */
#define RCSB_ENODEVICE (-2)  /* no CUDA device / driver */
#define RCSB_ECUDA (-3)      /* CUDA runtime error */
#define RCSB_ENOMEM (-4)
#define RCSB_EUNSUPPORTED (-5)

#define RCSB_HOST_PTRS 1u    /* array arguments are host pointers */
#define RCSB_IK_LEFT_INVERSE 2u   /* rcsb_ik_rmr: IkSolverRMR::solveLeftInverse (IkSolverRMR.cpp:205),
                                     the nq x nq left pseudo-inverse, instead of solveRightInverse */

/* task kinds of rcsb_ik_rmr, named after the reference's controlVariable strings
 * (src/RcsCore/TaskPosition3D.cpp:49, TaskEuler3D.cpp:52) */
#define RCSB_TASK_XYZ 1
#define RCSB_TASK_ABC 2
#define RCSB_TASK_XYZABC 3   /* TaskPose6D (src/RcsCore/TaskPose6D.cpp): XYZ followed by ABC of the same
                                effector, six entries of x_des */

typedef struct RcsbModel RcsbModel;

typedef struct
{
  int kind;         /* RCSB_TASK_* */
  int effector;     /* depth-first body index of the effector */
} RcsbTask;

/* ---- model ---------------------------------------------------------------------------- */

/* Flattens the graph (depth-first structure-of-arrays: parents, joint types and axes,
 * relative transforms, mass properties, coupling rules, joint metric) and uploads it to
 * `device`. The graph is not referenced afterwards. */
int rcsb_model_create(const RcsGraph* graph, int device, RcsbModel** model);
void rcsb_model_destroy(RcsbModel* model);
int rcsb_model_dof(const RcsbModel* model);
int rcsb_model_nj(const RcsbModel* model);
int rcsb_model_nbodies(const RcsbModel* model);
int rcsb_model_device(const RcsbModel* model);

/* ---- forward kinematics ---------------------------------------------------------------- */

/*
 * q [N x qdim], qd [N x qdim] or NULL.
 * bodySel: nSel depth-first body indices, or NULL with nSel == 0 for "all bodies".
 * Outputs (each may be NULL): A_BI [N x nSel x 12], A_JI [N x dof x 12] (all joints, by
 * jointIndex), xdot / omega [N x nSel x 3] (need qd), qOut [N x dof] the full state after
 * expansion and coupling.
 */
int rcsb_fk(const RcsbModel* model, long N, int qdim, const double* q, const double* qd,
            int nSel, const int* bodySel, double* A_BI, double* A_JI, double* xdot,
            double* omega, double* qOut, unsigned flags, void* stream);

/*
 * Forward kinematics plus, for each of nEff effectors, the translation Jacobian of the
 * body-fixed point effPt[e] (NULL = body origins) and the rotation Jacobian of the body:
 * JT, JR [N x nEff x 3 x nJ], world frame, columns of non-ancestor or constrained joints
 * exactly 0. A_BI as in rcsb_fk (NULL to skip the frames).
 */
int rcsb_fk_jacobians(const RcsbModel* model, long N, int qdim, const double* q, int nSel,
                      const int* bodySel, double* A_BI, int nEff, const int* effBody,
                      const double* effPt, double* JT, double* JR, unsigned flags,
                      void* stream);

/*
 * rcsb_fk_jacobians plus the joint-space mass matrix M [N x nJ x nJ]
 * (RcsGraph_computeMassMatrix, src/RcsCore/Rcs_dynamics.c:610) from the same tree walk: one
 * kernel launch for "FK + Jacobians + M(q)". Fused for models with nJ <= 8 and symmetric
 * inertia tensors (the 7-DoF arm); other models run rcsb_fk_jacobians and rcsb_kinetic_terms
 * one after the other. M == NULL is rcsb_fk_jacobians.
 */
int rcsb_fk_jacobians_mass(const RcsbModel* model, long N, int qdim, const double* q, int nSel,
                           const int* bodySel, double* A_BI, int nEff, const int* effBody,
                           const double* effPt, double* JT, double* JR, double* M, unsigned flags,
                           void* stream);

/*
 * Task Jacobians relative to a (possibly articulated) reference body / frame:
 * RcsGraph_3dPosJacobian (RCSB_JAC_POS) and RcsGraph_3dOmegaJacobian (RCSB_JAC_OMEGA),
 * src/RcsCore/Rcs_kinematics.c:566, :789. effector / refBdy / refFrame are depth-first body
 * indices, -1 = NULL (world). J [N x nJac x 3 x nJ].
 */
#define RCSB_JAC_POS 1
#define RCSB_JAC_OMEGA 2
typedef struct
{
  int kind;
  int effector, refBdy, refFrame;
} RcsbRelJac;
int rcsb_task_jacobians(const RcsbModel* model, long N, int qdim, const double* q, int nJac,
                        const RcsbRelJac* jac, double* J, unsigned flags, void* stream);

/* ---- dynamics --------------------------------------------------------------------------- */

/*
 * M [N x nJ x nJ], h [N x nJ], Fg [N x nJ], E [N]; any may be NULL. Sign conventions are
 * the reference's: M qdd = Fg + h + ..., i.e. h = -C(q,qd) qd and Fg = -G(q)
 * (src/RcsCore/Rcs_dynamics.c:718-731). qd may be NULL (zero velocities).
 */
int rcsb_kinetic_terms(const RcsbModel* model, long N, int qdim, const double* q,
                       const double* qd, double* M, double* h, double* Fg, double* E,
                       unsigned flags, void* stream);

/*
 * Joint accelerations qdd [N x nJ] of Rcs_directDynamics (src/RcsCore/Rcs_dynamics.c:673):
 * M qdd = F_gravity + F_ext - F_jnt + h, solved with the reference's Cholesky factorisation
 * (src/RcsCore/Rcs_linAlg.c:66, :133) inside the assembly kernel: M never leaves the SM.
 * Fext / Fjnt [N x nJ] may be NULL; E [N] optional. A failed factorisation gives zeros.
 */
int rcsb_direct_dynamics(const RcsbModel* model, long N, int qdim, const double* q,
                         const double* qd, const double* Fext, const double* Fjnt, double* qdd,
                         double* E, unsigned flags, void* stream);

/*
 * Centre of gravity cog [N x 3] (RcsGraph_COG, src/RcsCore/Rcs_kinematics.c:904) and its
 * Jacobian Jcog [N x 3 x nJ] (RcsGraph_COGJacobian, :973); either may be NULL. Same kernels as
 * rcsb_kinetic_terms: column k of sum_b m_b JT_b is the linear momentum of column k's composite
 * body under a unit joint rate.
 */
int rcsb_cog(const RcsbModel* model, long N, int qdim, const double* q, double* cog, double* Jcog,
             unsigned flags, void* stream);

/*
 * rcsb_kinetic_terms and rcsb_cog from ONE pass of the kinetic-terms kernels (the whole-body
 * evaluation of BASELINE.json configs[4] wants M together with the COG Jacobian rows of its
 * COGX / COGY tasks). Every output may be NULL.
 */
int rcsb_kinetic_terms_cog(const RcsbModel* model, long N, int qdim, const double* q,
                           const double* qd, double* M, double* h, double* Fg, double* E,
                           double* cog, double* Jcog, unsigned flags, void* stream);

/*
 * Whole-body evaluation of BASELINE.json configs[4] in one call: frames of the selected bodies
 * (A_BI as in rcsb_fk), the task Jacobians `jac` (J as in rcsb_task_jacobians), the mass matrix
 * M [N x nJ x nJ] and the COG Jacobian Jcog [N x 3 x nJ] of the same states. Every output may
 * be NULL. Reference per state: RcsGraph_setState (Rcs_graph.c:231), ControllerBase::computeJ
 * (ControllerBase.cpp:776) over XYZ / XYZABC / COG tasks, RcsGraph_computeMassMatrix
 * (Rcs_dynamics.c:610).
 */
int rcsb_wholebody(const RcsbModel* model, long N, int qdim, const double* q, int nSel,
                   const int* bodySel, double* A_BI, int nJac, const RcsbRelJac* jac, double* J,
                   double* M, double* Jcog, unsigned flags, void* stream);

/* ---- inverse kinematics ----------------------------------------------------------------- */

/*
 * `iters` resolved-motion-rate steps per state, all tasks active:
 *   dx = computeDX(x_des); dH = alpha * jointLimitGradient; dq = J# dx + N invWq (-dH);
 *   q += dq; setState.
 * q [N x dof] in/out. xDes [N x nx], nx = 3 per task (XYZ: position, ABC: x-y-z Euler
 * angles). Optional outputs of the LAST iteration: dx [N x nx] (task error before the
 * step), dq [N x dof], det [N] (determinant of J invWq J^T + lambda I; 0 = singular, the
 * step is then zero as in the reference).
 */
int rcsb_ik_rmr(const RcsbModel* model, int nTasks, const RcsbTask* tasks, long N, double* q,
                const double* xDes, double lambda, double alpha, int iters, double* dx,
                double* dq, double* det, unsigned flags, void* stream);

/*
 * The same solver with the remaining inputs and outputs of the reference's interface
 * (IkSolverRMR::solveRightInverse(dq_ts, dq_ns, dx, dH, activation, lambda),
 * src/RcsCore/IkSolverRMR.cpp:385-415, :415-596; solveLeftInverse :205-360):
 *   activation  one value per task, shared by all states of the call (the reference passes it per
 *               call). Tasks with activation <= 0 are dropped from J and dx
 *               (ControllerBase::computeJ / compressToActive, ControllerBase.cpp:776, :1204);
 *               xDes keeps its full layout [N x nx], nx over ALL tasks. NULL = all active.
 *   scaleDx     != 0: the rows of dx are multiplied by their task's activation
 *               (ControllerBase::computeDX(dx, x_des, a_des), ControllerBase.cpp:896-935);
 *               0: dx as computeDX(dx, x_des) leaves it (examples/ExampleIK.cpp:162).
 *   lambdaRows  right inverse: one regularisation value per ACTIVE row of J (the m x 1 lambda of
 *               MatNd_rwPinv, Rcs_MatNd.c:1833); NULL = the scalar `lambda` on every row.
 *   blending    left inverse: task-space weights Wx of TaskSpaceBlender (TaskSpaceBlender.cpp:
 *               1193 Binary, :1228 Linear, :1109 Approximate) as selected by
 *               IkSolverRMR::setActivationBlending (IkSolverRMR.cpp:800). The iterative modes
 *               (IndependentTasks / DependentTasks) are not implemented: RCSB_EUNSUPPORTED.
 *   dqTs, dqNs  [N x dof] task-space and null-space part of the LAST step (dq = dqTs + dqNs for
 *               the right inverse); either may be NULL.
 * dx (if given) holds the ACTIVE rows only, [N x nxActive]. Host arrays: activation, lambdaRows.
 * Calls that use scaleDx with activations != 1, lambdaRows, blending != Binary, dqTs or dqNs run
 * the general kernel (literal algebra of the reference); plain activation masks keep the fast
 * kernels.
 */
#define RCSB_BLEND_BINARY 0
#define RCSB_BLEND_LINEAR 1
#define RCSB_BLEND_APPROXIMATE 2
typedef struct
{
  const double* activation;   /* [nTasks] or NULL */
  int scaleDx;
  int blending;               /* RCSB_BLEND_*, left inverse only */
  const double* lambdaRows;   /* [active rows] or NULL */
  double* dqTs;
  double* dqNs;
} RcsbIkOptions;
int rcsb_ik_rmr_ex(const RcsbModel* model, int nTasks, const RcsbTask* tasks, long N, double* q,
                   const double* xDes, double lambda, double alpha, int iters, double* dx,
                   double* dq, double* det, const RcsbIkOptions* options, unsigned flags,
                   void* stream);

/*
 * Prioritised resolved-motion-rate steps: IkSolverPrioRMR::solve (src/RcsCore/IkSolverPrioRMR.cpp:
 * 178-420) iterated like examples/ExampleIK.cpp. priority[t] = 0: first level, 1: second level
 * (solved in the null space of the first), other values: the task is ignored, as are tasks with
 * activation[t] <= 0 (activation may be NULL). Per iteration
 *   dq1 = J1# dx1,  dq2 = (J2 N1)# (dx2 - J2 J1# dx1),  dq_ns = -N1 N2 invWq (alpha dH),
 *   q += dq1 + dq2 + dq_ns,
 * and all three are zero if one of the two projections is singular. Outputs of the LAST iteration
 * (each may be NULL): dq1, dq2, dqNs [N x dof], ok [N] (1.0 = both projections regular). xDes keeps
 * its full layout over all tasks. Models whose unconstrained joints are kinematically coupled are
 * not supported by this entry point (RCSB_EUNSUPPORTED); at most 4 three-dimensional tasks.
 */
int rcsb_ik_prio_rmr(const RcsbModel* model, int nTasks, const RcsbTask* tasks, const int* priority,
                     const double* activation, long N, double* q, const double* xDes, double lambda,
                     double alpha, int iters, double* dq1, double* dq2, double* dqNs, double* ok,
                     unsigned flags, void* stream);

/*
 * Resolved-motion-rate steps over ANY stack of the task kinds below, with reference bodies and
 * frames: the controller side of the reference for task sets like GenericHumanoid/cAction.xml
 * (hands XYZ, COGX / COGY, feet XYZABC relative to the Heel; BASELINE.json configs[4]), solved
 * with IkSolverRMR::solveRightInverse (IkSolverRMR.cpp:415) and looped like examples/ExampleIK.cpp.
 *   XYZ     TaskPosition3D (TaskPosition3D.cpp:127, :207): position of `effector` relative to refBdy,
 *           in the coordinates of refFrame; Jacobian RcsGraph_3dPosJacobian (Rcs_kinematics.c:566)
 *   ABC     TaskEuler3D (TaskEuler3D.cpp:110-262): x-y-z Euler angles of the rotation of `effector`
 *           relative to refBdy; Jacobian RcsGraph_3dOmegaJacobian (Rcs_kinematics.c:789)
 *   XYZABC  TaskPose6D: XYZ followed by ABC with the same bodies
 *   COGX / COGY / COGZ   TaskCOM1D (TaskCOM1D.cpp:105-125): one component of the centre of gravity of
 *           the subtree of `effector` (-1: the whole graph, RcsGraph_COG / RcsGraph_COGJacobian,
 *           Rcs_kinematics.c:904-1036), in world coordinates or, with refBdy, relative to that body
 *           and in its frame (TaskCOM3D.cpp:101-165)
 *   JOINT   TaskJoint (TaskJoint.cpp:227-265): the value of joint `effector` (a jointIndex; the
 *           joint must be unconstrained); with refBdy >= 0 (here: the jointIndex of an unconstrained
 *           reference joint, the XML's refJnt) x = q_joint + refGain q_ref and J has refGain in the
 *           reference joint's column (refGain: the XML default is 1)
 *   POLARX / POLARY / POLARZ   TaskPolar2D (controlVariable "POLAR" with axisDirection "X" / "Y" /
 *           "Z", the default; TaskPolar2D.cpp:131-290): the polar angles (phi, theta) of that axis
 *           of the effector in the frame of refBdy (world if -1): x = Vec3d_getPolarAngles of the
 *           axis' row of A_ER, dx = Vec3d_getPolarErrorFromDirection (Rcs_Vec3d.c:233), J = the two
 *           rows of RcsGraph_3dOmegaJacobian(effector, refBdy, effector) across the axis; refFrame
 *           is not used
 * effector / refBdy / refFrame: depth-first body indices, -1 = none; refFrame = -1 with a refBdy
 * means refFrame = refBdy, as the reference's XML constructor does (Task.cpp:160). xDes [N x nx],
 * nx = sum of the task dimensions (3, 3, 6, 1, 1, 2). Outputs of the last iteration (may be NULL):
 * dx [N x nx], dq [N x dof], det [N]. One thread per state, the reference's literal algebra;
 * up to 24 rows; models with kinematically coupled unconstrained joints are not supported here.
 */
#define RCSB_TASK_COGX 4
#define RCSB_TASK_COGY 5
#define RCSB_TASK_COGZ 6
#define RCSB_TASK_JOINT 7
#define RCSB_TASK_POLARX 8
#define RCSB_TASK_POLARY 9
#define RCSB_TASK_POLARZ 10
typedef struct
{
  int kind;         /* RCSB_TASK_* */
  int effector;
  int refBdy;
  int refFrame;
  double refGain;   /* JOINT tasks with a reference joint only (TaskJoint::refGain); ignored otherwise */
} RcsbTaskSpec;
int rcsb_ik_tasks_rmr(const RcsbModel* model, int nTasks, const RcsbTaskSpec* tasks, long N, double* q,
                      const double* xDes, double lambda, double alpha, int iters, double* dx, double* dq,
                      double* det, unsigned flags, void* stream);

/*
 * IkSolverConstraintRMR::solveRightInverse (src/RcsCore/IkSolverConstraintRMR.cpp:156-262), its
 * joint-limit active set (no collision model): the step of rcsb_ik_tasks_rmr; then every
 * unconstrained joint that this step would leave closer than 2 degrees to q_min / q_max, and that is
 * not a JOINT task of its own, becomes a joint constraint (a unit row of J; target 0, or a tenth of
 * 0.1 degrees back towards the margin if it is more than 0.1 degrees beyond it) and the step is
 * solved once more with these rows appended in joint order; q += dq, `iters` times. Arguments as
 * rcsb_ik_tasks_rmr (dx: the task rows); det is the determinant of the last solve; nViolations [N]
 * (may be NULL) = number of joint constraints of the last iteration. Up to 16 rows in all for models
 * with nJ <= 8, else up to 32; a state that would need more keeps the unconstrained step and
 * reports -1.
 */
int rcsb_ik_constraint_rmr(const RcsbModel* model, int nTasks, const RcsbTaskSpec* tasks, long N, double* q,
                           const double* xDes, double lambda, double alpha, int iters, double* dx, double* dq,
                           double* det, double* nViolations, unsigned flags, void* stream);

/* ---- multi-GPU summary statistics ------------------------------------------------------- */

/*
 * out5 = {count, sum, sum of squares, min, max} of x[0], x[stride], ..., x[(n - 1) stride];
 * x and out5 are device pointers on the current device, asynchronous on `stream`. This is the
 * payload every rank contributes to the one all-gather of the sharded path (states are
 * independent, so nothing else is exchanged).
 */
int rcsb_summary_stats(const double* x, long n, long stride, double* out5, void* stream);

/* ---- device groups: the GPUs of one node behind one handle ------------------------------ */

/*
 * States are independent and the model is replicated, so a batch is cut into contiguous slices,
 * one per device (slice r = [first, first + count) of rcsb_group_slice: sizes differ by at most one
 * state), each evaluated by the kernels of the single-device entry points on its own stream and
 * host thread. All array arguments of the group calls are HOST pointers of the WHOLE batch (pinned
 * memory from rcsb_host_alloc is fastest); the library moves every slice to its device and the
 * results back. The path has no data-path collective; its one exchange is an NCCL all-gather,
 * between the devices, of per-rank summary statistics {count, sum, sum of squares, min, max} of a
 * per-state result (named with each call). `stats` (may be NULL) receives the gathered
 * [nDevices x 5] table. `devices` = NULL means devices 0 .. nDevices-1. NCCL is loaded at run
 * time (libnccl.so.2); a group of one device does not need it.
 */
typedef struct RcsbGroup RcsbGroup;
int rcsb_group_create(const RcsGraph* graph, int nDevices, const int* devices, RcsbGroup** group);
void rcsb_group_destroy(RcsbGroup* group);
int rcsb_group_size(const RcsbGroup* group);
void rcsb_group_slice(const RcsbGroup* group, long N, int rank, long* first, long* count);
/* rcsb_fk_jacobians_mass over the group; statistics: world z of the first selected frame */
int rcsb_group_fk_jacobians_mass(RcsbGroup* group, long N, int qdim, const double* q, int nSel,
                                 const int* bodySel, double* A_BI, int nEff, const int* effBody,
                                 const double* effPt, double* JT, double* JR, double* M, double* stats);
/* rcsb_kinetic_terms over the group; statistics: total energy E (needs E) */
int rcsb_group_kinetic_terms(RcsbGroup* group, long N, int qdim, const double* q, const double* qd,
                             double* M, double* h, double* Fg, double* E, double* stats);
/* rcsb_ik_rmr over the group (the 4M-target strong-scaling case of BASELINE.json configs[3]);
 * statistics: determinant of the last step */
int rcsb_group_ik_rmr(RcsbGroup* group, int nTasks, const RcsbTask* tasks, long N, double* q,
                      const double* xDes, double lambda, double alpha, int iters, double* dx, double* dq,
                      double* det, double* stats);
/* rcsb_wholebody over the group (configs[4]); statistics: M[0][0] (needs M) */
int rcsb_group_wholebody(RcsbGroup* group, long N, int qdim, const double* q, int nSel,
                         const int* bodySel, double* A_BI, int nJac, const RcsbRelJac* jac, double* J,
                         double* M, double* Jcog, double* stats);

/* ---- utilities -------------------------------------------------------------------------- */

const char* rcsb_last_error(void);
int rcsb_device_count(void);
/* pinned host memory for RCSB_HOST_PTRS calls */
void* rcsb_host_alloc(size_t bytes);
void rcsb_host_free(void* p);
/*
 * Host side of the RCSB_HOST_PTRS path on multi-socket boxes: restricts the calling thread (and
 * the threads it creates later) to the allowed CPUs of the NUMA node `device` is attached to, and
 * makes that node the preferred one for the memory it allocates from now on (pinned buffers
 * included), as far as the process's cpuset allows. Call it once per rank before allocating the
 * host buffers. `info` (may be NULL) receives a one-line description of what was done.
 * Returns 0, or RCSB_EINVAL for an unknown device; an unavailable topology is not an error.
 */
int rcsb_bind_host_to_device(int device, char* info, size_t infoBytes);
/* number of kernels launched by this library in this process (all threads) */
unsigned long long rcsb_kernel_launch_count(void);

/* ---- single-state reference API (N = 1 through the kernels above) ---------------------- */

/* The wrappers below keep a device model per graph (rebuilt when the graph's parameters are
 * edited). RcsGraph_destroy() of this library releases it; for a graph owned by another library
 * (libRcsCore: same struct layout) call this before destroying the graph. */
void rcsb_graph_release_device_cache(RcsGraph* self);

/* src/RcsCore/Rcs_graph.c:231. q / q_dot: dof x 1, nJ x 1 or NULL (use graph->q). Fills
 * A_BI / A_JI / x_dot / omega of every body and joint. Returns true if a slave joint's
 * value had to be changed by more than 1e-8. Wrong dimensions are fatal (see
 * Rcs_setFatalMode). */
bool RcsGraph_setState(RcsGraph* self, const MatNd* q, const MatNd* q_dot);
/* src/RcsCore/Rcs_graph.c:301 */
void RcsGraph_computeForwardKinematics(RcsGraph* self, const MatNd* q, const MatNd* q_dot);
/* src/RcsCore/Rcs_graph.c:327: forward kinematics of one body (or its subtree) from explicit
 * q / q_dot; the other bodies keep their frames. */
void RcsGraph_computeBodyKinematics(RcsGraph* self, RcsBody* bdy, const MatNd* q, const MatNd* q_dot,
                                    bool subTree);
/* src/RcsCore/Rcs_kinematics.c:178, :200, :416. A_BI = NULL: columns in world coordinates,
 * otherwise rotated into the frame with rotation matrix A_BI (J <- A_BI J). */
void RcsGraph_worldPointJacobian(const RcsGraph* self, const RcsBody* body, const double I_pt[3],
                                 double A_BI[3][3], MatNd* J);
void RcsGraph_bodyPointJacobian(const RcsGraph* self, const RcsBody* body, const double B_pt[3],
                                double A_BI[3][3], MatNd* J);
void RcsGraph_rotationJacobian(const RcsGraph* self, const RcsBody* body, double A_BI[3][3],
                               MatNd* J);
/* src/RcsCore/Rcs_kinematics.c:566, :789 */
void RcsGraph_3dPosJacobian(const RcsGraph* self, const RcsBody* effector, const RcsBody* refBdy,
                            const RcsBody* refFrame, MatNd* J);
void RcsGraph_3dOmegaJacobian(const RcsGraph* self, const RcsBody* effector, const RcsBody* refBdy,
                              const RcsBody* refFrame, MatNd* J);
/* src/RcsCore/Rcs_dynamics.c:439 */
double RcsGraph_computeKineticTerms(const RcsGraph* self, MatNd* M, MatNd* h, MatNd* F_gravity);
/* src/RcsCore/Rcs_dynamics.c:673; returns the energy like the reference */
double Rcs_directDynamics(const RcsGraph* graph, const MatNd* F_ext, const MatNd* F_jnt,
                          MatNd* q_ddot);
/* src/RcsCore/Rcs_dynamics.c:610, :576 */
void RcsGraph_computeMassMatrix(const RcsGraph* self, MatNd* M);
void RcsGraph_computeGravityTorque(const RcsGraph* self, MatNd* T_gravity);
/* src/RcsCore/Rcs_kinematics.c:904, :973; RcsGraph_COG returns the mass it averaged over */
double RcsGraph_COG(const RcsGraph* self, double r_cog[3]);
void RcsGraph_COGJacobian(const RcsGraph* self, MatNd* J_cog);

#ifdef __cplusplus
}
#endif

#endif
