/*
 * Kernel argument blocks and launch entry points (host <-> .cu translation units).
 */
#ifndef RCSB_KERNELS_H
#define RCSB_KERNELS_H

#include <cuda_runtime.h>

#define RCSB_FK_THREADS 128
/* block size of the fused FK + Jacobians + mass-matrix variant (register-bound: smaller
 * blocks fill the register file more evenly) */
#ifndef RCSB_FKM_THREADS
#define RCSB_FKM_THREADS 128
#endif
#define RCSB_FK_BLOCK(mass) ((mass) ? RCSB_FKM_THREADS : RCSB_FK_THREADS)
#define RCSB_MAX_FRAME_SLOTS 8
#define RCSB_KT_WALK_THREADS 128
#define RCSB_KT_ASM_THREADS 128
#define RCSB_IK_THREADS 128

namespace rcsb
{

struct FkArgs
{
  const char* model;   /* device flat model */
  const char* plan;    /* device FK plan */
  int modelBytes;
  int planBytes;
  long N;
  int qdim;
  const double* q;
  const double* qd;
  double* A_BI;
  double* A_JI;
  double* xdot;
  double* omega;
  double* qOut;
  double* JT;
  double* JR;
  double* M;           /* N x nJ x nJ mass matrix (fused path, plans built with mass = 1) */
  double* qdOut;       /* N x dof joint rates after expansion and coupling (needs qd) */
  int massMode;        /* 0: no M; 1: per-column momentum accumulators; 2: serial chain, composites */
  int pruned;          /* 1: walk the plan's sub-tree only (body selections without velocities) */
  int noCoupling;      /* 1: take slave joints from q as given (RcsGraph_computeForwardKinematics) */
  int chainNc;         /* > 0: fkChainKernel<..., chainNc> (RcsbFkPlanHeader::chainNc; frames, Jacobians, M) or, with qd,
                          fkChainVelKernel (frames, x_dot, omega) */
  int velRecDoubles;   /* 3 x selected bodies: doubles per state in xdot / omega */
  int scrDoubles;      /* doubles of the Jacobian scratch: nScrRows x (RCSB_FK_THREADS + 1) */
  /* body-fixed points of up to RCSB_FK_INLINE_PTS effectors travel with the launch, not with the
   * cached plan: a control loop that moves its points (RcsGraph_worldPointJacobian) neither
   * rebuilds nor accumulates plans */
  int effPtInline;     /* 1: take the points from effPtArg */
  double effPtArg[3 * 8];
};
#define RCSB_FK_INLINE_PTS 8

struct KtArgs
{
  const char* model;
  const char* plan;    /* kinetic-terms plan (joint relation table, link slots) */
  int modelBytes;
  int planBytes;
  long N;
  int qdim;
  const double* q;
  const double* qd;
  double* M;
  double* h;
  double* Fg;
  double* E;
  const double* Fext;  /* N x nJ external / joint forces of the direct dynamics, or NULL */
  const double* Fjnt;
  double* qdd;         /* N x nJ joint accelerations (Rcs_directDynamics), or NULL */
  double* cog;         /* N x 3 centre of gravity (RcsGraph_COG) */
  double* Jcog;        /* N x 3 x nJ (RcsGraph_COGJacobian) */
  double* rec;         /* N x recDoubles per-state records (walk -> assembly) */
  int recDoubles;
  int planWalkBytes;   /* leading part of the plan staged by the walk kernel */
  int stageM;          /* assembly: stage M in shared memory (needed by the direct dynamics) */
  /* whole-body plans */
  double* A_BI;        /* N x nSel x 12 frames of the selected bodies (written by the walk) */
  double* Jt;          /* N x nJac x 3 x nJ task Jacobians (written by the assembly) */
  int wholeBody;       /* 1: the plan is a whole-body plan */
  int symmetric;       /* plan header's `symmetric` (shared-memory layout of the assembly kernel) */
};

#define RCSB_IK_CHAIN_ALLROT 1      /* every Jacobian column of the chain is a rotation and the program has no joint steps (offSegXf) */
#define RCSB_IK_CHAIN_XYZ_FIRST 2   /* the position task is the first task of the plan */
struct IkArgs
{
  const char* model;
  const char* plan;    /* IK plan: tasks, chains, joint metric */
  int modelBytes;
  int planBytes;
  long N;
  double* q;           /* N x dof in/out */
  const double* xDes;  /* N x nx */
  double lambda;
  double alpha;
  int iters;
  int leftInverse;     /* 1: IkSolverRMR::solveLeftInverse instead of solveRightInverse */
  int chainFlags;      /* RCSB_IK_CHAIN_*: what the plan builder knows about the single-chain program */
  double* dx;
  double* dq;
  double* det;
  /* extended solver inputs / outputs (rcsb_ik_rmr_ex; general kernel only) */
  int extended;        /* 1: any of the fields below is in use */
  int lambdaPerRow;    /* 1: lambdaRow[r] regularises row r of J W J^T (MatNd_addDiag, Rcs_MatNd.c:1872) */
  int blending;        /* left inverse: RCSB_BLEND_* task-space weights Wx (TaskSpaceBlender.cpp) */
  int constraint;      /* ikTasksKernel: 1 = IkSolverConstraintRMR's joint-limit active set */
  double lambdaRow[12];
  double taskScale[4]; /* dx rows of task t are multiplied by it (ControllerBase::computeDX with a_des) */
  double taskAct[4];   /* activation of task t (blending weights) */
  int taskLevel[4];    /* prioritised solver: priority level of plan task t (0, 1; else ignored) */
  double* dqTs;        /* N x dof: task-space part of the last step (prioritised solver: dq2) */
  double* dqNs;        /* N x dof: null-space part of the last step */
  double* nViol;       /* N: joint-limit constraints added in the last step (constraint solver), -1 = too many */
};

}   // namespace rcsb

extern "C" {
cudaError_t rcsb_launch_fk(const rcsb::FkArgs* a, int nSlots, int nScrRows, cudaStream_t stream);
/* enqueues the walk kernel and the assembly kernel (2 launches) */
cudaError_t rcsb_launch_kt(const rcsb::KtArgs* a, int nSlots, int nJ, int nJ2, int smCount,
                           cudaStream_t stream);
/* maximum dynamic shared memory of the assembly kernel for a record of recDoubles (plan checks) */
cudaError_t rcsb_launch_ik(const rcsb::IkArgs* a, int nSlots, int nx, int nJ, int dof, int chainNc,
                           int multiNu, int nEffs, int nTasks, cudaStream_t stream);
/* nxCap: rows the solve may reach (nx, or nx + candidate joint constraints for the constraint solver) */
cudaError_t rcsb_launch_ik_tasks(const rcsb::IkArgs* a, int nSlots, int nx, int nxCap, int nJ, int dof,
                                 int nFrameSlots, cudaStream_t stream);
cudaError_t rcsb_launch_ik_prio(const rcsb::IkArgs* a, int nSlots, int nx, int nJ, int dof,
                                cudaStream_t stream);
}

#endif
