/*
 * Batched resolved-motion-rate inverse kinematics for sm_100a: K damped-least-squares steps
 * per state, one thread per state.
 *
 * One iteration reproduces, per state, the loop of examples/ExampleIK.cpp:154-177:
 *   controller.computeDX(dx, x_des)            src/RcsCore/ControllerBase.cpp:896
 *     XYZ: x_des - x                           src/RcsCore/TaskGenericIK.cpp:80
 *     ABC: Euler error below 1 deg, else axis * angle
 *                                              src/RcsCore/TaskEuler3D.cpp:243-262,
 *                                              Rcs_Mat3d.c:956, :1044, :1152, :1227, :1613,
 *                                              external/Rcs_thirdPartyMath.c:90
 *   controller.computeJointlimitGradient(dH)   src/RcsCore/Rcs_kinematics.c:1496, dH *= alpha
 *   ikSolver.solveRightInverse(dq, dx, dH, a, lambda)
 *                                              src/RcsCore/IkSolverRMR.cpp:415-596
 *     J        stacked task Jacobians          ControllerBase.cpp:776, Rcs_kinematics.c:200, :416
 *     invWq    weightMetric (q_max - q_min)    Rcs_graph.c:1087
 *     J#     = invWq J^T (J invWq J^T + lambda 1)^-1   Rcs_MatNd.c:1833 (MatNd_rwPinv)
 *              via Banachiewicz Cholesky, in-place inverse of L, L^-T L^-1
 *                                              Rcs_linAlg.c:66-101, :210-227, :270-307
 *     dq     = J# dx + (1 - J# J) invWq (-dH); singular (a pivot <= 0): det = 0, dq = 0
 *   q += dq; RcsGraph_setState
 * All tasks are active (activation 1). Graphs with kinematically coupled joints are solved in
 * the reduced joint space of RcsGraph_coupledJointMatrix (Rcs_graph.c:2823, IkSolverRMR.cpp:
 * 424-443) by the general kernel.
 *
 * The small dense algebra (6 x 7 for the arm) stays in per-thread local arrays (L1), the
 * tree walk is the one of rcsb_fk.cu. Bound: FP64 pipe / L1, not HBM (176 B per solve).
 */
#include "rcsb_device.cuh"
#include "rcsb_kernels.h"
#include "rcsb_plan.h"

#include "rcsb.h"

#include <cfloat>

namespace rcsb
{

struct IkPlanView
{
  const RcsbIkPlanHeader* hdr;
  const int* taskKind;
  const int* taskBody;
  const int* bodyTaskStart;
  const int* bodyTaskList;
  const int* jntNeeded;
  const int* chainStart;
  const int* chain;
  const int* colJoint;
  const double* invWq;
  const double* jlGain;
  const double* q0;
  const int* colRed;
  const int* colCpl;
  const int* jntSlot;
  const int* slotTaskMask;
  const int* taskEff;
  const int* bodyEffIdx;
  const int4* walk;
  const int* progStart;
  const int4* prog;
  const int* chainType;
  const double* chainXf;
  const double* segXf;
  const int* chainJoint;
  const double* chainW;
  const double* chainGain;
  const double* chainQ0;
  const int* otherJoint;
  const double* otherW;
  const double* otherGain;
  const double* otherQ0;

  __device__ __forceinline__ void bind(const char* blob)
  {
    hdr = reinterpret_cast<const RcsbIkPlanHeader*>(blob);
    jntSlot = reinterpret_cast<const int*>(blob + hdr->offJntSlot);
    slotTaskMask = reinterpret_cast<const int*>(blob + hdr->offSlotTaskMask);
    taskEff = reinterpret_cast<const int*>(blob + hdr->offTaskEff);
    bodyEffIdx = reinterpret_cast<const int*>(blob + hdr->offBodyEffIdx);
    walk = reinterpret_cast<const int4*>(blob + hdr->offWalk);
    colRed = reinterpret_cast<const int*>(blob + hdr->offColRed);
    colCpl = reinterpret_cast<const int*>(blob + hdr->offColCpl);
    progStart = reinterpret_cast<const int*>(blob + hdr->offProgStart);
    prog = reinterpret_cast<const int4*>(blob + hdr->offProg);
    chainType = reinterpret_cast<const int*>(blob + hdr->offChainType);
    chainXf = reinterpret_cast<const double*>(blob + hdr->offChainXf);
    segXf = reinterpret_cast<const double*>(blob + hdr->offSegXf);
    chainJoint = reinterpret_cast<const int*>(blob + hdr->offChainJoint);
    chainW = reinterpret_cast<const double*>(blob + hdr->offChainW);
    chainGain = reinterpret_cast<const double*>(blob + hdr->offChainGain);
    chainQ0 = reinterpret_cast<const double*>(blob + hdr->offChainQ0);
    otherJoint = reinterpret_cast<const int*>(blob + hdr->offOtherJoint);
    otherW = reinterpret_cast<const double*>(blob + hdr->offOtherW);
    otherGain = reinterpret_cast<const double*>(blob + hdr->offOtherGain);
    otherQ0 = reinterpret_cast<const double*>(blob + hdr->offOtherQ0);
    taskKind = reinterpret_cast<const int*>(blob + hdr->offTaskKind);
    taskBody = reinterpret_cast<const int*>(blob + hdr->offTaskBody);
    bodyTaskStart = reinterpret_cast<const int*>(blob + hdr->offBodyTaskStart);
    bodyTaskList = reinterpret_cast<const int*>(blob + hdr->offBodyTaskList);
    jntNeeded = reinterpret_cast<const int*>(blob + hdr->offJntNeeded);
    chainStart = reinterpret_cast<const int*>(blob + hdr->offChainStart);
    chain = reinterpret_cast<const int*>(blob + hdr->offChain);
    colJoint = reinterpret_cast<const int*>(blob + hdr->offColJoint);
    invWq = reinterpret_cast<const double*>(blob + hdr->offInvWq);
    jlGain = reinterpret_cast<const double*>(blob + hdr->offJlGain);
    q0 = reinterpret_cast<const double*>(blob + hdr->offQ0);
  }
};

/* x-y-z Euler angles <-> rotation matrix (src/RcsCore/Rcs_Mat3d.c:956, :1152) */
__device__ __forceinline__ void toEuler(double ea[3], const double A[9])
{
  const double cy = sqrt(A[0] * A[0] + A[3] * A[3]);
  if (cy > 16.0 * (double) FLT_EPSILON)
  {
    ea[0] = -atan2(A[7], A[8]);
    ea[1] = -atan2(-A[6], cy);
    ea[2] = -atan2(A[3], A[0]);
  }
  else
  {
    ea[0] = -atan2(-A[5], A[4]);
    ea[1] = -atan2(-A[6], cy);
    ea[2] = 0.0;
  }
}

__device__ __forceinline__ void fromEuler(double A[9], const double ea[3])
{
  double sa, ca, sb, cb, sg, cg;
  sincos(ea[0], &sa, &ca);
  sincos(ea[1], &sb, &cb);
  sincos(ea[2], &sg, &cg);
  A[0] = cb * cg;
  A[1] = ca * sg + sa * sb * cg;
  A[2] = sa * sg - ca * sb * cg;
  A[3] = -cb * sg;
  A[4] = ca * cg - sa * sb * sg;
  A[5] = sa * cg + ca * sb * sg;
  A[6] = sb;
  A[7] = -sa * cb;
  A[8] = ca * cb;
}

/* Orientation error of TaskEuler3D::computeDX (src/RcsCore/TaskEuler3D.cpp:243-262) between
 * the current rotation Ac and the desired rotation Ad (both row-major A_KI). */
__device__ __forceinline__ void rotationTaskError(double dx[3], const double Ad[9],
                                                  const double Ac[9])
{
  /* trace(A_des A_curr^T), summed as the reference does (diagonal entries of the product first,
   * Rcs_Mat3d.c:1613-1632): the branches below switch on the result */
  const double d0 = Ad[0] * Ac[0] + Ad[1] * Ac[1] + Ad[2] * Ac[2];
  const double d1 = Ad[3] * Ac[3] + Ad[4] * Ac[4] + Ad[5] * Ac[5];
  const double d2 = Ad[6] * Ac[6] + Ad[7] * Ac[7] + Ad[8] * Ac[8];
  const double tr = d0 + d1 + d2;
  double cs = 0.5 * (tr - 1.0);
  cs = cs > 1.0 ? 1.0 : (cs < -1.0 ? -1.0 : cs);
  const double angle = acos(cs);

  if (angle < 1.0 * (M_PI / 180.0))
  {
    /* e = 0.5 (n x nd + s x sd + a x ad), rows of A_curr against rows of A_des */
    double e[3] = {0.0, 0.0, 0.0}, t[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
    {
      cross3(t, Ac + 3 * r, Ad + 3 * r);
      e[0] += t[0];
      e[1] += t[1];
      e[2] += t[2];
    }
    dx[0] = 0.5 * e[0];
    dx[1] = 0.5 * e[1];
    dx[2] = 0.5 * e[2];
    return;
  }

  /* axis of A_21 = A_des A_curr^T, expressed in the 1 (current) frame */
  double R[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
    {
      R[3 * i + j] = Ad[3 * i] * Ac[3 * j] + Ad[3 * i + 1] * Ac[3 * j + 1] + Ad[3 * i + 2] * Ac[3 * j + 2];
    }

  double ax[3];
  if (angle < M_PI)
  {
    ax[0] = R[5] - R[7];
    ax[1] = R[6] - R[2];
    ax[2] = R[1] - R[3];
    const double len2 = ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2];
    if (len2 > 0.0)
    {
      const double inv = rsqrt(len2);
      ax[0] *= inv;
      ax[1] *= inv;
      ax[2] *= inv;
    }
  }
  else
  {
    /* angle is pi: axis from the largest diagonal term
     * (external/Rcs_thirdPartyMath.c:107-150) */
    if (R[0] >= R[4] && R[0] >= R[8])
    {
      ax[0] = 0.5 * sqrt(1.0 + R[0] - R[4] - R[8]);
      const double hi = 0.5 / ax[0];
      ax[1] = hi * R[3];
      ax[2] = hi * R[6];
    }
    else if (R[4] > R[0] && R[4] >= R[8])
    {
      ax[1] = 0.5 * sqrt(1.0 + R[4] - R[0] - R[8]);
      const double hi = 0.5 / ax[1];
      ax[0] = hi * R[3];
      ax[2] = hi * R[7];
    }
    else
    {
      ax[2] = 0.5 * sqrt(1.0 + R[8] - R[0] - R[4]);
      const double hi = 0.5 / ax[2];
      ax[0] = hi * R[6];
      ax[1] = hi * R[7];
    }
  }

  /* rotate into the world frame (A_curr^T axis) and scale by the angle */
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    dx[i] = (Ac[i] * ax[0] + Ac[3 + i] * ax[1] + Ac[6 + i] * ax[2]) * angle;
  }
}

/* Out of line on purpose: taken for effector orientations in Euler gimbal lock only, it must not
 * cost the kernels around it registers. */
static __device__ __noinline__ void eulerRoundTrip(double A[9])
{
  double ea[3];
  toEuler(ea, A);
  fromEuler(A, ea);
}

/* Orientation error from the effector's rotation matrix itself. The reference rebuilds the
 * current rotation from its Euler angles (x = Mat3d_toEulerAngles(A_curr) in computeX, then
 * Mat3d_fromEulerAngles(x) in computeDX, TaskEuler3D.cpp:110, :243). Away from gimbal lock that
 * round trip is the identity up to rounding; in the gimbal branch of Mat3d_toEulerAngles
 * (cy <= 16 FLT_EPSILON, Rcs_Mat3d.c:956) the third angle is set to 0 and the rebuilt matrix
 * differs from A_curr by O(cy): there the round trip is made here too. */
__device__ __forceinline__ void rotationTaskErrorOfFrame(double dx[3], const double Ad[9],
                                                         const double Ac[9])
{
  const double cy2 = Ac[0] * Ac[0] + Ac[3] * Ac[3];
  const double lim = 16.0 * (double) FLT_EPSILON;
  if (cy2 <= lim * lim * 1.0000001)   /* warp-divergent and rare; exact test inside toEuler */
  {
    double Ar[9];
#pragma unroll
    for (int i = 0; i < 9; ++i)
    {
      Ar[i] = Ac[i];
    }
    if (sqrt(cy2) <= lim)
    {
      eulerRoundTrip(Ar);
    }
    rotationTaskError(dx, Ad, Ar);
    return;
  }
  rotationTaskError(dx, Ad, Ac);
}

/* The reference goes through the Euler angles of both frames (x = Mat3d_toEulerAngles(A_curr),
 * then Mat3d_fromEulerAngles of x and x_des inside computeDX). */
__device__ __forceinline__ void eulerTaskError(double dx[3], const double xDes[3],
                                               const double xCurr[3])
{
  double Ac[9], Ad[9];
  fromEuler(Ac, xCurr);
  fromEuler(Ad, xDes);
  rotationTaskError(dx, Ad, Ac);
}

/* TaskPolar2D::computeDX (TaskPolar2D.cpp:286): Vec3d_getPolarErrorFromDirection (Rcs_Vec3d.c:233-340)
 * of the desired polar angles (phi, theta) and row `direction` of A = A_ER. The current axis is turned
 * onto the desired one about their common normal by phi = acos(a_curr . a_des); the result is that
 * rotation vector in the effector frame without its component along the free axis. */
__device__ __noinline__ void polarTaskError(double om[2], const double polarDes[2], const double A[9],
                                            int direction)
{
  const double* ac = A + 3 * direction;
  double sp, cp, st, ct;
  sincos(polarDes[0], &sp, &cp);
  sincos(polarDes[1], &st, &ct);
  const double ad[3] = {ct * sp, st * sp, cp};
  const double cosPhi = ac[0] * ad[0] + ac[1] * ad[1] + ac[2] * ad[2];
  double n[3] = {ac[1] * ad[2] - ac[2] * ad[1], ac[2] * ad[0] - ac[0] * ad[2], ac[0] * ad[1] - ac[1] * ad[0]};
  const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  if (len == 0.0)
  {
    if (cosPhi > 0.0)
    {
      om[0] = 0.0;
      om[1] = 0.0;
      return;
    }
    /* opposite axes: any normal will do, the reference takes a_curr x (next row of A) */
    const double* o = A + 3 * ((direction + 1) % 3);
    n[0] = ac[1] * o[2] - ac[2] * o[1];
    n[1] = ac[2] * o[0] - ac[0] * o[2];
    n[2] = ac[0] * o[1] - ac[1] * o[0];
  }
  else
  {
    const double inv = 1.0 / len;
    n[0] *= inv;
    n[1] *= inv;
    n[2] *= inv;
  }
  const double phi = acos(fmin(1.0, fmax(-1.0, cosPhi)));   /* Math_acos */
  /* e_omega = A_ES (0, phi, 0), A_ES = A A_SR^T: column 1 of A_ES is A n. The products with the zero
   * entries of s_omega are kept out: they add exact zeros in the reference. */
  double e[3];
  for (int i = 0; i < 3; ++i)
  {
    e[i] = (A[3 * i] * n[0] + A[3 * i + 1] * n[1] + A[3 * i + 2] * n[2]) * phi;
  }
  om[0] = direction == 0 ? e[1] : e[0];
  om[1] = direction == 2 ? e[1] : e[2];
}

/* One state of the general kernels, everything the solvers need from the graph: forward
 * kinematics (RcsGraph_setState), the task errors dx (ControllerBase::computeDX), the stacked task
 * Jacobian J [nx x nJ] (ControllerBase::computeJ) and the joint-limit gradient term g = -alpha dH
 * in IK order (ControllerBase::computeJointlimitGradient). */
template <int MAXNX, int MAXNJ, int MAXDOF, int NSLOTS>
__device__ __forceinline__ void ikStateKinematics(const ModelView& m, const IkPlanView& p, const IkArgs& a,
                                                  long s, const double* qv, double* stk, double* ja,
                                                  double* ef, double* J, double* dx, double* g)
{
  const int nb = m.hdr->nb;
  const int nJ = p.hdr->nJ;
  const int nx = p.hdr->nx;
  const int nTasks = p.hdr->nTasks;
  /* ---- forward kinematics (RcsGraph_setState) ------------------------------------------ */
  Frame f, keep;
  setIdentity(f);
  keep = f;
  for (int b = 0; b < nb; ++b)
  {
    const int ld = m.bodyLoad[b];
    const int bflags = m.bodyFlags[b];
    if (ld == RCSB_LOAD_IDENT)
    {
      setIdentity(f);
    }
    else if (NSLOTS > 0 && ld >= 0)
    {
      const double* sp = stk + ld * 12;
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        f.o[i] = sp[i];
      }
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        f.R[i] = sp[3 + i];
      }
    }
    if (bflags & RCSB_BF_LEAF)
    {
      keep = f;
    }

    const int j0 = m.bodyJntStart[b];
    const int j1 = j0 + m.bodyJntCount[b];
    for (int j = j0; j < j1; ++j)
    {
      const int jf = m.jntFlags[j];
      if (jf & RCSB_JF_HAS_AJP)
      {
        applyRelF(f, m.jntAJP + 12 * j, jf);
      }
      const double qj = qv[j];
      const int dir = m.jntType[j] % 3;
      double ax[3];
      if (jf & RCSB_JF_ROT)
      {
        double sn, cs;
        sincos(qj, &sn, &cs);
        rotateAboutAxis(f, dir, sn, cs);
        axisOf(f, dir, ax);
      }
      else
      {
        axisOf(f, dir, ax);
        f.o[0] += qj * ax[0];
        f.o[1] += qj * ax[1];
        f.o[2] += qj * ax[2];
      }
      if (p.jntNeeded[j])
      {
        double* d = ja + 6 * m.jntJacobi[j];
        d[0] = ax[0];
        d[1] = ax[1];
        d[2] = ax[2];
        d[3] = f.o[0];
        d[4] = f.o[1];
        d[5] = f.o[2];
      }
    }
    if (bflags & RCSB_BF_HAS_ABP)
    {
      applyRelF(f, m.bodyABP + 12 * b, bflags);
    }
    if (NSLOTS > 0)
    {
      const int sv = m.bodySave[b];
      if (sv >= 0)
      {
        double* sp = stk + sv * 12;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          sp[i] = f.o[i];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i)
        {
          sp[3 + i] = f.R[i];
        }
      }
    }
    for (int k = p.bodyTaskStart[b]; k < p.bodyTaskStart[b + 1]; ++k)
    {
      double* d = ef + 12 * p.bodyTaskList[k];
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        d[i] = f.o[i];
      }
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        d[3 + i] = f.R[i];
      }
    }
    if (bflags & RCSB_BF_LEAF)
    {
      f = keep;
    }
  }

  /* ---- task error and stacked Jacobian ---------------------------------------------------- */
  for (int i = 0; i < nx * nJ; ++i)
  {
    J[i] = 0.0;
  }
  for (int t = 0; t < nTasks; ++t)
  {
    const double* e = ef + 12 * t;
    const double* xd = a.xDes + s * nx + 3 * t;
    const bool pos = p.taskKind[t] == RCSB_TASK_XYZ;
    if (pos)
    {
      dx[3 * t] = xd[0] - e[0];
      dx[3 * t + 1] = xd[1] - e[1];
      dx[3 * t + 2] = xd[2] - e[2];
    }
    else
    {
      double xc[3];
      const double xdl[3] = {xd[0], xd[1], xd[2]};
      toEuler(xc, e + 3);
      eulerTaskError(dx + 3 * t, xdl, xc);
    }
    if (a.extended && t < 4)
    {
      /* ControllerBase::computeDX with activations: dx_i *= a_i (ControllerBase.cpp:925) */
      dx[3 * t] *= a.taskScale[t];
      dx[3 * t + 1] *= a.taskScale[t];
      dx[3 * t + 2] *= a.taskScale[t];
    }
    for (int k = p.chainStart[t]; k < p.chainStart[t + 1]; ++k)
    {
      const int ent = p.chain[k];
      const int col = ent & 0xffff;
      const bool isRot = (ent >> 16) != 0;
      const double* jc = ja + 6 * col;
      double c0, c1, c2;
      if (pos)
      {
        if (isRot)
        {
          const double d0 = e[0] - jc[3], d1 = e[1] - jc[4], d2 = e[2] - jc[5];
          c0 = jc[1] * d2 - jc[2] * d1;
          c1 = jc[2] * d0 - jc[0] * d2;
          c2 = jc[0] * d1 - jc[1] * d0;
        }
        else
        {
          c0 = jc[0];
          c1 = jc[1];
          c2 = jc[2];
        }
      }
      else
      {
        c0 = isRot ? jc[0] : 0.0;
        c1 = isRot ? jc[1] : 0.0;
        c2 = isRot ? jc[2] : 0.0;
      }
      J[(3 * t) * nJ + col] = c0;
      J[(3 * t + 1) * nJ + col] = c1;
      J[(3 * t + 2) * nJ + col] = c2;
    }
  }

  /* ---- joint-limit gradient, (-alpha dH), in IK order --------------------------------------- */
  for (int i = 0; i < nJ; ++i)
  {
    const double delta = qv[p.colJoint[i]] - p.q0[i];
    g[i] = -(a.alpha * (p.jlGain[i] * delta));
  }
}

/* LEFT: IkSolverRMR::solveLeftInverse (IkSolverRMR.cpp:205-272) instead of solveRightInverse:
 * J_r = J A invWq_r (computeKinematics, :176-200; binary task blending = unit task weights),
 * J# = (J_r^T J_r + lambda 1)^-1 J_r^T (MatNd_rwPinv2, Rcs_MatNd.c:1920: explicit Cholesky inverse
 * of the nq x nq matrix), dq_r = invWq_r o (J# dx + (1 - J# J_r) invWq_r (-alpha A+ dH)), dq = A dq_r. */
template <int MAXNX, int MAXNJ, int MAXDOF, int NSLOTS, bool LEFT = false>
__global__ void __launch_bounds__(RCSB_IK_THREADS)
ikKernel(const IkArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  stageBlobsTma(sModel, a.model, a.modelBytes, sPlan, a.plan, a.planBytes);

  ModelView m;
  m.bind(sModel);
  IkPlanView p;
  p.bind(sPlan);

  const long s = (long) blockIdx.x * RCSB_IK_THREADS + threadIdx.x;
  if (s >= a.N)
  {
    return;
  }

  const int nb = m.hdr->nb;
  const int dof = m.hdr->dof;
  const int nJ = p.hdr->nJ;
  const int nx = p.hdr->nx;
  const int nTasks = p.hdr->nTasks;
  constexpr int MAXT = MAXNX / 3;

  double qv[MAXDOF];
  double ja[6 * MAXNJ];       /* axis, origin per Jacobian column */
  double ef[12 * MAXT];       /* effector frame per task */
  double J[MAXNX * MAXNJ];    /* nx x nJ */
  double P[MAXNJ * MAXNX];    /* nJ x nx: invWq J^T, then J# */
  constexpr int MAXS = LEFT ? MAXNJ : MAXNX;   /* size of the symmetric system */
  double A[MAXS * MAXS];      /* J invWq J^T + lambda 1 (LEFT: J_r^T J_r + lambda 1) -> L -> L^-1 */
  double Ai[MAXS * MAXS];     /* inverse */
  double dx[MAXNX], g[MAXNJ], Jg[MAXNX], dq[MAXNJ], row[MAXNX];
  double dqT[MAXNJ], dqN[MAXNJ];   /* task-space / null-space parts of dq (dq_ts, dq_ns) */
  double stk[(NSLOTS > 0 ? NSLOTS : 1) * 12];
  double det = 0.0;

  for (int j = 0; j < dof; ++j)
  {
    qv[j] = a.q[s * dof + j];
  }
  for (int i = 0; i < nJ; ++i)
  {
    dq[i] = 0.0;
    dqT[i] = 0.0;
    dqN[i] = 0.0;
  }
  for (int i = 0; i < nx; ++i)
  {
    dx[i] = 0.0;
  }

  const int nCpl = m.hdr->nCpl;
  const bool hasCpl = p.hdr->nqr < nJ;   /* unconstrained slave joints: reduced joint space */
  /* RcsGraph_updateSlaveJoints (Rcs_graph.c:2756): part of every RcsGraph_setState */
  auto updateSlaves = [&]() {
    for (int c = 0; c < nCpl; ++c)
    {
      qv[m.cpl[c].slave] = slaveAngle(m.cpl[c], qv[m.cpl[c].master]);
    }
  };
  const int nq = p.hdr->nqr;         /* columns of the (reduced) problem */
  double wv[MAXNJ], av[MAXNJ];       /* reduced joint metric; weights of the coupling matrix A */

  for (int it = 0; it < a.iters; ++it)
  {
    updateSlaves();
    ikStateKinematics<MAXNX, MAXNJ, MAXDOF, NSLOTS>(m, p, a, s, qv, stk, ja, ef, J, dx, g);

    /* ---- kinematically coupled joints: reduced joint space (IkSolverRMR.cpp:424-443,
     * RcsGraph_coupledJointMatrix, Rcs_graph.c:2823). Column r of A holds 1 at its master / free
     * joint and the coupling sensitivities at its slaves, divided by |column sum|; every joint
     * has exactly one entry av[c] in column colRed[c].  J <- J A, invWq <- A+ invWq,
     * dH <- A+ dH with A+ = (A^T A)^-1 A^T. ------------------------------------------------- */
    if (hasCpl)
    {
      double colSum[MAXNJ], den[MAXNJ];
      for (int r = 0; r < nq; ++r)
      {
        colSum[r] = 0.0;
        den[r] = 0.0;
      }
      for (int c = 0; c < nJ; ++c)
      {
        const int ci = p.colCpl[c];
        av[c] = ci < 0 ? 1.0 : slaveVelocity(m.cpl[ci], qv[m.cpl[ci].master], 1.0);
        colSum[p.colRed[c]] += av[c];
      }
      for (int c = 0; c < nJ; ++c)
      {
        av[c] = av[c] / fabs(colSum[p.colRed[c]]);
        den[p.colRed[c]] += av[c] * av[c];
      }
      for (int r = 0; r < nq; ++r)
      {
        den[r] = 1.0 / den[r];
        wv[r] = 0.0;
        dq[r] = 0.0;    /* used as scratch for the reduced gradient */
      }
      for (int c = 0; c < nJ; ++c)
      {
        const int r = p.colRed[c];
        const double ia = den[r] * av[c];
        wv[r] += ia * p.invWq[c];
        dq[r] += ia * g[c];
      }
      for (int r = 0; r < nq; ++r)
      {
        g[r] = dq[r];
      }
      for (int k = 0; k < nx; ++k)
      {
        /* row k of J A; den is free again and serves as the accumulator */
        double* Jk = J + k * nJ;
        for (int r = 0; r < nq; ++r)
        {
          den[r] = 0.0;
        }
        for (int c = 0; c < nJ; ++c)
        {
          den[p.colRed[c]] += Jk[c] * av[c];
        }
        for (int r = 0; r < nq; ++r)
        {
          Jk[r] = den[r];
        }
      }
    }
    else
    {
      for (int i = 0; i < nJ; ++i)
      {
        wv[i] = p.invWq[i];
      }
    }
    /* null-space gradient with the joint metric applied: invWq (-alpha dH) */
    for (int i = 0; i < nq; ++i)
    {
      g[i] = wv[i] * g[i];
    }

    /* ---- J# = invWq J^T (J invWq J^T + lambda 1)^-1 ------------------------------------------ */
    const int n = LEFT ? nq : nx;
    double wx[MAXNX];    /* LEFT: task-space blending weights Wx, one per row of J */
    if (LEFT)
    {
      for (int k = 0; k < nx; ++k)
      {
        for (int i = 0; i < nq; ++i)
        {
          J[k * nJ + i] *= wv[i];
        }
      }
      for (int k = 0; k < nx; ++k)
      {
        wx[k] = 1.0;     /* TaskSpaceBlender::computeBinary (TaskSpaceBlender.cpp:1193) */
        const int t = k / 3;
        if (a.extended && a.blending == 1 && t < 4)
        {
          wx[k] = a.taskAct[t];   /* computeLinear (:1228) */
        }
        else if (a.extended && a.blending == 2 && t < 4)
        {
          /* computeApproximate (:1109) on the weighted Jacobian J_r = J A invWq_r */
          const double act = a.taskAct[t];
          const double ci = fmin(fmax(act, 0.0), 1.0);
          double aBuf = 0.0;
          for (int i = 0; i < nq; ++i)
          {
            aBuf += J[k * nJ + i] * J[k * nJ + i];
          }
          double wi1 = 1.0;
          if (ci < 1.0)
          {
            wi1 = aBuf == 0.0 ? 0.0 : a.lambda * ci / (aBuf * (1.0 - ci));
          }
          const double wi2 = pow(0.01 * a.lambda, 1.0 - ci);
          wx[k] = act * wi2 + (1.0 - act) * wi1;
        }
      }
      /* J_r^T Wx J_r + lambda 1 (MatNd_rwPinv2, Rcs_MatNd.c:1920) */
      for (int r = 0; r < nq; ++r)
      {
        for (int c = 0; c < nq; ++c)
        {
          double sum = 0.0;
          for (int k = 0; k < nx; ++k)
          {
            sum += a.extended ? J[k * nJ + r] * wx[k] * J[k * nJ + c] : J[k * nJ + r] * J[k * nJ + c];
          }
          A[r * n + c] = sum + (r == c ? a.lambda : 0.0);
        }
      }
    }
    else
    {
      for (int i = 0; i < nq; ++i)
      {
        const double w = wv[i];
        for (int k = 0; k < nx; ++k)
        {
          P[i * nx + k] = w * J[k * nJ + i];
        }
      }
      for (int r = 0; r < nx; ++r)
      {
        for (int c = 0; c < nx; ++c)
        {
          double sum = 0.0;
          for (int i = 0; i < nq; ++i)
          {
            sum += J[r * nJ + i] * P[i * nx + c];
          }
          A[r * nx + c] = sum + (r == c ? ((a.extended && a.lambdaPerRow) ? a.lambdaRow[r] : a.lambda) : 0.0);
        }
      }
    }

    /* Banachiewicz Cholesky, lower triangle in place; det = prod L_ii^2 */
    det = 1.0;
    for (int i = 0; i < n && det != 0.0; ++i)
    {
      for (int k = 0; k <= i; ++k)
      {
        double sum = A[i * n + k];
        for (int t = 0; t < k; ++t)
        {
          sum -= A[i * n + t] * A[k * n + t];
        }
        if (i > k)
        {
          A[i * n + k] = sum / A[k * n + k];
        }
        else if (sum > 0.0)
        {
          A[i * n + i] = sqrt(sum);
          det *= A[i * n + i] * A[i * n + i];
        }
        else
        {
          det = 0.0;
          break;
        }
      }
    }

    if (det == 0.0)
    {
      /* singular: the reference returns zero displacements and leaves q untouched */
      for (int i = 0; i < nJ; ++i)
      {
        dq[i] = 0.0;
        dqT[i] = 0.0;
        dqN[i] = 0.0;
      }
      continue;
    }

    /* L <- L^-1 in place */
    for (int k = 0; k < n; ++k)
    {
      A[k * n + k] = 1.0 / A[k * n + k];
      for (int i = k + 1; i < n; ++i)
      {
        double sum = 0.0;
        for (int t = k; t < i; ++t)
        {
          sum += A[i * n + t] * A[t * n + k];
        }
        A[i * n + k] = -sum / A[i * n + i];
      }
    }
    /* inverse = L^-T L^-1 */
    for (int r = 0; r < n; ++r)
    {
      for (int c = r; c < n; ++c)
      {
        double sum = 0.0;
        for (int t = c; t < n; ++t)
        {
          sum += A[t * n + r] * A[t * n + c];
        }
        Ai[r * n + c] = sum;
        Ai[c * n + r] = sum;
      }
    }
    if (LEFT)
    {
      /* J# = inverse J_r^T */
      for (int i = 0; i < nq; ++i)
      {
        for (int c = 0; c < nx; ++c)
        {
          double sum = 0.0;
          for (int t = 0; t < nq; ++t)
          {
            sum += Ai[i * n + t] * J[c * nJ + t];
          }
          P[i * nx + c] = a.extended ? sum * wx[c] : sum;   /* (...)^-1 J_r^T Wx */
        }
      }
    }
    else
    {
      /* J# = (invWq J^T) inverse, row by row in place */
      for (int i = 0; i < nq; ++i)
      {
        for (int k = 0; k < nx; ++k)
        {
          row[k] = P[i * nx + k];
        }
        for (int c = 0; c < nx; ++c)
        {
          double sum = 0.0;
          for (int k = 0; k < nx; ++k)
          {
            sum += row[k] * Ai[k * nx + c];
          }
          P[i * nx + c] = sum;
        }
      }
    }

    /* ---- dq = J# dx + (1 - J# J) invWq (-alpha dH) ------------------------------------------- */
    for (int r = 0; r < nx; ++r)
    {
      double sum = 0.0;
      for (int i = 0; i < nq; ++i)
      {
        sum += J[r * nJ + i] * g[i];
      }
      Jg[r] = sum;
    }
    if (!hasCpl)
    {
      for (int i = 0; i < nJ; ++i)
      {
        double ts = 0.0, ns = 0.0;
        for (int c = 0; c < nx; ++c)
        {
          ts += P[i * nx + c] * dx[c];
          ns += P[i * nx + c] * Jg[c];
        }
        dq[i] = LEFT ? (ts + (g[i] - ns)) * wv[i] : ts + (g[i] - ns);
        dqT[i] = LEFT ? ts * wv[i] : ts;
        dqN[i] = LEFT ? (g[i] - ns) * wv[i] : g[i] - ns;
      }
    }
    else
    {
      /* task-space and null-space steps of the reduced problem, expanded with A separately as
       * the reference does (dq_ts = A dqr, dq_ns = A dqr; IkSolverRMR.cpp:541-590) */
      double tsr[MAXNJ], nsr[MAXNJ];
      for (int i = 0; i < nq; ++i)
      {
        double ts = 0.0, ns = 0.0;
        for (int c = 0; c < nx; ++c)
        {
          ts += P[i * nx + c] * dx[c];
          ns += P[i * nx + c] * Jg[c];
        }
        tsr[i] = ts;
        nsr[i] = g[i] - ns;
        if (LEFT)
        {
          /* split form (IkSolverRMR.cpp:278-360): each part goes back to the unweighted space and
           * through A on its own */
          dqT[i] = ts * wv[i];
          dqN[i] = nsr[i] * wv[i];
          tsr[i] = (ts + nsr[i]) * wv[i];   /* one reduced step, back in the unweighted space */
          nsr[i] = 0.0;
        }
      }
      double rT[MAXNJ], rN[MAXNJ];
      for (int i = 0; i < nq; ++i)
      {
        rT[i] = LEFT ? dqT[i] : tsr[i];
        rN[i] = LEFT ? dqN[i] : nsr[i];
      }
      for (int c = 0; c < nJ; ++c)
      {
        dq[c] = LEFT ? av[c] * tsr[p.colRed[c]] : av[c] * tsr[p.colRed[c]] + av[c] * nsr[p.colRed[c]];
        dqT[c] = av[c] * rT[p.colRed[c]];
        dqN[c] = av[c] * rN[p.colRed[c]];
      }
    }
    for (int i = 0; i < nJ; ++i)
    {
      qv[p.colJoint[i]] += dq[i];
    }
  }
  updateSlaves();   /* the state the last RcsGraph_setState leaves in the graph */

  /* ---- results ---------------------------------------------------------------------------------- */
  for (int j = 0; j < dof; ++j)
  {
    a.q[s * dof + j] = qv[j];
  }
  if (a.dx)
  {
    for (int i = 0; i < nx; ++i)
    {
      a.dx[s * nx + i] = dx[i];
    }
  }
  if (a.dq)
  {
    for (int j = 0; j < dof; ++j)
    {
      const int c = m.jntJacobi[j];
      a.dq[s * dof + j] = c >= 0 ? dq[c] : 0.0;
    }
  }
  if (a.det)
  {
    a.det[s] = det;
  }
  if (a.extended && (a.dqTs || a.dqNs))
  {
    for (int j = 0; j < dof; ++j)
    {
      const int c = m.jntJacobi[j];
      if (a.dqTs)
      {
        a.dqTs[s * dof + j] = c >= 0 ? dqT[c] : 0.0;
      }
      if (a.dqNs)
      {
        a.dqNs[s * dof + j] = c >= 0 ? dqN[c] : 0.0;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Prioritised solver                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* J# = W J^T (J W J^T + lambda 1)^-1 for a small row-major J [mr x n] (MatNd_rwPinv,
 * src/RcsCore/Rcs_MatNd.c:1833: explicit inverse from the Banachiewicz Cholesky factor,
 * Rcs_linAlg.c:66-300). P [n x mr] receives J#; A, Ai are mr x mr work arrays. Returns the
 * determinant (0: singular, P zeroed; mr == 0: 1, as the reference). */
__device__ __noinline__ double rwPinvSmall(const double* J, int mr, int n, const double* w, double lambda,
                                           double* P, double* A, double* Ai)
{
  if (mr == 0)
  {
    return 1.0;
  }
  for (int i = 0; i < n; ++i)
  {
    for (int k = 0; k < mr; ++k)
    {
      P[i * mr + k] = w[i] * J[k * n + i];
    }
  }
  for (int r = 0; r < mr; ++r)
  {
    for (int c = 0; c < mr; ++c)
    {
      double sum = 0.0;
      for (int i = 0; i < n; ++i)
      {
        sum += J[r * n + i] * P[i * mr + c];
      }
      A[r * mr + c] = sum + (r == c ? lambda : 0.0);
    }
  }
  double det = 1.0;
  for (int i = 0; i < mr && det != 0.0; ++i)
  {
    for (int k = 0; k <= i; ++k)
    {
      double sum = A[i * mr + k];
      for (int t = 0; t < k; ++t)
      {
        sum -= A[i * mr + t] * A[k * mr + t];
      }
      if (i > k)
      {
        A[i * mr + k] = sum / A[k * mr + k];
      }
      else if (sum > 0.0)
      {
        A[i * mr + i] = sqrt(sum);
        det *= A[i * mr + i] * A[i * mr + i];
      }
      else
      {
        det = 0.0;
        break;
      }
    }
  }
  if (det == 0.0)
  {
    for (int i = 0; i < n * mr; ++i)
    {
      P[i] = 0.0;
    }
    return 0.0;
  }
  for (int k = 0; k < mr; ++k)
  {
    A[k * mr + k] = 1.0 / A[k * mr + k];
    for (int i = k + 1; i < mr; ++i)
    {
      double sum = 0.0;
      for (int t = k; t < i; ++t)
      {
        sum += A[i * mr + t] * A[t * mr + k];
      }
      A[i * mr + k] = -sum / A[i * mr + i];
    }
  }
  for (int r = 0; r < mr; ++r)
  {
    for (int c = r; c < mr; ++c)
    {
      double sum = 0.0;
      for (int t = c; t < mr; ++t)
      {
        sum += A[t * mr + r] * A[t * mr + c];
      }
      Ai[r * mr + c] = sum;
      Ai[c * mr + r] = sum;
    }
  }
  /* P <- (W J^T) inverse, row by row (A is free again: one row of scratch) */
  for (int i = 0; i < n; ++i)
  {
    for (int k = 0; k < mr; ++k)
    {
      A[k] = P[i * mr + k];
    }
    for (int c = 0; c < mr; ++c)
    {
      double sum = 0.0;
      for (int k = 0; k < mr; ++k)
      {
        sum += A[k] * Ai[k * mr + c];
      }
      P[i * mr + c] = sum;
    }
  }
  return det;
}

/* N = 1 - J# J (MatNd_nullspace, Rcs_MatNd.c:3758), n x n */
__device__ __noinline__ void nullspaceSmall(double* N, const double* P, const double* J, int mr, int n)
{
  for (int i = 0; i < n; ++i)
  {
    for (int j = 0; j < n; ++j)
    {
      double sum = 0.0;
      for (int k = 0; k < mr; ++k)
      {
        sum += P[i * mr + k] * J[k * n + j];
      }
      N[i * n + j] = (i == j) ? (1.0 - sum) : (-sum);
    }
  }
}

/*
 * IkSolverPrioRMR::solve (src/RcsCore/IkSolverPrioRMR.cpp:178-420), iterated like the loop of
 * examples/ExampleIK.cpp: tasks of priority level 0 are solved first, tasks of level 1 in the null
 * space of the first, the joint-limit gradient in the null space of both:
 *   dq1 = J1# dx1,  N1 = 1 - J1# J1,  dq2 = (J2 N1)# (dx2 - J2 J1# dx1),
 *   dq_ns = -N1 (1 - (J2 N1)# J2 N1) invWq (alpha dH);   q += dq1 + dq2 + dq_ns.
 * A singular projection (det <= 0) zeroes all three. One thread per state, literal algebra of the
 * reference in thread-local arrays (models without kinematically coupled unconstrained joints).
 */
template <int MAXNX, int MAXNJ, int MAXDOF, int NSLOTS>
__global__ void __launch_bounds__(RCSB_IK_THREADS)
ikPrioKernel(const IkArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  stageBlobsTma(sModel, a.model, a.modelBytes, sPlan, a.plan, a.planBytes);

  ModelView m;
  m.bind(sModel);
  IkPlanView p;
  p.bind(sPlan);

  const long s = (long) blockIdx.x * RCSB_IK_THREADS + threadIdx.x;
  if (s >= a.N)
  {
    return;
  }
  const int dof = m.hdr->dof;
  const int nJ = p.hdr->nJ;
  const int nx = p.hdr->nx;
  const int nTasks = p.hdr->nTasks;
  constexpr int MAXT = MAXNX / 3;

  double qv[MAXDOF];
  double ja[6 * MAXNJ];
  double ef[12 * MAXT];
  double J[MAXNX * MAXNJ];
  double dx[MAXNX], g[MAXNJ];
  double stk[(NSLOTS > 0 ? NSLOTS : 1) * 12];
  double J1[MAXNX * MAXNJ], J2[MAXNX * MAXNJ], P1[MAXNJ * MAXNX], P2[MAXNJ * MAXNX];
  double N1[MAXNJ * MAXNJ], N2[MAXNJ * MAXNJ];
  double A[MAXNX * MAXNX], Ai[MAXNX * MAXNX];
  double dx1[MAXNX], dx2[MAXNX], dq1[MAXNJ], dq2[MAXNJ], dqn[MAXNJ], tmp[MAXNJ > MAXNX ? MAXNJ : MAXNX];
  bool ok = true;

  for (int j = 0; j < dof; ++j)
  {
    qv[j] = a.q[s * dof + j];
  }
  for (int i = 0; i < nJ; ++i)
  {
    dq1[i] = dq2[i] = dqn[i] = 0.0;
  }
  const int nCpl = m.hdr->nCpl;
  auto updateSlaves = [&]() {
    for (int c = 0; c < nCpl; ++c)
    {
      qv[m.cpl[c].slave] = slaveAngle(m.cpl[c], qv[m.cpl[c].master]);
    }
  };

  for (int it = 0; it < a.iters; ++it)
  {
    updateSlaves();
    ikStateKinematics<MAXNX, MAXNJ, MAXDOF, NSLOTS>(m, p, a, s, qv, stk, ja, ef, J, dx, g);

    /* rows of the two priority levels */
    int n1 = 0, n2 = 0;
    for (int t = 0; t < nTasks; ++t)
    {
      const int lvl = t < 4 ? a.taskLevel[t] : -1;
      for (int r = 0; r < 3; ++r)
      {
        const double* src = J + (3 * t + r) * nJ;
        if (lvl == 0)
        {
          for (int i = 0; i < nJ; ++i)
          {
            J1[n1 * nJ + i] = src[i];
          }
          dx1[n1++] = dx[3 * t + r];
        }
        else if (lvl == 1)
        {
          for (int i = 0; i < nJ; ++i)
          {
            J2[n2 * nJ + i] = src[i];
          }
          dx2[n2++] = dx[3 * t + r];
        }
      }
    }

    const double det1 = rwPinvSmall(J1, n1, nJ, p.invWq, a.lambda, P1, A, Ai);
    for (int i = 0; i < nJ; ++i)
    {
      double sum = 0.0;
      for (int k = 0; k < n1; ++k)
      {
        sum += P1[i * n1 + k] * dx1[k];
      }
      dq1[i] = sum;
    }
    nullspaceSmall(N1, P1, J1, n1, nJ);
    /* J2 N1, in place of J (free from here on) */
    for (int r = 0; r < n2; ++r)
    {
      for (int c = 0; c < nJ; ++c)
      {
        double sum = 0.0;
        for (int i = 0; i < nJ; ++i)
        {
          sum += J2[r * nJ + i] * N1[i * nJ + c];
        }
        J[r * nJ + c] = sum;
      }
    }
    const double det2 = rwPinvSmall(J, n2, nJ, p.invWq, a.lambda, P2, A, Ai);
    ok = det1 > 0.0 && det2 > 0.0;
    /* (J2 J1#) dx1, the matrix first as in the reference */
    for (int r = 0; r < n2; ++r)
    {
      double acc = 0.0;
      for (int c = 0; c < n1; ++c)
      {
        double sum = 0.0;
        for (int i = 0; i < nJ; ++i)
        {
          sum += J2[r * nJ + i] * P1[i * n1 + c];
        }
        acc += sum * dx1[c];
      }
      tmp[r] = dx2[r] - acc;
    }
    for (int i = 0; i < nJ; ++i)
    {
      double sum = 0.0;
      for (int k = 0; k < n2; ++k)
      {
        sum += P2[i * n2 + k] * tmp[k];
      }
      dq2[i] = sum;
    }
    /* null space of both levels: N1 (1 - (J2 N1)# J2 N1) applied to invWq (-alpha dH) */
    nullspaceSmall(N2, P2, J, n2, nJ);
    for (int i = 0; i < nJ; ++i)
    {
      tmp[i] = p.invWq[i] * (-g[i]);       /* dHA = invWq o (alpha dH); g = -alpha dH */
    }
    for (int i = 0; i < nJ; ++i)
    {
      double sum = 0.0;
      for (int c = 0; c < nJ; ++c)
      {
        double nn = 0.0;
        for (int k = 0; k < nJ; ++k)
        {
          nn += N1[i * nJ + k] * N2[k * nJ + c];
        }
        sum += nn * tmp[c];
      }
      dqn[i] = -1.0 * sum;
    }
    if (!ok)
    {
      for (int i = 0; i < nJ; ++i)
      {
        dq1[i] = dq2[i] = dqn[i] = 0.0;
      }
    }
    for (int i = 0; i < nJ; ++i)
    {
      qv[p.colJoint[i]] += (dq1[i] + dq2[i]) + dqn[i];
    }
  }
  updateSlaves();

  for (int j = 0; j < dof; ++j)
  {
    a.q[s * dof + j] = qv[j];
    const int c = m.jntJacobi[j];
    if (a.dq)
    {
      a.dq[s * dof + j] = c >= 0 ? dq1[c] : 0.0;
    }
    if (a.dqTs)
    {
      a.dqTs[s * dof + j] = c >= 0 ? dq2[c] : 0.0;
    }
    if (a.dqNs)
    {
      a.dqNs[s * dof + j] = c >= 0 ? dqn[c] : 0.0;
    }
  }
  if (a.dx)
  {
    for (int i = 0; i < nx; ++i)
    {
      a.dx[s * nx + i] = dx[i];
    }
  }
  if (a.det)
  {
    a.det[s] = ok ? 1.0 : 0.0;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* General task stacks                                                                         */
/* ------------------------------------------------------------------------------------------ */

struct IkTasksPlanView
{
  const RcsbIkTasksPlanHeader* hdr;
  const int* task;
  const int* bodySlotStart;
  const int* bodySlotList;
  const unsigned long long* slotMask;
  const unsigned char* bodyInCog;
  const int* colJoint;
  const double* invWq;
  const double* jlGain;
  const double* q0;

  __device__ __forceinline__ void bind(const char* blob)
  {
    hdr = reinterpret_cast<const RcsbIkTasksPlanHeader*>(blob);
    task = reinterpret_cast<const int*>(blob + hdr->offTask);
    bodySlotStart = reinterpret_cast<const int*>(blob + hdr->offBodySlotStart);
    bodySlotList = reinterpret_cast<const int*>(blob + hdr->offBodySlotList);
    slotMask = reinterpret_cast<const unsigned long long*>(blob + hdr->offSlotMask);
    bodyInCog = reinterpret_cast<const unsigned char*>(blob + hdr->offBodyInCog);
    colJoint = reinterpret_cast<const int*>(blob + hdr->offColJoint);
    invWq = reinterpret_cast<const double*>(blob + hdr->offInvWq);
    jlGain = reinterpret_cast<const double*>(blob + hdr->offJlGain);
    q0 = reinterpret_cast<const double*>(blob + hdr->offQ0);
  }
};

/*
 * rcsb_ik_tasks_rmr: one thread per state. Per iteration: forward kinematics of the whole tree
 * (frames of the bodies the tasks refer to, axis and origin of every Jacobian column, centre of
 * gravity and its Jacobian if a COG task asks for it), the task values, errors and Jacobian rows of
 * every task in the reference's formulations (see rcsb.h), then dq = J# dx + (1 - J# J) invWq
 * (-alpha dH) with J# from MatNd_rwPinv (IkSolverRMR.cpp:415-596), q += dq.
 */
template <int MAXNX, int MAXNJ, int MAXDOF, int MAXSLOT, int NSLOTS, bool CONSTRAINT>
__global__ void __launch_bounds__(RCSB_IK_THREADS)
ikTasksKernel(const IkArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  stageBlobsTma(sModel, a.model, a.modelBytes, sPlan, a.plan, a.planBytes);

  ModelView m;
  m.bind(sModel);
  IkTasksPlanView p;
  p.bind(sPlan);

  const long s = (long) blockIdx.x * RCSB_IK_THREADS + threadIdx.x;
  if (s >= a.N)
  {
    return;
  }
  const int nb = m.hdr->nb;
  const int dof = m.hdr->dof;
  const int nJ = p.hdr->nJ;
  const int nx = p.hdr->nx;
  const int nTasks = p.hdr->nTasks;
  const bool wantCog = p.hdr->cogRoot >= -1 && p.hdr->cogRoot != -2;

  double qv[MAXDOF];
  double ja[6 * MAXNJ];            /* axis, origin of every Jacobian column */
  double fr[12 * MAXSLOT];         /* frames of the task bodies; slot 0 = world */
  double J[MAXNX * MAXNJ], P[MAXNJ * MAXNX], A[MAXNX * MAXNX], Ai[MAXNX * MAXNX];
  double Jc[3 * MAXNJ];            /* sum of m_b JT_b(COM_b) over the COG bodies */
  double dx[MAXNX], g[MAXNJ], Jg[MAXNX], dq[MAXNJ];
  double stk[(NSLOTS > 0 ? NSLOTS : 1) * 12];
  double det = 0.0;
  int nViol = 0;

  for (int j = 0; j < dof; ++j)
  {
    qv[j] = a.q[s * dof + j];
  }
  for (int i = 0; i < nJ; ++i)
  {
    dq[i] = 0.0;
  }
  for (int i = 0; i < nx; ++i)
  {
    dx[i] = 0.0;
  }
  /* dq = P dx + (g - P (J g)) for the first `rows` rows of J and dx */
  auto solveStep = [&](int rows) -> double {
    const double d = rwPinvSmall(J, rows, nJ, p.invWq, a.lambda, P, A, Ai);
    if (d == 0.0)
    {
      for (int i = 0; i < nJ; ++i)
      {
        dq[i] = 0.0;
      }
      return d;
    }
    for (int r = 0; r < rows; ++r)
    {
      double sum = 0.0;
      for (int i = 0; i < nJ; ++i)
      {
        sum += J[r * nJ + i] * g[i];
      }
      Jg[r] = sum;
    }
    for (int i = 0; i < nJ; ++i)
    {
      double ts = 0.0, ns = 0.0;
      for (int c = 0; c < rows; ++c)
      {
        ts += P[i * rows + c] * dx[c];
        ns += P[i * rows + c] * Jg[c];
      }
      dq[i] = ts + (g[i] - ns);
    }
    return d;
  };
  {
    Frame id;
    setIdentity(id);
    for (int i = 0; i < 3; ++i)
    {
      fr[i] = id.o[i];
    }
    for (int i = 0; i < 9; ++i)
    {
      fr[3 + i] = id.R[i];
    }
  }
  const int nCpl = m.hdr->nCpl;
  auto updateSlaves = [&]() {
    for (int c = 0; c < nCpl; ++c)
    {
      qv[m.cpl[c].slave] = slaveAngle(m.cpl[c], qv[m.cpl[c].master]);
    }
  };

  for (int it = 0; it < a.iters; ++it)
  {
    updateSlaves();
    /* ---- forward kinematics --------------------------------------------------------------- */
    double mTot = 0.0, mc[3] = {0.0, 0.0, 0.0};
    if (wantCog)
    {
      for (int i = 0; i < 3 * nJ; ++i)
      {
        Jc[i] = 0.0;
      }
    }
    Frame f, keep;
    setIdentity(f);
    keep = f;
    for (int b = 0; b < nb; ++b)
    {
      const int ld = m.bodyLoad[b];
      const int bflags = m.bodyFlags[b];
      if (ld == RCSB_LOAD_IDENT)
      {
        setIdentity(f);
      }
      else if (NSLOTS > 0 && ld >= 0)
      {
        const double* sp = stk + ld * 12;
        for (int i = 0; i < 3; ++i)
        {
          f.o[i] = sp[i];
        }
        for (int i = 0; i < 9; ++i)
        {
          f.R[i] = sp[3 + i];
        }
      }
      if (bflags & RCSB_BF_LEAF)
      {
        keep = f;
      }
      const int j0 = m.bodyJntStart[b];
      const int j1 = j0 + m.bodyJntCount[b];
      for (int j = j0; j < j1; ++j)
      {
        const int jf = m.jntFlags[j];
        if (jf & RCSB_JF_HAS_AJP)
        {
          applyRelF(f, m.jntAJP + 12 * j, jf);
        }
        const double qj = qv[j];
        const int dir = m.jntType[j] % 3;
        double ax[3];
        if (jf & RCSB_JF_ROT)
        {
          double sn, cs;
          sincos(qj, &sn, &cs);
          rotateAboutAxis(f, dir, sn, cs);
          axisOf(f, dir, ax);
        }
        else
        {
          axisOf(f, dir, ax);
          f.o[0] += qj * ax[0];
          f.o[1] += qj * ax[1];
          f.o[2] += qj * ax[2];
        }
        const int col = m.jntJacobi[j];
        if (col >= 0)
        {
          double* d = ja + 6 * col;
          d[0] = ax[0];
          d[1] = ax[1];
          d[2] = ax[2];
          d[3] = f.o[0];
          d[4] = f.o[1];
          d[5] = f.o[2];
        }
      }
      if (bflags & RCSB_BF_HAS_ABP)
      {
        applyRelF(f, m.bodyABP + 12 * b, bflags);
      }
      if (NSLOTS > 0)
      {
        const int sv = m.bodySave[b];
        if (sv >= 0)
        {
          double* sp = stk + sv * 12;
          for (int i = 0; i < 3; ++i)
          {
            sp[i] = f.o[i];
          }
          for (int i = 0; i < 9; ++i)
          {
            sp[3 + i] = f.R[i];
          }
        }
      }
      for (int k = p.bodySlotStart[b]; k < p.bodySlotStart[b + 1]; ++k)
      {
        double* d = fr + 12 * p.bodySlotList[k];
        for (int i = 0; i < 3; ++i)
        {
          d[i] = f.o[i];
        }
        for (int i = 0; i < 9; ++i)
        {
          d[3 + i] = f.R[i];
        }
      }
      /* centre of gravity and sum_b m_b JT_b(COM_b) (RcsGraph_COG_Body / RcsGraph_COGJacobian_Body,
       * Rcs_kinematics.c:938, :1004): the body's ancestor joints have all been visited */
      if (wantCog && (bflags & RCSB_BF_HAS_MASS) && p.bodyInCog[b])
      {
        const double mass = m.bodyMass[b];
        const double* cl = m.bodyCom + 3 * b;
        const double* R = f.R;
        const double c[3] = {f.o[0] + (R[0] * cl[0] + R[3] * cl[1] + R[6] * cl[2]),
                             f.o[1] + (R[1] * cl[0] + R[4] * cl[1] + R[7] * cl[2]),
                             f.o[2] + (R[2] * cl[0] + R[5] * cl[1] + R[8] * cl[2])};
        mTot += mass;
        mc[0] += mass * c[0];
        mc[1] += mass * c[1];
        mc[2] += mass * c[2];
        for (int j = m.bodyLastJnt[b]; j >= 0; j = m.jntPrev[j])
        {
          const int col = m.jntJacobi[j];
          if (col < 0)
          {
            continue;
          }
          const double* jc = ja + 6 * col;
          if (m.jntFlags[j] & RCSB_JF_ROT)
          {
            const double d0 = c[0] - jc[3], d1 = c[1] - jc[4], d2 = c[2] - jc[5];
            Jc[col] += mass * (jc[1] * d2 - jc[2] * d1);
            Jc[nJ + col] += mass * (jc[2] * d0 - jc[0] * d2);
            Jc[2 * nJ + col] += mass * (jc[0] * d1 - jc[1] * d0);
          }
          else
          {
            Jc[col] += mass * jc[0];
            Jc[nJ + col] += mass * jc[1];
            Jc[2 * nJ + col] += mass * jc[2];
          }
        }
      }
      if (bflags & RCSB_BF_LEAF)
      {
        f = keep;
      }
    }

    /* ---- task values, errors and Jacobian rows --------------------------------------------- */
    for (int i = 0; i < nx * nJ; ++i)
    {
      J[i] = 0.0;
    }
    const double invM = mTot > 0.0 ? 1.0 / mTot : 0.0;
    for (int t = 0; t < nTasks; ++t)
    {
      const int* tk = p.task + 8 * t;
      const int kind = tk[0], row = tk[1], sE = tk[3], sR = tk[4], sF = tk[5], aux = tk[6];
      const double* xd = a.xDes + s * nx + row;
      const double* fE = fr + 12 * sE;
      const double* fR = fr + 12 * sR;
      const double* fF = fr + 12 * sF;
      double* Jr = J + row * nJ;
      if (kind == RCSB_TASK_JOINT)
      {
        double x = qv[tk[7]];
        Jr[aux] = 1.0;
        if (tk[3] > 0)
        {
          /* TaskJoint with a reference joint: x = q + refGain q_ref (TaskJoint.cpp:227-236, :256-265) */
          const double gain = __hiloint2double(tk[5], tk[4]);
          x += gain * qv[tk[3] - 1];
          Jr[m.jntJacobi[tk[3] - 1]] = gain;
        }
        dx[row] = xd[0] - x;
        continue;
      }
      if (kind >= RCSB_TASK_COGX && kind <= RCSB_TASK_COGZ)
      {
        if (sR == 0)
        {
          dx[row] = xd[0] - mc[aux] * invM;
          for (int k = 0; k < nJ; ++k)
          {
            Jr[k] = Jc[aux * nJ + k] * invM;
          }
          continue;
        }
        /* relative to a reference body (TaskCOM3D::computeX / computeJ): x = A_R (c - o_R);
         * J = A_R J_cog for a fixed reference, else A_R (J_cog - JT_R + (c - o_R) x JR_R[:, k]) */
        const double* A = fR + 3 + 3 * aux;          /* row `aux` of A_R */
        const double c[3] = {mc[0] * invM, mc[1] * invM, mc[2] * invM};
        const double r[3] = {c[0] - fR[0], c[1] - fR[1], c[2] - fR[2]};
        dx[row] = xd[0] - (A[0] * r[0] + A[1] * r[1] + A[2] * r[2]);
        const unsigned long long mRc = p.slotMask[sR];
        for (int k = 0; k < nJ; ++k)
        {
          double v[3] = {Jc[k] * invM, Jc[nJ + k] * invM, Jc[2 * nJ + k] * invM};
          if ((mRc >> k) & 1ull)
          {
            const double* jc = ja + 6 * k;
            if (m.jntFlags[p.colJoint[k]] & RCSB_JF_ROT)
            {
              /* JT_R column: a x (o_R - o_k); MatNd_columnCrossProductSelf(JR_R, r): r x a (the frame
               * of the reference turns under the centre of gravity) */
              const double d0 = fR[0] - jc[3], d1 = fR[1] - jc[4], d2 = fR[2] - jc[5];
              v[0] -= jc[1] * d2 - jc[2] * d1;
              v[1] -= jc[2] * d0 - jc[0] * d2;
              v[2] -= jc[0] * d1 - jc[1] * d0;
              v[0] += r[1] * jc[2] - r[2] * jc[1];
              v[1] += r[2] * jc[0] - r[0] * jc[2];
              v[2] += r[0] * jc[1] - r[1] * jc[0];
            }
            else
            {
              v[0] -= jc[0];
              v[1] -= jc[1];
              v[2] -= jc[2];
            }
          }
          Jr[k] = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
        }
        continue;
      }
      const unsigned long long mE = p.slotMask[sE], mR = p.slotMask[sR], mF = p.slotMask[sF];
      if (kind == RCSB_TASK_XYZ)
      {
        /* TaskPosition3D::computeX: I_r = o_E - o_R, rotated into refFrame (or refBdy) */
        double r[3] = {fE[0], fE[1], fE[2]};
        const double* Arot = nullptr;
        if (sR > 0)
        {
          r[0] -= fR[0];
          r[1] -= fR[1];
          r[2] -= fR[2];
          Arot = (sF > 0 && sF != sR) ? fF + 3 : fR + 3;
        }
        else if (sF > 0)
        {
          Arot = fF + 3;
        }
        double x[3] = {r[0], r[1], r[2]};
        if (Arot)
        {
          x[0] = Arot[0] * r[0] + Arot[1] * r[1] + Arot[2] * r[2];
          x[1] = Arot[3] * r[0] + Arot[4] * r[1] + Arot[5] * r[2];
          x[2] = Arot[6] * r[0] + Arot[7] * r[1] + Arot[8] * r[2];
        }
        dx[row] = xd[0] - x[0];
        dx[row + 1] = xd[1] - x[1];
        dx[row + 2] = xd[2] - x[2];
        /* RcsGraph_3dPosJacobian: JT_2, or A (JT_2 - JT_1 + r_12 x JR_3) */
        for (int k = 0; k < nJ; ++k)
        {
          const double* jc = ja + 6 * k;
          const bool isRot = (m.jntFlags[p.colJoint[k]] & RCSB_JF_ROT) != 0;
          auto colT = [&](const double* fo, bool on, double (&c)[3]) {
            c[0] = c[1] = c[2] = 0.0;
            if (on)
            {
              if (isRot)
              {
                const double d0 = fo[0] - jc[3], d1 = fo[1] - jc[4], d2 = fo[2] - jc[5];
                c[0] = jc[1] * d2 - jc[2] * d1;
                c[1] = jc[2] * d0 - jc[0] * d2;
                c[2] = jc[0] * d1 - jc[1] * d0;
              }
              else
              {
                c[0] = jc[0];
                c[1] = jc[1];
                c[2] = jc[2];
              }
            }
          };
          double v[3];
          colT(fE, sE > 0 && ((mE >> k) & 1ull), v);
          if (sR > 0)
          {
            double j1[3];
            colT(fR, (mR >> k) & 1ull, j1);
            v[0] -= j1[0];
            v[1] -= j1[1];
            v[2] -= j1[2];
            if (sF > 0 && isRot && ((mF >> k) & 1ull))
            {
              const double w0 = jc[0], w1 = jc[1], w2 = jc[2];
              v[0] += r[1] * w2 - r[2] * w1;
              v[1] += -r[0] * w2 + r[2] * w0;
              v[2] += r[0] * w1 - r[1] * w0;
            }
          }
          if (Arot)
          {
            const double x0 = v[0], y0 = v[1], z0 = v[2];
            v[0] = Arot[0] * x0 + Arot[1] * y0 + Arot[2] * z0;
            v[1] = Arot[3] * x0 + Arot[4] * y0 + Arot[5] * z0;
            v[2] = Arot[6] * x0 + Arot[7] * y0 + Arot[8] * z0;
          }
          Jr[k] = v[0];
          Jr[nJ + k] = v[1];
          Jr[2 * nJ + k] = v[2];
        }
        continue;
      }
      /* RCSB_TASK_ABC / POLAR*: A_ER = A_E A_R^T (Task::computeRelativeRotationMatrix, Task.cpp:1204);
       * ABC: its Euler angles, error as TaskEuler3D::computeDX */
      {
        double Aer[9];
        const double* AE = fE + 3;
        const double* AR = fR + 3;
        for (int i = 0; i < 3; ++i)
        {
          for (int j = 0; j < 3; ++j)
          {
            Aer[3 * i + j] = AE[3 * i] * AR[3 * j] + AE[3 * i + 1] * AR[3 * j + 1] + AE[3 * i + 2] * AR[3 * j + 2];
          }
        }
        const bool polar = kind >= RCSB_TASK_POLARX && kind <= RCSB_TASK_POLARZ;
        if (polar)
        {
          /* TaskPolar2D: x = polar angles of row `aux` of A_ER, J = two rows of
           * RcsGraph_3dOmegaJacobian(ef, refBdy, ef) (TaskPolar2D.cpp:207-243); sF is the effector */
          const double xdl[2] = {xd[0], xd[1]};
          polarTaskError(dx + row, xdl, Aer, aux);
        }
        else
        {
          double xc[3];
          const double xdl[3] = {xd[0], xd[1], xd[2]};
          toEuler(xc, Aer);
          eulerTaskError(dx + row, xdl, xc);
        }
        /* RcsGraph_3dOmegaJacobian: JR_2, A_1 JR_2 (fixed reference), or A_3 (JR_2 - JR_1) */
        const bool fixedRef = sR > 0 && sF == sR && mR == 0ull;
        const bool relative = !(sR == 0 && sF == 0) && !fixedRef;
        const double* Arot = fixedRef ? fR + 3 : ((relative && sF > 0) ? fF + 3 : nullptr);
        for (int k = 0; k < nJ; ++k)
        {
          const double* jc = ja + 6 * k;
          const bool isRot = (m.jntFlags[p.colJoint[k]] & RCSB_JF_ROT) != 0;
          double sgn = (sE > 0 && isRot && ((mE >> k) & 1ull)) ? 1.0 : 0.0;
          if (relative && sR > 0 && isRot && ((mR >> k) & 1ull))
          {
            sgn -= 1.0;
          }
          double v[3] = {sgn * jc[0], sgn * jc[1], sgn * jc[2]};
          if (Arot)
          {
            const double x0 = v[0], y0 = v[1], z0 = v[2];
            v[0] = Arot[0] * x0 + Arot[1] * y0 + Arot[2] * z0;
            v[1] = Arot[3] * x0 + Arot[4] * y0 + Arot[5] * z0;
            v[2] = Arot[6] * x0 + Arot[7] * y0 + Arot[8] * z0;
          }
          if (polar)
          {
            Jr[k] = aux == 0 ? v[1] : v[0];
            Jr[nJ + k] = aux == 2 ? v[1] : v[2];
          }
          else
          {
            Jr[k] = v[0];
            Jr[nJ + k] = v[1];
            Jr[2 * nJ + k] = v[2];
          }
        }
      }
    }

    /* ---- dq = J# dx + (1 - J# J) invWq (-alpha dH) ------------------------------------------ */
    for (int i = 0; i < nJ; ++i)
    {
      const double delta = qv[p.colJoint[i]] - p.q0[i];
      g[i] = p.invWq[i] * (-(a.alpha * (p.jlGain[i] * delta)));
    }
    det = solveStep(nx);
    nViol = 0;
    if (det == 0.0)
    {
      continue;
    }
    if (CONSTRAINT)
    {
      /* IkSolverConstraintRMR::solveRightInverse (IkSolverConstraintRMR.cpp:156-262): every free joint
       * that the step would leave closer than 2 degrees to a limit (and that is not a task of its
       * own) becomes a joint constraint: a unit row with a target velocity of zero, or of a tenth of
       * 0.1 degrees back towards the limit margin if it is deeper than that; then ONE more solve with
       * these rows appended in joint order. */
      const double margin = (3.14159265358979323846 / 180.0) * 2.0;
      const double baumgarte = (3.14159265358979323846 / 180.0) * 0.1;
      int rows = nx;
      for (int j = 0; j < dof; ++j)
      {
        const int col = m.jntJacobi[j];
        if (col < 0)
        {
          continue;
        }
        bool ownTask = false;
        for (int t = 0; t < nTasks; ++t)
        {
          ownTask = ownTask || (p.task[8 * t] == RCSB_TASK_JOINT && p.task[8 * t + 7] == j);
        }
        if (ownTask)
        {
          continue;
        }
        const double x = qv[j] + dq[col];
        const double lo = m.jntQMin[j] + margin, hi = m.jntQMax[j] - margin;
        double target;
        if (x < lo)
        {
          target = (lo - x > baumgarte) ? 0.1 * baumgarte : 0.0;
        }
        else if (x > hi)
        {
          target = (x - hi > baumgarte) ? -0.1 * baumgarte : 0.0;
        }
        else
        {
          continue;
        }
        ++nViol;
        if (rows < MAXNX)
        {
          for (int k = 0; k < nJ; ++k)
          {
            J[rows * nJ + k] = 0.0;
          }
          J[rows * nJ + col] = 1.0;
          dx[rows] = target;
          ++rows;
        }
      }
      if (nViol > 0)
      {
        if (nx + nViol > MAXNX)
        {
          nViol = -1;      /* does not fit: the state keeps the unconstrained step and says so */
        }
        else
        {
          det = solveStep(rows);
        }
      }
    }
    for (int i = 0; i < nJ; ++i)
    {
      qv[p.colJoint[i]] += dq[i];
    }
  }
  updateSlaves();

  for (int j = 0; j < dof; ++j)
  {
    a.q[s * dof + j] = qv[j];
    if (a.dq)
    {
      const int c = m.jntJacobi[j];
      a.dq[s * dof + j] = c >= 0 ? dq[c] : 0.0;
    }
  }
  if (a.dx)
  {
    for (int i = 0; i < nx; ++i)
    {
      a.dx[s * nx + i] = dx[i];
    }
  }
  if (a.det)
  {
    a.det[s] = det;
  }
  if (CONSTRAINT && a.nViol)
  {
    a.nViol[s] = (double) nViol;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Single-chain fast path                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* One step of the root-to-effector program: a constant relative transform, or the motion of a
 * joint that is not a Jacobian column (constrained: its value never changes during the solve). */
template <bool PERM>
__device__ __forceinline__ void chainStep(Frame& f, const int4 st, const double* __restrict__ xf,
                                          const double* __restrict__ qRow)
{
  if (st.x == 0)
  {
    applyRelF(f, xf + 12 * st.y, st.z);
    return;
  }
  if (PERM && st.x == 2)
  {
    /* rows shifted by st.y (1 or 2): new row i = old row (i + st.y) mod 3 */
    double t[9];
#pragma unroll
    for (int i = 0; i < 9; ++i)
    {
      t[i] = f.R[i];
    }
    if (st.y == 1)
    {
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        f.R[i] = t[(i + 3) % 9];
      }
    }
    else
    {
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        f.R[i] = t[(i + 6) % 9];
      }
    }
    return;
  }
  const double qj = qRow[st.z];
  if (st.y < 3)
  {
    double sn, cs;
    sincos(qj, &sn, &cs);
    if (PERM)
    {
      rotateRows01(f, sn, cs);
    }
    else
    {
      rotateAboutAxis(f, st.w, sn, cs);
    }
  }
  else
  {
    double ax[3];
    axisOf(f, st.w, ax);
    f.o[0] += qj * ax[0];
    f.o[1] += qj * ax[1];
    f.o[2] += qj * ax[2];
  }
}

/*
 * All tasks on one effector, NC <= 8 Jacobian columns on its chain: everything the iteration
 * needs stays in registers (column axes and Jacobian columns 6 NC doubles, the lower triangle
 * of J invWq J^T + lambda 1, NX (NX + 1) / 2 doubles), all loops over columns and task rows
 * are unrolled at compile time, bodies off the chain are never visited.
 *
 * Same step as the general kernel, evaluated as
 *     dq = g + invWq J^T y,   (J invWq J^T + lambda 1) y = dx - J g,   g = invWq (-alpha dH)
 * which equals J# dx + (1 - J# J) g of IkSolverRMR.cpp:415-596 without forming J# and the
 * null-space projector; y comes from the same Cholesky factor (forward / backward
 * substitution instead of the explicit inverse L^-T L^-1). The orientation error uses the
 * effector's rotation matrix directly instead of rebuilding it from its Euler angles
 * (TaskEuler3D.cpp:110, :243): identical up to rounding away from gimbal lock.
 */
#ifndef RCSB_IK_CHAIN_MIN_BLOCKS
#define RCSB_IK_CHAIN_MIN_BLOCKS 2
#endif
/* LEFT: the step of IkSolverRMR::solveLeftInverse (IkSolverRMR.cpp:205-272) instead: with
 * J_r = J invWq, dq = invWq o (g + (J_r^T J_r + lambda 1)^-1 J_r^T (dx - J_r g)), the NC x NC
 * system factored like the NX x NX one of the right inverse. */
/* MODE 0: column types, the order of the two tasks and the chain program are read from the plan at
 * run time. MODE 1 / 2: every column is a rotation, the program has no joint steps and the position
 * task comes first / second (what the plan builder reports in IkArgs::chainFlags): no per-column type
 * branches and no row selects — the merges behind those branches were 18 % of the executed
 * instructions as register moves —, the sines / cosines of all columns are evaluated together before
 * the walk (sincosBatch), and the program is not interpreted: every segment is ONE relative transform
 * folded on the host (RcsbIkPlanHeader::offSegXf), so the walk is straight-line code. */
template <int NC, int NX, bool LEFT = false, int MODE = 0>
__global__ void __launch_bounds__(RCSB_IK_THREADS, RCSB_IK_CHAIN_MIN_BLOCKS)
ikChainKernel(const IkArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  double* sAd = reinterpret_cast<double*>(sPlan + a.planBytes);   /* [9][RCSB_IK_THREADS] */
  stageBlobsTma(sModel, a.model, a.modelBytes, sPlan, a.plan, a.planBytes);

  ModelView m;
  m.bind(sModel);
  IkPlanView p;
  p.bind(sPlan);

  const int tid = threadIdx.x;
  const long s = (long) blockIdx.x * RCSB_IK_THREADS + tid;
  if (s >= a.N)
  {
    return;
  }

  constexpr int NT = NX / 3;
  constexpr int NA = NX * (NX + 1) / 2;
  constexpr bool ALLROT = MODE != 0;
  const int dof = m.hdr->dof;
  const bool xyzFirst = MODE == 0 ? (p.taskKind[0] == RCSB_TASK_XYZ) : (MODE == 1);
  const bool hasAbc = (NT == 2) || !xyzFirst;
  const bool hasXyz = (NT == 2) || xyzFirst;
  double* qRow = a.q + s * dof;
  const double* xd = a.xDes + s * NX;

  double qc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
  {
    qc[c] = qRow[p.chainJoint[c]];
  }

  /* desired position / rotation; the rotation is parked in shared memory */
  double xPos[3] = {0.0, 0.0, 0.0};
  if (hasXyz)
  {
    const int o = (NT == 2 && !xyzFirst) ? 3 : 0;
    xPos[0] = xd[o];
    xPos[1] = xd[o + 1];
    xPos[2] = xd[o + 2];
  }
  if (hasAbc)
  {
    const int o = (NT == 2 && xyzFirst) ? 3 : 0;
    const double ea[3] = {xd[o], xd[o + 1], xd[o + 2]};
    double Ad[9];
    fromEuler(Ad, ea);
#pragma unroll
    for (int i = 0; i < 9; ++i)
    {
      sAd[i * RCSB_IK_THREADS + tid] = Ad[i];
    }
  }

  double dx[NX], dq[NC];
#pragma unroll
  for (int i = 0; i < NX; ++i)
  {
    dx[i] = 0.0;
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
  {
    dq[c] = 0.0;
  }
  double det = 0.0;
  unsigned long long singular = 0ull;

  for (int it = 0; it < a.iters; ++it)
  {
    /* ---- forward kinematics along the chain ------------------------------------------------ */
    Frame f;
    setIdentity(f);
    double ax[NC][3], jo[NC][3];
    bool isRot[NC];
    double snAll[NC], csAll[NC];
    if constexpr (ALLROT)
    {
      if (!sincosBatch<NC>(qc, snAll, csAll))
      {
        /* an argument beyond the fast range of the reduction (or not finite): the library's path */
#pragma unroll
        for (int c = 0; c < NC; ++c)
        {
          sincos(qc[c], &snAll[c], &csAll[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      if constexpr (ALLROT)
      {
        applyRel(f, p.segXf + 12 * c);   /* the segment's steps as one transform, no interpretation */
      }
      else
      {
        for (int k = p.progStart[c]; k < p.progStart[c + 1]; ++k)
        {
          chainStep<true>(f, p.prog[k], p.chainXf, qRow);
        }
      }
      const int ctype = ALLROT ? 0 : p.chainType[c];
      isRot[c] = ALLROT || (ctype & 15) < 3;
      if (isRot[c])
      {
        double sn, cs;
        if (ALLROT)
        {
          sn = snAll[c];
          cs = csAll[c];
        }
        else
        {
          sincos(qc[c], &sn, &cs);
        }
        rotateRows01(f, sn, cs);
        ax[c][0] = f.R[6];         /* the axis of a rotation is row 2 of the shifted frame */
        ax[c][1] = f.R[7];
        ax[c][2] = f.R[8];
      }
      else
      {
        axisOf(f, ctype >> 4, ax[c]);
        f.o[0] += qc[c] * ax[c][0];
        f.o[1] += qc[c] * ax[c][1];
        f.o[2] += qc[c] * ax[c][2];
      }
      jo[c][0] = f.o[0];
      jo[c][1] = f.o[1];
      jo[c][2] = f.o[2];
    }
    if constexpr (ALLROT)
    {
      applyRel(f, p.segXf + 12 * NC);
    }
    else
    {
      for (int k = p.progStart[NC]; k < p.progStart[NC + 1]; ++k)
      {
        chainStep<true>(f, p.prog[k], p.chainXf, qRow);
      }
    }

    /* ---- task error (ControllerBase::computeDX) ---------------------------------------------- */
    double ePos[3] = {0.0, 0.0, 0.0}, eRot[3] = {0.0, 0.0, 0.0};
    if (hasXyz)
    {
      ePos[0] = xPos[0] - f.o[0];
      ePos[1] = xPos[1] - f.o[1];
      ePos[2] = xPos[2] - f.o[2];
    }
    if (hasAbc)
    {
      double Ad[9];
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        Ad[i] = sAd[i * RCSB_IK_THREADS + tid];
      }
      rotationTaskErrorOfFrame(eRot, Ad, f.R);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
      dx[i] = xyzFirst ? ePos[i] : eRot[i];
      if (NT == 2)
      {
        dx[3 + i] = xyzFirst ? eRot[i] : ePos[i];
      }
    }

    /* ---- Jacobian columns in place: jo <- translation rows, ax <- rotation rows ------------- */
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      if (isRot[c])
      {
        const double d0 = f.o[0] - jo[c][0], d1 = f.o[1] - jo[c][1], d2 = f.o[2] - jo[c][2];
        jo[c][0] = ax[c][1] * d2 - ax[c][2] * d1;
        jo[c][1] = ax[c][2] * d0 - ax[c][0] * d2;
        jo[c][2] = ax[c][0] * d1 - ax[c][1] * d0;
      }
      else
      {
        jo[c][0] = ax[c][0];
        jo[c][1] = ax[c][1];
        jo[c][2] = ax[c][2];
        ax[c][0] = ax[c][1] = ax[c][2] = 0.0;
      }
    }

    if (LEFT)
    {
      constexpr int NB = NC * (NC + 1) / 2;
      double B[NB], rl[NX], gl[NC], bl[NC], invl[NC];
#pragma unroll
      for (int i = 0; i < NX; ++i)
      {
        rl[i] = 0.0;
      }
      /* columns of J_r = J invWq in place; r = dx - J_r g */
#pragma unroll
      for (int c = 0; c < NC; ++c)
      {
        const double w = p.chainW[c];
        gl[c] = w * (-(a.alpha * (p.chainGain[c] * (qc[c] - p.chainQ0[c]))));
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          jo[c][i] *= w;
          ax[c][i] *= w;
          rl[i] += (xyzFirst ? jo[c][i] : ax[c][i]) * gl[c];
          if (NT == 2)
          {
            rl[3 + i] += (xyzFirst ? ax[c][i] : jo[c][i]) * gl[c];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NX; ++i)
      {
        rl[i] = dx[i] - rl[i];
      }
      /* B = J_r^T J_r + lambda 1 (lower triangle), b = J_r^T r */
#pragma unroll
      for (int c = 0; c < NC; ++c)
      {
#pragma unroll
        for (int d = 0; d <= c; ++d)
        {
          double sum = 0.0;
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            sum += jo[c][i] * jo[d][i];
          }
          if (NT == 2)
          {
#pragma unroll
            for (int i = 0; i < 3; ++i)
            {
              sum += ax[c][i] * ax[d][i];
            }
          }
          else if (!xyzFirst)
          {
            /* a single orientation task: its rows are the rotation rows */
            sum = ax[c][0] * ax[d][0] + ax[c][1] * ax[d][1] + ax[c][2] * ax[d][2];
          }
          B[c * (c + 1) / 2 + d] = sum + (c == d ? a.lambda : 0.0);
        }
        double sb = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          sb += (xyzFirst ? jo[c][i] : ax[c][i]) * rl[i];
          if (NT == 2)
          {
            sb += (xyzFirst ? ax[c][i] : jo[c][i]) * rl[3 + i];
          }
        }
        bl[c] = sb;
      }
      /* Banachiewicz Cholesky of the NC x NC matrix, det = prod of the pivots */
      bool okl = true;
      det = 1.0;
#pragma unroll
      for (int i = 0; i < NC; ++i)
      {
#pragma unroll
        for (int k = 0; k <= i; ++k)
        {
          double sum = B[i * (i + 1) / 2 + k];
#pragma unroll
          for (int t = 0; t < k; ++t)
          {
            sum -= B[i * (i + 1) / 2 + t] * B[k * (k + 1) / 2 + t];
          }
          if (k < i)
          {
            B[i * (i + 1) / 2 + k] = sum * invl[k];
          }
          else
          {
            okl = okl && (sum > 0.0);
            const double piv = okl ? sum : 1.0;
            invl[i] = rsqrt(piv);
            B[i * (i + 1) / 2 + i] = piv * invl[i];
            det *= piv;
          }
        }
      }
      if (!okl)
      {
        det = 0.0;
        singular |= 1ull << it;
#pragma unroll
        for (int c = 0; c < NC; ++c)
        {
          dq[c] = 0.0;
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < NC; ++i)
      {
        double sum = bl[i];
#pragma unroll
        for (int t = 0; t < i; ++t)
        {
          sum -= B[i * (i + 1) / 2 + t] * bl[t];
        }
        bl[i] = sum * invl[i];
      }
#pragma unroll
      for (int i = NC - 1; i >= 0; --i)
      {
        double sum = bl[i];
#pragma unroll
        for (int t = i + 1; t < NC; ++t)
        {
          sum -= B[t * (t + 1) / 2 + i] * bl[t];
        }
        bl[i] = sum * invl[i];
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
      {
        dq[c] = (gl[c] + bl[c]) * p.chainW[c];
        qc[c] += dq[c];
      }
      continue;
    }

    /* ---- A = J invWq J^T + lambda 1 (lower triangle), r = dx - J g ---------------------------- */
    double A[NA], r[NX], g[NC];
#pragma unroll
    for (int i = 0; i < NA; ++i)
    {
      A[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      r[i] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      const double w = p.chainW[c];
      g[c] = w * (-(a.alpha * (p.chainGain[c] * (qc[c] - p.chainQ0[c]))));
      double u[NX];
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        u[i] = xyzFirst ? jo[c][i] : ax[c][i];
        if (NT == 2)
        {
          u[3 + i] = xyzFirst ? ax[c][i] : jo[c][i];
        }
      }
#pragma unroll
      for (int i = 0; i < NX; ++i)
      {
        const double wu = w * u[i];
#pragma unroll
        for (int k = 0; k <= i; ++k)
        {
          A[i * (i + 1) / 2 + k] += u[k] * wu;
        }
        r[i] += u[i] * g[c];
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      A[i * (i + 1) / 2 + i] += a.lambda;
      r[i] = dx[i] - r[i];
    }

    /* ---- Banachiewicz Cholesky (Rcs_linAlg.c:66-101), det = prod L_ii^2 ------------------------ */
    double invd[NX];
    bool ok = true;
    det = 1.0;
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
#pragma unroll
      for (int k = 0; k <= i; ++k)
      {
        double sum = A[i * (i + 1) / 2 + k];
#pragma unroll
        for (int t = 0; t < k; ++t)
        {
          sum -= A[i * (i + 1) / 2 + t] * A[k * (k + 1) / 2 + t];
        }
        if (k < i)
        {
          A[i * (i + 1) / 2 + k] = sum * invd[k];
        }
        else
        {
          /* L_ii = sqrt(pivot) and 1 / L_ii from one reciprocal square root (the sqrt + divide
           * pair sits on the critical path of the factorisation); det = prod L_ii^2 = prod pivot */
          ok = ok && (sum > 0.0);
          const double piv = ok ? sum : 1.0;
          invd[i] = rsqrt(piv);
          A[i * (i + 1) / 2 + i] = piv * invd[i];
          det *= piv;
        }
      }
    }
    if (!ok)
    {
      /* singular: zero displacement, q untouched (Rcs_MatNd.c:1833, IkSolverRMR.cpp:470-480) */
      det = 0.0;
      singular |= 1ull << it;
#pragma unroll
      for (int c = 0; c < NC; ++c)
      {
        dq[c] = 0.0;
      }
      continue;
    }

    /* ---- y = (L L^T)^-1 r ------------------------------------------------------------------------ */
    double y[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      double sum = r[i];
#pragma unroll
      for (int t = 0; t < i; ++t)
      {
        sum -= A[i * (i + 1) / 2 + t] * y[t];
      }
      y[i] = sum * invd[i];
    }
#pragma unroll
    for (int i = NX - 1; i >= 0; --i)
    {
      double sum = y[i];
#pragma unroll
      for (int t = i + 1; t < NX; ++t)
      {
        sum -= A[t * (t + 1) / 2 + i] * y[t];
      }
      y[i] = sum * invd[i];
    }

    /* ---- dq = g + invWq J^T y; q += dq ----------------------------------------------------------- */
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        sum += (xyzFirst ? jo[c][i] : ax[c][i]) * y[i];
      }
      if (NT == 2)
      {
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          sum += (xyzFirst ? ax[c][i] : jo[c][i]) * y[3 + i];
        }
      }
      dq[c] = g[c] + p.chainW[c] * sum;
      qc[c] += dq[c];
    }
  }

  /* ---- results ---------------------------------------------------------------------------------- */
  const bool lastSingular = a.iters > 0 && ((singular >> (a.iters - 1)) & 1ull);
  if (a.dq)
  {
    for (int j = 0; j < dof; ++j)
    {
      a.dq[s * dof + j] = 0.0;
    }
  }
  /* unconstrained joints off the chain have zero Jacobian columns: they only follow the
   * joint-limit gradient, dq_i = g_i, in every non-singular iteration */
  for (int k = 0; k < p.hdr->nOther; ++k)
  {
    const int j = p.otherJoint[k];
    double qj = qRow[j], step = 0.0;
    for (int it = 0; it < a.iters; ++it)
    {
      step = 0.0;
      if (!((singular >> it) & 1ull))
      {
        step = p.otherW[k] * (-(a.alpha * (p.otherGain[k] * (qj - p.otherQ0[k]))));
        qj += step;
      }
    }
    qRow[j] = qj;
    if (a.dq)
    {
      a.dq[s * dof + j] = step;
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
  {
    qRow[p.chainJoint[c]] = qc[c];
    if (a.dq)
    {
      a.dq[s * dof + p.chainJoint[c]] = lastSingular ? 0.0 : dq[c];
    }
  }
  if (a.dx)
  {
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      a.dx[s * NX + i] = dx[i];
    }
  }
  if (a.det)
  {
    a.det[s] = det;
  }
}

/* In-place Cholesky factor of the packed lower triangle A (row i at i (i + 1) / 2), as the
 * reference's MatNd_choleskyDecompositionSelf (Rcs_linAlg.c:66-101), followed by the solution of
 * (L L^T) y = r by substitution. Returns false on a non-positive pivot (det = 0). */
template <int NX>
__device__ __forceinline__ bool choleskySolve(double (&A)[NX * (NX + 1) / 2], const double (&r)[NX],
                                              double (&y)[NX], double& det)
{
  double invd[NX];
  bool ok = true;
  det = 1.0;
#pragma unroll
  for (int i = 0; i < NX; ++i)
  {
#pragma unroll
    for (int k = 0; k <= i; ++k)
    {
      double sum = A[i * (i + 1) / 2 + k];
#pragma unroll
      for (int t = 0; t < k; ++t)
      {
        sum -= A[i * (i + 1) / 2 + t] * A[k * (k + 1) / 2 + t];
      }
      if (k < i)
      {
        A[i * (i + 1) / 2 + k] = sum * invd[k];
      }
      else
      {
        ok = ok && (sum > 0.0);
        const double piv = ok ? sum : 1.0;
        invd[i] = rsqrt(piv);
        A[i * (i + 1) / 2 + i] = piv * invd[i];
        det *= piv;
      }
    }
  }
  if (!ok)
  {
    det = 0.0;
    return false;
  }
#pragma unroll
  for (int i = 0; i < NX; ++i)
  {
    double sum = r[i];
#pragma unroll
    for (int t = 0; t < i; ++t)
    {
      sum -= A[i * (i + 1) / 2 + t] * y[t];
    }
    y[i] = sum * invd[i];
  }
#pragma unroll
  for (int i = NX - 1; i >= 0; --i)
  {
    double sum = y[i];
#pragma unroll
    for (int t = i + 1; t < NX; ++t)
    {
      sum -= A[t * (t + 1) / 2 + i] * y[t];
    }
    y[i] = sum * invd[i];
  }
  return true;
}

/*
 * Single chain with more columns than the register file takes (9 .. RCSB_IK_LOOP_MAX_NC, e.g.
 * a humanoid's hand: 13 columns): the same step as ikChainKernel with the per-column data
 * (axis / origin, then the Jacobian column; joint value; step) in shared memory
 * [column][element][thread] and ordinary loops over the columns; the NX x NX system stays in
 * registers.
 */
#define RCSB_IK_LOOP_MAX_NC 32
#define RCSB_IK_LOOP_THREADS 64
template <int NX>
__global__ void __launch_bounds__(RCSB_IK_LOOP_THREADS)
ikChainLoopKernel(const IkArgs a, int nc)
{
  extern __shared__ __align__(16) char smem[];
  constexpr int T = RCSB_IK_LOOP_THREADS;
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  double* sAd = reinterpret_cast<double*>(sPlan + a.planBytes);   /* [9][T] */
  double* sCol = sAd + 9 * T;                                     /* [nc][8][T]: 0..5 column, 6 q, 7 dq */
  stageBlobsTma(sModel, a.model, a.modelBytes, sPlan, a.plan, a.planBytes);

  ModelView m;
  m.bind(sModel);
  IkPlanView p;
  p.bind(sPlan);

  const int tid = threadIdx.x;
  const long s = (long) blockIdx.x * T + tid;
  if (s >= a.N)
  {
    return;
  }

  constexpr int NT = NX / 3;
  constexpr int NA = NX * (NX + 1) / 2;
  const int dof = m.hdr->dof;
  const bool xyzFirst = p.taskKind[0] == RCSB_TASK_XYZ;
  const bool hasAbc = (NT == 2) || !xyzFirst;
  const bool hasXyz = (NT == 2) || xyzFirst;
  double* qRow = a.q + s * dof;
  const double* xd = a.xDes + s * NX;
  double* col = sCol + tid;
  auto at = [&](int c, int i) -> double& { return col[(c * 8 + i) * T]; };

  for (int c = 0; c < nc; ++c)
  {
    at(c, 6) = qRow[p.chainJoint[c]];
    at(c, 7) = 0.0;
  }
  double xPos[3] = {0.0, 0.0, 0.0};
  if (hasXyz)
  {
    const int o = (NT == 2 && !xyzFirst) ? 3 : 0;
    xPos[0] = xd[o];
    xPos[1] = xd[o + 1];
    xPos[2] = xd[o + 2];
  }
  if (hasAbc)
  {
    const int o = (NT == 2 && xyzFirst) ? 3 : 0;
    const double ea[3] = {xd[o], xd[o + 1], xd[o + 2]};
    double Ad[9];
    fromEuler(Ad, ea);
#pragma unroll
    for (int i = 0; i < 9; ++i)
    {
      sAd[i * T + tid] = Ad[i];
    }
  }

  double dx[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i)
  {
    dx[i] = 0.0;
  }
  double det = 0.0;
  unsigned long long singular = 0ull;

  for (int it = 0; it < a.iters; ++it)
  {
    /* ---- forward kinematics along the chain ------------------------------------------------ */
    Frame f;
    setIdentity(f);
    for (int c = 0; c < nc; ++c)
    {
      for (int k = p.progStart[c]; k < p.progStart[c + 1]; ++k)
      {
        chainStep<false>(f, p.prog[k], p.chainXf, qRow);
      }
      const int ctype = p.chainType[c];
      const double qc = at(c, 6);
      double axc[3];
      if ((ctype & 15) < 3)
      {
        double sn, cs;
        sincos(qc, &sn, &cs);
        rotateAboutAxis(f, ctype >> 4, sn, cs);   /* long chains keep the rows in place */
        axisOf(f, ctype >> 4, axc);
      }
      else
      {
        axisOf(f, ctype >> 4, axc);
        f.o[0] += qc * axc[0];
        f.o[1] += qc * axc[1];
        f.o[2] += qc * axc[2];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        at(c, i) = axc[i];
        at(c, 3 + i) = f.o[i];
      }
    }
    for (int k = p.progStart[nc]; k < p.progStart[nc + 1]; ++k)
    {
      chainStep<false>(f, p.prog[k], p.chainXf, qRow);
    }

    /* ---- task error (ControllerBase::computeDX) ---------------------------------------------- */
    double ePos[3] = {0.0, 0.0, 0.0}, eRot[3] = {0.0, 0.0, 0.0};
    if (hasXyz)
    {
      ePos[0] = xPos[0] - f.o[0];
      ePos[1] = xPos[1] - f.o[1];
      ePos[2] = xPos[2] - f.o[2];
    }
    if (hasAbc)
    {
      double Ad[9];
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        Ad[i] = sAd[i * T + tid];
      }
      rotationTaskErrorOfFrame(eRot, Ad, f.R);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
      dx[i] = xyzFirst ? ePos[i] : eRot[i];
      if (NT == 2)
      {
        dx[3 + i] = xyzFirst ? eRot[i] : ePos[i];
      }
    }

    /* ---- Jacobian columns in row order; A = J invWq J^T + lambda 1; r = dx - J g ------------- */
    double A[NA], r[NX];
#pragma unroll
    for (int i = 0; i < NA; ++i)
    {
      A[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      r[i] = 0.0;
    }
    for (int c = 0; c < nc; ++c)
    {
      double ja[3] = {at(c, 0), at(c, 1), at(c, 2)};
      double jt[3];
      if ((p.chainType[c] & 15) < 3)
      {
        const double d0 = f.o[0] - at(c, 3), d1 = f.o[1] - at(c, 4), d2 = f.o[2] - at(c, 5);
        jt[0] = ja[1] * d2 - ja[2] * d1;
        jt[1] = ja[2] * d0 - ja[0] * d2;
        jt[2] = ja[0] * d1 - ja[1] * d0;
      }
      else
      {
        jt[0] = ja[0];
        jt[1] = ja[1];
        jt[2] = ja[2];
        ja[0] = ja[1] = ja[2] = 0.0;
      }
      double u[NX];
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        u[i] = xyzFirst ? jt[i] : ja[i];
        if (NT == 2)
        {
          u[3 + i] = xyzFirst ? ja[i] : jt[i];
        }
      }
      const double w = p.chainW[c];
      const double g = w * (-(a.alpha * (p.chainGain[c] * (at(c, 6) - p.chainQ0[c]))));
      at(c, 7) = g;
#pragma unroll
      for (int i = 0; i < NX; ++i)
      {
        at(c, i) = u[i];
        const double wu = w * u[i];
#pragma unroll
        for (int k = 0; k <= i; ++k)
        {
          A[i * (i + 1) / 2 + k] += u[k] * wu;
        }
        r[i] += u[i] * g;
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      A[i * (i + 1) / 2 + i] += a.lambda;
      r[i] = dx[i] - r[i];
    }

    double y[NX];
    if (!choleskySolve<NX>(A, r, y, det))
    {
      /* singular: zero displacement, q untouched */
      singular |= 1ull << it;
      for (int c = 0; c < nc; ++c)
      {
        at(c, 7) = 0.0;
      }
      continue;
    }

    /* ---- dq = g + invWq J^T y; q += dq ----------------------------------------------------------- */
    for (int c = 0; c < nc; ++c)
    {
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < NX; ++i)
      {
        sum += at(c, i) * y[i];
      }
      const double dqc = at(c, 7) + p.chainW[c] * sum;
      at(c, 7) = dqc;
      at(c, 6) += dqc;
    }
  }

  /* ---- results ---------------------------------------------------------------------------------- */
  const bool lastSingular = a.iters > 0 && ((singular >> (a.iters - 1)) & 1ull);
  if (a.dq)
  {
    for (int j = 0; j < dof; ++j)
    {
      a.dq[s * dof + j] = 0.0;
    }
  }
  for (int k = 0; k < p.hdr->nOther; ++k)
  {
    const int j = p.otherJoint[k];
    double qj = qRow[j], step = 0.0;
    for (int it = 0; it < a.iters; ++it)
    {
      step = 0.0;
      if (!((singular >> it) & 1ull))
      {
        step = p.otherW[k] * (-(a.alpha * (p.otherGain[k] * (qj - p.otherQ0[k]))));
        qj += step;
      }
    }
    qRow[j] = qj;
    if (a.dq)
    {
      a.dq[s * dof + j] = step;
    }
  }
  for (int c = 0; c < nc; ++c)
  {
    qRow[p.chainJoint[c]] = at(c, 6);
    if (a.dq)
    {
      a.dq[s * dof + p.chainJoint[c]] = (lastSingular || a.iters == 0) ? 0.0 : at(c, 7);
    }
  }
  if (a.dx)
  {
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      a.dx[s * NX + i] = dx[i];
    }
  }
  if (a.det)
  {
    a.det[s] = det;
  }
}

template <int NX>
static cudaError_t launchChainLoop(const IkArgs& a, int nc, cudaStream_t stream)
{
  const size_t smemBytes = (size_t) a.modelBytes + (size_t) a.planBytes +
                           (size_t) (9 + 8 * nc) * RCSB_IK_LOOP_THREADS * sizeof(double);
  if (smemBytes > 227 * 1024)
  {
    return cudaErrorInvalidValue;
  }
  const dim3 grid((unsigned int) ((a.N + RCSB_IK_LOOP_THREADS - 1) / RCSB_IK_LOOP_THREADS));
  auto k = ikChainLoopKernel<NX>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int) smemBytes);
  if (e != cudaSuccess)
  {
    return e;
  }
  k<<<grid, RCSB_IK_LOOP_THREADS, smemBytes, stream>>>(a, nc);
  return cudaGetLastError();
}

/*
 * Tasks on several effectors (no coupled joints), e.g. both hands of a humanoid: the step of
 * the chain kernels over the union of the tasks' chain joints. Only the bodies above some
 * effector are walked (pruned depth-first list with its own frame-stack program); per union slot the joint axis / origin / value / step live in shared
 * memory [slot][element][thread], the effector frames and desired rotations likewise; the
 * NX x NX system stays in registers. A slot contributes to the rows of the tasks whose chain
 * it lies on (slotTaskMask).
 */
#define RCSB_IK_MULTI_THREADS 64
template <int NX, int NSLOTS>
__global__ void __launch_bounds__(RCSB_IK_MULTI_THREADS)
ikMultiKernel(const IkArgs a, int nU, int nEffs)
{
  extern __shared__ __align__(16) char smem[];
  constexpr int T = RCSB_IK_MULTI_THREADS;
  constexpr int NT = NX / 3;
  constexpr int NA = NX * (NX + 1) / 2;
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  double* sAd = reinterpret_cast<double*>(sPlan + a.planBytes);   /* [NT][9][T] desired rotations */
  double* sEff = sAd + NT * 9 * T;                                /* [nEffs][12][T] effector frames */
  double* sCol = sEff + nEffs * 12 * T;                           /* [nU][8][T] */
  stageBlobsTma(sModel, a.model, a.modelBytes, sPlan, a.plan, a.planBytes);

  ModelView m;
  m.bind(sModel);
  IkPlanView p;
  p.bind(sPlan);

  const int tid = threadIdx.x;
  const long s = (long) blockIdx.x * T + tid;
  if (s >= a.N)
  {
    return;
  }
  const int dof = m.hdr->dof;
  double* qRow = a.q + s * dof;
  const double* xd = a.xDes + s * NX;
  double* col = sCol + tid;
  double* eff = sEff + tid;
  auto at = [&](int c, int i) -> double& { return col[(c * 8 + i) * T]; };

  for (int c = 0; c < nU; ++c)
  {
    at(c, 6) = qRow[p.chainJoint[c]];
    at(c, 7) = 0.0;
  }
#pragma unroll
  for (int t = 0; t < NT; ++t)
  {
    if (p.taskKind[t] == RCSB_TASK_ABC)
    {
      const double ea[3] = {xd[3 * t], xd[3 * t + 1], xd[3 * t + 2]};
      double Ad[9];
      fromEuler(Ad, ea);
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        sAd[(t * 9 + i) * T + tid] = Ad[i];
      }
    }
  }

  double dx[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i)
  {
    dx[i] = 0.0;
  }
  double det = 0.0;
  unsigned long long singular = 0ull;
  double stk[(NSLOTS > 0 ? NSLOTS : 1) * 12];

  /* column of the stacked task Jacobian for union slot c (rows of tasks off its chain are 0) */
  auto column = [&](int c, double (&u)[NX]) {
    const int mask = p.slotTaskMask[c];
    const bool isRot = p.chainType[c] < 3;
    const double ax[3] = {at(c, 0), at(c, 1), at(c, 2)};
#pragma unroll
    for (int t = 0; t < NT; ++t)
    {
      double v[3] = {0.0, 0.0, 0.0};
      if ((mask >> t) & 1)
      {
        if (p.taskKind[t] == RCSB_TASK_XYZ)
        {
          if (isRot)
          {
            const double* e = eff + p.taskEff[t] * 12 * T;
            const double d0 = e[0] - at(c, 3), d1 = e[T] - at(c, 4), d2 = e[2 * T] - at(c, 5);
            v[0] = ax[1] * d2 - ax[2] * d1;
            v[1] = ax[2] * d0 - ax[0] * d2;
            v[2] = ax[0] * d1 - ax[1] * d0;
          }
          else
          {
            v[0] = ax[0];
            v[1] = ax[1];
            v[2] = ax[2];
          }
        }
        else if (isRot)
        {
          v[0] = ax[0];
          v[1] = ax[1];
          v[2] = ax[2];
        }
      }
      u[3 * t] = v[0];
      u[3 * t + 1] = v[1];
      u[3 * t + 2] = v[2];
    }
  };

  for (int it = 0; it < a.iters; ++it)
  {
    /* ---- forward kinematics (RcsGraph_setState) ------------------------------------------ */
    Frame f;
    setIdentity(f);
    for (int wi = 0; wi < p.hdr->nWalk; ++wi)
    {
      const int4 wd = p.walk[wi];
      const int b = wd.x;
      const int ld = wd.y;
      const int bflags = m.bodyFlags[b];
      if (ld == RCSB_LOAD_IDENT)
      {
        setIdentity(f);
      }
      else if (NSLOTS > 0 && ld >= 0)
      {
        const double* sp = stk + ld * 12;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          f.o[i] = sp[i];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i)
        {
          f.R[i] = sp[3 + i];
        }
      }
      const int j0 = m.bodyJntStart[b];
      const int j1 = j0 + m.bodyJntCount[b];
      for (int j = j0; j < j1; ++j)
      {
        const int jf = m.jntFlags[j];
        if (jf & RCSB_JF_HAS_AJP)
        {
          applyRelF(f, m.jntAJP + 12 * j, jf);
        }
        const int slot = p.jntSlot[j];
        const double qj = slot >= 0 ? at(slot, 6) : qRow[j];
        const int dir = m.jntType[j] % 3;
        double ax[3];
        if (jf & RCSB_JF_ROT)
        {
          double sn, cs;
          sincos(qj, &sn, &cs);
          rotateAboutAxis(f, dir, sn, cs);
          axisOf(f, dir, ax);
        }
        else
        {
          axisOf(f, dir, ax);
          f.o[0] += qj * ax[0];
          f.o[1] += qj * ax[1];
          f.o[2] += qj * ax[2];
        }
        if (slot >= 0)
        {
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            at(slot, i) = ax[i];
            at(slot, 3 + i) = f.o[i];
          }
        }
      }
      if (bflags & RCSB_BF_HAS_ABP)
      {
        applyRelF(f, m.bodyABP + 12 * b, bflags);
      }
      if (NSLOTS > 0)
      {
        const int sv = wd.z;
        if (sv >= 0)
        {
          double* sp = stk + sv * 12;
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            sp[i] = f.o[i];
          }
#pragma unroll
          for (int i = 0; i < 9; ++i)
          {
            sp[3 + i] = f.R[i];
          }
        }
      }
      const int ei = p.bodyEffIdx[b];
      if (ei >= 0)
      {
        double* e = eff + ei * 12 * T;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          e[i * T] = f.o[i];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i)
        {
          e[(3 + i) * T] = f.R[i];
        }
      }
    }

    /* ---- task errors (ControllerBase::computeDX) -------------------------------------------- */
#pragma unroll
    for (int t = 0; t < NT; ++t)
    {
      const double* e = eff + p.taskEff[t] * 12 * T;
      if (p.taskKind[t] == RCSB_TASK_XYZ)
      {
        dx[3 * t] = xd[3 * t] - e[0];
        dx[3 * t + 1] = xd[3 * t + 1] - e[T];
        dx[3 * t + 2] = xd[3 * t + 2] - e[2 * T];
      }
      else
      {
        double Ad[9], Ac[9], er[3];
#pragma unroll
        for (int i = 0; i < 9; ++i)
        {
          Ad[i] = sAd[(t * 9 + i) * T + tid];
          Ac[i] = e[(3 + i) * T];
        }
        rotationTaskErrorOfFrame(er, Ad, Ac);
        dx[3 * t] = er[0];
        dx[3 * t + 1] = er[1];
        dx[3 * t + 2] = er[2];
      }
    }

    /* ---- A = J invWq J^T + lambda 1, r = dx - J g ---------------------------------------------- */
    double A[NA], r[NX];
#pragma unroll
    for (int i = 0; i < NA; ++i)
    {
      A[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      r[i] = 0.0;
    }
    for (int c = 0; c < nU; ++c)
    {
      double u[NX];
      column(c, u);
      const double w = p.chainW[c];
      const double g = w * (-(a.alpha * (p.chainGain[c] * (at(c, 6) - p.chainQ0[c]))));
      at(c, 7) = g;
#pragma unroll
      for (int i = 0; i < NX; ++i)
      {
        const double wu = w * u[i];
#pragma unroll
        for (int k = 0; k <= i; ++k)
        {
          A[i * (i + 1) / 2 + k] += u[k] * wu;
        }
        r[i] += u[i] * g;
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      A[i * (i + 1) / 2 + i] += a.lambda;
      r[i] = dx[i] - r[i];
    }

    double y[NX];
    if (!choleskySolve<NX>(A, r, y, det))
    {
      singular |= 1ull << it;
      for (int c = 0; c < nU; ++c)
      {
        at(c, 7) = 0.0;
      }
      continue;
    }

    for (int c = 0; c < nU; ++c)
    {
      double u[NX];
      column(c, u);
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < NX; ++i)
      {
        sum += u[i] * y[i];
      }
      const double dqc = at(c, 7) + p.chainW[c] * sum;
      at(c, 7) = dqc;
      at(c, 6) += dqc;
    }
  }

  /* ---- results ---------------------------------------------------------------------------------- */
  const bool lastSingular = a.iters > 0 && ((singular >> (a.iters - 1)) & 1ull);
  if (a.dq)
  {
    for (int j = 0; j < dof; ++j)
    {
      a.dq[s * dof + j] = 0.0;
    }
  }
  for (int k = 0; k < p.hdr->nOther; ++k)
  {
    const int j = p.otherJoint[k];
    double qj = qRow[j], step = 0.0;
    for (int it = 0; it < a.iters; ++it)
    {
      step = 0.0;
      if (!((singular >> it) & 1ull))
      {
        step = p.otherW[k] * (-(a.alpha * (p.otherGain[k] * (qj - p.otherQ0[k]))));
        qj += step;
      }
    }
    qRow[j] = qj;
    if (a.dq)
    {
      a.dq[s * dof + j] = step;
    }
  }
  for (int c = 0; c < nU; ++c)
  {
    qRow[p.chainJoint[c]] = at(c, 6);
    if (a.dq)
    {
      a.dq[s * dof + p.chainJoint[c]] = (lastSingular || a.iters == 0) ? 0.0 : at(c, 7);
    }
  }
  if (a.dx)
  {
#pragma unroll
    for (int i = 0; i < NX; ++i)
    {
      a.dx[s * NX + i] = dx[i];
    }
  }
  if (a.det)
  {
    a.det[s] = det;
  }
}

template <int NX>
static cudaError_t launchMulti(const IkArgs& a, int nSlots, int nU, int nEffs, cudaStream_t stream)
{
  const size_t smemBytes = (size_t) a.modelBytes + (size_t) a.planBytes +
                           (size_t) ((NX / 3) * 9 + nEffs * 12 + nU * 8) * RCSB_IK_MULTI_THREADS *
                             sizeof(double);
  if (smemBytes > 227 * 1024)
  {
    return cudaErrorInvalidValue;
  }
  const dim3 grid((unsigned int) ((a.N + RCSB_IK_MULTI_THREADS - 1) / RCSB_IK_MULTI_THREADS));
#define RCSB_LAUNCH(NS)                                                                        \
  {                                                                                            \
    auto k = ikMultiKernel<NX, NS>;                                                            \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) smemBytes);                                     \
    if (e != cudaSuccess)                                                                      \
    {                                                                                          \
      return e;                                                                                \
    }                                                                                          \
    k<<<grid, RCSB_IK_MULTI_THREADS, smemBytes, stream>>>(a, nU, nEffs);                       \
    return cudaGetLastError();                                                                 \
  }
  if (nSlots <= 0)
  {
    RCSB_LAUNCH(0)
  }
  else
  {
    RCSB_LAUNCH(RCSB_MAX_FRAME_SLOTS)
  }
#undef RCSB_LAUNCH
}

template <int NC, int NX, bool LEFT = false, int MODE = 0>
static cudaError_t launchChain(const IkArgs& a, cudaStream_t stream)
{
  const size_t smemBytes = (size_t) a.modelBytes + (size_t) a.planBytes +
                           9 * RCSB_IK_THREADS * sizeof(double);
  const dim3 grid((unsigned int) ((a.N + RCSB_IK_THREADS - 1) / RCSB_IK_THREADS));
  auto k = ikChainKernel<NC, NX, LEFT, MODE>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int) smemBytes);
  if (e != cudaSuccess)
  {
    return e;
  }
  k<<<grid, RCSB_IK_THREADS, smemBytes, stream>>>(a);
  return cudaGetLastError();
}

template <int NX, bool LEFT = false, int MODE = 0>
static cudaError_t launchChainNc(const IkArgs& a, int nc, cudaStream_t stream)
{
  switch (nc)
  {
    case 1: return launchChain<1, NX, LEFT, MODE>(a, stream);
    case 2: return launchChain<2, NX, LEFT, MODE>(a, stream);
    case 3: return launchChain<3, NX, LEFT, MODE>(a, stream);
    case 4: return launchChain<4, NX, LEFT, MODE>(a, stream);
    case 5: return launchChain<5, NX, LEFT, MODE>(a, stream);
    case 6: return launchChain<6, NX, LEFT, MODE>(a, stream);
    case 7: return launchChain<7, NX, LEFT, MODE>(a, stream);
    case 8: return launchChain<8, NX, LEFT, MODE>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

/* chains of rotations without joint steps take the specialised instances (IkArgs::chainFlags) */
template <int NX, bool LEFT = false>
static cudaError_t launchChainMode(const IkArgs& a, int nc, cudaStream_t stream)
{
  if (a.chainFlags & RCSB_IK_CHAIN_ALLROT)
  {
    return (a.chainFlags & RCSB_IK_CHAIN_XYZ_FIRST) ? launchChainNc<NX, LEFT, 1>(a, nc, stream)
                                                    : launchChainNc<NX, LEFT, 2>(a, nc, stream);
  }
  return launchChainNc<NX, LEFT, 0>(a, nc, stream);
}

template <int MAXNX, int MAXNJ, int MAXDOF, bool LEFT = false>
static cudaError_t launchIk(const IkArgs& a, int nSlots, cudaStream_t stream)
{
  const size_t smemBytes = (size_t) a.modelBytes + (size_t) a.planBytes;
  const dim3 grid((unsigned int) ((a.N + RCSB_IK_THREADS - 1) / RCSB_IK_THREADS));
#define RCSB_LAUNCH(NS)                                                                        \
  {                                                                                            \
    auto k = ikKernel<MAXNX, MAXNJ, MAXDOF, NS, LEFT>;                                         \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) smemBytes);                                     \
    if (e != cudaSuccess)                                                                      \
    {                                                                                          \
      return e;                                                                                \
    }                                                                                          \
    k<<<grid, RCSB_IK_THREADS, smemBytes, stream>>>(a);                                        \
    return cudaGetLastError();                                                                 \
  }
  if (nSlots <= 0)
  {
    RCSB_LAUNCH(0)
  }
  else
  {
    RCSB_LAUNCH(RCSB_MAX_FRAME_SLOTS)
  }
#undef RCSB_LAUNCH
}

template <int MAXNX, int MAXNJ, int MAXDOF>
static cudaError_t launchIkPrio(const IkArgs& a, int nSlots, cudaStream_t stream)
{
  const size_t smemBytes = (size_t) a.modelBytes + (size_t) a.planBytes;
  const dim3 grid((unsigned int) ((a.N + RCSB_IK_THREADS - 1) / RCSB_IK_THREADS));
#define RCSB_LAUNCH(NS)                                                                        \
  {                                                                                            \
    auto k = ikPrioKernel<MAXNX, MAXNJ, MAXDOF, NS>;                                           \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) smemBytes);                                     \
    if (e != cudaSuccess)                                                                      \
    {                                                                                          \
      return e;                                                                                \
    }                                                                                          \
    k<<<grid, RCSB_IK_THREADS, smemBytes, stream>>>(a);                                        \
    return cudaGetLastError();                                                                 \
  }
  if (nSlots <= 0)
  {
    RCSB_LAUNCH(0)
  }
  else
  {
    RCSB_LAUNCH(RCSB_MAX_FRAME_SLOTS)
  }
#undef RCSB_LAUNCH
}

}   // namespace rcsb

template <int MAXNX, int MAXNJ, int MAXDOF, int MAXSLOT, bool CONSTRAINT>
static cudaError_t launchIkTasks(const rcsb::IkArgs& a, int nSlots, cudaStream_t stream)
{
  const size_t smemBytes = (size_t) a.modelBytes + (size_t) a.planBytes;
  const dim3 grid((unsigned int) ((a.N + RCSB_IK_THREADS - 1) / RCSB_IK_THREADS));
  if (nSlots <= 0)
  {
    auto k = rcsb::ikTasksKernel<MAXNX, MAXNJ, MAXDOF, MAXSLOT, 0, CONSTRAINT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemBytes);
    if (e != cudaSuccess)
    {
      return e;
    }
    k<<<grid, RCSB_IK_THREADS, smemBytes, stream>>>(a);
    return cudaGetLastError();
  }
  auto k = rcsb::ikTasksKernel<MAXNX, MAXNJ, MAXDOF, MAXSLOT, RCSB_MAX_FRAME_SLOTS, CONSTRAINT>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemBytes);
  if (e != cudaSuccess)
  {
    return e;
  }
  k<<<grid, RCSB_IK_THREADS, smemBytes, stream>>>(a);
  return cudaGetLastError();
}

extern "C" cudaError_t rcsb_launch_ik_tasks(const rcsb::IkArgs* a, int nSlots, int nx, int nxCap, int nJ,
                                            int dof, int nFrameSlots, cudaStream_t stream)
{
  using namespace rcsb;
  if (a->N <= 0)
  {
    return cudaSuccess;
  }
  if (a->constraint)
  {
    /* rows: the tasks' plus one per joint that can hit a limit; the large instantiation stops at 32
     * rows and reports states that would need more */
    if (nx > 24 || nFrameSlots > 24)
    {
      return cudaErrorInvalidValue;
    }
    if (nxCap <= 16 && nJ <= 8 && dof <= 12 && nFrameSlots <= 12)
    {
      return launchIkTasks<16, 8, 12, 12, true>(*a, nSlots, stream);
    }
    if (nJ <= 40 && dof <= 72)
    {
      return launchIkTasks<32, 40, 72, 24, true>(*a, nSlots, stream);
    }
    return cudaErrorInvalidValue;
  }
  if (nx <= 12 && nJ <= 8 && dof <= 12 && nFrameSlots <= 12)
  {
    return launchIkTasks<12, 8, 12, 12, false>(*a, nSlots, stream);
  }
  if (nx <= 24 && nJ <= 40 && dof <= 72 && nFrameSlots <= 24)
  {
    return launchIkTasks<24, 40, 72, 24, false>(*a, nSlots, stream);
  }
  return cudaErrorInvalidValue;
}

extern "C" cudaError_t rcsb_launch_ik_prio(const rcsb::IkArgs* a, int nSlots, int nx, int nJ, int dof,
                                           cudaStream_t stream)
{
  using namespace rcsb;
  if (a->N <= 0)
  {
    return cudaSuccess;
  }
  if (nx <= 6 && nJ <= 8 && dof <= 8)
  {
    return launchIkPrio<6, 8, 8>(*a, nSlots, stream);
  }
  if (nx <= 12 && nJ <= 40 && dof <= 72)
  {
    return launchIkPrio<12, 40, 72>(*a, nSlots, stream);
  }
  return cudaErrorInvalidValue;
}

extern "C" cudaError_t rcsb_launch_ik(const rcsb::IkArgs* a, int nSlots, int nx, int nJ, int dof,
                                      int chainNc, int multiNu, int nEffs, int nTasks,
                                      cudaStream_t stream)
{
  using namespace rcsb;
  if (a->N <= 0)
  {
    return cudaSuccess;
  }
  if (a->leftInverse)
  {
    /* solveLeftInverse: the register kernel for short single chains, else the general kernel */
    if (chainNc >= 1 && chainNc <= 8 && (nx == 3 || nx == 6))
    {
      return nx == 3 ? launchChainMode<3, true>(*a, chainNc, stream) : launchChainMode<6, true>(*a, chainNc, stream);
    }
    if (nx <= 6 && nJ <= 8 && dof <= 8)
    {
      return launchIk<6, 8, 8, true>(*a, nSlots, stream);
    }
    if (nx <= 12 && nJ <= 40 && dof <= 72)
    {
      return launchIk<12, 40, 72, true>(*a, nSlots, stream);
    }
    return cudaErrorInvalidValue;
  }
  if (chainNc >= 1 && chainNc <= 8 && (nx == 3 || nx == 6))
  {
    return nx == 3 ? launchChainMode<3>(*a, chainNc, stream) : launchChainMode<6>(*a, chainNc, stream);
  }
  if (chainNc > 8 && chainNc <= RCSB_IK_LOOP_MAX_NC && (nx == 3 || nx == 6))
  {
    const cudaError_t e = nx == 3 ? launchChainLoop<3>(*a, chainNc, stream)
                                  : launchChainLoop<6>(*a, chainNc, stream);
    if (e != cudaErrorInvalidValue)
    {
      return e;
    }
    cudaGetLastError();   /* does not fit in shared memory: general kernel */
  }
  if (multiNu > 0 && nx == 3 * nTasks && nTasks >= 1 && nTasks <= 4)
  {
    cudaError_t e = cudaErrorInvalidValue;
    switch (nTasks)
    {
      /* nEffs carries the frame-stack depth of the pruned walk in its upper half */
      case 1: e = launchMulti<3>(*a, nEffs >> 16, multiNu, nEffs & 0xffff, stream); break;
      case 2: e = launchMulti<6>(*a, nEffs >> 16, multiNu, nEffs & 0xffff, stream); break;
      case 3: e = launchMulti<9>(*a, nEffs >> 16, multiNu, nEffs & 0xffff, stream); break;
      default: e = launchMulti<12>(*a, nEffs >> 16, multiNu, nEffs & 0xffff, stream); break;
    }
    if (e != cudaErrorInvalidValue)
    {
      return e;
    }
    cudaGetLastError();
  }
  if (nx <= 6 && nJ <= 8 && dof <= 8)
  {
    return launchIk<6, 8, 8>(*a, nSlots, stream);
  }
  if (nx <= 12 && nJ <= 40 && dof <= 72)
  {
    return launchIk<12, 40, 72>(*a, nSlots, stream);
  }
  return cudaErrorInvalidValue;
}
