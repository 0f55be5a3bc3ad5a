/*
 * Batched kinetic terms M(q), h(q, qd), F_gravity(q) and total energy for sm_100a.
 *
 * Reference result reproduced per state (RcsGraph_computeKineticTerms,
 * src/RcsCore/Rcs_dynamics.c:439-573):
 *   M   = sum_b J_b^T diag(m_b 1, I_b^world) J_b            J_b = [JT at the COM; JR]
 *   h   = - sum_b J_b^T (M_b Jdot_b qd + [0; omega_b x I_b omega_b])
 *   F_g = sum_b JT_b^T (0, 0, -9.81 m_b)
 *   E   = sum_b m_b 9.81 z_b + 1/2 qd^T M qd
 * with the reference's conventions: bodies with m == 0 are skipped; qd is the IK-space
 * velocity (constrained joints do not contribute to Jdot qd) while omega_b is the body's
 * angular velocity from the forward kinematics (all joints), Rcs_dynamics.c:449-451, :536.
 *
 * The reference builds a 6 x nJ Jacobian and a 3 nJ x nJ Hessian per body (O(nb nJ^3)). Here
 * everything is expressed in world coordinates about the world origin (composite rigid body
 * algorithm for M, recursive bias accelerations for h: Jdot qd is the acceleration at
 * qdd = 0, so no Hessian is needed), in two kernels:
 *
 *   ktWalkKernel      one thread per state, one depth-first walk in registers: forward
 *                     kinematics, velocity / bias-acceleration recursion, and for every link
 *                     (body with an unconstrained joint) the composite inertia (m, m c, inertia
 *                     about the world origin) and bias wrench of its subtree, accumulated on a
 *                     shared-memory stack indexed by link depth and closed when the walk
 *                     leaves the subtree. Joint axes / origins / rates and the closed
 *                     composites stream through a 32-record shared-memory window into the
 *                     state's record in global memory as full 256-byte rows.
 *   ktAssembleKernel  a group of 8 / 16 / 32 lanes per state (lane = Jacobian column): reads the
 *                     record (coalesced), forms per column the momentum (p, L) of its
 *                     composite under a unit joint rate (L also with the transposed inertia:
 *                     input tensors need not be symmetric), h and F_gravity by projecting the
 *                     composite wrench / weight on the joint, and M row by row,
 *                     M[i][k] = <joint i, momentum of column k> for i an ancestor of k; rows
 *                     leave as coalesced stores straight from registers.
 *
 * Bound: HBM (M stream + the record round trip); see DESIGN.md for the byte counts.
 */
#include "rcsb_device.cuh"
#include "rcsb_kernels.h"
#include "rcsb_plan.h"

#ifndef RCSB_KT_PREFETCH
#define RCSB_KT_PREFETCH 3
#endif

namespace rcsb
{

struct KtPlanView
{
  const RcsbKtPlanHeader* hdr;
  const int* bodyLink;
  const int* bodyEmit;
  const int* linkParent;
  const int* colJoint;
  const int* colLink;
  const int* colIsRot;
  const int* colRec;
  const int* runLink;
  const int* runRec;
  const int4* pairs;
  const int* linkHome;
  const int2* extraRuns;
  const int* colHome;
  /* whole-body plans */
  const int* bodyOut;
  const int* bodyFrameRec;
  const int* taskBodyRec;
  const int* jacTasks;
  const unsigned char* taskColOn;
  const int* rootLinks;
  const int* jacEntries;
  const int2* linkSteps;

  __device__ __forceinline__ void bind(const char* blob)
  {
    hdr = reinterpret_cast<const RcsbKtPlanHeader*>(blob);
    bodyLink = reinterpret_cast<const int*>(blob + hdr->offBodyLink);
    bodyEmit = reinterpret_cast<const int*>(blob + hdr->offBodyEmit);
    linkParent = reinterpret_cast<const int*>(blob + hdr->offLinkParent);
    colJoint = reinterpret_cast<const int*>(blob + hdr->offColJoint);
    colLink = reinterpret_cast<const int*>(blob + hdr->offColLink);
    colIsRot = reinterpret_cast<const int*>(blob + hdr->offColIsRot);
    colRec = reinterpret_cast<const int*>(blob + hdr->offColRec);
    runLink = reinterpret_cast<const int*>(blob + hdr->offRunLink);
    runRec = reinterpret_cast<const int*>(blob + hdr->offRunRec);
    pairs = reinterpret_cast<const int4*>(blob + hdr->offPairs);
    linkHome = reinterpret_cast<const int*>(blob + hdr->offLinkHome);
    extraRuns = reinterpret_cast<const int2*>(blob + hdr->offExtraRuns);
    colHome = reinterpret_cast<const int*>(blob + hdr->offColHome);
    bodyOut = reinterpret_cast<const int*>(blob + hdr->offBodyOut);
    bodyFrameRec = reinterpret_cast<const int*>(blob + hdr->offBodyFrameRec);
    taskBodyRec = reinterpret_cast<const int*>(blob + hdr->offTaskBodyRec);
    jacTasks = reinterpret_cast<const int*>(blob + hdr->offJacTasks);
    taskColOn = reinterpret_cast<const unsigned char*>(blob + hdr->offTaskColOn);
    rootLinks = reinterpret_cast<const int*>(blob + hdr->offRootLinks);
    jacEntries = reinterpret_cast<const int*>(blob + hdr->offJacEntries);
    linkSteps = reinterpret_cast<const int2*>(blob + hdr->offLinkSteps);
  }
};

/* running kinematic state of the depth-first walk */
struct Kin
{
  Frame f;
  double wik[3];   /* angular velocity from the IK-space rates */
  double wfl[3];   /* angular velocity from all joint rates (BODY->omega) */
  double wd[3];    /* bias angular acceleration */
  double ao[3];    /* bias acceleration of the point of the frame currently at f.o */
};

__device__ __forceinline__ void addOffsetAccel(Kin& k, const double d[3])
{
  /* point moves by d within the same rigid frame: a += wd x d + w x (w x d) */
  double t[3], u[3], v[3];
  cross3(t, k.wd, d);
  cross3(u, k.wik, d);
  cross3(v, k.wik, u);
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    k.ao[i] += t[i] + v[i];
  }
}

/* <spatial axis of a joint (w, v), momentum (p, L) about the world origin> = w . L + v . p
 * (rotation about a through o: w = a, v = o x a, i.e. a . (L - o x p); translation along a: w = 0,
 * v = a, i.e. a . p) */
__device__ __forceinline__ double pairing(const double* __restrict__ sj, const double p[3], const double L[3])
{
  return (sj[0] * L[0] + sj[1] * L[1] + sj[2] * L[2]) + (sj[3] * p[0] + sj[4] * p[1] + sj[5] * p[2]);
}


/* ------------------------------------------------------------------------------------------ */
/* kernel 1: depth-first walk, one thread per state                                            */
/* ------------------------------------------------------------------------------------------ */

/* WB (whole-body plans, rcsb_wholebody): the walk also stores the frames of the selected bodies
 * (A_BI: three 32-byte stores per frame like fkKernel) and emits the frames of the bodies the task
 * Jacobians refer to into the record. */
template <bool VEL, int NSLOTS, bool WB>
__global__ void __launch_bounds__(RCSB_KT_WALK_THREADS)
ktWalkKernel(const KtArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  double* sWin = reinterpret_cast<double*>(sPlan + a.planWalkBytes);
  double* sAcc = sWin + (RCSB_KT_WALK_THREADS / 32) * 32 * 33;

  stageBlobsTma(sModel, a.model, a.modelBytes, sPlan, a.plan, a.planWalkBytes);

  ModelView m;
  m.bind(sModel);
  KtPlanView p;
  p.bind(sPlan);   /* only hdr, bodyLink, bodyEmit (and bodyOut, bodyFrameRec) are staged */

  constexpr int T = RCSB_KT_WALK_THREADS;
  constexpr int LREC = RCSB_KT_LINK_REC_N(VEL), CREC = RCSB_KT_COL_REC_N(VEL);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nb = m.hdr->nb;
  const int dof = m.hdr->dof;
  const long s = (long) blockIdx.x * T + tid;
  const long s0 = s - lane;
  if (s0 >= a.N)
  {
    return;   /* whole warp idle */
  }
  const long sc = s < a.N ? s : a.N - 1;   /* idle lanes shadow the last state */
  const long nvl = a.N - s0;

  RecordWriter w;
  w.win = sWin + (tid >> 5) * 32 * 33;
  w.dst = a.rec + s0 * a.recDoubles;
  w.recDoubles = a.recDoubles;
  w.nv = nvl > 32 ? 32 : (int) nvl;
  w.lane = lane;
  w.p = 0;

  StateReader rd;
  rd.full = (a.qdim == dof);
  rd.q = a.q + sc * a.qdim;
  rd.qd = VEL ? a.qd + sc * a.qdim : nullptr;

  constexpr int SLOT_DOUBLES = VEL ? 24 : 12;
  double stk[(NSLOTS > 0 ? NSLOTS : 1) * SLOT_DOUBLES];
  const double grav = 9.81;

  /* mass properties of the bodies visited since the last emit (all of one link) */
  double* acc = sAcc + tid;
#pragma unroll
  for (int i = 0; i < LREC; ++i)
  {
    acc[i * T] = 0.0;
  }
  auto emitRun = [&]() {
    double v[LREC];
#pragma unroll
    for (int i = 0; i < LREC; ++i)
    {
      v[i] = acc[i * T];
      acc[i * T] = 0.0;
    }
    w.emitN(v);
  };

  /* `keep` holds the running state while a leaf body is evaluated. The compiler answers with two
   * register sets for the walk state and 2 x 48 moves per body (11 % of the kernel's instructions,
   * ncu source page). Measured alternative (round 2): every start state of a body read from a
   * frame-stack slot in local memory by predicated in-place loads, the leaf's parent parked there
   * too - 250 -> 166 registers, 428 M -> 393 M warp instructions per 262,144 humanoid states, but the
   * same 0.97 ms with velocities (the walk is latency-bound, not issue-bound), 1.57 -> 1.66 ms for
   * the whole-body walk and 3 % more DRAM traffic (the stack lines are written back); 192-thread
   * blocks, which the smaller footprint allows, are slower still (1.30 ms: less L1 next to the larger
   * shared-memory carve-out). Not kept. */
  Kin k, keep;
  setIdentity(k.f);
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    k.wik[i] = k.wfl[i] = k.wd[i] = k.ao[i] = 0.0;
  }
  keep = k;
  double V = 0.0;
  double freeM[4] = {0.0, 0.0, 0.0, 0.0};   /* m, m c of massive bodies no joint moves */
  /* q / q_dot of the next RCSB_KT_PREFETCH joints are in flight while a joint is processed: the
   * rows of a warp's 32 states (2 x 520 B each for the humanoid) do not stay in L1 next to
   * those of the other warps, so a load is an L2 round trip (with distance 1 the joint loop
   * waited on it for 17 % of the kernel's stall samples) */
  double nextQ[RCSB_KT_PREFETCH], nextQd[RCSB_KT_PREFETCH];
#pragma unroll
  for (int i = 0; i < RCSB_KT_PREFETCH; ++i)
  {
    nextQ[i] = nextQd[i] = 0.0;
    if (i < dof)
    {
      nextQ[i] = rd.fetchQ(m, i);
      if (VEL)
      {
        nextQd[i] = rd.fetchQd(m, i);
      }
    }
  }

  for (int b = 0; b < nb; ++b)
  {
    if (p.bodyEmit[b])
    {
      emitRun();
    }

    const int ld = m.bodyLoad[b];
    const int bflags = m.bodyFlags[b];
    if (ld == RCSB_LOAD_IDENT)
    {
      setIdentity(k.f);
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        k.wik[i] = k.wfl[i] = k.wd[i] = k.ao[i] = 0.0;
      }
    }
    else if (NSLOTS > 0 && ld >= 0)
    {
      const double* sp = stk + ld * SLOT_DOUBLES;
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        k.f.o[i] = sp[i];
      }
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        k.f.R[i] = sp[3 + i];
      }
      if (VEL)
      {
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          k.wik[i] = sp[12 + i];
          k.wfl[i] = sp[15 + i];
          k.wd[i] = sp[18 + i];
          k.ao[i] = sp[21 + i];
        }
      }
    }
    if (bflags & RCSB_BF_LEAF)
    {
      keep = k;
    }

    const int j0 = m.bodyJntStart[b];
    const int j1 = j0 + m.bodyJntCount[b];
    for (int j = j0; j < j1; ++j)
    {
      const int jf = m.jntFlags[j];
      if (jf & RCSB_JF_HAS_AJP)
      {
        double d[3];
        applyRelD(k.f, m.jntAJP + 12 * j, jf, d);
        if (VEL)
        {
          addOffsetAccel(k, d);
        }
      }

      /* joints are visited in jointIndex order: the loads of the next joint are in flight
       * while this one is processed */
      const double rawQ = nextQ[0], rawQd = nextQd[0];
#pragma unroll
      for (int i = 0; i + 1 < RCSB_KT_PREFETCH; ++i)
      {
        nextQ[i] = nextQ[i + 1];
        nextQd[i] = nextQd[i + 1];
      }
      if (j + RCSB_KT_PREFETCH < dof)
      {
        nextQ[RCSB_KT_PREFETCH - 1] = rd.fetchQ(m, j + RCSB_KT_PREFETCH);
        if (VEL)
        {
          nextQd[RCSB_KT_PREFETCH - 1] = rd.fetchQd(m, j + RCSB_KT_PREFETCH);
        }
      }
      const double qj = rd.finishQ(m, j, rawQ);
      const int dir = m.jntType[j] % 3;
      const int col = m.jntJacobi[j];
      double qdFull = 0.0, qdIk = 0.0;
      if (VEL)
      {
        qdFull = rd.finishQd(m, j, rawQ, rawQd);
        qdIk = col >= 0 ? qdFull : 0.0;
      }
      double ax[3];

      if (jf & RCSB_JF_ROT)
      {
        double sn, cs;
        sincos(qj, &sn, &cs);
        rotateAboutAxis(k.f, dir, sn, cs);
        axisOf(k.f, dir, ax);
        if (VEL)
        {
          /* d/dt (a qd) at qdd = 0: qd (w x a) */
          double t[3];
          cross3(t, k.wik, ax);
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            k.wd[i] += qdIk * t[i];
            k.wik[i] += qdIk * ax[i];
            k.wfl[i] += qdFull * ax[i];
          }
        }
      }
      else
      {
        axisOf(k.f, dir, ax);
        const double d[3] = {qj * ax[0], qj * ax[1], qj * ax[2]};
        if (VEL)
        {
          /* sliding origin: rigid offset terms + Coriolis 2 w x (a qd) */
          double t[3];
          addOffsetAccel(k, d);
          cross3(t, k.wik, ax);
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            k.ao[i] += 2.0 * qdIk * t[i];
          }
        }
        k.f.o[0] += d[0];
        k.f.o[1] += d[1];
        k.f.o[2] += d[2];
      }

      if (col >= 0)
      {
        if (VEL)
        {
          const double v[7] = {ax[0], ax[1], ax[2], k.f.o[0], k.f.o[1], k.f.o[2], qdIk};
          w.emitN(v);
        }
        else
        {
          const double v[6] = {ax[0], ax[1], ax[2], k.f.o[0], k.f.o[1], k.f.o[2]};
          w.emitN(v);
        }
        (void) CREC;
      }
    }

    if (bflags & RCSB_BF_HAS_ABP)
    {
      double d[3];
      applyRelD(k.f, m.bodyABP + 12 * b, bflags, d);
      if (VEL)
      {
        addOffsetAccel(k, d);
      }
    }

    if (NSLOTS > 0)
    {
      const int sv = m.bodySave[b];
      if (sv >= 0)
      {
        double* sp = stk + sv * SLOT_DOUBLES;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          sp[i] = k.f.o[i];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i)
        {
          sp[3 + i] = k.f.R[i];
        }
        if (VEL)
        {
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            sp[12 + i] = k.wik[i];
            sp[15 + i] = k.wfl[i];
            sp[18 + i] = k.wd[i];
            sp[21 + i] = k.ao[i];
          }
        }
      }
    }

    if (WB)
    {
      const int out = p.bodyOut[b];
      if (out >= 0 && a.A_BI && s < a.N)
      {
        double* dst = a.A_BI + (s * p.hdr->nSel + out) * 12;
        if ((reinterpret_cast<uintptr_t>(a.A_BI) & 31u) == 0)
        {
          storeFrame<true>(dst, k.f);
        }
        else
        {
          storeFrame<false>(dst, k.f);
        }
      }
      if (p.bodyFrameRec[b])
      {
        const double v[12] = {k.f.o[0], k.f.o[1], k.f.o[2], k.f.R[0], k.f.R[1], k.f.R[2],
                              k.f.R[3], k.f.R[4], k.f.R[5], k.f.R[6], k.f.R[7], k.f.R[8]};
        w.emitN(v);
      }
    }

    /* ---- mass properties of the body ---------------------------------------------- */
    if (bflags & RCSB_BF_HAS_MASS)
    {
      const double mass = m.bodyMass[b];
      const double* cl = m.bodyCom + 3 * b;
      const double* Ib = m.bodyInertia + 9 * b;
      const double* R = k.f.R;
      /* COM offset and position in world coordinates: r = R^T c_local */
      const double r[3] = {R[0] * cl[0] + R[3] * cl[1] + R[6] * cl[2],
                           R[1] * cl[0] + R[4] * cl[1] + R[7] * cl[2],
                           R[2] * cl[0] + R[5] * cl[1] + R[8] * cl[2]};
      const double c[3] = {k.f.o[0] + r[0], k.f.o[1] + r[1], k.f.o[2] + r[2]};
      V += mass * grav * c[2];

      if (p.bodyLink[b] < 0)
      {
        freeM[0] += mass;
        freeM[1] += mass * c[0];
        freeM[2] += mass * c[1];
        freeM[3] += mass * c[2];
      }
      else   /* some unconstrained joint moves the body */
      {
        /* world inertia about the COM: R^T I_b R (Mat3d_similarityTransform) */
        double tmp[9], Iw[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj)
          {
            tmp[3 * i + jj] = Ib[3 * i] * R[jj] + Ib[3 * i + 1] * R[3 + jj] + Ib[3 * i + 2] * R[6 + jj];
          }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj)
          {
            Iw[3 * i + jj] = R[i] * tmp[jj] + R[3 + i] * tmp[3 + jj] + R[6 + i] * tmp[6 + jj];
          }

        /* the inertia tensor is used as given (9 entries): the reference does not
         * symmetrise it and some graph files are not symmetric */
        double* cp = acc;
        const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
        cp[0] += mass;
        cp[T] += mass * c[0];
        cp[2 * T] += mass * c[1];
        cp[3 * T] += mass * c[2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj)
          {
            cp[(4 + 3 * i + jj) * T] += Iw[3 * i + jj] + mass * ((i == jj ? cc : 0.0) - c[i] * c[jj]);
          }

        if (VEL)
        {
          /* bias wrench about the world origin */
          double t[3], u[3], v[3];
          cross3(t, k.wd, r);
          cross3(u, k.wik, r);
          cross3(v, k.wik, u);
          const double F[3] = {mass * (k.ao[0] + t[0] + v[0]), mass * (k.ao[1] + t[1] + v[1]),
                               mass * (k.ao[2] + t[2] + v[2])};
          const double* wf = k.wfl;
          const double Iwf[3] = {Iw[0] * wf[0] + Iw[1] * wf[1] + Iw[2] * wf[2],
                                 Iw[3] * wf[0] + Iw[4] * wf[1] + Iw[5] * wf[2],
                                 Iw[6] * wf[0] + Iw[7] * wf[1] + Iw[8] * wf[2]};
          const double Iwd[3] = {Iw[0] * k.wd[0] + Iw[1] * k.wd[1] + Iw[2] * k.wd[2],
                                 Iw[3] * k.wd[0] + Iw[4] * k.wd[1] + Iw[5] * k.wd[2],
                                 Iw[6] * k.wd[0] + Iw[7] * k.wd[1] + Iw[8] * k.wd[2]};
          double gy[3], cf[3];
          cross3(gy, wf, Iwf);
          cross3(cf, c, F);
          cp[13 * T] += F[0];
          cp[14 * T] += F[1];
          cp[15 * T] += F[2];
          cp[16 * T] += Iwd[0] + gy[0] + cf[0];
          cp[17 * T] += Iwd[1] + gy[1] + cf[1];
          cp[18 * T] += Iwd[2] + gy[2] + cf[2];
        }
      }
    }

    if (bflags & RCSB_BF_LEAF)
    {
      k = keep;
    }
  }

  if (p.bodyEmit[nb])
  {
    emitRun();
  }
  const double tail[5] = {V, freeM[0], freeM[1], freeM[2], freeM[3]};
  w.emitN(tail);
  w.finish();
}

/* ------------------------------------------------------------------------------------------ */
/* kernel 2: assembly, LPS lanes per state                                                     */
/* ------------------------------------------------------------------------------------------ */

template <bool VEL, int LPS>
__global__ void __launch_bounds__(RCSB_KT_ASM_THREADS)
ktAssembleKernel(const KtArgs a, int groupStride)
{
  extern __shared__ __align__(16) char smem[];
  char* sPlan = smem;
  double* sBuf = reinterpret_cast<double*>(sPlan + a.planBytes);

  stageBlobsTma(sPlan, a.plan, a.planBytes, nullptr, nullptr, 0);
  KtPlanView p;
  p.bind(sPlan);

  constexpr int GPB = RCSB_KT_ASM_THREADS / LPS;   /* state groups per block */
  const int tid = threadIdx.x;
  const int gl = tid & (LPS - 1);
  const int grp = tid / LPS;
  const int nJ = p.hdr->nJ;
  const int nLinks = p.hdr->nLinks;
  const int R = a.recDoubles;
  const bool sym = p.hdr->symmetric != 0;
  double* buf = sBuf + (long) grp * groupStride;   /* layout: RcsbKtPlanHeader, asm* fields */
  const RcsbKtAsmLayout lay = rcsbKtAsmLayout(R, nJ, a.stageM, sym ? 1 : 0);
  constexpr int LREC = RCSB_KT_LINK_REC_N(VEL);
  constexpr int SS = RCSB_KT_S_STRIDE;
  const int FS = RCSB_KT_F_STRIDE(sym ? 1 : 0);
  double* scol = buf + lay.S;                      /* SS per column: spatial axis (w, v), rate */
  double* fb = buf + lay.F;                        /* FS per column: p, L (, L with I^T) */
  double* tail = buf + lay.tail;                   /* V, m and m c of the bodies no joint moves */
  double* Ms = buf;                                /* nJ x nJ, over the consumed record (stageM) */
  double* rhs = buf + lay.rhs;                     /* nJ: right-hand side of the direct dynamics */
  const bool stage = a.stageM != 0;
  const double grav = 9.81;

  /* The record of a state (R doubles, a multiple of 32 bytes, contiguous in global memory) comes
   * in as ONE bulk asynchronous copy (TMA, cp.async.bulk) straight into the group's region,
   * completion on the group's mbarrier: no per-lane loads, no staging registers. The copy of
   * the group's next state is issued as soon as the record region is dead: after the column
   * phase when M goes straight to global memory, at the end of the iteration when the region
   * is reused (matrix staging of the direct dynamics, COG sums). */
  __shared__ __align__(8) unsigned long long recBar[GPB];
  const unsigned int barAddr = (unsigned int) __cvta_generic_to_shared(&recBar[grp]);
  const unsigned int bufAddr = (unsigned int) __cvta_generic_to_shared(buf);
  const unsigned int recBytes = (unsigned int) R * 8u;
  if (gl == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto fetch = [&](long state) {
    /* every lane of the group has finished with the region (warp barrier before the call);
     * order those generic-proxy accesses before the asynchronous write */
    if (gl == 0)
    {
      const double* src = a.rec + (state < a.N ? state : a.N - 1) * R;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(recBytes)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(bufAddr), "l"(src), "r"(recBytes), "r"(barAddr)
                   : "memory");
    }
  };
  const bool earlyFetch = !stage && !a.cog && !a.Jcog;
  const bool midFetch = !stage && !earlyFetch;   /* COG outputs: the composites are read once more after the columns */
  unsigned int phase = 0;

  /* all groups of a warp iterate together (warp-level barriers inside) */
  const long nGroupsTotal = (a.N + GPB - 1) / GPB * GPB;
  const long sFirst = (long) blockIdx.x * GPB + grp;
  const long sStep = (long) gridDim.x * GPB;
  if (sFirst < nGroupsTotal)
  {
    fetch(sFirst);
  }
  for (long s = sFirst; s < nGroupsTotal; s += sStep)
  {
    const bool valid = s < a.N;
    const long sc = valid ? s : a.N - 1;

    {
      unsigned int done;
      do
      {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(barAddr), "r"(phase)
                     : "memory");
      } while (!done);
      phase ^= 1u;
    }
    __syncwarp();

    /* ---- composites in place: extra runs into their link, links into their ancestors ------ */
    if (gl < 5)
    {
      tail[gl] = buf[p.hdr->recV + gl];
    }
    if (p.hdr->nExtraRuns > 0)
    {
      for (int i = gl; i < LREC; i += LPS)
      {
        for (int r = 0; r < p.hdr->nExtraRuns; ++r)
        {
          const int2 x = p.extraRuns[r];
          buf[x.y + i] += buf[x.x + i];
        }
      }
      __syncwarp();
    }
    /* links into their ancestors, element i of every composite with one lane: the links come in
     * reverse depth-first order (children before parents). Along a chain (the parent is the link
     * before) the running sum stays in a register; only at a branching point it is added to the
     * parent's composite in shared memory. (Round 1 did a shared-memory read-modify-write per link:
     * a third of the kernel's stall samples sat on that dependency chain.) */
    /* The value of the next link is requested one step ahead: the step's stores never touch it (a
     * parent that comes next receives the sum through `carry`, not through shared memory), and the
     * chain of a lane is then an add per link instead of a shared-memory round trip per link (the
     * loop held 19 % of the kernel's stall samples). */
    for (int i = gl; i < LREC; i += LPS)
    {
      double carry = 0.0;
      int2 st = p.linkSteps[0];
      double cur = nLinks > 0 ? buf[st.x + i] : 0.0;
      for (int l = 0; l < nLinks; ++l)
      {
        const int2 nx = p.linkSteps[l + 1 < nLinks ? l + 1 : l];
        const double nv = buf[nx.x + i];
        const double v = cur + carry;
        buf[st.x + i] = v;
        carry = 0.0;
        if (st.y == -1)
        {
          carry = v;
        }
        else if (st.y >= 0)
        {
          buf[st.y + i] += v;
        }
        st = nx;
        cur = nv;
      }
    }
    __syncwarp();

    /* ---- per column: momentum of its composite under a unit joint rate, h, F_gravity ---- */
    for (int k = gl; k < nJ; k += LPS)
    {
      const double* ja = buf + p.colRec[k];
      const double* cp = buf + p.colHome[k];
      const bool isRot = p.colIsRot[k] != 0;
      const double ax[3] = {ja[0], ja[1], ja[2]};
      const double o[3] = {ja[3], ja[4], ja[5]};
      const double mC = cp[0];
      const double hC[3] = {cp[1], cp[2], cp[3]};
      double pv[3], Lv[3], Lt[3];
      if (isRot)
      {
        const double hr[3] = {hC[0] - mC * o[0], hC[1] - mC * o[1], hC[2] - mC * o[2]};
        double axo[3], t[3];
        cross3(pv, ax, hr);            /* p = a x (h - m o) */
        cross3(axo, ax, o);
        cross3(t, hC, axo);            /* h x (a x o) */
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          Lv[i] = cp[4 + 3 * i] * ax[0] + cp[4 + 3 * i + 1] * ax[1] + cp[4 + 3 * i + 2] * ax[2] - t[i];
          Lt[i] = cp[4 + i] * ax[0] + cp[4 + 3 + i] * ax[1] + cp[4 + 6 + i] * ax[2] - t[i];
        }
      }
      else
      {
        pv[0] = mC * ax[0];
        pv[1] = mC * ax[1];
        pv[2] = mC * ax[2];
        cross3(Lv, hC, ax);            /* L = h x a */
        Lt[0] = Lv[0];
        Lt[1] = Lv[1];
        Lt[2] = Lv[2];
      }
      /* spatial axis of the joint about the world origin */
      double* sk = scol + SS * k;
      if (isRot)
      {
        sk[0] = ax[0]; sk[1] = ax[1]; sk[2] = ax[2];
        sk[3] = o[1] * ax[2] - o[2] * ax[1];
        sk[4] = o[2] * ax[0] - o[0] * ax[2];
        sk[5] = o[0] * ax[1] - o[1] * ax[0];
      }
      else
      {
        sk[0] = 0.0; sk[1] = 0.0; sk[2] = 0.0;
        sk[3] = ax[0]; sk[4] = ax[1]; sk[5] = ax[2];
      }
      if (VEL)
      {
        sk[6] = ja[6];
      }
      double* f = fb + FS * k;
      f[0] = pv[0]; f[1] = pv[1]; f[2] = pv[2];
      f[3] = Lv[0]; f[4] = Lv[1]; f[5] = Lv[2];
      if (!sym)
      {
        f[6] = Lt[0]; f[7] = Lt[1]; f[8] = Lt[2];
      }

      double hval = 0.0, gval = 0.0;
      if ((a.h || a.qdd) && VEL)
      {
        /* a . (N - o x F) for a rotation, a . F for a translation */
        const double* Fb = cp + 13;
        const double* Nb = cp + 16;
        hval = -((sk[0] * Nb[0] + sk[1] * Nb[1] + sk[2] * Nb[2]) + (sk[3] * Fb[0] + sk[4] * Fb[1] + sk[5] * Fb[2]));
      }
      if (a.Fg || a.qdd)
      {
        if (isRot)
        {
          /* a . ((h - m o) x (0, 0, -g)) */
          const double hr0 = hC[0] - mC * o[0];
          const double hr1 = hC[1] - mC * o[1];
          gval = ax[0] * (hr1 * -grav) + ax[1] * (hr0 * grav);
        }
        else
        {
          gval = ax[2] * (mC * -grav);
        }
      }
      if (a.h && valid)
      {
        a.h[s * nJ + k] = hval;
      }
      if (a.Fg && valid)
      {
        a.Fg[s * nJ + k] = gval;
      }
      if (a.qdd)
      {
        /* b = F_gravity + F_ext - F_jnt + h, in this order (Rcs_dynamics.c:718-731) */
        double b = gval;
        if (a.Fext)
        {
          b += a.Fext[sc * nJ + k];
        }
        if (a.Fjnt)
        {
          b -= a.Fjnt[sc * nJ + k];
        }
        rhs[k] = b + hval;
      }
    }
    __syncwarp();

    /* ---- task Jacobians (whole-body plans): RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian
     * (Rcs_kinematics.c:566, :789) from the spatial joint axes and the frames of the task bodies in
     * the record. Column k of a body-point Jacobian is w_k x p + v_k, of a rotation Jacobian w_k
     * (both joint types). The rows are zero-filled with coalesced stores; the plan lists the
     * (task, column) pairs that are not structurally zero, one per lane and pass. Modes as in
     * combineKernel (rcsb_taskjac.cu). */
    if (a.wholeBody && a.Jt)
    {
      const int nJac = p.hdr->nJac;
      double* Jg = a.Jt + s * (long) nJac * 3 * nJ;
      if (valid)
      {
        for (int e = gl; e < nJac * 3 * nJ; e += LPS)
        {
          Jg[e] = 0.0;
        }
      }
      __syncwarp();   /* orders the zeros before the entries below (same addresses, other lanes) */
      const int nEnt = p.hdr->nJacEntries;
      for (int e = gl; e < nEnt; e += LPS)
      {
        const int ent = p.jacEntries[e];
        const int kc = ent & 0xff, t = (ent >> 8) & 0xff;
        const int* tk = p.jacTasks + 8 * t;
        const int kind = tk[0], u2 = tk[2], u1 = tk[3], uA = tk[5];
        const double* sj = scol + SS * kc;
        const double w0 = sj[0], w1 = sj[1], w2 = sj[2];
        double v[3] = {0.0, 0.0, 0.0};
        if (kind == 1)   /* RCSB_JAC_POS: JT_2 - JT_1 + r_12 x JR_3 */
        {
          if (ent & (1 << 16))
          {
            const double* o2 = buf + p.taskBodyRec[u2];
            v[0] = (w1 * o2[2] - w2 * o2[1]) + sj[3];
            v[1] = (w2 * o2[0] - w0 * o2[2]) + sj[4];
            v[2] = (w0 * o2[1] - w1 * o2[0]) + sj[5];
          }
          if (ent & (1 << 17))
          {
            const double* o1 = buf + p.taskBodyRec[u1];
            v[0] -= (w1 * o1[2] - w2 * o1[1]) + sj[3];
            v[1] -= (w2 * o1[0] - w0 * o1[2]) + sj[4];
            v[2] -= (w0 * o1[1] - w1 * o1[0]) + sj[5];
          }
          if (ent & (1 << 18))
          {
            const double* o1 = buf + p.taskBodyRec[u1];
            double r[3] = {-o1[0], -o1[1], -o1[2]};
            if (u2 >= 0)
            {
              const double* o2 = buf + p.taskBodyRec[u2];
              r[0] += o2[0];
              r[1] += o2[1];
              r[2] += o2[2];
            }
            v[0] += r[1] * w2 - r[2] * w1;
            v[1] += -r[0] * w2 + r[2] * w0;
            v[2] += r[0] * w1 - r[1] * w0;
          }
        }
        else             /* RCSB_JAC_OMEGA: JR_2 - JR_1 */
        {
          const double sgn = ((ent & (1 << 16)) ? 1.0 : 0.0) - ((ent & (1 << 17)) ? 1.0 : 0.0);
          v[0] = sgn * w0;
          v[1] = sgn * w1;
          v[2] = sgn * w2;
        }
        if (uA >= 0)
        {
          const double* Rm = buf + p.taskBodyRec[uA] + 3;
          const double x = v[0], y = v[1], z = v[2];
          v[0] = Rm[0] * x + Rm[1] * y + Rm[2] * z;
          v[1] = Rm[3] * x + Rm[4] * y + Rm[5] * z;
          v[2] = Rm[6] * x + Rm[7] * y + Rm[8] * z;
        }
        if (valid)
        {
          double* o = Jg + (long) t * 3 * nJ + kc;
          o[0] = v[0];
          o[nJ] = v[1];
          o[2 * nJ] = v[2];
        }
      }
      __syncwarp();   /* the frames are read from the record region */
    }

    if (earlyFetch && s + sStep < nGroupsTotal)
    {
      fetch(s + sStep);   /* the record region is dead from here on */
    }

    /* ---- centre of gravity and its Jacobian (RcsGraph_COG / RcsGraph_COGJacobian,
     * Rcs_kinematics.c:904, :973): column k of sum_b m_b JT_b is the linear momentum p of
     * column k's composite under a unit joint rate ---------------------------------------- */
    if (a.cog || a.Jcog)
    {
      double mT = tail[1], mc0 = tail[2], mc1 = tail[3], mc2 = tail[4];
      for (int l = 0; l < p.hdr->nRoots; ++l)
      {
        const double* cp = buf + p.rootLinks[l];
        mT += cp[0];
        mc0 += cp[1];
        mc1 += cp[2];
        mc2 += cp[3];
      }
      const double inv = mT > 0.0 ? 1.0 / mT : 0.0;
      if (a.cog && gl == 0 && valid)
      {
        a.cog[s * 3] = mc0 * inv;
        a.cog[s * 3 + 1] = mc1 * inv;
        a.cog[s * 3 + 2] = mc2 * inv;
      }
      if (a.Jcog && valid)
      {
        for (int e = gl; e < 3 * nJ; e += LPS)
        {
          const int r = e / nJ, k = e - r * nJ;
          a.Jcog[s * 3 * nJ + e] = fb[FS * k + r] * inv;
        }
      }
      __syncwarp();   /* the composites are overwritten by the matrix staging below */
      if (midFetch && s + sStep < nGroupsTotal)
      {
        fetch(s + sStep);   /* the record region is dead from here on: the copy runs under the matrix phase */
      }
    }

    /* ---- mass matrix: only the structurally non-zero pairs (ancestor i, column k) ----------
     * M[i][k] = <joint i, momentum of column k>; M[k][i] the same with the transposed
     * inertia. Staged in shared memory (over the consumed runs), rows leave as coalesced
     * stores. */
    double tk = 0.0;
    const bool wantM = a.M || a.qdd;
    if (wantM || (a.E && VEL))
    {
      double* Mg = (a.M && valid) ? a.M + s * nJ * nJ : nullptr;
      if (stage)
      {
        for (int e = gl; e < nJ * nJ; e += LPS)
        {
          Ms[e] = 0.0;
        }
        __syncwarp();
      }
      else
      {
        /* no staging: zeros straight to global memory (coalesced), the non-zero pairs follow
         * after the barrier, which orders the two stores to the same address; L2 merges them.
         * The barrier is taken by every group of the warp, valid state or not. */
        if (Mg)
        {
          for (int e = gl; e < nJ * nJ; e += LPS)
          {
            Mg[e] = 0.0;
          }
        }
        __syncwarp();
      }
      const int nPairs = p.hdr->nPairs;
      for (int e = gl; e < nPairs; e += LPS)
      {
        const int4 pr = p.pairs[e];
        const double* ji = scol + pr.x;
        const double* fk = fb + pr.y;
        const bool diag = (pr.w & 2) != 0;
        const double pk[3] = {fk[0], fk[1], fk[2]};
        const double Lk[3] = {fk[3], fk[4], fk[5]};
        const double upper = pairing(ji, pk, Lk);
        double lower = upper;
        if (!sym && !diag)
        {
          const double Lkt[3] = {fk[6], fk[7], fk[8]};
          lower = pairing(ji, pk, Lkt);
        }
        if (stage)
        {
          Ms[pr.z & 0xffff] = upper;
          Ms[pr.z >> 16] = lower;
        }
        else if (Mg)
        {
          Mg[pr.z & 0xffff] = upper;
          Mg[pr.z >> 16] = lower;
        }
        if (VEL)
        {
          const double qq = ji[6] * scol[(pr.w >> 8) + 6];
          tk += diag ? qq * upper : qq * (upper + lower);
        }
      }
      if (stage)
      {
        __syncwarp();
        if (Mg)
        {
          for (int e = gl; e < nJ * nJ; e += LPS)
          {
            Mg[e] = Ms[e];
          }
        }
      }
    }

    /* ---- direct dynamics: M qdd = b by the reference's Cholesky solve, in shared memory ----
     * (Rcs_directDynamics, Rcs_dynamics.c:673; MatNd_choleskySolve, Rcs_linAlg.c:133; factor
     * :66-101). Column-oriented evaluation of the same sums: for column k every lane owns a
     * row i > k and forms A[i][k] - sum_{t<k} L[i][t] L[k][t] in the reference's order. */
    if (a.qdd)
    {
      __syncwarp();
      /* no early exit: several groups share a warp and its barriers */
      bool ok = true;
      for (int k = 0; k < nJ; ++k)
      {
        /* diagonal (every lane computes it: no broadcast needed) */
        double d = 0.0;
        for (int t = 0; t < k; ++t)
        {
          const double l = Ms[k * nJ + t];
          d += l * l;
        }
        const double piv = Ms[k * nJ + k] - d;
        const bool good = piv > 0.0;
        ok = ok && good;
        const double lkk = sqrt(good ? piv : 1.0);
        __syncwarp();
        if (gl == 0)
        {
          Ms[k * nJ + k] = lkk;
        }
        for (int i = k + 1 + gl; i < nJ; i += LPS)
        {
          double dot = 0.0;
          for (int t = 0; t < k; ++t)
          {
            dot += Ms[i * nJ + t] * Ms[k * nJ + t];
          }
          Ms[i * nJ + k] = (Ms[i * nJ + k] - dot) / lkk;
        }
        __syncwarp();
      }
      {
        /* forward solve L z = b, backward solve L^T x = z (lane 0 holds the running
         * element; the other lanes update the rows below / above); run unconditionally, the
         * result of a failed factorisation is discarded below */
        for (int i = 0; i < nJ; ++i)
        {
          const double xi = rhs[i] / Ms[i * nJ + i];
          __syncwarp();
          if (gl == 0)
          {
            rhs[i] = xi;
          }
          for (int r = i + 1 + gl; r < nJ; r += LPS)
          {
            rhs[r] -= Ms[r * nJ + i] * xi;
          }
          __syncwarp();
        }
        for (int i = nJ - 1; i >= 0; --i)
        {
          const double xi = rhs[i] / Ms[i * nJ + i];
          __syncwarp();
          if (gl == 0)
          {
            rhs[i] = xi;
          }
          for (int r = gl; r < i; r += LPS)
          {
            rhs[r] -= Ms[i * nJ + r] * xi;
          }
          __syncwarp();
        }
      }
      if (valid)
      {
        for (int k = gl; k < nJ; k += LPS)
        {
          a.qdd[s * nJ + k] = ok ? rhs[k] : 0.0;   /* failed factorisation: zeros, as the reference */
        }
      }
    }

    if (a.E)
    {
      if (VEL)
      {
#pragma unroll
        for (int off = LPS / 2; off > 0; off >>= 1)
        {
          tk += __shfl_xor_sync(0xffffffffu, tk, off);
        }
      }
      if (gl == 0 && valid)
      {
        a.E[s] = tail[0] + 0.5 * tk;
      }
    }
    __syncwarp();
    if (!earlyFetch && !midFetch && s + sStep < nGroupsTotal)
    {
      fetch(s + sStep);
    }
  }
}

template <bool VEL, bool WB>
static cudaError_t launchWalk(const KtArgs& a, int nSlots, cudaStream_t stream)
{
  const size_t smemBytes = (size_t) a.modelBytes + (size_t) a.planWalkBytes +
                           (size_t) (RCSB_KT_WALK_THREADS / 32) * 32 * 33 * sizeof(double) +
                           (size_t) RCSB_KT_LINK_REC_N(VEL) * RCSB_KT_WALK_THREADS * sizeof(double);
  const dim3 grid((unsigned int) ((a.N + RCSB_KT_WALK_THREADS - 1) / RCSB_KT_WALK_THREADS));
#define RCSB_LAUNCH(NS)                                                                        \
  {                                                                                            \
    auto k = ktWalkKernel<VEL, NS, WB>;                                                            \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) smemBytes);                                     \
    if (e != cudaSuccess)                                                                      \
    {                                                                                          \
      return e;                                                                                \
    }                                                                                          \
    k<<<grid, RCSB_KT_WALK_THREADS, smemBytes, stream>>>(a);                                   \
    return cudaGetLastError();                                                                 \
  }
  if (nSlots <= 0)
  {
    RCSB_LAUNCH(0)
  }
  else if (nSlots <= 2)
  {
    RCSB_LAUNCH(2)
  }
  else
  {
    RCSB_LAUNCH(RCSB_MAX_FRAME_SLOTS)
  }
#undef RCSB_LAUNCH
}

template <bool VEL, int LPS>
static cudaError_t launchAssemble(const KtArgs& a, int nJ, int, int smCount, cudaStream_t stream)
{
  constexpr int GPB = RCSB_KT_ASM_THREADS / LPS;
  /* group buffers 4 doubles apart modulo 16: the groups of a warp read different banks */
  const int asmDoubles = rcsbKtAsmLayout(a.recDoubles, nJ, a.stageM, a.symmetric).doubles;
  int groupStride = ((asmDoubles + 15) & ~15) + 4;
  const size_t smemBytes = (size_t) a.planBytes + (size_t) GPB * groupStride * sizeof(double);
  auto k = ktAssembleKernel<VEL, LPS>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int) smemBytes);
  if (e != cudaSuccess)
  {
    return e;
  }
  int perSm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k, RCSB_KT_ASM_THREADS, smemBytes);
  if (e != cudaSuccess)
  {
    return e;
  }
  long blocks = (a.N + GPB - 1) / GPB;
  const long resident = (long) smCount * (perSm > 0 ? perSm : 1);
  blocks = blocks < resident ? blocks : resident;
  k<<<(unsigned int) blocks, RCSB_KT_ASM_THREADS, smemBytes, stream>>>(a, groupStride);
  return cudaGetLastError();
}

}   // namespace rcsb

extern "C" cudaError_t rcsb_launch_kt(const rcsb::KtArgs* a, int nSlots, int nJ, int asmDoubles,
                                      int smCount, cudaStream_t stream)
{
  using namespace rcsb;
  if (a->N <= 0)
  {
    return cudaSuccess;
  }
  const bool vel = a->qd != nullptr;
  cudaError_t e;
  if (a->wholeBody)
  {
    e = vel ? launchWalk<true, true>(*a, nSlots, stream) : launchWalk<false, true>(*a, nSlots, stream);
  }
  else
  {
    e = vel ? launchWalk<true, false>(*a, nSlots, stream) : launchWalk<false, false>(*a, nSlots, stream);
  }
  if (e != cudaSuccess)
  {
    return e;
  }
  if (nJ <= 8)
  {
    return vel ? launchAssemble<true, 8>(*a, nJ, asmDoubles, smCount, stream)
               : launchAssemble<false, 8>(*a, nJ, asmDoubles, smCount, stream);
  }
  if (nJ <= 16)
  {
    return vel ? launchAssemble<true, 16>(*a, nJ, asmDoubles, smCount, stream)
               : launchAssemble<false, 16>(*a, nJ, asmDoubles, smCount, stream);
  }
  return vel ? launchAssemble<true, 32>(*a, nJ, asmDoubles, smCount, stream)
             : launchAssemble<false, 32>(*a, nJ, asmDoubles, smCount, stream);
}
