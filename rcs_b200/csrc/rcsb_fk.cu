/*
 * Batched forward kinematics + world-point / rotation Jacobians for sm_100a.
 *
 * What it computes, per state (reference, one state per call on one RcsGraph):
 *   RcsGraph_setState / RcsGraph_internalBodyKinematics   src/RcsCore/Rcs_graph.c:61-296
 *   RcsGraph_bodyPointJacobian / worldPointJacobian       src/RcsCore/Rcs_kinematics.c:80-215
 *   RcsGraph_rotationJacobian                             src/RcsCore/Rcs_kinematics.c:416-458
 *
 * How: one thread per state walks the flattened tree depth first. The walk is the same for
 * every state, so all model reads are warp-uniform shared-memory broadcasts (128-bit) and
 * there is no divergence. The running frame lives in registers. Leaf bodies are evaluated
 * on a register copy; a body whose parent frame is not in registers reloads it from a small
 * per-thread frame stack (only bodies that need it are parked there). Body frames leave as
 * three 32-byte stores per frame (STG.E.256, L1 no-allocate: full sectors, nothing to
 * merge, nothing displaced). Joint axes and origins on an effector's chain are parked in
 * padded shared memory; once the effector point is known the owning thread turns the
 * origins into Jacobian columns in place, and the warp streams the 3 x nJ records of its
 * 32 consecutive states to global memory as one contiguous, coalesced run.
 *
 * Bound: HBM (output stream); see DESIGN.md for the byte counts.
 */
#include <cstdlib>
#include "rcsb_device.cuh"
#include "rcsb_kernels.h"
#include "rcsb_plan.h"

namespace rcsb
{

struct PlanView
{
  const RcsbFkPlanHeader* hdr;
  const int* bodyOut;
  const int* jntScr;
  const int* bodyEffStart;
  const int* bodyEffList;
  const double* effPt;
  const int* jtab;
  const int* jtabR;
  const int* effChainStart;
  const int* effChain;
  const int* bodyPathMask;
  const int4* massList;
  const int4* visit;
  const int* jntNext;
  const double* chainG;
  const double* chainX;
  const int* chainOutStart;
  const int4* chainOut;
  const double* chainMass;
  const int* chainJoint;

  __device__ __forceinline__ void bind(const char* blob)
  {
    hdr = reinterpret_cast<const RcsbFkPlanHeader*>(blob);
    bodyOut = reinterpret_cast<const int*>(blob + hdr->offBodyOut);
    jntScr = reinterpret_cast<const int*>(blob + hdr->offJntScr);
    bodyEffStart = reinterpret_cast<const int*>(blob + hdr->offBodyEffStart);
    bodyEffList = reinterpret_cast<const int*>(blob + hdr->offBodyEffList);
    effPt = reinterpret_cast<const double*>(blob + hdr->offEffPt);
    jtab = reinterpret_cast<const int*>(blob + hdr->offJTab);
    jtabR = reinterpret_cast<const int*>(blob + hdr->offJTabR);
    effChainStart = reinterpret_cast<const int*>(blob + hdr->offEffChainStart);
    effChain = reinterpret_cast<const int*>(blob + hdr->offEffChain);
    bodyPathMask = reinterpret_cast<const int*>(blob + hdr->offBodyPathMask);
    massList = reinterpret_cast<const int4*>(blob + hdr->offMassList);
    visit = reinterpret_cast<const int4*>(blob + hdr->offVisit);
    jntNext = reinterpret_cast<const int*>(blob + hdr->offJntNext);
    chainG = reinterpret_cast<const double*>(blob + hdr->offChainG);
    chainX = reinterpret_cast<const double*>(blob + hdr->offChainX);
    chainOutStart = reinterpret_cast<const int*>(blob + hdr->offChainOutStart);
    chainOut = reinterpret_cast<const int4*>(blob + hdr->offChainOut);
    chainMass = reinterpret_cast<const double*>(blob + hdr->offChainMass);
    chainJoint = reinterpret_cast<const int*>(blob + hdr->offChainJoint);
  }
};

/* Jacobian records and mass matrices of a warp's 32 states leave through the shared-memory scratch
 * (shared by fkKernel and fkChainKernel). On entry the scratch holds, per chain joint, axis and
 * origin (MASS != 2) or the spatial axis (w, v) (MASS == 2), and the effector points. */
template <int MASS>
__device__ __forceinline__ void fkTail(const FkArgs& a, const PlanView& p, double* scr, const int scrStride,
                                       const int nChain, const int tid, const long s,
                                       const double (&Mv)[MASS ? 36 : 1])
{
  /* ---- Jacobians --------------------------------------------------------------------------- */
  const int rec = p.hdr->jacRecord;
  if (rec > 0 && (a.JT || a.JR))
  {
    const int rs = nChain * scrStride;
    const bool inPlace = p.hdr->inPlace != 0;

    if (inPlace)
    {
      /* own column: joint origin -> axis x (I_p - origin), or the axis for a translation */
      for (int e = 0; e < p.hdr->nEff; ++e)
      {
        const double* pe = scr + (6 * nChain + 3 * e) * scrStride + tid;
        const double px = pe[0], py = pe[scrStride], pz = pe[2 * scrStride];
        for (int k = p.effChainStart[e]; k < p.effChainStart[e + 1]; ++k)
        {
          const int ent = p.effChain[k];
          double* col = scr + (ent & 0xffff) * scrStride + tid;
          const double a0 = col[0], a1 = col[rs], a2 = col[2 * rs];
          if (MASS == 2)
          {
            /* (w, v) storage: column = w x I_p + v for both joint types */
            col[3 * rs] += a1 * pz - a2 * py;
            col[4 * rs] += a2 * px - a0 * pz;
            col[5 * rs] += a0 * py - a1 * px;
          }
          else if (ent >> 16)
          {
            const double d0 = px - col[3 * rs], d1 = py - col[4 * rs], d2 = pz - col[5 * rs];
            col[3 * rs] = a1 * d2 - a2 * d1;
            col[4 * rs] = a2 * d0 - a0 * d2;
            col[5 * rs] = a0 * d1 - a1 * d0;
          }
          else
          {
            col[3 * rs] = a0;
            col[4 * rs] = a1;
            col[5 * rs] = a2;
          }
        }
      }
    }

    __syncwarp();
    const int lane = tid & 31;
    const long w0 = s - lane;                      /* first state of this warp */
    long nv = a.N - w0;
    nv = nv < 0 ? 0 : (nv > 32 ? 32 : nv);
    const int total = (int) nv * rec;
    const float invRec = p.hdr->invJacRecord;
    const double* scrWarp = scr + (tid - lane);
    double* JTw = a.JT ? a.JT + w0 * rec : nullptr;
    double* JRw = a.JR ? a.JR + w0 * rec : nullptr;

    if (inPlace)
    {
      /* one record row per warp instruction: lane r owns element r of the record, so its
       * scratch row is fixed (structural zeros read the zero row) and the loop over the
       * warp's states is a shared-memory read and a global store with constant strides, no
       * index arithmetic; the 32 records of the warp are contiguous in global memory */
      const double* zero = scrWarp + p.hdr->zeroRow * scrStride;
      for (int r0 = 0; r0 < rec; r0 += 32)
      {
        const int rem = r0 + lane;
        if (rem < rec)
        {
          const int tT = p.jtab[rem], tR = p.jtabR[rem];
          const double* srcT = tT >= 0 ? scrWarp + tT * scrStride : zero;
          const double* srcR = tR >= 0 ? scrWarp + tR * scrStride : zero;
          if (JTw && JRw)
          {
            double* dT = JTw + rem;
            double* dR = JRw + rem;
#pragma unroll 4
            for (int sl = 0; sl < (int) nv; ++sl)
            {
              dT[(long) sl * rec] = srcT[sl];
              dR[(long) sl * rec] = srcR[sl];
            }
          }
          else
          {
            double* d = JTw ? JTw + rem : JRw + rem;
            const double* src = JTw ? srcT : srcR;
#pragma unroll 4
            for (int sl = 0; sl < (int) nv; ++sl)
            {
              d[(long) sl * rec] = src[sl];
            }
          }
        }
      }
    }
    else
    {
      for (int e = lane; e < total; e += 32)
      {
        const int sl = __float2int_rd(((float) e + 0.5f) * invRec);
        const int rem = e - sl * rec;
        const int t = p.jtab[rem];
        double vT = 0.0, vR = 0.0;
        if (t >= 0)
        {
          const int r = RCSB_JTAB_ROW(t);
          const double* ja = scrWarp + RCSB_JTAB_SLOT(t) * scrStride + sl;
          if (MASS == 2)
          {
            const int r1 = r == 2 ? 0 : r + 1;
            const int r2 = r == 0 ? 2 : r - 1;
            const double* pe = scrWarp + (6 * nChain + 3 * RCSB_JTAB_EFF(t)) * scrStride + sl;
            vT = ja[(3 + r) * rs] + (ja[r1 * rs] * pe[r2 * scrStride] - ja[r2 * rs] * pe[r1 * scrStride]);
            vR = ja[r * rs];
          }
          else if (RCSB_JTAB_ISROT(t))
          {
            /* column = axis x (I_p - joint origin), component r */
            const int r1 = r == 2 ? 0 : r + 1;
            const int r2 = r == 0 ? 2 : r - 1;
            const double* pe = scrWarp + (6 * nChain + 3 * RCSB_JTAB_EFF(t)) * scrStride + sl;
            const double a1 = ja[r1 * rs], a2 = ja[r2 * rs];
            const double d1 = pe[r1 * scrStride] - ja[(3 + r1) * rs];
            const double d2 = pe[r2 * scrStride] - ja[(3 + r2) * rs];
            vT = a1 * d2 - a2 * d1;
            vR = ja[r * rs];
          }
          else
          {
            vT = ja[r * rs];
          }
        }
        if (JTw)
        {
          JTw[e] = vT;
        }
        if (JRw)
        {
          JRw[e] = vR;
        }
      }
    }
  }

  if (MASS && a.M)
  {
    /* stage the matrices over the scratch rows the Jacobians have left, then copy out: the
     * warp's 32 matrices are contiguous in global memory */
    __syncwarp();
    {
      const int nJ = p.hdr->nJ;
      double* st = scr + p.hdr->mStageRow * scrStride + tid;
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
#pragma unroll
        for (int i = 0; i <= k; ++i)
        {
          if (k < nJ)
          {
            st[(i * nJ + k) * scrStride] = Mv[k * (k + 1) / 2 + i];
            st[(k * nJ + i) * scrStride] = Mv[k * (k + 1) / 2 + i];
          }
        }
      }
    }
    __syncwarp();
    const int lane = tid & 31;
    const long w0 = s - lane;
    long nv = a.N - w0;
    nv = nv < 0 ? 0 : (nv > 32 ? 32 : nv);
    const int nJ2 = p.hdr->nJ * p.hdr->nJ;
    const double* stw = scr + p.hdr->mStageRow * scrStride + (tid - lane);
    double* Mw = a.M + w0 * nJ2;
    for (int r0 = 0; r0 < nJ2; r0 += 32)
    {
      const int rem = r0 + lane;
      if (rem < nJ2)
      {
        const double* src = stw + rem * scrStride;
        double* d = Mw + rem;
#pragma unroll 4
        for (int sl = 0; sl < (int) nv; ++sl)
        {
          d[(long) sl * nJ2] = src[sl];
        }
      }
    }
  }
}

/* MASS: additionally the joint-space mass matrix (RcsGraph_computeMassMatrix, Rcs_dynamics.c:
 * 610) for models with at most 8 Jacobian columns.
 * MASS = 1 (any tree): for every massive body and every column k that moves it, the body's
 * momentum under a unit rate of joint k, (p, L) about the world origin, is added to the
 * column's accumulator F[k] (registers); after the walk M[i][k] = <joint i, F[k]> for i an
 * ancestor-or-self of k.
 * MASS = 2 (serial chains whose body frames are all written, e.g. an arm): the walk stays the
 * plain FK walk. Afterwards the thread reads the frames of the massive bodies back (its own
 * stores, still in L2), tip to root, sums their inertia about the world origin into one
 * composite (10 doubles), and at every column k takes F[k] = composite x joint k and
 * M[i][k] = <joint i, F[k]>, i <= k. Joints are kept in shared memory as (w, v): angular part
 * and velocity of the point at the world origin under a unit rate (rotation: a, o x a;
 * translation: 0, a), so that M[i][k] = w_i . L + v_i . p and a Jacobian column is w x p + v.
 * 48 accumulator doubles less than MASS = 1: 3 blocks per SM instead of 2. */
#ifndef RCSB_FKM_CHAIN_BLOCKS
#define RCSB_FKM_CHAIN_BLOCKS 3
#endif
#ifndef RCSB_FKM_MIN_BLOCKS
#define RCSB_FKM_MIN_BLOCKS 1
#endif
/* with velocities: 3 resident blocks per SM pay off for shallow frame stacks (arm: 0.58 ->
 * 0.45 ms per 1M states), not for deep ones (humanoid: spills) */
/* PRUNED: the caller selected some bodies only; the walk visits them, the effector bodies and
 * their ancestors (plan tables visit / jntNext) instead of the whole tree. */
template <bool VEL, bool ALIGNED32, int NSLOTS, int MASS, bool PRUNED = false>
__global__ void __launch_bounds__(RCSB_FK_BLOCK(MASS), MASS == 2 ? RCSB_FKM_CHAIN_BLOCKS : (MASS ? RCSB_FKM_MIN_BLOCKS : (VEL ? (NSLOTS <= 2 ? 3 : 1) : 4)))
fkKernel(const FkArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  double* scr = reinterpret_cast<double*>(sPlan + a.planBytes);
  /* VEL: two 32-record windows per warp behind the scratch (x_dot, omega) */
  double* velWin = scr + a.scrDoubles;

  /* model + plan by two TMA bulk copies (fused launch 0.500 -> 0.475 ms per 1M states against
   * per-thread copy loops and a block barrier) */
  stageBlobsTma(sModel, a.model, a.modelBytes, sPlan, a.plan, a.planBytes);

  ModelView m;
  m.bind(sModel);
  PlanView p;
  p.bind(sPlan);

  constexpr int THREADS = RCSB_FK_BLOCK(MASS);
  const int tid = threadIdx.x;
  const int scrStride = THREADS + 1;           /* odd stride: conflict-free transposed reads */
  if (p.hdr->jacRecord > 0)
  {
    scr[p.hdr->zeroRow * scrStride + tid] = 0.0;   /* source of the Jacobians' structural zeros */
  }
  const long s = (long) blockIdx.x * THREADS + tid;
  const bool active = s < a.N;
  const long sc = active ? s : a.N - 1;        /* idle lanes shadow the last state, no stores */

  const int nb = m.hdr->nb;
  const int dof = m.hdr->dof;
  const int nSel = p.hdr->nSel;
  const int nChain = p.hdr->nChain;

  StateReader rd;
  rd.full = (a.qdim == dof);
  rd.q = a.q + sc * a.qdim;
  rd.qd = VEL ? a.qd + sc * a.qdim : nullptr;
  rd.couple = (a.noCoupling == 0);

  constexpr int SLOT_DOUBLES = VEL ? 18 : 12;
  double stk[(NSLOTS > 0 ? NSLOTS : 1) * SLOT_DOUBLES];
  double F[MASS == 1 ? 8 : 1][6];
  if (MASS == 1)
  {
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
#pragma unroll
      for (int i = 0; i < 6; ++i)
      {
        F[k][i] = 0.0;
      }
    }
  }

  /* x_dot / omega of the selected bodies stream out through shared-memory windows as full
   * 256-byte rows per state (scalar 24-byte stores per body cost 3x the kernel time); only
   * when the selection follows the walk order */
  const bool velStream = VEL && p.hdr->velStream != 0 && (a.xdot || a.omega);
  RecordWriter wx, wo;
  if (VEL)
  {
    const int lane = tid & 31;
    const long w0 = s - lane;
    long nvl = a.N - w0;
    nvl = nvl < 0 ? 0 : (nvl > 32 ? 32 : nvl);
    wx.win = velWin + (tid >> 5) * 2 * 32 * 33;
    wo.win = wx.win + 32 * 33;
    wx.dst = a.xdot ? a.xdot + w0 * nSel * 3 : nullptr;
    wo.dst = a.omega ? a.omega + w0 * nSel * 3 : nullptr;
    wx.recDoubles = wo.recDoubles = nSel * 3;
    wx.nv = wo.nv = (int) nvl;
    wx.lane = wo.lane = lane;
    wx.p = wo.p = 0;
  }

  Frame f;
  double xd[3] = {0.0, 0.0, 0.0}, om[3] = {0.0, 0.0, 0.0};
  setIdentity(f);
  /* joint values are fetched two joints ahead of their use */
  double nextQ = 0.0, nextQd = 0.0, next2Q = 0.0, next2Qd = 0.0;
  const int jFirst = PRUNED ? p.hdr->firstJnt : (dof > 0 ? 0 : -1);
  auto jointAfter = [&](int j) { return PRUNED ? p.jntNext[j] : (j + 1 < dof ? j + 1 : -1); };
  if (jFirst >= 0)
  {
    nextQ = rd.fetchQ(m, jFirst);
    if (VEL)
    {
      nextQd = rd.fetchQd(m, jFirst);
    }
    const int jSecond = jointAfter(jFirst);
    if (jSecond >= 0)
    {
      next2Q = rd.fetchQ(m, jSecond);
      if (VEL)
      {
        next2Qd = rd.fetchQd(m, jSecond);
      }
    }
  }

  /* the model lives in shared memory next to the Jacobian scratch the kernel writes, so the
   * compiler cannot hoist its loads: the per-body and per-joint descriptors are fetched one
   * step ahead by hand */
  const int nWalk = PRUNED ? p.hdr->nVisit : nb;
  int nBody = 0, nLd, nFlags, nSave = -1, nJ0, nCnt;
  if (PRUNED)
  {
    const int4 v = p.visit[0];
    nBody = v.x;
    nLd = v.y;
    nSave = v.z;
    nFlags = (m.bodyFlags[nBody] & ~RCSB_BF_LEAF) | (v.w ? RCSB_BF_LEAF : 0);
  }
  else
  {
    nLd = m.bodyLoad[0];
    nFlags = m.bodyFlags[0];
  }
  nJ0 = m.bodyJntStart[nBody];
  nCnt = m.bodyJntCount[nBody];
  int nJf = jFirst >= 0 ? m.jntFlags[jFirst] : 0, nType = jFirst >= 0 ? m.jntType[jFirst] : 0;

  for (int vi = 0; vi < nWalk; ++vi)
  {
    /* ---- parent frame ---------------------------------------------------------------- */
    const int b = PRUNED ? nBody : vi;
    const int ld = nLd;
    const int bflags = nFlags;
    const int svp = nSave;
    const int j0 = nJ0;
    const int j1 = j0 + nCnt;
    if (vi + 1 < nWalk)
    {
      if (PRUNED)
      {
        const int4 v = p.visit[vi + 1];
        nBody = v.x;
        nLd = v.y;
        nSave = v.z;
        nFlags = (m.bodyFlags[nBody] & ~RCSB_BF_LEAF) | (v.w ? RCSB_BF_LEAF : 0);
      }
      else
      {
        nBody = vi + 1;
        nLd = m.bodyLoad[vi + 1];
        nFlags = m.bodyFlags[vi + 1];
      }
      nJ0 = m.bodyJntStart[nBody];
      nCnt = m.bodyJntCount[nBody];
    }
    if (ld == RCSB_LOAD_IDENT)
    {
      setIdentity(f);
      if (VEL)
      {
        xd[0] = xd[1] = xd[2] = 0.0;
        om[0] = om[1] = om[2] = 0.0;
      }
    }
    else if (NSLOTS > 0 && ld >= 0)
    {
      const double* sp = stk + ld * SLOT_DOUBLES;
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
      if (VEL)
      {
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          xd[i] = sp[12 + i];
          om[i] = sp[15 + i];
        }
      }
    }

    /* The body is evaluated in place when the walk continues below it, and on a copy of the
     * parent frame when it is a leaf (the walk then carries on from the parent): two inlined
     * instances behind a warp-uniform branch instead of predicated register copies. */
    auto evalBody = [&](Frame& f, double (&xd)[3], double (&om)[3])
    {
    double op[3], w[3], c[3], vt[3];
    if (VEL)
    {
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        op[i] = f.o[i];
        w[i] = c[i] = vt[i] = 0.0;
      }
    }

    /* ---- joints of the body ------------------------------------------------------------ */
    for (int j = j0; j < j1; ++j)
    {
      const int jf = nJf;
      const int jtype = nType;
      const int jn = jointAfter(j);
      if (jn >= 0)
      {
        nJf = m.jntFlags[jn];
        nType = m.jntType[jn];
      }
      if (jf & RCSB_JF_HAS_AJP)
      {
        applyRelF(f, m.jntAJP + 12 * j, jf);
      }

      /* joints are visited in jointIndex order: the loads of the next joint are in flight
       * while this one is processed */
      const double rawQ = nextQ, rawQd = nextQd;
      nextQ = next2Q;
      nextQd = next2Qd;
      const int jn2 = jn >= 0 ? jointAfter(jn) : -1;
      if (jn2 >= 0)
      {
        next2Q = rd.fetchQ(m, jn2);
        if (VEL)
        {
          next2Qd = rd.fetchQd(m, jn2);
        }
      }
      const double qj = rd.finishQ(m, j, rawQ);
      const int dir = jtype % 3;
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

      if (VEL)
      {
        /* x_dot += sum_trans a qd + sum_rot (a qd) x (r_b - o_j), omega += sum_rot a qd;
         * r_b is only known after the last joint, so the cross products are split into
         * w x r_b - sum (a qd) x o_j. */
        const double qdj = rd.finishQd(m, j, rawQ, rawQd);
        if (a.qdOut && active)
        {
          a.qdOut[s * dof + j] = qdj;
        }
        const double aq[3] = {ax[0] * qdj, ax[1] * qdj, ax[2] * qdj};
        if (jf & RCSB_JF_ROT)
        {
          double t[3];
          cross3(t, aq, f.o);
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            w[i] += aq[i];
            c[i] += t[i];
          }
        }
        else
        {
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            vt[i] += aq[i];
          }
        }
      }

      if (a.qOut && active)
      {
        a.qOut[s * dof + j] = qj;
      }
      if (a.A_JI && active)
      {
        storeFrame<ALIGNED32>(a.A_JI + (s * dof + j) * 12, f);
      }

      const int slot = p.jntScr[j];
      if (slot >= 0)
      {
        double* col = scr + slot * scrStride + tid;
        const int rs = nChain * scrStride;
        if (MASS == 2)
        {
          const bool rot = (jf & RCSB_JF_ROT) != 0;
          col[0] = rot ? ax[0] : 0.0;
          col[rs] = rot ? ax[1] : 0.0;
          col[2 * rs] = rot ? ax[2] : 0.0;
          col[3 * rs] = rot ? f.o[1] * ax[2] - f.o[2] * ax[1] : ax[0];
          col[4 * rs] = rot ? f.o[2] * ax[0] - f.o[0] * ax[2] : ax[1];
          col[5 * rs] = rot ? f.o[0] * ax[1] - f.o[1] * ax[0] : ax[2];
        }
        else
        {
          col[0] = ax[0];
          col[rs] = ax[1];
          col[2 * rs] = ax[2];
          col[3 * rs] = f.o[0];
          col[4 * rs] = f.o[1];
          col[5 * rs] = f.o[2];
        }
      }
    }

    /* ---- body frame ---------------------------------------------------------------------- */
    if (bflags & RCSB_BF_HAS_ABP)
    {
      applyRelF(f, m.bodyABP + 12 * b, bflags);
    }

    if (VEL)
    {
      if (ld != RCSB_LOAD_IDENT)
      {
        const double d[3] = {f.o[0] - op[0], f.o[1] - op[1], f.o[2] - op[2]};
        double t[3];
        cross3(t, om, d);
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          xd[i] += t[i];
        }
      }
      double t[3];
      cross3(t, w, f.o);
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        xd[i] += vt[i] + (t[i] - c[i]);
        om[i] += w[i];
      }
    }

    if (NSLOTS > 0)
    {
      const int sv = PRUNED ? svp : m.bodySave[b];
      if (sv >= 0)
      {
        double* sp = stk + sv * SLOT_DOUBLES;
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
        if (VEL)
        {
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            sp[12 + i] = xd[i];
            sp[15 + i] = om[i];
          }
        }
      }
    }

    /* ---- mass properties: momentum of the body under unit joint rates ---------------------- */
    if (MASS == 1 && (bflags & RCSB_BF_HAS_MASS))
    {
      const double mass = m.bodyMass[b];
      const double* cl = m.bodyCom + 3 * b;
      const double* Ib = m.bodyInertia + 9 * b;
      const double* R = f.R;
      const double c[3] = {f.o[0] + (R[0] * cl[0] + R[3] * cl[1] + R[6] * cl[2]),
                           f.o[1] + (R[1] * cl[0] + R[4] * cl[1] + R[7] * cl[2]),
                           f.o[2] + (R[2] * cl[0] + R[5] * cl[1] + R[8] * cl[2])};
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
      const int mask = p.bodyPathMask[b];
      const int rotMask = p.hdr->rotMask;
      const int rs = nChain * scrStride;
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
        if ((mask >> k) & 1)
        {
          const double* col = scr + k * scrStride + tid;
          const double ak[3] = {col[0], col[rs], col[2 * rs]};
          double pk[3], Lk[3];
          if ((rotMask >> k) & 1)
          {
            const double d[3] = {c[0] - col[3 * rs], c[1] - col[4 * rs], c[2] - col[5 * rs]};
            double jt[3];
            cross3(jt, ak, d);                 /* velocity of the COM: a x (c - o) */
            pk[0] = mass * jt[0];
            pk[1] = mass * jt[1];
            pk[2] = mass * jt[2];
            double cp[3];
            cross3(cp, c, pk);
#pragma unroll
            for (int i = 0; i < 3; ++i)
            {
              Lk[i] = Iw[3 * i] * ak[0] + Iw[3 * i + 1] * ak[1] + Iw[3 * i + 2] * ak[2] + cp[i];
            }
          }
          else
          {
            pk[0] = mass * ak[0];
            pk[1] = mass * ak[1];
            pk[2] = mass * ak[2];
            cross3(Lk, c, pk);
          }
#pragma unroll
          for (int i = 0; i < 3; ++i)
          {
            F[k][i] += pk[i];
            F[k][3 + i] += Lk[i];
          }
        }
      }
    }

    const int out = p.bodyOut[b];
    if (out >= 0 && active)
    {
      if (a.A_BI)
      {
        storeFrame<ALIGNED32>(a.A_BI + (s * nSel + out) * 12, f);
      }
      if (VEL && !velStream)
      {
        if (a.xdot)
        {
          double* d = a.xdot + (s * nSel + out) * 3;
          d[0] = xd[0]; d[1] = xd[1]; d[2] = xd[2];
        }
        if (a.omega)
        {
          double* d = a.omega + (s * nSel + out) * 3;
          d[0] = om[0]; d[1] = om[1]; d[2] = om[2];
        }
      }
    }
    if (VEL && velStream && out >= 0)
    {
      /* warp-uniform: every lane (shadow lanes too) feeds its window column */
      if (a.xdot)
      {
        wx.emitN(xd);
      }
      if (a.omega)
      {
        wo.emitN(om);
      }
    }

    /* effector points attached to this body: I_p = o + R^T B_p */
    for (int k = p.bodyEffStart[b]; k < p.bodyEffStart[b + 1]; ++k)
    {
      const int e = p.bodyEffList[k];
      const double* pt = a.effPtInline ? a.effPtArg + 3 * e : p.effPt + 3 * e;
      double* col = scr + (6 * nChain + 3 * e) * scrStride + tid;
      col[0] = f.o[0] + (f.R[0] * pt[0] + f.R[3] * pt[1] + f.R[6] * pt[2]);
      col[scrStride] = f.o[1] + (f.R[1] * pt[0] + f.R[4] * pt[1] + f.R[7] * pt[2]);
      col[2 * scrStride] = f.o[2] + (f.R[2] * pt[0] + f.R[5] * pt[1] + f.R[8] * pt[2]);
    }

    };
    if (bflags & RCSB_BF_LEAF)
    {
      Frame g = f;
      double gx[3] = {xd[0], xd[1], xd[2]}, go[3] = {om[0], om[1], om[2]};
      evalBody(g, gx, go);
    }
    else
    {
      evalBody(f, xd, om);
    }
  }

  if (VEL && velStream)
  {
    if (a.xdot)
    {
      wx.finish();
    }
    if (a.omega)
    {
      wo.finish();
    }
  }

  /* ---- mass matrix entries (upper triangle, registers): computed before the Jacobian pass
   * overwrites the joint origins, staged after it over the same scratch rows ------------------ */
  double Mv[MASS ? 36 : 1];
  if (MASS == 2 && a.M)
  {
    const int nJ = p.hdr->nJ;
    const int rs = nChain * scrStride;
    const int nMass = p.hdr->nMass;
    const double* frames = a.A_BI + sc * nSel * 12;
    /* composite about the world origin: mass, first moment, inertia (xx, xy, xz, yy, yz, zz) */
    double cm = 0.0, ch[3] = {0.0, 0.0, 0.0}, cI[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int li = 0;
    int4 ent = p.massList[0];
#pragma unroll
    for (int k = 7; k >= 0; --k)
    {
      if (k < nJ)
      {
        while (li < nMass && ent.z == k + 1)
        {
          /* own stores of the walk, read back from L2 */
          const double* fp = frames + ent.y * 12;
          double o[3], R[9];
          if (ALIGNED32)
          {
            load4cg(fp, o[0], o[1], o[2], R[0]);
            load4cg(fp + 4, R[1], R[2], R[3], R[4]);
            load4cg(fp + 8, R[5], R[6], R[7], R[8]);
          }
          else
          {
#pragma unroll
            for (int i = 0; i < 3; ++i)
            {
              o[i] = __ldcg(fp + i);
            }
#pragma unroll
            for (int i = 0; i < 9; ++i)
            {
              R[i] = __ldcg(fp + 3 + i);
            }
          }
          const int b = ent.x;
          const double mass = m.bodyMass[b];
          const double* cl = m.bodyCom + 3 * b;
          const double* Ib = m.bodyInertia + 9 * b;
          ++li;
          if (li < nMass)
          {
            ent = p.massList[li];
          }
          const double c[3] = {o[0] + (R[0] * cl[0] + R[3] * cl[1] + R[6] * cl[2]),
                               o[1] + (R[1] * cl[0] + R[4] * cl[1] + R[7] * cl[2]),
                               o[2] + (R[2] * cl[0] + R[5] * cl[1] + R[8] * cl[2])};
          /* R^T I_b R (Mat3d_similarityTransform), symmetric: upper triangle only */
          double t[9];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
            {
              t[3 * i + jj] = Ib[3 * i] * R[jj] + Ib[3 * i + 1] * R[3 + jj] + Ib[3 * i + 2] * R[6 + jj];
            }
          const double mc[3] = {mass * c[0], mass * c[1], mass * c[2]};
          cm += mass;
          ch[0] += mc[0];
          ch[1] += mc[1];
          ch[2] += mc[2];
          cI[0] += (R[0] * t[0] + R[3] * t[3] + R[6] * t[6]) + (mc[1] * c[1] + mc[2] * c[2]);
          cI[1] += (R[0] * t[1] + R[3] * t[4] + R[6] * t[7]) - mc[0] * c[1];
          cI[2] += (R[0] * t[2] + R[3] * t[5] + R[6] * t[8]) - mc[0] * c[2];
          cI[3] += (R[1] * t[1] + R[4] * t[4] + R[7] * t[7]) + (mc[0] * c[0] + mc[2] * c[2]);
          cI[4] += (R[1] * t[2] + R[4] * t[5] + R[7] * t[8]) - mc[1] * c[2];
          cI[5] += (R[2] * t[2] + R[5] * t[5] + R[8] * t[8]) + (mc[0] * c[0] + mc[1] * c[1]);
        }
        /* momentum of the composite under a unit rate of column k: p = m v + w x h,
         * L = I w + h x v */
        const double* col = scr + k * scrStride + tid;
        const double w[3] = {col[0], col[rs], col[2 * rs]};
        const double v[3] = {col[3 * rs], col[4 * rs], col[5 * rs]};
        const double pk[3] = {cm * v[0] + (w[1] * ch[2] - w[2] * ch[1]),
                              cm * v[1] + (w[2] * ch[0] - w[0] * ch[2]),
                              cm * v[2] + (w[0] * ch[1] - w[1] * ch[0])};
        const double Lk[3] = {cI[0] * w[0] + cI[1] * w[1] + cI[2] * w[2] + (ch[1] * v[2] - ch[2] * v[1]),
                              cI[1] * w[0] + cI[3] * w[1] + cI[4] * w[2] + (ch[2] * v[0] - ch[0] * v[2]),
                              cI[2] * w[0] + cI[4] * w[1] + cI[5] * w[2] + (ch[0] * v[1] - ch[1] * v[0])};
        Mv[k * (k + 1) / 2 + k] = (w[0] * Lk[0] + w[1] * Lk[1] + w[2] * Lk[2]) +
                                  (v[0] * pk[0] + v[1] * pk[1] + v[2] * pk[2]);
#pragma unroll
        for (int i = 0; i < k; ++i)
        {
          const double* ci = scr + i * scrStride + tid;
          Mv[k * (k + 1) / 2 + i] = (ci[0] * Lk[0] + ci[rs] * Lk[1] + ci[2 * rs] * Lk[2]) +
                                    (ci[3 * rs] * pk[0] + ci[4 * rs] * pk[1] + ci[5 * rs] * pk[2]);
        }
      }
    }
  }
  if (MASS == 1 && a.M)
  {
    const int nJ = p.hdr->nJ;
    const int rs = nChain * scrStride;
    const int rotMask = p.hdr->rotMask;
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      const int rel = k < nJ ? (p.hdr->relMask[k] | (1 << k)) : 0;
#pragma unroll
      for (int i = 0; i <= k; ++i)
      {
        double val = 0.0;
        if ((rel >> i) & 1)
        {
          const double* col = scr + i * scrStride + tid;
          const double ai[3] = {col[0], col[rs], col[2 * rs]};
          if ((rotMask >> i) & 1)
          {
            /* a . (L - o x p) */
            const double oi[3] = {col[3 * rs], col[4 * rs], col[5 * rs]};
            const double pk[3] = {F[k][0], F[k][1], F[k][2]};
            double t[3];
            cross3(t, oi, pk);
            val = ai[0] * (F[k][3] - t[0]) + ai[1] * (F[k][4] - t[1]) + ai[2] * (F[k][5] - t[2]);
          }
          else
          {
            val = ai[0] * F[k][0] + ai[1] * F[k][1] + ai[2] * F[k][2];
          }
        }
        Mv[k * (k + 1) / 2 + i] = val;
      }
    }
  }

  fkTail<MASS>(a, p, scr, scrStride, nChain, tid, s, Mv);
}


/*
 * Serial chains of rotations (RcsbFkPlanHeader::chainNc): the same results as fkKernel<MASS = 0 / 2>
 * from straight-line code. NC columns at compile time; per column ONE relative transform folded on
 * the host (G_c) and one rotation about row 2 of the row-shifted base; the frames of the bodies of
 * a level are side outputs X_b o base (they do not touch the running frame, so nothing merges); the
 * sines / cosines of all columns come from one interleaved batch (sincosBatch). The flat model is
 * not needed on the device at all - the plan carries every constant - and is not staged.
 * MASS = 2: the matrix is built tip to root like in fkKernel, but without reading any frame back: the
 * running base is taken back through the columns by the inverse transforms (R = G^T R', o = o' -
 * R^T g, rotation by -q), and the massive bodies of a level enter as ONE composite whose mass, first
 * moment and inertia in base coordinates are constants of the plan (offChainMass).
 */
template <bool ALIGNED32, int MASS, int NC>
__global__ void __launch_bounds__(RCSB_FK_BLOCK(MASS), MASS ? RCSB_FKM_CHAIN_BLOCKS : 4)
fkChainKernel(const FkArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sPlan = smem;
  constexpr int THREADS = RCSB_FK_BLOCK(MASS);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  /* scratch: one private region per warp, row r of lane l at region[r * 33 + l] (fkKernel interleaves
   * the warps of a block in 129-double rows); `scr` is shifted so that the shared expressions
   * scr + row * scrStride + tid address it. A private region can be reused by its warp behind a warp
   * barrier: the results leave it as bulk stores (below). */
  constexpr int scrStride = 33;
  const int regionDoubles = a.scrDoubles;
  double* scrW = reinterpret_cast<double*>(sPlan + a.planBytes) + (tid >> 5) * regionDoubles;
  double* scr = scrW - (tid - lane);
  const long s = (long) blockIdx.x * THREADS + tid;
  const bool active = s < a.N;
  const long sc = active ? s : a.N - 1;

  /* every joint is a column (dof == nJ: column c is joint c): the joint values are requested before
   * the plan arrives */
  double qv[NC], sn[NC], cs[NC];
  {
    const double* qRow = a.q + sc * a.qdim;
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      qv[c] = qRow[c];
    }
  }
  stageBlobsTma(sPlan, a.plan, a.planBytes, nullptr, nullptr, 0);
  PlanView p;
  p.bind(sPlan);
  if (p.hdr->jacRecord > 0)
  {
    scr[p.hdr->zeroRow * scrStride + tid] = 0.0;
  }
  const int nSel = p.hdr->nSel;
  const int nChain = p.hdr->nChain;
  const int rs = nChain * scrStride;
  /* bodies per level: the bounds of the output loops, fetched once */
  int outStart[NC + 2];
#pragma unroll
  for (int L = 0; L < NC + 2; ++L)
  {
    outStart[L] = p.chainOutStart[L];
  }

  if (!sincosBatch<NC>(qv, sn, cs))
  {
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      sincos(qv[c], &sn[c], &cs[c]);
    }
  }

  Frame f;
  setIdentity(f);
  /* frames of the bodies of level L (and the points of the effectors attached to them) */
  auto outputs = [&](const int k0, const int k1) {
    int4 eNext = p.chainOut[k0 < k1 ? k0 : 0];
    for (int k = k0; k < k1; ++k)
    {
      const int4 e = eNext;
      eNext = p.chainOut[k + 1 < k1 ? k + 1 : k];   /* the next descriptor is in flight during this body */
      if (!(e.z & 1))
      {
        continue;
      }
      Frame t = f;
      applyRelF(t, p.chainX + 12 * k, e.z);
      if (e.y >= 0 && active && a.A_BI)
      {
        storeFrame<ALIGNED32>(a.A_BI + (s * nSel + e.y) * 12, t);
      }
      for (int i = p.bodyEffStart[e.x]; i < p.bodyEffStart[e.x + 1]; ++i)
      {
        const int ef = p.bodyEffList[i];
        const double* pt = a.effPtInline ? a.effPtArg + 3 * ef : p.effPt + 3 * ef;
        double* col = scr + (6 * nChain + 3 * ef) * scrStride + tid;
        col[0] = t.o[0] + (t.R[0] * pt[0] + t.R[3] * pt[1] + t.R[6] * pt[2]);
        col[scrStride] = t.o[1] + (t.R[1] * pt[0] + t.R[4] * pt[1] + t.R[7] * pt[2]);
        col[2 * scrStride] = t.o[2] + (t.R[2] * pt[0] + t.R[5] * pt[1] + t.R[8] * pt[2]);
      }
    }
  };

  outputs(outStart[0], outStart[1]);
#pragma unroll
  for (int c = 0; c < NC; ++c)
  {
    applyRel(f, p.chainG + 12 * c);
    rotateRows01(f, sn[c], cs[c]);
    const int slot = p.jntScr[c];
    if (slot >= 0)
    {
      /* the column's axis is row 2 of the shifted base */
      double* col = scr + slot * scrStride + tid;
      col[0] = f.R[6];
      col[rs] = f.R[7];
      col[2 * rs] = f.R[8];
      if (MASS == 2)
      {
        col[3 * rs] = f.o[1] * f.R[8] - f.o[2] * f.R[7];
        col[4 * rs] = f.o[2] * f.R[6] - f.o[0] * f.R[8];
        col[5 * rs] = f.o[0] * f.R[7] - f.o[1] * f.R[6];
      }
      else
      {
        col[3 * rs] = f.o[0];
        col[4 * rs] = f.o[1];
        col[5 * rs] = f.o[2];
      }
    }
    outputs(outStart[c + 1], outStart[c + 2]);
  }

  double Mv[MASS ? 36 : 1];
  if (MASS == 2 && a.M)
  {
    /* composite about the world origin: mass, first moment, inertia (xx, xy, xz, yy, yz, zz) */
    double cm = 0.0, ch[3] = {0.0, 0.0, 0.0}, cI[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = NC - 1; c >= 0; --c)
    {
      /* f: base of level c + 1 */
      const double* K = p.chainMass + 10 * (c + 1);
      const double mL = K[0];
      const double* R = f.R;
      const double H[3] = {R[0] * K[1] + R[3] * K[2] + R[6] * K[3], R[1] * K[1] + R[4] * K[2] + R[7] * K[3],
                           R[2] * K[1] + R[5] * K[2] + R[8] * K[3]};
      const double Ib[9] = {K[4], K[5], K[6], K[5], K[7], K[8], K[6], K[8], K[9]};
      double t[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int jj = 0; jj < 3; ++jj)
        {
          t[3 * i + jj] = Ib[3 * i] * R[jj] + Ib[3 * i + 1] * R[3 + jj] + Ib[3 * i + 2] * R[6 + jj];
        }
      const double* o = f.o;
      const double oo = mL * (o[0] * o[0] + o[1] * o[1] + o[2] * o[2]) + 2.0 * (o[0] * H[0] + o[1] * H[1] + o[2] * H[2]);
      cm += mL;
      ch[0] += mL * o[0] + H[0];
      ch[1] += mL * o[1] + H[1];
      ch[2] += mL * o[2] + H[2];
      cI[0] += (R[0] * t[0] + R[3] * t[3] + R[6] * t[6]) + (oo - (mL * o[0] * o[0] + 2.0 * o[0] * H[0]));
      cI[1] += (R[0] * t[1] + R[3] * t[4] + R[6] * t[7]) - (mL * o[0] * o[1] + (o[0] * H[1] + H[0] * o[1]));
      cI[2] += (R[0] * t[2] + R[3] * t[5] + R[6] * t[8]) - (mL * o[0] * o[2] + (o[0] * H[2] + H[0] * o[2]));
      cI[3] += (R[1] * t[1] + R[4] * t[4] + R[7] * t[7]) + (oo - (mL * o[1] * o[1] + 2.0 * o[1] * H[1]));
      cI[4] += (R[1] * t[2] + R[4] * t[5] + R[7] * t[8]) - (mL * o[1] * o[2] + (o[1] * H[2] + H[1] * o[2]));
      cI[5] += (R[2] * t[2] + R[5] * t[5] + R[8] * t[8]) + (oo - (mL * o[2] * o[2] + 2.0 * o[2] * H[2]));

      /* momentum of the composite under a unit rate of column c: p = m v + w x h, L = I w + h x v */
      const double w[3] = {R[6], R[7], R[8]};
      const double v[3] = {o[1] * w[2] - o[2] * w[1], o[2] * w[0] - o[0] * w[2], o[0] * w[1] - o[1] * w[0]};
      const double pk[3] = {cm * v[0] + (w[1] * ch[2] - w[2] * ch[1]), cm * v[1] + (w[2] * ch[0] - w[0] * ch[2]),
                            cm * v[2] + (w[0] * ch[1] - w[1] * ch[0])};
      const double Lk[3] = {cI[0] * w[0] + cI[1] * w[1] + cI[2] * w[2] + (ch[1] * v[2] - ch[2] * v[1]),
                            cI[1] * w[0] + cI[3] * w[1] + cI[4] * w[2] + (ch[2] * v[0] - ch[0] * v[2]),
                            cI[2] * w[0] + cI[4] * w[1] + cI[5] * w[2] + (ch[0] * v[1] - ch[1] * v[0])};
      Mv[c * (c + 1) / 2 + c] = (w[0] * Lk[0] + w[1] * Lk[1] + w[2] * Lk[2]) + (v[0] * pk[0] + v[1] * pk[1] + v[2] * pk[2]);
#pragma unroll
      for (int i = 0; i < c; ++i)
      {
        const double* ci = scr + i * scrStride + tid;
        Mv[c * (c + 1) / 2 + i] = (ci[0] * Lk[0] + ci[rs] * Lk[1] + ci[2 * rs] * Lk[2]) +
                                  (ci[3 * rs] * pk[0] + ci[4 * rs] * pk[1] + ci[5 * rs] * pk[2]);
      }

      if (c > 0)
      {
        /* back to the base of level c: undo the rotation about row 2, then G_c */
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
          const double ru = f.R[k], rv = f.R[3 + k];
          f.R[k] = cs[c] * ru - sn[c] * rv;
          f.R[3 + k] = sn[c] * ru + cs[c] * rv;
        }
        const double2* G2 = reinterpret_cast<const double2*>(p.chainG + 12 * c);
        const double2 g0 = G2[0], g1 = G2[1], g2 = G2[2], g3 = G2[3], g4 = G2[4], g5 = G2[5];
        const double ga[3] = {g0.x, g0.y, g1.x};
        const double A[9] = {g1.y, g2.x, g2.y, g3.x, g3.y, g4.x, g4.y, g5.x, g5.y};
        double n[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj)
          {
            n[3 * i + jj] = A[i] * f.R[jj] + A[3 + i] * f.R[3 + jj] + A[6 + i] * f.R[6 + jj];
          }
#pragma unroll
        for (int i = 0; i < 9; ++i)
        {
          f.R[i] = n[i];
        }
        f.o[0] -= f.R[0] * ga[0] + f.R[3] * ga[1] + f.R[6] * ga[2];
        f.o[1] -= f.R[1] * ga[0] + f.R[4] * ga[1] + f.R[7] * ga[2];
        f.o[2] -= f.R[2] * ga[0] + f.R[5] * ga[1] + f.R[8] * ga[2];
      }
    }
  }
  /* ---- results of a full warp with one effector: Jacobians and M are laid out in the warp's region
   * exactly as its 32 records lie in global memory ([state][element]) and leave as three bulk stores
   * (cp.async.bulk.global.shared::cta), issued by lane 0: 42 + NC^2 shared-memory stores per thread
   * instead of the transposing copy loops of fkTail (a quarter of the kernel's instructions) ---- */
  constexpr int REC = 3 * NC, NJ2 = NC * NC;
  const long w0 = s - lane;
  const bool bulk = a.JT && a.JR && a.N - w0 >= 32 && p.hdr->nEff == 1 && p.hdr->jacRecord == REC &&
                    p.hdr->inPlace != 0 && 64 * REC <= regionDoubles && 32 * NJ2 <= regionDoubles &&
                    ((reinterpret_cast<uintptr_t>(a.JT) | reinterpret_cast<uintptr_t>(a.JR) |
                      reinterpret_cast<uintptr_t>(a.M)) & 15u) == 0;
  if (!bulk)
  {
    fkTail<MASS>(a, p, scr, scrStride, nChain, tid, s, Mv);
    return;
  }
  double jt[REC], jr[REC];
  {
    const double* pe = scr + (6 * nChain) * scrStride + tid;
    const double px = pe[0], py = pe[scrStride], pz = pe[2 * scrStride];
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      const int tT = p.jtab[c];              /* row of the column's translation part, -1: not on the chain */
      if (tT >= 0)
      {
        const double* col = scr + (tT - 3 * nChain) * scrStride + tid;
        const double w[3] = {col[0], col[rs], col[2 * rs]};
        const double u[3] = {col[3 * rs], col[4 * rs], col[5 * rs]};
        if (MASS == 2)
        {
          /* (w, v): column = w x I_p + v */
          jt[c] = u[0] + (w[1] * pz - w[2] * py);
          jt[NC + c] = u[1] + (w[2] * px - w[0] * pz);
          jt[2 * NC + c] = u[2] + (w[0] * py - w[1] * px);
        }
        else
        {
          /* (axis, origin): column = axis x (I_p - origin) */
          const double d0 = px - u[0], d1 = py - u[1], d2 = pz - u[2];
          jt[c] = w[1] * d2 - w[2] * d1;
          jt[NC + c] = w[2] * d0 - w[0] * d2;
          jt[2 * NC + c] = w[0] * d1 - w[1] * d0;
        }
        jr[c] = w[0];
        jr[NC + c] = w[1];
        jr[2 * NC + c] = w[2];
      }
      else
      {
        jt[c] = jt[NC + c] = jt[2 * NC + c] = 0.0;
        jr[c] = jr[NC + c] = jr[2 * NC + c] = 0.0;
      }
    }
  }
  __syncwarp();   /* every lane has read its rows: the region is free */
  {
    double* stT = scrW + lane * REC;
    double* stR = scrW + 32 * REC + lane * REC;
#pragma unroll
    for (int e = 0; e < REC; ++e)
    {
      stT[e] = jt[e];
      stR[e] = jr[e];
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0)
  {
    const unsigned int srcT = (unsigned int) __cvta_generic_to_shared(scrW);
    const unsigned int srcR = (unsigned int) __cvta_generic_to_shared(scrW + 32 * REC);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.JT + w0 * REC), "r"(srcT),
                 "n"(32 * REC * 8)
                 : "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.JR + w0 * REC), "r"(srcR),
                 "n"(32 * REC * 8)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  if (MASS == 2 && a.M)
  {
    if (lane == 0)
    {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   /* the Jacobians have left the region */
    }
    __syncwarp();
    double* stM = scrW + lane * NJ2;
#pragma unroll
    for (int k = 0; k < NC; ++k)
    {
#pragma unroll
      for (int i = 0; i <= k; ++i)
      {
        stM[i * NC + k] = Mv[k * (k + 1) / 2 + i];
        stM[k * NC + i] = Mv[k * (k + 1) / 2 + i];
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
    {
      const unsigned int srcM = (unsigned int) __cvta_generic_to_shared(scrW);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.M + w0 * NJ2), "r"(srcM),
                   "n"(32 * NJ2 * 8)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0)
  {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   /* the region outlives the copies */
  }
}


/*
 * fkChainKernel's walk with velocities (RcsGraph_setState(q, q_dot) on a serial chain of rotations:
 * frames, x_dot, omega of the selected bodies; no Jacobians). The base carries its twist along: a
 * constant transform moves the origin inside the same rigid frame (v += omega x (o' - o)), the
 * column's rotation adds axis * q_dot to omega (the origin lies on the axis), and a body of the level
 * has x_dot = v + omega x (o_b - o), omega_b = omega (Rcs_graph.c:150-226 sums the same terms joint by
 * joint). Full warps stage x_dot [state][body][3] in their scratch region and hand it to the TMA engine
 * as one bulk store, then omega (one value per level, kept in registers) the same way; the last,
 * partial warp of a batch and unaligned outputs store directly.
 */
template <bool ALIGNED32, int NC>
__global__ void __launch_bounds__(RCSB_FK_THREADS, 4)
fkChainVelKernel(const FkArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sPlan = smem;
  constexpr int THREADS = RCSB_FK_THREADS;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  double* region = reinterpret_cast<double*>(sPlan + a.planBytes) + (tid >> 5) * a.scrDoubles;
  const long s = (long) blockIdx.x * THREADS + tid;
  const bool active = s < a.N;
  const long sc = active ? s : a.N - 1;
  const long w0 = s - lane;

  double qv[NC], qdv[NC], sn[NC], cs[NC];
  {
    const double* qRow = a.q + sc * a.qdim;
    const double* qdRow = a.qd + sc * a.qdim;
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      qv[c] = qRow[c];
      qdv[c] = qdRow[c];
    }
  }
  stageBlobsTma(sPlan, a.plan, a.planBytes, nullptr, nullptr, 0);
  PlanView p;
  p.bind(sPlan);
  const int nSel = p.hdr->nSel;
  const int rec = 3 * nSel;
  int outStart[NC + 2];
#pragma unroll
  for (int L = 0; L < NC + 2; ++L)
  {
    outStart[L] = p.chainOutStart[L];
  }
  if (!sincosBatch<NC>(qv, sn, cs))
  {
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
      sincos(qv[c], &sn[c], &cs[c]);
    }
  }
  const bool bulk = a.N - w0 >= 32 && 32 * rec <= a.scrDoubles &&
                    ((reinterpret_cast<uintptr_t>(a.xdot) | reinterpret_cast<uintptr_t>(a.omega)) & 15u) == 0;

  Frame f;
  setIdentity(f);
  double v[3] = {0.0, 0.0, 0.0}, om[3] = {0.0, 0.0, 0.0};
  double omL[NC + 1][3];
  auto outputs = [&](const int k0, const int k1) {
    int4 eNext = p.chainOut[k0 < k1 ? k0 : 0];
    for (int k = k0; k < k1; ++k)
    {
      const int4 e = eNext;
      eNext = p.chainOut[k + 1 < k1 ? k + 1 : k];
      if (e.y < 0)
      {
        continue;
      }
      Frame t = f;
      applyRelF(t, p.chainX + 12 * k, e.z);
      if (active && a.A_BI)
      {
        storeFrame<ALIGNED32>(a.A_BI + (s * nSel + e.y) * 12, t);
      }
      if (a.xdot)
      {
        const double d[3] = {t.o[0] - f.o[0], t.o[1] - f.o[1], t.o[2] - f.o[2]};
        const double x[3] = {v[0] + (om[1] * d[2] - om[2] * d[1]), v[1] + (om[2] * d[0] - om[0] * d[2]),
                             v[2] + (om[0] * d[1] - om[1] * d[0])};
        if (bulk)
        {
          double* st = region + lane * rec + 3 * e.y;
          st[0] = x[0];
          st[1] = x[1];
          st[2] = x[2];
        }
        else if (active)
        {
          double* dst = a.xdot + (s * nSel + e.y) * 3;
          dst[0] = x[0];
          dst[1] = x[1];
          dst[2] = x[2];
        }
      }
      if (a.omega && !bulk && active)
      {
        double* dst = a.omega + (s * nSel + e.y) * 3;
        dst[0] = om[0];
        dst[1] = om[1];
        dst[2] = om[2];
      }
    }
  };

  omL[0][0] = omL[0][1] = omL[0][2] = 0.0;
  outputs(outStart[0], outStart[1]);
#pragma unroll
  for (int c = 0; c < NC; ++c)
  {
    const double o0[3] = {f.o[0], f.o[1], f.o[2]};
    applyRel(f, p.chainG + 12 * c);
    {
      const double d[3] = {f.o[0] - o0[0], f.o[1] - o0[1], f.o[2] - o0[2]};
      v[0] += om[1] * d[2] - om[2] * d[1];
      v[1] += om[2] * d[0] - om[0] * d[2];
      v[2] += om[0] * d[1] - om[1] * d[0];
    }
    rotateRows01(f, sn[c], cs[c]);
    om[0] += f.R[6] * qdv[c];
    om[1] += f.R[7] * qdv[c];
    om[2] += f.R[8] * qdv[c];
    omL[c + 1][0] = om[0];
    omL[c + 1][1] = om[1];
    omL[c + 1][2] = om[2];
    outputs(outStart[c + 1], outStart[c + 2]);
  }

  if (!bulk)
  {
    return;
  }
  const unsigned int src = (unsigned int) __cvta_generic_to_shared(region);
  const unsigned int bytes = 32u * (unsigned int) rec * 8u;
  if (a.xdot)
  {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
    {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.xdot + w0 * rec), "r"(src),
                   "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (a.omega)
  {
    if (lane == 0)
    {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
#pragma unroll
    for (int L = 0; L < NC + 1; ++L)
    {
      for (int k = outStart[L]; k < outStart[L + 1]; ++k)
      {
        const int slot = p.chainOut[k].y;
        if (slot >= 0)
        {
          double* st = region + lane * rec + 3 * slot;
          st[0] = omL[L][0];
          st[1] = omL[L][1];
          st[2] = omL[L][2];
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
    {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.omega + w0 * rec), "r"(src),
                   "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0)
  {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <bool VEL, bool ALIGNED32, int MASS, bool PRUNED = false>
static cudaError_t launchSlots(const FkArgs& a, int nSlots, dim3 grid, size_t smemBytes,
                               cudaStream_t stream)
{
#define RCSB_LAUNCH(NS)                                                                        \
  {                                                                                            \
    auto k = fkKernel<VEL, ALIGNED32, NS, MASS, PRUNED>;                                       \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) smemBytes);                                     \
    if (e != cudaSuccess)                                                                      \
    {                                                                                          \
      return e;                                                                                \
    }                                                                                          \
    k<<<grid, RCSB_FK_BLOCK(MASS), smemBytes, stream>>>(a);                                        \
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

template <bool ALIGNED32, int MASS>
static cudaError_t launchFkChain(const FkArgs& a, int nScrRows, cudaStream_t stream)
{
  constexpr int threads = RCSB_FK_BLOCK(MASS);
  /* one scratch region per warp: nScrRows rows of 33 doubles, 16-byte multiple (FkArgs::scrDoubles) */
  FkArgs b = a;
  b.scrDoubles = (nScrRows * 33 + 1) & ~1;
  const size_t smemBytes = (size_t) a.planBytes + (size_t) (threads / 32) * b.scrDoubles * sizeof(double);
  const dim3 grid((unsigned int) ((a.N + threads - 1) / threads));
#define RCSB_LAUNCH(NC)                                                                        \
  case NC:                                                                                     \
  {                                                                                            \
    auto k = fkChainKernel<ALIGNED32, MASS, NC>;                                               \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) smemBytes);                                     \
    if (e != cudaSuccess)                                                                      \
    {                                                                                          \
      return e;                                                                                \
    }                                                                                          \
    k<<<grid, threads, smemBytes, stream>>>(b);                                                \
    return cudaGetLastError();                                                                 \
  }
  switch (a.chainNc)
  {
    RCSB_LAUNCH(1)
    RCSB_LAUNCH(2)
    RCSB_LAUNCH(3)
    RCSB_LAUNCH(4)
    RCSB_LAUNCH(5)
    RCSB_LAUNCH(6)
    RCSB_LAUNCH(7)
    RCSB_LAUNCH(8)
    default: return cudaErrorInvalidValue;
  }
#undef RCSB_LAUNCH
}

template <bool ALIGNED32>
static cudaError_t launchFkChainVel(const FkArgs& a, cudaStream_t stream)
{
  constexpr int threads = RCSB_FK_THREADS;
  /* one region per warp for the 32 x_dot / omega records of its states (FkArgs::velRecDoubles each) */
  FkArgs b = a;
  b.scrDoubles = (32 * a.velRecDoubles + 1) & ~1;
  const size_t smemBytes = (size_t) a.planBytes + (size_t) (threads / 32) * b.scrDoubles * sizeof(double);
  const dim3 grid((unsigned int) ((a.N + threads - 1) / threads));
#define RCSB_LAUNCH(NC)                                                                        \
  case NC:                                                                                     \
  {                                                                                            \
    auto k = fkChainVelKernel<ALIGNED32, NC>;                                                  \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) smemBytes);                                     \
    if (e != cudaSuccess)                                                                      \
    {                                                                                          \
      return e;                                                                                \
    }                                                                                          \
    k<<<grid, threads, smemBytes, stream>>>(b);                                                \
    return cudaGetLastError();                                                                 \
  }
  switch (a.chainNc)
  {
    RCSB_LAUNCH(1)
    RCSB_LAUNCH(2)
    RCSB_LAUNCH(3)
    RCSB_LAUNCH(4)
    RCSB_LAUNCH(5)
    RCSB_LAUNCH(6)
    RCSB_LAUNCH(7)
    RCSB_LAUNCH(8)
    default: return cudaErrorInvalidValue;
  }
#undef RCSB_LAUNCH
}

}   // namespace rcsb

extern "C" cudaError_t rcsb_launch_fk(const rcsb::FkArgs* a, int nSlots, int nScrRows,
                                      cudaStream_t stream)
{
  using namespace rcsb;
  if (a->N <= 0)
  {
    return cudaSuccess;
  }
  const bool velOut = a->qd != nullptr;
  const int threads = RCSB_FK_BLOCK(a->M != nullptr);
  const size_t smemBytes = (size_t) a->modelBytes + (size_t) a->planBytes +
                           (size_t) nScrRows * (threads + 1) * sizeof(double) +
                           (velOut ? (size_t) (threads / 32) * 2 * 32 * 33 * sizeof(double) : 0);
  const dim3 grid((unsigned int) ((a->N + threads - 1) / threads));
#ifdef RCSB_TUNING_KNOBS
  /* occupancy experiments: unused shared memory lowers the resident blocks per SM */
  static const size_t padBytes = getenv("RCSB_FK_PAD_SMEM") ? (size_t) atol(getenv("RCSB_FK_PAD_SMEM")) : 0;
  const_cast<size_t&>(smemBytes) += padBytes;
#endif
  const bool vel = (a->qd != nullptr);
  const bool al = ((reinterpret_cast<uintptr_t>(a->A_BI) | reinterpret_cast<uintptr_t>(a->A_JI)) &
                   31u) == 0;

  if (a->chainNc > 0 && a->qd)
  {
    /* the same walk with velocities: frames, x_dot, omega (no Jacobians) */
    /* (large body selections would not leave room for four blocks' record regions: fkKernel<VEL>) */
    if ((size_t) a->planBytes + (size_t) (RCSB_FK_THREADS / 32) * ((32 * a->velRecDoubles + 1) & ~1) * 8 <= (size_t) 56 * 1024)
    {
      return al ? launchFkChainVel<true>(*a, stream) : launchFkChainVel<false>(*a, stream);
    }
  }
  else if (a->chainNc > 0)
  {
    /* serial chain of rotations: straight-line walk, nothing of the model interpreted */
    if (a->M)
    {
      return al ? launchFkChain<true, 2>(*a, nScrRows, stream) : launchFkChain<false, 2>(*a, nScrRows, stream);
    }
    return al ? launchFkChain<true, 0>(*a, nScrRows, stream) : launchFkChain<false, 0>(*a, nScrRows, stream);
  }
  if (a->M && a->massMode == 2)
  {
    return al ? launchSlots<false, true, 2>(*a, nSlots, grid, smemBytes, stream)
              : launchSlots<false, false, 2>(*a, nSlots, grid, smemBytes, stream);
  }
  if (a->M)
  {
    return al ? launchSlots<false, true, 1>(*a, nSlots, grid, smemBytes, stream)
              : launchSlots<false, false, 1>(*a, nSlots, grid, smemBytes, stream);
  }
  if (vel)
  {
    return al ? launchSlots<true, true, 0>(*a, nSlots, grid, smemBytes, stream)
              : launchSlots<true, false, 0>(*a, nSlots, grid, smemBytes, stream);
  }
  if (a->pruned)
  {
    return al ? launchSlots<false, true, 0, true>(*a, nSlots, grid, smemBytes, stream)
              : launchSlots<false, false, 0, true>(*a, nSlots, grid, smemBytes, stream);
  }
  return al ? launchSlots<false, true, 0>(*a, nSlots, grid, smemBytes, stream)
            : launchSlots<false, false, 0>(*a, nSlots, grid, smemBytes, stream);
}
