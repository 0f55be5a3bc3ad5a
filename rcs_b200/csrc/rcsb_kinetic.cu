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
 * The reference builds a 6 x nJ Jacobian and a 3 nJ x nJ Hessian per body (O(nb nJ^3)).
 * Here one thread per state walks the tree once, all quantities in world coordinates:
 *   pass 1  forward kinematics + velocity / bias-acceleration recursion; per link the
 *           composite inertia (m, m c, inertia about the world origin) and the bias wrench
 *           (m a_c, I wd + w x I w) are accumulated; joint axes / origins are kept per
 *           Jacobian column;
 *   pass 2  links are summed into their parent links (reverse depth-first order);
 *   pass 3  per column: momentum (p, L) of its composite under a unit joint rate (L also with
 *           the transposed inertia: input tensors need not be symmetric), h and
 *           F_gravity by projecting the composite wrench / weight on the joint axis;
 *   pass 4  M row by row: M[i][k] = <axis of the ancestor, (p, L) of the descendant>.
 * Jdot qd equals the bias acceleration (qdd = 0) of the rigid-body recursion, so no Hessian
 * is needed. Per-state intermediates live in a per-thread workspace in global memory that is
 * sized to stay resident in the 126 MB L2 (persistent grid, element k of thread t at
 * scratch[k * T + t], i.e. coalesced); results are staged through shared memory so that the
 * rows of 32 consecutive states leave as coalesced 8-byte stores.
 *
 * Bound: HBM for the M stream at the roofline; see DESIGN.md.
 */
#include "rcsb_device.cuh"
#include "rcsb_kernels.h"
#include "rcsb_plan.h"

namespace rcsb
{

struct KtPlanView
{
  const RcsbKtPlanHeader* hdr;
  const int* bodyLink;
  const int* linkParent;
  const int* colJoint;
  const int* colLink;
  const int* colIsRot;
  const unsigned char* rel;

  __device__ __forceinline__ void bind(const char* blob)
  {
    hdr = reinterpret_cast<const RcsbKtPlanHeader*>(blob);
    bodyLink = reinterpret_cast<const int*>(blob + hdr->offBodyLink);
    linkParent = reinterpret_cast<const int*>(blob + hdr->offLinkParent);
    colJoint = reinterpret_cast<const int*>(blob + hdr->offColJoint);
    colLink = reinterpret_cast<const int*>(blob + hdr->offColLink);
    colIsRot = reinterpret_cast<const int*>(blob + hdr->offColIsRot);
    rel = reinterpret_cast<const unsigned char*>(blob + hdr->offRel);
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

/* <motion axis of a joint, momentum (p, L)>: rot: a . (L - o x p), trans: a . p */
__device__ __forceinline__ double pairing(bool isRot, const double a[3], const double o[3],
                                          const double p[3], const double L[3])
{
  if (isRot)
  {
    double t[3];
    cross3(t, o, p);
    return a[0] * (L[0] - t[0]) + a[1] * (L[1] - t[1]) + a[2] * (L[2] - t[2]);
  }
  return a[0] * p[0] + a[1] * p[1] + a[2] * p[2];
}

/* Rows of 32 consecutive states staged in `tile` (tile[sl * tileStride + k]) -> global,
 * dst element (state, k) at dst[state * perState + k]; consecutive lanes write consecutive
 * doubles of a row. */
__device__ __forceinline__ void flushRows(const double* tile, int tileStride, double* dst,
                                          long perState, int cnt, float invCnt, int nv, int lane)
{
  const int total = nv * cnt;
  for (int e = lane; e < total; e += 32)
  {
    const int sl = __float2int_rd(((float) e + 0.5f) * invCnt);
    const int k = e - sl * cnt;
    dst[(long) sl * perState + k] = tile[sl * tileStride + k];
  }
}

template <bool VEL, int NSLOTS>
__global__ void __launch_bounds__(RCSB_KT_THREADS)
ktKernel(const KtArgs a)
{
  extern __shared__ __align__(16) char smem[];
  char* sModel = smem;
  char* sPlan = smem + a.modelBytes;
  double* tiles = reinterpret_cast<double*>(sPlan + a.planBytes);

  stageBlob(sModel, a.model, a.modelBytes);
  stageBlob(sPlan, a.plan, a.planBytes);
  __syncthreads();

  ModelView m;
  m.bind(sModel);
  KtPlanView p;
  p.bind(sPlan);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nb = m.hdr->nb;
  const int dof = m.hdr->dof;
  const int nJ = p.hdr->nJ;
  const int nLinks = p.hdr->nLinks;
  const int tileStride = nJ | 1;
  double* tile = tiles + (tid >> 5) * 32 * tileStride;
  const float invNJ = p.hdr->invNJ;

  const long T = (long) gridDim.x * RCSB_KT_THREADS;
  const long gtid = (long) blockIdx.x * RCSB_KT_THREADS + tid;
  double* scr = a.scratch + gtid;
  const long ss = a.scratchStride;
  double* sComp = scr + (long) p.hdr->scrComp * ss;
  double* sWr = scr + (long) p.hdr->scrWr * ss;
  double* sJa = scr + (long) p.hdr->scrJa * ss;
  double* sPl = scr + (long) p.hdr->scrPl * ss;
  double* sQd = scr + (long) p.hdr->scrQd * ss;

  constexpr int SLOT_DOUBLES = VEL ? 24 : 12;
  double stk[(NSLOTS > 0 ? NSLOTS : 1) * SLOT_DOUBLES];
  const double grav = 9.81;

  for (long s0 = gtid - lane; s0 < a.N; s0 += T)
  {
    const long s = s0 + lane;
    const bool active = s < a.N;
    const long sc = active ? s : a.N - 1;
    long nvl = a.N - s0;
    const int nv = nvl > 32 ? 32 : (int) nvl;

    StateReader rd;
    rd.full = (a.qdim == dof);
    rd.q = a.q + sc * a.qdim;
    rd.qd = VEL ? a.qd + sc * a.qdim : nullptr;

    for (int i = 0; i < 19 * nLinks; ++i)
    {
      sComp[(long) i * ss] = 0.0;   /* comp and wr are contiguous */
    }

    /* ---- pass 1: walk ------------------------------------------------------------------ */
    Kin k, keep;
    setIdentity(k.f);
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
      k.wik[i] = k.wfl[i] = k.wd[i] = k.ao[i] = 0.0;
    }
    keep = k;
    double V = 0.0;

    for (int b = 0; b < nb; ++b)
    {
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
          applyRelD(k.f, m.jntAJP + 12 * j, d);
          if (VEL)
          {
            addOffsetAccel(k, d);
          }
        }

        const double qj = rd.jointQ(m, j);
        const int dir = m.jntType[j] % 3;
        const int col = m.jntJacobi[j];
        double qdFull = 0.0, qdIk = 0.0;
        if (VEL)
        {
          qdFull = rd.jointQd(m, j);
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
          double* ja = sJa + (long) (6 * col) * ss;
          ja[0] = ax[0];
          ja[ss] = ax[1];
          ja[2 * ss] = ax[2];
          ja[3 * ss] = k.f.o[0];
          ja[4 * ss] = k.f.o[1];
          ja[5 * ss] = k.f.o[2];
          sQd[(long) col * ss] = qdIk;
        }
      }

      if (bflags & RCSB_BF_HAS_ABP)
      {
        double d[3];
        applyRelD(k.f, m.bodyABP + 12 * b, d);
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

        const int link = p.bodyLink[b];
        if (link >= 0)
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
          double* cp = sComp + (long) (13 * link) * ss;
          const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
          cp[0] += mass;
          cp[ss] += mass * c[0];
          cp[2 * ss] += mass * c[1];
          cp[3 * ss] += mass * c[2];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
            {
              cp[(long) (4 + 3 * i + jj) * ss] +=
                Iw[3 * i + jj] + mass * ((i == jj ? cc : 0.0) - c[i] * c[jj]);
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
            double* wp = sWr + (long) (6 * link) * ss;
            wp[0] += F[0];
            wp[ss] += F[1];
            wp[2 * ss] += F[2];
            wp[3 * ss] += Iwd[0] + gy[0] + cf[0];
            wp[4 * ss] += Iwd[1] + gy[1] + cf[1];
            wp[5 * ss] += Iwd[2] + gy[2] + cf[2];
          }
        }
      }

      if (bflags & RCSB_BF_LEAF)
      {
        k = keep;
      }
    }

    /* ---- pass 2: composites ---------------------------------------------------------- */
    for (int l = nLinks - 1; l > 0; --l)
    {
      const int pl = p.linkParent[l];
      if (pl >= 0)
      {
        double* src = sComp + (long) (13 * l) * ss;
        double* dst = sComp + (long) (13 * pl) * ss;
#pragma unroll
        for (int i = 0; i < 13; ++i)
        {
          dst[(long) i * ss] += src[(long) i * ss];
        }
        if (VEL)
        {
          double* ws = sWr + (long) (6 * l) * ss;
          double* wdst = sWr + (long) (6 * pl) * ss;
#pragma unroll
          for (int i = 0; i < 6; ++i)
          {
            wdst[(long) i * ss] += ws[(long) i * ss];
          }
        }
      }
    }

    /* ---- pass 3: per column momentum, h, F_gravity ------------------------------------ */
    for (int ci = 0; ci < nJ; ++ci)
    {
      const int li = p.colLink[ci];
      const bool isRot = p.colIsRot[ci] != 0;
      const double* ja = sJa + (long) (6 * ci) * ss;
      const double* cp = sComp + (long) (13 * li) * ss;
      const double ax[3] = {ja[0], ja[ss], ja[2 * ss]};
      const double o[3] = {ja[3 * ss], ja[4 * ss], ja[5 * ss]};
      const double mC = cp[0];
      const double hC[3] = {cp[ss], cp[2 * ss], cp[3 * ss]};
      double pv[3], Lv[3], Lt[3];
      if (isRot)
      {
        double Ic[9];
#pragma unroll
        for (int i = 0; i < 9; ++i)
        {
          Ic[i] = cp[(long) (4 + i) * ss];
        }
        const double hr[3] = {hC[0] - mC * o[0], hC[1] - mC * o[1], hC[2] - mC * o[2]};
        double axo[3], t[3];
        cross3(pv, ax, hr);            /* p = a x (h - m o) */
        cross3(axo, ax, o);
        cross3(t, hC, axo);            /* h x (a x o) */
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
          Lv[i] = Ic[3 * i] * ax[0] + Ic[3 * i + 1] * ax[1] + Ic[3 * i + 2] * ax[2] - t[i];
          Lt[i] = Ic[i] * ax[0] + Ic[3 + i] * ax[1] + Ic[6 + i] * ax[2] - t[i];
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
      double* pl = sPl + (long) (9 * ci) * ss;
      pl[0] = pv[0];
      pl[ss] = pv[1];
      pl[2 * ss] = pv[2];
      pl[3 * ss] = Lv[0];
      pl[4 * ss] = Lv[1];
      pl[5 * ss] = Lv[2];
      pl[6 * ss] = Lt[0];
      pl[7 * ss] = Lt[1];
      pl[8 * ss] = Lt[2];

      double hval = 0.0;
      if (VEL)
      {
        const double* wp = sWr + (long) (6 * li) * ss;
        const double F[3] = {wp[0], wp[ss], wp[2 * ss]};
        const double Nn[3] = {wp[3 * ss], wp[4 * ss], wp[5 * ss]};
        hval = -pairing(isRot, ax, o, F, Nn);
      }
      tile[lane * tileStride + ci] = hval;
    }
    __syncwarp();
    if (a.h)
    {
      flushRows(tile, tileStride, a.h + s0 * nJ, nJ, nJ, invNJ, nv, lane);
    }
    __syncwarp();

    if (a.Fg)
    {
      for (int ci = 0; ci < nJ; ++ci)
      {
        const int li = p.colLink[ci];
        const double* ja = sJa + (long) (6 * ci) * ss;
        const double* cp = sComp + (long) (13 * li) * ss;
        const double ax[3] = {ja[0], ja[ss], ja[2 * ss]};
        const double mC = cp[0];
        double g;
        if (p.colIsRot[ci])
        {
          /* a . ((h - m o) x (0, 0, -g)) */
          const double hr0 = cp[ss] - mC * ja[3 * ss];
          const double hr1 = cp[2 * ss] - mC * ja[4 * ss];
          g = ax[0] * (hr1 * -grav) + ax[1] * (hr0 * grav);
        }
        else
        {
          g = ax[2] * (mC * -grav);
        }
        tile[lane * tileStride + ci] = g;
      }
      __syncwarp();
      flushRows(tile, tileStride, a.Fg + s0 * nJ, nJ, nJ, invNJ, nv, lane);
      __syncwarp();
    }

    /* ---- pass 4: mass matrix rows + kinetic energy -------------------------------------- */
    double Tkin = 0.0;
    if (a.M || (a.E && VEL))
    {
      for (int ri = 0; ri < nJ; ++ri)
      {
        const bool rowRot = p.colIsRot[ri] != 0;
        const double* ja = sJa + (long) (6 * ri) * ss;
        const double* pl = sPl + (long) (9 * ri) * ss;
        const double ai[3] = {ja[0], ja[ss], ja[2 * ss]};
        const double oi[3] = {ja[3 * ss], ja[4 * ss], ja[5 * ss]};
        const double pi[3] = {pl[0], pl[ss], pl[2 * ss]};
        /* with the transposed inertia: M[row][col] = a_row^T I a_col = a_col . (I^T a_row) */
        const double Li[3] = {pl[6 * ss], pl[7 * ss], pl[8 * ss]};
        const unsigned char* rel = p.rel + ri * nJ;
        double rowDot = 0.0;

        for (int ck = 0; ck < nJ; ++ck)
        {
          const int r = rel[ck];
          double val = 0.0;
          if (r == 1)
          {
            /* column joint is moved by the row joint: row axis against column momentum */
            const double* pk = sPl + (long) (9 * ck) * ss;
            const double pp[3] = {pk[0], pk[ss], pk[2 * ss]};
            const double LL[3] = {pk[3 * ss], pk[4 * ss], pk[5 * ss]};
            val = pairing(rowRot, ai, oi, pp, LL);
          }
          else if (r == 2)
          {
            const double* jk = sJa + (long) (6 * ck) * ss;
            const double ak[3] = {jk[0], jk[ss], jk[2 * ss]};
            const double ok[3] = {jk[3 * ss], jk[4 * ss], jk[5 * ss]};
            val = pairing(p.colIsRot[ck] != 0, ak, ok, pi, Li);
          }
          if (VEL && r != 0)
          {
            rowDot += val * sQd[(long) ck * ss];
          }
          tile[lane * tileStride + ck] = val;
        }
        if (VEL)
        {
          Tkin += 0.5 * sQd[(long) ri * ss] * rowDot;
        }
        __syncwarp();
        if (a.M)
        {
          flushRows(tile, tileStride, a.M + (s0 * nJ + ri) * nJ, (long) nJ * nJ, nJ, invNJ, nv,
                    lane);
        }
        __syncwarp();
      }
    }

    if (a.E && active)
    {
      a.E[s] = V + Tkin;
    }
  }
}

template <bool VEL>
static cudaError_t launchKt(const KtArgs& a, int nSlots, int blocks, size_t smemBytes,
                            cudaStream_t stream)
{
#define RCSB_LAUNCH(NS)                                                                        \
  {                                                                                            \
    auto k = ktKernel<VEL, NS>;                                                                \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int) smemBytes);                                     \
    if (e != cudaSuccess)                                                                      \
    {                                                                                          \
      return e;                                                                                \
    }                                                                                          \
    k<<<blocks, RCSB_KT_THREADS, smemBytes, stream>>>(a);                                      \
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

}   // namespace rcsb

extern "C" cudaError_t rcsb_launch_kt(const rcsb::KtArgs* a, int nSlots, int nJ, int blocks,
                                      cudaStream_t stream)
{
  using namespace rcsb;
  if (a->N <= 0)
  {
    return cudaSuccess;
  }
  const size_t smemBytes = (size_t) a->modelBytes + (size_t) a->planBytes +
                           (size_t) (RCSB_KT_THREADS / 32) * 32 * (nJ | 1) * sizeof(double);
  return a->qd ? launchKt<true>(*a, nSlots, blocks, smemBytes, stream)
               : launchKt<false>(*a, nSlots, blocks, smemBytes, stream);
}
