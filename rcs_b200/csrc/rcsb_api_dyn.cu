/*
 * C ABI of rcs_b200 (include/rcsb.h), part 2: kinetic terms and resolved-motion-rate IK —
 * plan construction (host) and launches. No arithmetic on states happens here.
 */
#include "rcsb_priv.h"

using namespace rcsbhost;

/* ---- kinetic-terms plan ------------------------------------------------------------------- */

int rcsbhost::buildKtPlan(RcsbModel* m, const WbSpec* wb, bool vel, DevicePlan** out)
{
  std::lock_guard<std::recursive_mutex> lock(m->mtx);
  const int LREC = RCSB_KT_LINK_REC_N(vel), CREC = RCSB_KT_COL_REC_N(vel);
  if (!wb && m->ktPlanBuilt[vel ? 1 : 0])
  {
    *out = &m->ktPlan[vel ? 1 : 0];
    return RCSB_OK;
  }
  if (wb)
  {
    auto it = m->wbPlans.find(wb->key);
    if (it != m->wbPlans.end())
    {
      *out = &it->second;
      return RCSB_OK;
    }
  }

  const int nb = m->hdr.nb, dof = m->hdr.dof, nJ = m->hdr.nJ;
  const int* bodyParent = m->iarr(RCSB_A_BODY_PARENT);
  const int* bodyJntStart = m->iarr(RCSB_A_BODY_JNT_START);
  const int* bodyJntCount = m->iarr(RCSB_A_BODY_JNT_COUNT);
  const int* jntJacobi = m->iarr(RCSB_A_JNT_JACOBI);
  const int* jntPrev = m->iarr(RCSB_A_JNT_PREV);
  const int* jntFlags = m->iarr(RCSB_A_JNT_FLAGS);
  const int* jntBody = m->iarr(RCSB_A_JNT_BODY);

  /* links: bodies owning at least one unconstrained joint, numbered depth first */
  std::vector<int> bodyLink(nb, -1), linkOfBody(nb, -1), linkParent;
  int nLinks = 0;
  for (int b = 0; b < nb; ++b)
  {
    bool owns = false;
    for (int j = bodyJntStart[b]; j < bodyJntStart[b] + bodyJntCount[b]; ++j)
    {
      owns = owns || (jntJacobi[j] >= 0);
    }
    const int inherited = bodyParent[b] >= 0 ? bodyLink[bodyParent[b]] : -1;
    if (owns)
    {
      linkOfBody[b] = nLinks;
      bodyLink[b] = nLinks++;
      linkParent.push_back(inherited);
    }
    else
    {
      bodyLink[b] = inherited;
    }
  }

  std::vector<int> colJoint(nJ > 0 ? nJ : 1, 0), colLink(nJ > 0 ? nJ : 1, 0),
    colIsRot(nJ > 0 ? nJ : 1, 0);
  for (int j = 0; j < dof; ++j)
  {
    const int c = jntJacobi[j];
    if (c >= 0)
    {
      colJoint[c] = j;
      colLink[c] = linkOfBody[jntBody[j]];
      colIsRot[c] = (jntFlags[j] & RCSB_JF_ROT) ? 1 : 0;
    }
  }

  /* relation table from the backward joint chains (jnt->prev links) */
  std::vector<unsigned char> rel((size_t) (nJ > 0 ? nJ * nJ : 1), 0);
  for (int k = 0; k < nJ; ++k)
  {
    rel[(size_t) k * nJ + k] = 1;
    for (int j = jntPrev[colJoint[k]]; j >= 0; j = jntPrev[j])
    {
      const int i = jntJacobi[j];
      if (i >= 0)
      {
        rel[(size_t) i * nJ + k] = 1;   /* column k is moved by row joint i */
        rel[(size_t) k * nJ + i] = 2;   /* column i is a proper ancestor of row joint k */
      }
    }
  }

  /* whole-body plans: output slots of the selected frames, task bodies whose frames go into the
   * record, and which columns move them */
  std::vector<int> bodyOut((size_t) nb, -1), bodyFrameRec((size_t) nb, 0), taskBodyRec;
  std::vector<unsigned char> taskColOn;
  int nSelOut = 0;
  if (wb)
  {
    const int* bodyLastJnt = m->iarr(RCSB_A_BODY_LASTJNT);
    if (wb->nSel == 0)
    {
      for (int b = 0; b < nb; ++b)
      {
        bodyOut[(size_t) b] = b;
      }
      nSelOut = nb;
    }
    else
    {
      for (int i = 0; i < wb->nSel; ++i)
      {
        const int b = wb->bodySel[i];
        if (b < 0 || b >= nb || bodyOut[(size_t) b] >= 0)
        {
          rcsb_set_error("rcsb_wholebody: bodySel[%d] = %d is out of range or selected twice", i, b);
          return RCSB_EINVAL;
        }
        bodyOut[(size_t) b] = i;
      }
      nSelOut = wb->nSel;
    }
    taskBodyRec.assign(wb->taskBodies.size(), 0);
    taskColOn.assign(wb->taskBodies.size() * (size_t) (nJ > 0 ? nJ : 1), 0);
    for (size_t u = 0; u < wb->taskBodies.size(); ++u)
    {
      const int b = wb->taskBodies[u];
      bodyFrameRec[(size_t) b] = 1;
      for (int j = bodyLastJnt[b]; j >= 0; j = jntPrev[j])
      {
        if (jntJacobi[j] >= 0)
        {
          taskColOn[u * (size_t) nJ + (size_t) jntJacobi[j]] = 1;
        }
      }
    }
  }

  /* walk bookkeeping: the accumulator is emitted as a run whenever the walk moves to a body of
   * another link; every record of the state gets its offset in the order of production */
  std::vector<int> bodyEmit((size_t) nb + 1, 0), runLink, runRec;
  std::vector<int> colRec(nJ > 0 ? nJ : 1, 0);
  int rec = 0;
  {
    int cur = -1;
    for (int b = 0; b < nb; ++b)
    {
      if (bodyLink[b] != cur)
      {
        if (cur >= 0)   /* every visit of a link leaves a run, so that each link has one */
        {
          bodyEmit[b] = 1;
          runLink.push_back(cur);
          runRec.push_back(rec);
          rec += LREC;
        }
        cur = bodyLink[b];
      }
      for (int j = bodyJntStart[b]; j < bodyJntStart[b] + bodyJntCount[b]; ++j)
      {
        if (jntJacobi[j] >= 0)
        {
          colRec[jntJacobi[j]] = rec;
          rec += CREC;
        }
      }
      if (bodyFrameRec[(size_t) b])
      {
        for (size_t u = 0; u < taskBodyRec.size(); ++u)
        {
          if (wb->taskBodies[u] == b)
          {
            taskBodyRec[u] = rec;
          }
        }
        rec += 12;   /* the body's frame, after its joints and A_BP */
      }
    }
    if (cur >= 0)
    {
      bodyEmit[nb] = 1;
      runLink.push_back(cur);
      runRec.push_back(rec);
      rec += LREC;
    }
  }
  const int nRuns = (int) runLink.size();

  if (nJ > 255 || rec + 5 > 65535)
  {
    rcsb_set_error("rcsb_kinetic_terms: model too large (nJ %d, record %d doubles)", nJ, rec + 5);
    return RCSB_EUNSUPPORTED;
  }

  /* shared-memory layout of the assembly kernel */
  const int recDoubles = (rec + 5 + 3) & ~3;
  std::vector<int> linkHome((size_t) (nLinks > 0 ? nLinks : 1), -1), extraRuns;
  for (int r = 0; r < nRuns; ++r)
  {
    if (linkHome[runLink[r]] < 0)
    {
      linkHome[runLink[r]] = runRec[r];
    }
    else
    {
      extraRuns.push_back(runRec[r]);
      extraRuns.push_back(linkHome[runLink[r]]);
    }
  }
  std::vector<int> colHome((size_t) (nJ > 0 ? nJ : 1), 0);
  for (int k = 0; k < nJ; ++k)
  {
    colHome[k] = linkHome[colLink[k]];
  }

  bool symmetric = true;
  {
    const double* I = m->darr(RCSB_A_BODY_INERTIA);
    for (int b = 0; b < nb; ++b)
    {
      symmetric = symmetric && I[9 * b + 1] == I[9 * b + 3] && I[9 * b + 2] == I[9 * b + 6] &&
                  I[9 * b + 5] == I[9 * b + 7];
    }
  }
  /* structurally non-zero entries of the upper triangle: (ancestor column i, column k) */
  std::vector<int> pairs;
  const int fStride = RCSB_KT_F_STRIDE(symmetric ? 1 : 0);
  for (int k = 0; k < nJ; ++k)
  {
    for (int i = 0; i <= k; ++i)
    {
      if (rel[(size_t) i * nJ + k] == 1)
      {
        pairs.push_back(RCSB_KT_S_STRIDE * i);
        pairs.push_back(fStride * k);
        pairs.push_back((i * nJ + k) | ((k * nJ + i) << 16));
        pairs.push_back((i == k ? 2 : 0) | ((RCSB_KT_S_STRIDE * k) << 8));
      }
    }
  }
  /* composites: links in reverse depth-first order, each added to its parent (see offLinkSteps) */
  std::vector<int> linkSteps;
  for (int l = nLinks - 1; l >= 0; --l)
  {
    const int pl = linkParent[(size_t) l];
    linkSteps.push_back(linkHome[(size_t) l]);
    linkSteps.push_back(pl < 0 ? -2 : (pl == l - 1 ? -1 : linkHome[(size_t) pl]));
  }
  /* links without a parent link: their composites add up to the whole mechanism (centre of gravity) */
  std::vector<int> rootLinks;
  for (int l = 0; l < nLinks; ++l)
  {
    if (linkParent[(size_t) l] < 0)
    {
      rootLinks.push_back(linkHome[(size_t) l]);
    }
  }
  /* whole-body plans: the (task Jacobian, column) pairs that are not structurally zero */
  std::vector<int> jacEntries;
  if (wb)
  {
    const int nJacT = (int) wb->jacTasks.size() / 8;
    for (int t = 0; t < nJacT; ++t)
    {
      const int* tk = wb->jacTasks.data() + 8 * t;
      const int kind = tk[0], mode = tk[1], u2 = tk[2], u1 = tk[3], u3 = tk[4];
      for (int k = 0; k < nJ; ++k)
      {
        auto on = [&](int u) { return u >= 0 && taskColOn[(size_t) u * nJ + (size_t) k] != 0; };
        const bool rot = colIsRot[k] != 0;
        int o2 = 0, o1 = 0, o3 = 0;
        if (kind == RCSB_JAC_POS)
        {
          o2 = on(u2) ? 1 : 0;
          o1 = (mode == 1 && on(u1)) ? 1 : 0;
          o3 = (mode == 1 && u3 >= 0 && rot && on(u3)) ? 1 : 0;
        }
        else
        {
          o2 = (rot && on(u2)) ? 1 : 0;
          o1 = (mode == 1 && u1 >= 0 && rot && on(u1)) ? 1 : 0;
        }
        if (o2 || o1 || o3)
        {
          jacEntries.push_back(k | (t << 8) | (o2 << 16) | (o1 << 17) | (o3 << 18));
        }
      }
    }
    if (nJ > 255 || nJacT > 255)
    {
      rcsb_set_error("rcsb_wholebody: too many columns / task Jacobians (%d, %d; limit 255)", nJ, nJacT);
      return RCSB_EUNSUPPORTED;
    }
  }

  RcsbKtPlanHeader ph;
  memset(&ph, 0, sizeof(ph));
  ph.nLinks = nLinks;
  ph.nJ = nJ;
  ph.nRuns = nRuns;
  ph.nPairs = (int) pairs.size() / 4;
  ph.symmetric = symmetric ? 1 : 0;
  ph.vel = vel ? 1 : 0;
  ph.nRoots = (int) rootLinks.size();
  ph.nJacEntries = (int) jacEntries.size();
  ph.recV = rec;
  ph.recDoubles = recDoubles;
  ph.nExtraRuns = (int) extraRuns.size() / 2;
  int off = align16((int) sizeof(ph));
  auto place = [&](int& field, size_t bytes) {
    field = off;
    off = align16(off + (int) (bytes > 0 ? bytes : 4));
  };
  place(ph.offBodyLink, (size_t) nb * 4);
  place(ph.offBodyEmit, ((size_t) nb + 1) * 4);
  if (wb)
  {
    ph.wholeBody = 1;
    ph.nSel = nSelOut;
    ph.nTaskBodies = (int) wb->taskBodies.size();
    ph.nJac = (int) wb->jacTasks.size() / 8;
    place(ph.offBodyOut, (size_t) nb * 4);
    place(ph.offBodyFrameRec, (size_t) nb * 4);
  }
  ph.walkBytes = off;
  if (wb)
  {
    place(ph.offTaskBodyRec, taskBodyRec.size() * 4);
    place(ph.offJacTasks, wb->jacTasks.size() * 4);
    place(ph.offTaskColOn, taskColOn.size());
  }
  place(ph.offLinkParent, (size_t) nLinks * 4);
  place(ph.offColJoint, colJoint.size() * 4);
  place(ph.offColLink, colLink.size() * 4);
  place(ph.offColIsRot, colIsRot.size() * 4);
  place(ph.offColRec, colRec.size() * 4);
  place(ph.offRunLink, (size_t) nRuns * 4);
  place(ph.offRunRec, (size_t) nRuns * 4);
  place(ph.offPairs, pairs.size() * 4);
  place(ph.offLinkHome, linkHome.size() * 4);
  place(ph.offExtraRuns, extraRuns.size() * 4);
  place(ph.offColHome, colHome.size() * 4);
  place(ph.offLinkSteps, linkSteps.size() * 4);
  place(ph.offRootLinks, rootLinks.size() * 4);
  place(ph.offJacEntries, jacEntries.size() * 4);
  ph.totalBytes = off;

  std::vector<char> blob((size_t) ph.totalBytes, 0);
  auto put = [&](int offset, const void* src, size_t bytes) {
    if (bytes > 0)
    {
      memcpy(blob.data() + offset, src, bytes);
    }
  };
  memcpy(blob.data(), &ph, sizeof(ph));
  put(ph.offBodyLink, bodyLink.data(), (size_t) nb * 4);
  put(ph.offBodyEmit, bodyEmit.data(), ((size_t) nb + 1) * 4);
  put(ph.offLinkParent, linkParent.data(), (size_t) nLinks * 4);
  put(ph.offColJoint, colJoint.data(), colJoint.size() * 4);
  put(ph.offColLink, colLink.data(), colLink.size() * 4);
  put(ph.offColIsRot, colIsRot.data(), colIsRot.size() * 4);
  put(ph.offColRec, colRec.data(), colRec.size() * 4);
  put(ph.offRunLink, runLink.data(), (size_t) nRuns * 4);
  put(ph.offRunRec, runRec.data(), (size_t) nRuns * 4);
  put(ph.offPairs, pairs.data(), pairs.size() * 4);
  put(ph.offLinkHome, linkHome.data(), linkHome.size() * 4);
  put(ph.offExtraRuns, extraRuns.data(), extraRuns.size() * 4);
  put(ph.offColHome, colHome.data(), colHome.size() * 4);
  put(ph.offLinkSteps, linkSteps.data(), linkSteps.size() * 4);
  put(ph.offRootLinks, rootLinks.data(), rootLinks.size() * 4);
  put(ph.offJacEntries, jacEntries.data(), jacEntries.size() * 4);
  if (wb)
  {
    put(ph.offBodyOut, bodyOut.data(), (size_t) nb * 4);
    put(ph.offBodyFrameRec, bodyFrameRec.data(), (size_t) nb * 4);
    put(ph.offTaskBodyRec, taskBodyRec.data(), taskBodyRec.size() * 4);
    put(ph.offJacTasks, wb->jacTasks.data(), wb->jacTasks.size() * 4);
    put(ph.offTaskColOn, taskColOn.data(), taskColOn.size());
  }

  DevicePlan dp;
  RCSB_CUDA(cudaSetDevice(m->device));
  RCSB_CUDA(cudaMalloc(&dp.dev, blob.size()));
  /* the first launch may follow on a non-blocking stream: make sure the plan has landed */
  RCSB_CUDA(cudaMemcpy(dp.dev, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  RCSB_CUDA(cudaDeviceSynchronize());
  dp.bytes = ph.totalBytes;
  dp.aux0 = ph.recDoubles;
  dp.aux1 = nJ;
  dp.nScrRows = ph.walkBytes;
  dp.nSel = nSelOut;
  dp.nEff = ph.nJac;
  dp.aux2 = ph.symmetric;
  if (wb)
  {
    auto ins = m->wbPlans.emplace(wb->key, dp);
    *out = &ins.first->second;
  }
  else
  {
    m->ktPlan[vel ? 1 : 0] = dp;
    m->ktPlanBuilt[vel ? 1 : 0] = true;
    *out = &m->ktPlan[vel ? 1 : 0];
  }
  return RCSB_OK;
}

namespace
{

/* States per pass: the records of one pass (recDoubles x 8 B per state) are written by the walk
 * kernel and read back by the assembly kernel right after. Larger passes are faster (measured
 * on B200, humanoid: 37,888 states 110.6, 75,776: 112.7, 262,144: 116.9 M states/s; arm:
 * 75,776: 657, 1,048,576: 752 M states/s), so a pass takes the whole batch up to a record
 * buffer of 2 GiB. RCSB_KT_CHUNK overrides. */
long ktChunk(const RcsbModel* m, int recDoubles)
{
  const char* env = getenv("RCSB_KT_CHUNK");
  long chunk = env ? atol(env) : 0;
  if (chunk <= 0)
  {
    (void) m;
    chunk = (2L << 30) / ((long) recDoubles * (long) sizeof(double));
  }
  /* a multiple of both block sizes of the walk kernel (128, 192) */
  chunk = chunk < 384 ? 384 : chunk;
  return (chunk + 383) / 384 * 384;
}

}   // namespace

int rcsbhost::ktDevice(RcsbModel* m, long N, int qdim, const double* q, const double* qd, double* M,
                       double* h, double* Fg, double* E, const KtExtra& x, cudaStream_t stream, int ws)
{
  double* cog = x.cog;
  double* Jcog = x.Jcog;
  const DevicePlan& plan = x.plan ? *x.plan : m->ktPlan[qd ? 1 : 0];
  const int recDoubles = plan.aux0;
  const long chunk = ktChunk(m, recDoubles) < N ? ktChunk(m, recDoubles) : N;
  const size_t need = (size_t) chunk * (size_t) recDoubles;
  RCSB_CUDA(cudaSetDevice(m->device));
  {
    /* one record buffer per staging stream: kernels of the two streams may overlap */
    std::lock_guard<std::recursive_mutex> lock(m->mtx);
    if (m->ktScratchDoubles[ws] < need)
    {
      if (m->ktScratch[ws])
      {
        RCSB_CUDA(cudaFree(m->ktScratch[ws]));
        m->ktScratch[ws] = nullptr;
        m->ktScratchDoubles[ws] = 0;
      }
      RCSB_CUDA(cudaMalloc(&m->ktScratch[ws], need * sizeof(double)));
      m->ktScratchDoubles[ws] = need;
    }
  }

  const long nJ = m->hdr.nJ;
  for (long s0 = 0; s0 < N; s0 += chunk)
  {
    const long n = (N - s0 < chunk) ? (N - s0) : chunk;
    rcsb::KtArgs a;
    memset(&a, 0, sizeof(a));
    a.model = m->dev;
    a.plan = plan.dev;
    a.modelBytes = m->hdr.totalBytes;
    a.planBytes = plan.bytes;
    a.planWalkBytes = plan.nScrRows;
    a.wholeBody = x.plan ? 1 : 0;
    a.symmetric = plan.aux2;
    a.A_BI = x.A_BI ? x.A_BI + s0 * (long) plan.nSel * 12 : nullptr;
    a.Jt = x.Jt ? x.Jt + s0 * (long) plan.nEff * 3 * nJ : nullptr;
    a.N = n;
    a.qdim = qdim;
    a.q = q + s0 * qdim;
    a.qd = qd ? qd + s0 * qdim : nullptr;
    a.M = M ? M + s0 * nJ * nJ : nullptr;
    a.h = h ? h + s0 * nJ : nullptr;
    a.Fg = Fg ? Fg + s0 * nJ : nullptr;
    a.E = E ? E + s0 : nullptr;
    a.cog = cog ? cog + s0 * 3 : nullptr;
    a.Jcog = Jcog ? Jcog + s0 * 3 * nJ : nullptr;
    a.Fext = x.Fext ? x.Fext + s0 * nJ : nullptr;
    a.Fjnt = x.Fjnt ? x.Fjnt + s0 * nJ : nullptr;
    a.qdd = x.qdd ? x.qdd + s0 * nJ : nullptr;
    a.stageM = x.qdd ? 1 : 0;   /* the Cholesky solve of the direct dynamics works on the staged M */
    a.rec = m->ktScratch[ws];
    a.recDoubles = recDoubles;
    RCSB_CUDA(rcsb_launch_kt(&a, m->hdr.nSlots, m->hdr.nJ, plan.aux1, m->smCount, stream));
    g_launches.fetch_add(2);
  }
  return RCSB_OK;
}

static int ktCommon(const RcsbModel* cm, long N, int qdim, const double* q, const double* qd,
                    double* M, double* h, double* Fg, double* E, const KtExtra& x, unsigned flags,
                    void* stream)
{
  RcsbModel* m = const_cast<RcsbModel*>(cm);
  int rc = checkModel(m, N, qdim, q);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  DevicePlan* basePlan = nullptr;
  rc = buildKtPlan(m, nullptr, qd != nullptr, &basePlan);
  if (rc != RCSB_OK || N == 0)
  {
    return rc;
  }

  if (!(flags & RCSB_HOST_PTRS))
  {
    /* the record buffer is shared: device-pointer calls on one model must use one stream at a
     * time (documented in rcsb.h) */
    return ktDevice(m, N, qdim, q, qd, M, h, Fg, E, x, static_cast<cudaStream_t>(stream), 0);
  }

  const size_t nJ = (size_t) m->hdr.nJ;
  std::vector<StagedArray> arr;
  enum { I_Q, I_QD, I_M, I_H, I_G, I_E, I_C, I_JC, I_FE, I_FJ, I_QDD };
  arr.push_back({q, nullptr, (size_t) qdim, nullptr});
  arr.push_back({qd, nullptr, qd ? (size_t) qdim : 0, nullptr});
  arr.push_back({nullptr, M, M ? nJ * nJ : 0, nullptr});
  arr.push_back({nullptr, h, h ? nJ : 0, nullptr});
  arr.push_back({nullptr, Fg, Fg ? nJ : 0, nullptr});
  arr.push_back({nullptr, E, (size_t) (E ? 1 : 0), nullptr});
  arr.push_back({nullptr, x.cog, (size_t) (x.cog ? 3 : 0), nullptr});
  arr.push_back({nullptr, x.Jcog, x.Jcog ? 3 * nJ : 0, nullptr});
  arr.push_back({x.Fext, nullptr, x.Fext ? nJ : 0, nullptr});
  arr.push_back({x.Fjnt, nullptr, x.Fjnt ? nJ : 0, nullptr});
  arr.push_back({nullptr, x.qdd, x.qdd ? nJ : 0, nullptr});
  return runStaged(m, N, arr, {}, [&](long n, cudaStream_t st, int slot) {
    auto dv = [&](int i, const void* host) { return host ? arr[(size_t) i].dev : nullptr; };
    KtExtra d;
    d.cog = dv(I_C, x.cog);
    d.Jcog = dv(I_JC, x.Jcog);
    d.Fext = dv(I_FE, x.Fext);
    d.Fjnt = dv(I_FJ, x.Fjnt);
    d.qdd = dv(I_QDD, x.qdd);
    return ktDevice(m, n, qdim, dv(I_Q, q), dv(I_QD, qd), dv(I_M, M), dv(I_H, h), dv(I_G, Fg),
                    dv(I_E, E), d, st, slot);
  });
}

extern "C" int rcsb_kinetic_terms(const RcsbModel* cm, long N, int qdim, const double* q,
                                  const double* qd, double* M, double* h, double* Fg, double* E,
                                  unsigned flags, void* stream)
{
  return ktCommon(cm, N, qdim, q, qd, M, h, Fg, E, KtExtra(), flags, stream);
}

extern "C" int rcsb_cog(const RcsbModel* cm, long N, int qdim, const double* q, double* cog,
                        double* Jcog, unsigned flags, void* stream)
{
  KtExtra x;
  x.cog = cog;
  x.Jcog = Jcog;
  return ktCommon(cm, N, qdim, q, nullptr, nullptr, nullptr, nullptr, nullptr, x, flags, stream);
}

extern "C" int rcsb_kinetic_terms_cog(const RcsbModel* cm, long N, int qdim, const double* q,
                                      const double* qd, double* M, double* h, double* Fg,
                                      double* E, double* cog, double* Jcog, unsigned flags,
                                      void* stream)
{
  KtExtra x;
  x.cog = cog;
  x.Jcog = Jcog;
  return ktCommon(cm, N, qdim, q, qd, M, h, Fg, E, x, flags, stream);
}

extern "C" int rcsb_direct_dynamics(const RcsbModel* cm, long N, int qdim, const double* q,
                                    const double* qd, const double* Fext, const double* Fjnt,
                                    double* qdd, double* E, unsigned flags, void* stream)
{
  if (N > 0 && !qdd)
  {
    rcsb_set_error("rcsb_direct_dynamics: qdd is NULL");
    return RCSB_EINVAL;
  }
  KtExtra x;
  x.Fext = Fext;
  x.Fjnt = Fjnt;
  x.qdd = qdd;
  return ktCommon(cm, N, qdim, q, qd, nullptr, nullptr, nullptr, E, x, flags, stream);
}

/* ------------------------------------------------------------------------------------------ */
/* resolved-motion-rate IK                                                                     */
/* ------------------------------------------------------------------------------------------ */

namespace
{

int buildIkPlan(RcsbModel* m, int nTasksIn, const RcsbTask* tasksIn, const DevicePlan** out)
{
  const int nb = m->hdr.nb, dof = m->hdr.dof, nJ = m->hdr.nJ;
  if (nTasksIn <= 0 || !tasksIn)
  {
    rcsb_set_error("rcsb_ik_rmr: no tasks");
    return RCSB_EINVAL;
  }
  /* a 6-D pose task is a position task followed by an orientation task of the same effector */
  std::vector<RcsbTask> expanded;
  for (int t = 0; t < nTasksIn; ++t)
  {
    if (tasksIn[t].kind == RCSB_TASK_XYZABC)
    {
      expanded.push_back({RCSB_TASK_XYZ, tasksIn[t].effector});
      expanded.push_back({RCSB_TASK_ABC, tasksIn[t].effector});
    }
    else
    {
      expanded.push_back(tasksIn[t]);
    }
  }
  const int nTasks = (int) expanded.size();
  const RcsbTask* tasks = expanded.data();
  for (int t = 0; t < nTasks; ++t)
  {
    if (tasks[t].kind != RCSB_TASK_XYZ && tasks[t].kind != RCSB_TASK_ABC)
    {
      rcsb_set_error("task %d: unknown kind %d", t, tasks[t].kind);
      return RCSB_EINVAL;
    }
    if (tasks[t].effector < 0 || tasks[t].effector >= nb)
    {
      rcsb_set_error("task %d: effector %d out of range [0, %d)", t, tasks[t].effector, nb);
      return RCSB_EINVAL;
    }
  }

  const int* jntJacobi = m->iarr(RCSB_A_JNT_JACOBI);
  const int* jntPrev = m->iarr(RCSB_A_JNT_PREV);
  const int* jntFlags = m->iarr(RCSB_A_JNT_FLAGS);
  const int* jntCpl = m->iarr(RCSB_A_JNT_CPL);
  const int* bodyLastJnt = m->iarr(RCSB_A_BODY_LASTJNT);

  /* RcsGraph_countCoupledJoints (Rcs_graph.c:2804): slaves that are not constrained. They put
   * the solver into the reduced joint space of RcsGraph_coupledJointMatrix (Rcs_graph.c:2823). */
  const RcsbCoupling* cpl = reinterpret_cast<const RcsbCoupling*>(m->host.data() + m->hdr.off[RCSB_A_CPL]);
  std::vector<int> colRed((size_t) (nJ > 0 ? nJ : 1), 0), colCpl((size_t) (nJ > 0 ? nJ : 1), -1);
  int nqr = 0;
  {
    std::vector<int> redOfCol((size_t) (nJ > 0 ? nJ : 1), -1);
    for (int j = 0; j < dof; ++j)
    {
      if (jntJacobi[j] >= 0 && jntCpl[j] < 0)
      {
        redOfCol[jntJacobi[j]] = 0;   /* free or master column */
      }
    }
    for (int c = 0; c < nJ; ++c)
    {
      if (redOfCol[c] == 0)
      {
        redOfCol[c] = nqr++;
      }
    }
    for (int j = 0; j < dof; ++j)
    {
      const int c = jntJacobi[j];
      if (c < 0)
      {
        continue;
      }
      if (jntCpl[j] < 0)
      {
        colRed[c] = redOfCol[c];
      }
      else
      {
        const int mc = jntJacobi[cpl[jntCpl[j]].master];
        if (mc < 0 || redOfCol[mc] < 0)
        {
          rcsb_set_error("rcsb_ik_rmr: slave joint %d is coupled to a constrained or slave joint", j);
          return RCSB_EUNSUPPORTED;
        }
        colRed[c] = redOfCol[mc];
        colCpl[c] = jntCpl[j];
      }
    }
  }

  std::string key(reinterpret_cast<const char*>(tasks), sizeof(RcsbTask) * (size_t) nTasks);
  std::lock_guard<std::recursive_mutex> lock(m->mtx);
  auto it = m->ikPlans.find(key);
  if (it != m->ikPlans.end())
  {
    *out = &it->second;
    return RCSB_OK;
  }

  std::vector<int> kind(nTasks), body(nTasks), needed(dof, 0), chainStart(nTasks + 1, 0), chain;
  std::vector<int> taskStart(nb + 1, 0), taskList(nTasks, 0);
  for (int t = 0; t < nTasks; ++t)
  {
    kind[t] = tasks[t].kind;
    body[t] = tasks[t].effector;
    taskStart[body[t] + 1]++;
    for (int j = bodyLastJnt[body[t]]; j >= 0; j = jntPrev[j])
    {
      if (jntJacobi[j] >= 0)
      {
        needed[j] = 1;
        chain.push_back(jntJacobi[j] | ((jntFlags[j] & RCSB_JF_ROT) ? (1 << 16) : 0));
      }
    }
    chainStart[t + 1] = (int) chain.size();
  }
  for (int b = 0; b < nb; ++b)
  {
    taskStart[b + 1] += taskStart[b];
  }
  {
    std::vector<int> fill(taskStart.begin(), taskStart.end() - 1);
    for (int t = 0; t < nTasks; ++t)
    {
      taskList[fill[body[t]]++] = t;
    }
  }
  if (chain.empty())
  {
    chain.push_back(0);
  }

  const double* qmin = m->darr(RCSB_A_JNT_QMIN);
  const double* qmax = m->darr(RCSB_A_JNT_QMAX);
  const double* q0 = m->darr(RCSB_A_JNT_Q0);
  const double* wjl = m->darr(RCSB_A_JNT_WJL);
  const double* wmetric = m->darr(RCSB_A_JNT_WMETRIC);
  std::vector<int> colJoint(nJ > 0 ? nJ : 1, 0);
  std::vector<double> invWq(nJ > 0 ? nJ : 1, 0.0), jlGain(nJ > 0 ? nJ : 1, 0.0),
    q0c(nJ > 0 ? nJ : 1, 0.0);
  for (int j = 0; j < dof; ++j)
  {
    const int c = jntJacobi[j];
    if (c >= 0)
    {
      const double range = qmax[j] - qmin[j];
      colJoint[c] = j;
      invWq[c] = wmetric[j] * range;
      jlGain[c] = range > 0.0 ? wjl[j] / (range * range) : 0.0;
      q0c[c] = q0[j];
    }
  }

  /* ---- single-chain program (fast path) -------------------------------------------------- */
  const int* bodyParent = m->iarr(RCSB_A_BODY_PARENT);
  const int* bodyJntStart = m->iarr(RCSB_A_BODY_JNT_START);
  const int* bodyJntCount = m->iarr(RCSB_A_BODY_JNT_COUNT);
  const int* bodyFlags = m->iarr(RCSB_A_BODY_FLAGS);
  const int* jntType = m->iarr(RCSB_A_JNT_TYPE);
  std::vector<int> progStart, prog, chainType, chainJoint, otherJoint;
  std::vector<double> chainXf, segXf;
  std::vector<double> chainW, chainGain, chainQ0, otherW, otherGain, otherQ0;
  int chainNc = 0;
  {
    bool eligible = (nTasks <= 2);
    for (int t = 1; t < nTasks; ++t)
    {
      eligible = eligible && (body[t] == body[0]) && (kind[t] != kind[0]);
    }
    if (eligible)
    {
      std::vector<int> path;
      for (int b = body[0]; b >= 0; b = bodyParent[b])
      {
        path.push_back(b);
      }
      /* events along the chain: relative transforms, joints with a fixed value, columns */
      struct Event
      {
        int kind;       /* 0 transform, 1 joint that is not a column, 2 column, 3 row shift */
        int joint;
        int in, out;    /* transform / row shift: shift of the frame before / after it; joints:
                           `in` = physical row that holds the joint's axis */
        double A[12];
      };
      std::vector<Event> ev;
      const double* ajp = m->darr(RCSB_A_JNT_AJP);
      const double* abp = m->darr(RCSB_A_BODY_ABP);
      /* the register kernel (at most 8 columns, fully unrolled) gains from the shifted-row
       * representation; the loop kernel for longer chains does not (measured on the humanoid's
       * hand: 188 -> 181 M solves/s) and keeps the rows in place */
      int pathCols = 0;
      for (int b : path)
      {
        for (int j = bodyJntStart[b]; j < bodyJntStart[b] + bodyJntCount[b]; ++j)
        {
          pathCols += jntJacobi[j] >= 0 ? 1 : 0;
        }
      }
      const bool permute = pathCols <= 8;
      int shift = 0, lastXf = -1;
      auto addXf = [&](const double* A) {
        Event e;
        e.kind = 0;
        e.joint = -1;
        e.in = e.out = shift;
        if (A)
        {
          memcpy(e.A, A, sizeof(e.A));
        }
        else
        {
          const double I[12] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
          memcpy(e.A, I, sizeof(e.A));
        }
        lastXf = (int) ev.size();
        ev.push_back(e);
      };
      auto wantShift = [&](int need) {
        if (permute && need != shift)
        {
          bool fold = lastXf >= 0;
          if (fold)
          {
            /* a transform without rotation costs no arithmetic on the matrix: folding the shift
             * into it would turn it into a full 3 x 3 product, a register-only row shift is cheaper */
            const double* A = ev[(size_t) lastXf].A;
            fold = !(A[3] == 1.0 && A[4] == 0.0 && A[5] == 0.0 && A[6] == 0.0 && A[7] == 1.0 &&
                     A[8] == 0.0 && A[9] == 0.0 && A[10] == 0.0 && A[11] == 1.0);
          }
          if (!fold)
          {
            /* nothing to fold the permutation into: a register-only row shift */
            Event e;
            e.kind = 3;
            e.joint = -1;
            e.in = shift;
            e.out = need;
            ev.push_back(e);
          }
          else
          {
            ev[(size_t) lastXf].out = need;
          }
          shift = need;
        }
      };
      for (auto it = path.rbegin(); it != path.rend(); ++it)
      {
        const int b = *it;
        for (int j = bodyJntStart[b]; j < bodyJntStart[b] + bodyJntCount[b]; ++j)
        {
          if (jntFlags[j] & RCSB_JF_HAS_AJP)
          {
            addXf(ajp + 12 * j);
          }
          const int dir = jntType[j] % 3;
          if (jntFlags[j] & RCSB_JF_ROT)
          {
            wantShift((dir + 1) % 3);   /* axis -> row 2, the rows it mixes -> rows 0 and 1 */
            lastXf = -1;                /* the frame changes: later shifts cannot be folded back */
          }
          Event e;
          e.kind = jntJacobi[j] >= 0 ? 2 : 1;
          e.joint = j;
          e.in = e.out = (dir - shift + 3) % 3;   /* a translation only reads its axis row */
          ev.push_back(e);
        }
        if (bodyFlags[b] & RCSB_BF_HAS_ABP)
        {
          addXf(abp + 12 * b);
        }
      }
      wantShift(0);   /* the effector frame comes out with its rows in place */

      progStart.push_back(0);
      for (const Event& e : ev)
      {
        if (e.kind == 0)
        {
          double X[12];
          bool rotIdent = true, noTrans = true;
          for (int i = 0; i < 3; ++i)
          {
            X[i] = e.A[(i + e.in) % 3];
            noTrans = noTrans && X[i] == 0.0;
            for (int k = 0; k < 3; ++k)
            {
              X[3 + 3 * i + k] = e.A[3 + 3 * ((i + e.out) % 3) + (k + e.in) % 3];
              rotIdent = rotIdent && X[3 + 3 * i + k] == (i == k ? 1.0 : 0.0);
            }
          }
          const int step[4] = {0, (int) chainXf.size() / 12,
                               (rotIdent ? RCSB_REL_ROT_IDENT : 0) | (noTrans ? RCSB_REL_NO_TRANS : 0), 0};
          chainXf.insert(chainXf.end(), X, X + 12);
          prog.insert(prog.end(), step, step + 4);
        }
        else if (e.kind == 3)
        {
          const int step[4] = {2, (e.out - e.in + 3) % 3, 0, 0};
          prog.insert(prog.end(), step, step + 4);
        }
        else if (e.kind == 2)
        {
          chainType.push_back(jntType[e.joint] | (e.in << 4));
          chainJoint.push_back(e.joint);
          progStart.push_back((int) prog.size() / 4);
        }
        else
        {
          const int step[4] = {1, jntType[e.joint], e.joint, e.in};
          prog.insert(prog.end(), step, step + 4);
        }
      }
      progStart.push_back((int) prog.size() / 4);
      chainNc = (int) chainJoint.size();
      if (chainNc < 1 || chainNc > 32 || m->hdr.nCpl > 0)
      {
        chainNc = 0;   /* long chains and graphs with coupled joints take the general kernel */
      }
    }
    if (chainNc > 0)
    {
      std::vector<char> onChain((size_t) dof, 0);
      for (int c = 0; c < chainNc; ++c)
      {
        const int col = jntJacobi[chainJoint[c]];
        onChain[(size_t) chainJoint[c]] = 1;
        chainW.push_back(invWq[col]);
        chainGain.push_back(jlGain[col]);
        chainQ0.push_back(q0c[col]);
      }
      for (int j = 0; j < dof; ++j)
      {
        const int col = jntJacobi[j];
        if (col >= 0 && !onChain[(size_t) j])
        {
          otherJoint.push_back(j);
          otherW.push_back(invWq[col]);
          otherGain.push_back(jlGain[col]);
          otherQ0.push_back(q0c[col]);
        }
      }
    }
    else
    {
      progStart.assign(2, 0);
      prog.clear();
      chainXf.clear();
      chainType.clear();
      chainJoint.clear();
    }
    /* every segment of the program folded into one transform (RcsbIkPlanHeader::offSegXf): applying
     * A then B is applying {A.org + A.rot^T B.org, B.rot A.rot}; exact whenever one of the two is a
     * pure translation, a permutation or the identity */
    if (chainNc > 0 && chainNc <= 8)
    {
      bool fold = true;
      for (int c = 0; c < chainNc; ++c)
      {
        fold = fold && (chainType[(size_t) c] & 15) < 3;
      }
      for (size_t k = 0; fold && k < prog.size() / 4; ++k)
      {
        fold = prog[4 * k] != 1;
      }
      for (int c = 0; fold && c <= chainNc; ++c)
      {
        double X[12] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int k = progStart[(size_t) c]; k < progStart[(size_t) c + 1]; ++k)
        {
          double B[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
          if (prog[4 * (size_t) k] == 0)
          {
            memcpy(B, chainXf.data() + 12 * (size_t) prog[4 * (size_t) k + 1], sizeof(B));
          }
          else
          {
            const int sh = prog[4 * (size_t) k + 1];   /* new row i = old row (i + sh) mod 3 */
            for (int i = 0; i < 3; ++i)
            {
              B[3 + 3 * i + (i + sh) % 3] = 1.0;
            }
          }
          double Y[12];
          for (int i = 0; i < 3; ++i)
          {
            Y[i] = X[i] + (X[3 + i] * B[0] + X[6 + i] * B[1] + X[9 + i] * B[2]);
            for (int j = 0; j < 3; ++j)
            {
              Y[3 + 3 * i + j] = B[3 + 3 * i] * X[3 + j] + B[3 + 3 * i + 1] * X[6 + j] + B[3 + 3 * i + 2] * X[9 + j];
            }
          }
          memcpy(X, Y, sizeof(X));
        }
        segXf.insert(segXf.end(), X, X + 12);
      }
      if (!fold)
      {
        segXf.clear();
      }
    }
    if (prog.empty())
    {
      prog.assign(4, 0);
    }
  }

  /* ---- several effectors, no couplings (ikMultiKernel) ------------------------------------ */
  std::vector<int> jntSlot((size_t) dof, -1), slotTaskMask, taskEff((size_t) nTasks, 0),
    bodyEffIdx((size_t) nb, -1);
  int multiNu = 0, nEffs = 0;
  if (chainNc == 0 && m->hdr.nCpl == 0 && nTasks <= 4)
  {
    for (int t = 0; t < nTasks; ++t)
    {
      if (bodyEffIdx[(size_t) body[t]] < 0)
      {
        bodyEffIdx[(size_t) body[t]] = nEffs++;
      }
      taskEff[(size_t) t] = bodyEffIdx[(size_t) body[t]];
    }
    for (int j = 0; j < dof; ++j)
    {
      if (needed[j])
      {
        jntSlot[(size_t) j] = multiNu++;
      }
    }
    if (multiNu >= 1 && multiNu <= 48)
    {
      slotTaskMask.assign((size_t) multiNu, 0);
      for (int t = 0; t < nTasks; ++t)
      {
        for (int j = bodyLastJnt[body[t]]; j >= 0; j = jntPrev[j])
        {
          if (jntJacobi[j] >= 0)
          {
            slotTaskMask[(size_t) jntSlot[(size_t) j]] |= 1 << t;
          }
        }
      }
      chainType.clear();
      chainJoint.clear();
      chainW.clear();
      chainGain.clear();
      chainQ0.clear();
      otherJoint.clear();
      otherW.clear();
      otherGain.clear();
      otherQ0.clear();
      for (int j = 0; j < dof; ++j)
      {
        const int col = jntJacobi[j];
        if (col < 0)
        {
          continue;
        }
        if (jntSlot[(size_t) j] >= 0)
        {
          chainType.push_back(jntType[j]);
          chainJoint.push_back(j);
          chainW.push_back(invWq[col]);
          chainGain.push_back(jlGain[col]);
          chainQ0.push_back(q0c[col]);
        }
        else
        {
          otherJoint.push_back(j);
          otherW.push_back(invWq[col]);
          otherGain.push_back(jlGain[col]);
          otherQ0.push_back(q0c[col]);
        }
      }
    }
    else
    {
      multiNu = 0;
    }
  }
  if (slotTaskMask.empty())
  {
    slotTaskMask.assign(1, 0);
  }
  /* pruned walk: ancestors-or-self of the effectors in depth-first order, with a frame-stack
   * program of their own (a body whose parent is not the previously visited one reloads it) */
  std::vector<int> walk;
  int walkSlots = 0;
  if (multiNu > 0)
  {
    std::vector<char> wanted((size_t) nb, 0);
    for (int t = 0; t < nTasks; ++t)
    {
      for (int b = body[t]; b >= 0; b = bodyParent[b])
      {
        wanted[(size_t) b] = 1;
      }
    }
    std::vector<int> list;
    for (int b = 0; b < nb; ++b)
    {
      if (wanted[(size_t) b])
      {
        list.push_back(b);
      }
    }
    std::vector<int> saved;   /* bodies whose frame is parked, innermost last */
    auto isAncestor = [&](int anc, int b) {
      for (b = bodyParent[b]; b >= 0; b = bodyParent[b])
      {
        if (b == anc)
        {
          return true;
        }
      }
      return false;
    };
    for (size_t i = 0; i < list.size(); ++i)
    {
      const int b = list[i];
      while (!saved.empty() && !isAncestor(saved.back(), b))
      {
        saved.pop_back();
      }
      int load = RCSB_LOAD_IDENT;
      if (bodyParent[b] >= 0)
      {
        if (i > 0 && list[i - 1] == bodyParent[b])
        {
          load = RCSB_LOAD_CONT;
        }
        else
        {
          load = -3;
          for (size_t k = 0; k < saved.size(); ++k)
          {
            if (saved[k] == bodyParent[b])
            {
              load = (int) k;
            }
          }
        }
      }
      /* park the frame if a later body (not the next one) hangs under this one */
      bool later = false;
      for (size_t k = i + 2; k < list.size(); ++k)
      {
        later = later || bodyParent[list[k]] == b;
      }
      int save = -1;
      if (later)
      {
        saved.push_back(b);
        save = (int) saved.size() - 1;
        walkSlots = (int) saved.size() > walkSlots ? (int) saved.size() : walkSlots;
      }
      const int w4[4] = {b, load, save, 0};
      walk.insert(walk.end(), w4, w4 + 4);
      if (load == -3 || walkSlots > RCSB_MAX_FRAME_SLOTS)
      {
        multiNu = 0;   /* should not happen; the general kernel takes over */
      }
    }
  }
  if (walk.empty())
  {
    walk.assign(4, 0);
  }

  RcsbIkPlanHeader ph;
  memset(&ph, 0, sizeof(ph));
  ph.nTasks = nTasks;
  ph.nx = 3 * nTasks;
  ph.nJ = nJ;
  ph.nqr = nqr;
  ph.multiNu = multiNu;
  ph.nEffs = nEffs;
  ph.nWalk = multiNu > 0 ? (int) walk.size() / 4 : 0;
  ph.walkSlots = walkSlots;
  ph.chainNc = chainNc;
  ph.nOther = (int) otherJoint.size();
  int off = align16((int) sizeof(ph));
  auto place = [&](int& field, size_t bytes) {
    field = off;
    off = align16(off + (int) bytes);
  };
  place(ph.offTaskKind, kind.size() * 4);
  place(ph.offTaskBody, body.size() * 4);
  place(ph.offBodyTaskStart, taskStart.size() * 4);
  place(ph.offBodyTaskList, taskList.size() * 4);
  place(ph.offJntNeeded, (needed.empty() ? 1 : needed.size()) * 4);
  place(ph.offChainStart, chainStart.size() * 4);
  place(ph.offChain, chain.size() * 4);
  place(ph.offColJoint, colJoint.size() * 4);
  place(ph.offInvWq, invWq.size() * 8);
  place(ph.offJlGain, jlGain.size() * 8);
  place(ph.offQ0, q0c.size() * 8);
  auto atLeast1 = [](size_t n) { return n > 0 ? n : (size_t) 1; };
  place(ph.offColRed, colRed.size() * 4);
  place(ph.offColCpl, colCpl.size() * 4);
  place(ph.offProgStart, progStart.size() * 4);
  place(ph.offProg, prog.size() * 4);
  place(ph.offChainXf, atLeast1(chainXf.size()) * 8);
  ph.offSegXf = 0;
  if (!segXf.empty())
  {
    place(ph.offSegXf, segXf.size() * 8);
  }
  place(ph.offChainType, atLeast1(chainType.size()) * 4);
  place(ph.offChainJoint, atLeast1(chainJoint.size()) * 4);
  place(ph.offChainW, atLeast1(chainW.size()) * 8);
  place(ph.offChainGain, atLeast1(chainGain.size()) * 8);
  place(ph.offChainQ0, atLeast1(chainQ0.size()) * 8);
  place(ph.offOtherJoint, atLeast1(otherJoint.size()) * 4);
  place(ph.offOtherW, atLeast1(otherW.size()) * 8);
  place(ph.offOtherGain, atLeast1(otherGain.size()) * 8);
  place(ph.offOtherQ0, atLeast1(otherQ0.size()) * 8);
  place(ph.offJntSlot, jntSlot.size() * 4);
  place(ph.offSlotTaskMask, slotTaskMask.size() * 4);
  place(ph.offTaskEff, taskEff.size() * 4);
  place(ph.offBodyEffIdx, bodyEffIdx.size() * 4);
  place(ph.offWalk, walk.size() * 4);
  ph.totalBytes = off;

  std::vector<char> blob((size_t) ph.totalBytes, 0);
  memcpy(blob.data(), &ph, sizeof(ph));
  memcpy(blob.data() + ph.offTaskKind, kind.data(), kind.size() * 4);
  memcpy(blob.data() + ph.offTaskBody, body.data(), body.size() * 4);
  memcpy(blob.data() + ph.offBodyTaskStart, taskStart.data(), taskStart.size() * 4);
  memcpy(blob.data() + ph.offBodyTaskList, taskList.data(), taskList.size() * 4);
  if (!needed.empty())
  {
    memcpy(blob.data() + ph.offJntNeeded, needed.data(), needed.size() * 4);
  }
  memcpy(blob.data() + ph.offChainStart, chainStart.data(), chainStart.size() * 4);
  memcpy(blob.data() + ph.offChain, chain.data(), chain.size() * 4);
  memcpy(blob.data() + ph.offColJoint, colJoint.data(), colJoint.size() * 4);
  memcpy(blob.data() + ph.offInvWq, invWq.data(), invWq.size() * 8);
  memcpy(blob.data() + ph.offJlGain, jlGain.data(), jlGain.size() * 8);
  memcpy(blob.data() + ph.offQ0, q0c.data(), q0c.size() * 8);
  auto put = [&](int offset, const void* src, size_t bytes) {
    if (bytes > 0)
    {
      memcpy(blob.data() + offset, src, bytes);
    }
  };
  put(ph.offColRed, colRed.data(), colRed.size() * 4);
  put(ph.offColCpl, colCpl.data(), colCpl.size() * 4);
  put(ph.offProgStart, progStart.data(), progStart.size() * 4);
  put(ph.offProg, prog.data(), prog.size() * 4);
  put(ph.offChainXf, chainXf.data(), chainXf.size() * 8);
  if (!segXf.empty())
  {
    put(ph.offSegXf, segXf.data(), segXf.size() * 8);
  }
  put(ph.offChainType, chainType.data(), chainType.size() * 4);
  put(ph.offChainJoint, chainJoint.data(), chainJoint.size() * 4);
  put(ph.offChainW, chainW.data(), chainW.size() * 8);
  put(ph.offChainGain, chainGain.data(), chainGain.size() * 8);
  put(ph.offChainQ0, chainQ0.data(), chainQ0.size() * 8);
  put(ph.offOtherJoint, otherJoint.data(), otherJoint.size() * 4);
  put(ph.offOtherW, otherW.data(), otherW.size() * 8);
  put(ph.offOtherGain, otherGain.data(), otherGain.size() * 8);
  put(ph.offOtherQ0, otherQ0.data(), otherQ0.size() * 8);
  put(ph.offJntSlot, jntSlot.data(), jntSlot.size() * 4);
  put(ph.offSlotTaskMask, slotTaskMask.data(), slotTaskMask.size() * 4);
  put(ph.offTaskEff, taskEff.data(), taskEff.size() * 4);
  put(ph.offBodyEffIdx, bodyEffIdx.data(), bodyEffIdx.size() * 4);
  put(ph.offWalk, walk.data(), walk.size() * 4);

  DevicePlan dp;
  RCSB_CUDA(cudaSetDevice(m->device));
  RCSB_CUDA(cudaMalloc(&dp.dev, blob.size()));
  RCSB_CUDA(cudaMemcpy(dp.dev, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  RCSB_CUDA(cudaDeviceSynchronize());   /* pageable source: the plan must have landed before the first launch */
  dp.bytes = ph.totalBytes;
  dp.aux0 = ph.nx;
  dp.aux1 = ph.chainNc;
  dp.nEff = ph.multiNu;
  dp.nSel = ph.nEffs;
  dp.nScrRows = nTasks | (walkSlots << 8);
  {
    /* the specialised chain instances: all columns rotations, no joint steps in the program */
    dp.aux2 = (!segXf.empty() ? RCSB_IK_CHAIN_ALLROT : 0) | (kind[0] == RCSB_TASK_XYZ ? RCSB_IK_CHAIN_XYZ_FIRST : 0);
  }
  auto ins = m->ikPlans.emplace(key, dp);
  *out = &ins.first->second;
  return RCSB_OK;
}

/* extended inputs of rcsb_ik_rmr_ex, resolved on the host */
struct IkEx
{
  bool general = false;       /* needs the general kernel */
  int lambdaPerRow = 0;
  int blending = 0;
  double lambdaRow[12] = {0.0};
  double taskScale[4] = {1.0, 1.0, 1.0, 1.0};
  double taskAct[4] = {1.0, 1.0, 1.0, 1.0};
  double* dqTs = nullptr;
  double* dqNs = nullptr;
  /* activation mask: xDes rows have nxAll entries, the plan's tasks read the columns xMap[0..nx) */
  int nxAll = 0;
  std::vector<int> xMap;
};

/* x_des rows of the active tasks, gathered from the full rows */
__global__ void __launch_bounds__(256)
ikGatherKernel(const double* __restrict__ src, double* __restrict__ dst, long N, int nxAll, int nxAct,
               const int* __restrict__ map)
{
  const long total = N * nxAct;
  for (long e = (long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long) gridDim.x * blockDim.x)
  {
    const long s = e / nxAct;
    const int k = (int) (e - s * nxAct);
    dst[e] = src[s * nxAll + map[k]];
  }
}

int ikDevice(RcsbModel* m, const DevicePlan* plan, long N, double* q, const double* xDes,
             double lambda, double alpha, int iters, double* dx, double* dq, double* det,
             bool leftInverse, cudaStream_t stream, const IkEx* ex = nullptr, int ws = 0)
{
  rcsb::IkArgs a;
  memset(&a, 0, sizeof(a));
  RCSB_CUDA(cudaSetDevice(m->device));
  if (ex && !ex->xMap.empty())
  {
    /* compressToActive on x_des: one small gather kernel into a per-stream scratch buffer */
    const int nxAct = (int) ex->xMap.size();
    const size_t need = (size_t) N * nxAct * sizeof(double) + 256;
    {
      std::lock_guard<std::recursive_mutex> lock(m->mtx);
      if (m->ikScratchBytes[ws] < need)
      {
        if (m->ikScratch[ws])
        {
          RCSB_CUDA(cudaFree(m->ikScratch[ws]));
          m->ikScratch[ws] = nullptr;
          m->ikScratchBytes[ws] = 0;
        }
        RCSB_CUDA(cudaMalloc(&m->ikScratch[ws], need));
        m->ikScratchBytes[ws] = need;
      }
    }
    int* dMap = reinterpret_cast<int*>(m->ikScratch[ws]);
    double* dX = reinterpret_cast<double*>(m->ikScratch[ws] + 256);
    RCSB_CUDA(cudaMemcpyAsync(dMap, ex->xMap.data(), sizeof(int) * (size_t) nxAct, cudaMemcpyHostToDevice, stream));
    const long total = N * nxAct;
    long blocks = (total + 255) / 256;
    const long cap = (long) m->smCount * 8;
    blocks = blocks < cap ? blocks : cap;
    ikGatherKernel<<<(unsigned int) (blocks > 0 ? blocks : 1), 256, 0, stream>>>(xDes, dX, N, ex->nxAll, nxAct, dMap);
    RCSB_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    xDes = dX;
  }
  if (ex)
  {
    a.extended = 1;
    a.lambdaPerRow = ex->lambdaPerRow;
    a.blending = ex->blending;
    memcpy(a.lambdaRow, ex->lambdaRow, sizeof(a.lambdaRow));
    memcpy(a.taskScale, ex->taskScale, sizeof(a.taskScale));
    memcpy(a.taskAct, ex->taskAct, sizeof(a.taskAct));
    a.dqTs = ex->dqTs;
    a.dqNs = ex->dqNs;
  }
  a.model = m->dev;
  a.plan = plan->dev;
  a.modelBytes = m->hdr.totalBytes;
  a.planBytes = plan->bytes;
  a.N = N;
  a.q = q;
  a.xDes = xDes;
  a.lambda = lambda;
  a.alpha = alpha;
  a.iters = iters;
  a.leftInverse = leftInverse ? 1 : 0;
  a.chainFlags = plan->aux2;
  a.dx = dx;
  a.dq = dq;
  a.det = det;
  const char* noFast = getenv("RCSB_IK_GENERAL");   /* force the general kernel (tests) */
  const bool general = (noFast && noFast[0] == '1') || iters > 64 || (ex && ex->general);
  const int chainNc = general ? 0 : plan->aux1;
  const int multiNu = general ? 0 : plan->nEff;
  cudaError_t e = rcsb_launch_ik(&a, m->hdr.nSlots, plan->aux0, m->hdr.nJ, m->hdr.dof, chainNc,
                                 multiNu, plan->nSel | ((plan->nScrRows >> 8) << 16), plan->nScrRows & 0xff, stream);
  if (e == cudaErrorInvalidValue)
  {
    cudaGetLastError();
    rcsb_set_error("rcsb_ik_rmr: problem too large (nx %d, nJ %d, dof %d; limits 12 / 40 / 72)",
                   plan->aux0, m->hdr.nJ, m->hdr.dof);
    return RCSB_EUNSUPPORTED;
  }
  RCSB_CUDA(e);
  g_launches.fetch_add(1);
  return RCSB_OK;
}

}   // namespace

/* unconstrained joints that are kinematically coupled slaves (RcsGraph_countCoupledJoints, Rcs_graph.c:2804) */
static int RcsGraph_countCoupledColumns(const RcsbModel* m)
{
  const int* jntJacobi = m->iarr(RCSB_A_JNT_JACOBI);
  const int* jntCpl = m->iarr(RCSB_A_JNT_CPL);
  int n = 0;
  for (int j = 0; j < m->hdr.dof; ++j)
  {
    n += (jntJacobi[j] >= 0 && jntCpl[j] >= 0) ? 1 : 0;
  }
  return n;
}

static int taskDim(int kind)
{
  return kind == RCSB_TASK_XYZABC ? 6 : 3;
}

extern "C" int rcsb_ik_rmr_ex(const RcsbModel* cm, int nTasks, const RcsbTask* tasks, long N, double* q,
                              const double* xDes, double lambda, double alpha, int iters, double* dx,
                              double* dq, double* det, const RcsbIkOptions* opt, unsigned flags,
                              void* stream)
{
  RcsbModel* m = const_cast<RcsbModel*>(cm);
  if (!m)
  {
    rcsb_set_error("model is NULL");
    return RCSB_EINVAL;
  }
  if (N < 0 || (N > 0 && (!q || !xDes)) || iters < 0 || nTasks <= 0 || !tasks)
  {
    rcsb_set_error("rcsb_ik_rmr: invalid batch (N = %ld, q = %p, xDes = %p, iters = %d, nTasks = %d)", N,
                   (void*) q, (const void*) xDes, iters, nTasks);
    return RCSB_EINVAL;
  }
  const bool left = (flags & RCSB_IK_LEFT_INVERSE) != 0;

  /* ---- activation mask and the extended inputs ------------------------------------------- */
  std::vector<RcsbTask> active;
  IkEx ex;
  bool useEx = false;
  if (opt)
  {
    if (opt->blending != RCSB_BLEND_BINARY && opt->blending != RCSB_BLEND_LINEAR &&
        opt->blending != RCSB_BLEND_APPROXIMATE)
    {
      rcsb_set_error("rcsb_ik_rmr_ex: blending mode %d is not implemented (Binary, Linear, Approximate are)",
                     opt->blending);
      return RCSB_EUNSUPPORTED;
    }
    int planTask = 0, col = 0;
    bool allActive = true;
    for (int t = 0; t < nTasks; ++t)
    {
      const int dim = taskDim(tasks[t].kind);
      const double act = opt->activation ? opt->activation[t] : 1.0;
      if (act > 0.0)
      {
        active.push_back(tasks[t]);
        for (int k = 0; k < dim; ++k)
        {
          ex.xMap.push_back(col + k);
        }
        for (int k = 0; k < dim / 3; ++k, ++planTask)
        {
          if (planTask < 4)
          {
            ex.taskScale[planTask] = opt->scaleDx ? act : 1.0;
            ex.taskAct[planTask] = act;
          }
          ex.general = ex.general || (opt->scaleDx && act != 1.0);
        }
      }
      else
      {
        allActive = false;
      }
      col += dim;
    }
    ex.nxAll = col;
    if (allActive)
    {
      ex.xMap.clear();   /* nothing to gather */
    }
    if (active.empty())
    {
      rcsb_set_error("rcsb_ik_rmr_ex: no task is active");
      return RCSB_EINVAL;
    }
    const int nxAct = allActive ? col : (int) ex.xMap.size();
    if (opt->lambdaRows)
    {
      if (left)
      {
        rcsb_set_error("rcsb_ik_rmr_ex: the left inverse takes a scalar lambda (IkSolverRMR.cpp:205)");
        return RCSB_EINVAL;
      }
      if (nxAct > 12)
      {
        rcsb_set_error("rcsb_ik_rmr_ex: per-row lambda supports at most 12 active rows (%d)", nxAct);
        return RCSB_EUNSUPPORTED;
      }
      ex.lambdaPerRow = 1;
      memcpy(ex.lambdaRow, opt->lambdaRows, sizeof(double) * (size_t) nxAct);
      ex.general = true;
    }
    if (left && opt->blending != RCSB_BLEND_BINARY)
    {
      ex.blending = opt->blending;
      ex.general = true;
    }
    ex.dqTs = opt->dqTs;
    ex.dqNs = opt->dqNs;
    ex.general = ex.general || opt->dqTs || opt->dqNs;
    if (ex.general && planTask > 4)
    {
      rcsb_set_error("rcsb_ik_rmr_ex: the extended solver takes at most 4 three-dimensional tasks (%d)",
                     planTask);
      return RCSB_EUNSUPPORTED;
    }
    useEx = ex.general || !ex.xMap.empty();
    tasks = active.data();
    nTasks = (int) active.size();
  }

  const DevicePlan* plan = nullptr;
  int rc = buildIkPlan(m, nTasks, tasks, &plan);
  if (rc != RCSB_OK || N == 0)
  {
    return rc;
  }
  const IkEx* exp = useEx ? &ex : nullptr;

  if (!(flags & RCSB_HOST_PTRS))
  {
    return ikDevice(m, plan, N, q, xDes, lambda, alpha, iters, dx, dq, det, left,
                    static_cast<cudaStream_t>(stream), exp, 0);
  }

  const size_t dof = (size_t) m->hdr.dof, nx = (size_t) plan->aux0;
  const size_t nxIn = (exp && !ex.xMap.empty()) ? (size_t) ex.nxAll : nx;
  std::vector<StagedArray> arr;
  enum { I_QIN, I_X, I_QOUT, I_DX, I_DQ, I_DET, I_TS, I_NS };
  arr.push_back({q, nullptr, dof, nullptr});
  arr.push_back({xDes, nullptr, nxIn, nullptr});
  arr.push_back({nullptr, q, dof, nullptr});          /* in-out: shares the input buffer */
  arr.push_back({nullptr, dx, dx ? nx : 0, nullptr});
  arr.push_back({nullptr, dq, dq ? dof : 0, nullptr});
  arr.push_back({nullptr, det, (size_t) (det ? 1 : 0), nullptr});
  arr.push_back({nullptr, ex.dqTs, ex.dqTs ? dof : 0, nullptr});
  arr.push_back({nullptr, ex.dqNs, ex.dqNs ? dof : 0, nullptr});
  return runStaged(m, N, arr, {-1, -1, I_QIN, -1, -1, -1, -1, -1}, [&](long n, cudaStream_t st, int slot) {
    auto dv = [&](int i, const void* host) { return host ? arr[(size_t) i].dev : nullptr; };
    IkEx local = ex;
    local.dqTs = dv(I_TS, ex.dqTs);
    local.dqNs = dv(I_NS, ex.dqNs);
    return ikDevice(m, plan, n, arr[I_QIN].dev, arr[I_X].dev, lambda, alpha, iters,
                    dv(I_DX, dx), dv(I_DQ, dq), dv(I_DET, det), left, st, useEx ? &local : nullptr, slot);
  });
}

static int RcsGraph_countCoupledColumns(const RcsbModel* m);

/* ---- general task stacks (rcsb_ik_tasks_rmr) ------------------------------------------------ */

static int buildIkTasksPlan(RcsbModel* m, int nTasksIn, const RcsbTaskSpec* tasksIn, const DevicePlan** out)
{
  const int nb = m->hdr.nb, dof = m->hdr.dof, nJ = m->hdr.nJ;
  const int* jntJacobi = m->iarr(RCSB_A_JNT_JACOBI);
  const int* jntPrev = m->iarr(RCSB_A_JNT_PREV);
  const int* bodyLastJnt = m->iarr(RCSB_A_BODY_LASTJNT);
  const int* bodyParent = m->iarr(RCSB_A_BODY_PARENT);
  if (nJ > 64)
  {
    rcsb_set_error("rcsb_ik_tasks_rmr: more than 64 Jacobian columns (%d)", nJ);
    return RCSB_EUNSUPPORTED;
  }
  if (RcsGraph_countCoupledColumns(m) > 0)
  {
    rcsb_set_error("rcsb_ik_tasks_rmr: kinematically coupled unconstrained joints are not supported");
    return RCSB_EUNSUPPORTED;
  }
  std::string key("T");
  key.append(reinterpret_cast<const char*>(tasksIn), sizeof(RcsbTaskSpec) * (size_t) nTasksIn);
  std::lock_guard<std::recursive_mutex> lock(m->mtx);
  auto it = m->ikPlans.find(key);
  if (it != m->ikPlans.end())
  {
    *out = &it->second;
    return RCSB_OK;
  }

  /* frame slots: 0 = world, then one per distinct body */
  std::vector<int> slotBody(1, -1), slotOfBody((size_t) nb, 0);
  auto slot = [&](int b) {
    if (b < 0)
    {
      return 0;
    }
    if (slotOfBody[(size_t) b] == 0)
    {
      slotOfBody[(size_t) b] = (int) slotBody.size();
      slotBody.push_back(b);
    }
    return slotOfBody[(size_t) b];
  };
  std::vector<int> task;
  int row = 0, cogRoot = -2;
  for (int t = 0; t < nTasksIn; ++t)
  {
    const RcsbTaskSpec& ts = tasksIn[t];
    auto bodyOk = [&](int b) { return b >= -1 && b < nb; };
    if (ts.kind == RCSB_TASK_JOINT)
    {
      if (ts.effector < 0 || ts.effector >= dof || jntJacobi[ts.effector] < 0)
      {
        rcsb_set_error("task %d: joint %d is out of range or constrained", t, ts.effector);
        return RCSB_EINVAL;
      }
      /* reference joint (TaskJoint.cpp:70-82: it must exist and be unconstrained): its jointIndex + 1
       * in slot 3, the gain's bits in slots 4 / 5 */
      int refJoint = 0, gainBits[2] = {0, 0};
      if (ts.refBdy >= 0)
      {
        if (ts.refBdy >= dof || jntJacobi[ts.refBdy] < 0 || ts.refBdy == ts.effector)
        {
          rcsb_set_error("task %d: reference joint %d is out of range, constrained or the joint itself", t,
                         ts.refBdy);
          return RCSB_EINVAL;
        }
        refJoint = ts.refBdy + 1;
        memcpy(gainBits, &ts.refGain, sizeof(double));
      }
      const int rec[8] = {ts.kind, row, 1, refJoint, gainBits[0], gainBits[1], jntJacobi[ts.effector], ts.effector};
      task.insert(task.end(), rec, rec + 8);
      row += 1;
      continue;
    }
    if (ts.kind >= RCSB_TASK_COGX && ts.kind <= RCSB_TASK_COGZ)
    {
      if (!bodyOk(ts.effector) || !bodyOk(ts.refBdy) || (ts.refFrame >= 0 && ts.refFrame != ts.refBdy))
      {
        rcsb_set_error("task %d: COG tasks take a subtree root (or -1), optionally a reference body, and no "
                       "separate reference frame", t);
        return ts.refFrame >= 0 ? RCSB_EUNSUPPORTED : RCSB_EINVAL;
      }
      if (cogRoot != -2 && cogRoot != ts.effector)
      {
        rcsb_set_error("task %d: all COG tasks of a call must refer to the same subtree", t);
        return RCSB_EUNSUPPORTED;
      }
      cogRoot = ts.effector;
      /* TaskCOM3D with refBody (TaskCOM3D.cpp:101-105, :128-165): slot of the reference body */
      const int rec[8] = {ts.kind, row, 1, 0, slot(ts.refBdy), 0, ts.kind - RCSB_TASK_COGX, 0};
      task.insert(task.end(), rec, rec + 8);
      row += 1;
      continue;
    }
    const bool polar = ts.kind >= RCSB_TASK_POLARX && ts.kind <= RCSB_TASK_POLARZ;
    if (ts.kind != RCSB_TASK_XYZ && ts.kind != RCSB_TASK_ABC && ts.kind != RCSB_TASK_XYZABC && !polar)
    {
      rcsb_set_error("task %d: unknown kind %d", t, ts.kind);
      return RCSB_EINVAL;
    }
    if (!bodyOk(ts.effector) || !bodyOk(ts.refBdy) || !bodyOk(ts.refFrame))
    {
      rcsb_set_error("task %d: body index out of range", t);
      return RCSB_EINVAL;
    }
    if (polar)
    {
      /* TaskPolar2D: the Jacobian is expressed in the effector's own frame (TaskPolar2D.cpp:213), a
       * refFrame attribute plays no role */
      if (ts.effector < 0)
      {
        rcsb_set_error("task %d: a POLAR task needs an effector", t);
        return RCSB_EINVAL;
      }
      const int sE = slot(ts.effector), sR = slot(ts.refBdy);
      const int rec[8] = {ts.kind, row, 2, sE, sR, sE, ts.kind - RCSB_TASK_POLARX, 0};
      task.insert(task.end(), rec, rec + 8);
      row += 2;
      continue;
    }
    const int frameBody = (ts.refFrame < 0 && ts.refBdy >= 0) ? ts.refBdy : ts.refFrame;   /* Task.cpp:160 */
    const int sE = slot(ts.effector), sR = slot(ts.refBdy), sF = slot(frameBody);
    if (ts.kind != RCSB_TASK_ABC)
    {
      const int rec[8] = {RCSB_TASK_XYZ, row, 3, sE, sR, sF, 0, 0};
      task.insert(task.end(), rec, rec + 8);
      row += 3;
    }
    if (ts.kind != RCSB_TASK_XYZ)
    {
      const int rec[8] = {RCSB_TASK_ABC, row, 3, sE, sR, sF, 0, 0};
      task.insert(task.end(), rec, rec + 8);
      row += 3;
    }
  }
  const int nTasks = (int) task.size() / 8, nx = row, nSlotsF = (int) slotBody.size();
  if (nx > 24 || nSlotsF > 24)
  {
    rcsb_set_error("rcsb_ik_tasks_rmr: at most 24 task rows and 23 distinct bodies (%d, %d)", nx, nSlotsF - 1);
    return RCSB_EUNSUPPORTED;
  }

  std::vector<unsigned long long> slotMask((size_t) nSlotsF, 0ull);
  std::vector<int> slotStart((size_t) nb + 1, 0), slotList((size_t) (nSlotsF > 1 ? nSlotsF - 1 : 1), 0);
  for (int sIdx = 1; sIdx < nSlotsF; ++sIdx)
  {
    const int b = slotBody[(size_t) sIdx];
    slotStart[(size_t) b + 1]++;
    for (int j = bodyLastJnt[b]; j >= 0; j = jntPrev[j])
    {
      if (jntJacobi[j] >= 0)
      {
        slotMask[(size_t) sIdx] |= 1ull << jntJacobi[j];
      }
    }
  }
  for (int b = 0; b < nb; ++b)
  {
    slotStart[(size_t) b + 1] += slotStart[(size_t) b];
  }
  {
    std::vector<int> fill(slotStart.begin(), slotStart.end() - 1);
    for (int sIdx = 1; sIdx < nSlotsF; ++sIdx)
    {
      slotList[(size_t) fill[(size_t) slotBody[(size_t) sIdx]]++] = sIdx;
    }
  }
  std::vector<unsigned char> inCog((size_t) nb, 0);
  for (int b = 0; b < nb && cogRoot != -2; ++b)
  {
    bool in = cogRoot == -1;
    for (int a = b; a >= 0 && !in; a = bodyParent[a])
    {
      in = (a == cogRoot);
    }
    inCog[(size_t) b] = in ? 1 : 0;
  }
  std::vector<int> colJoint((size_t) (nJ > 0 ? nJ : 1), 0);
  std::vector<double> invWq((size_t) (nJ > 0 ? nJ : 1), 0.0), jlGain(invWq), q0c(invWq);
  {
    const double* qMin = m->darr(RCSB_A_JNT_QMIN);
    const double* qMax = m->darr(RCSB_A_JNT_QMAX);
    const double* q0 = m->darr(RCSB_A_JNT_Q0);
    const double* wJL = m->darr(RCSB_A_JNT_WJL);
    const double* wM = m->darr(RCSB_A_JNT_WMETRIC);
    for (int j = 0; j < dof; ++j)
    {
      const int c = jntJacobi[j];
      if (c < 0)
      {
        continue;
      }
      colJoint[(size_t) c] = j;
      const double range = qMax[j] - qMin[j];
      invWq[(size_t) c] = wM[j] * range;                              /* RcsGraph_getInvWq, Rcs_graph.c:1087 */
      jlGain[(size_t) c] = range > 0.0 ? wJL[j] / (range * range) : 0.0;   /* Rcs_kinematics.c:1496 */
      q0c[(size_t) c] = q0[j];
    }
  }

  RcsbIkTasksPlanHeader ph;
  memset(&ph, 0, sizeof(ph));
  ph.nTasks = nTasks;
  ph.nx = nx;
  ph.nJ = nJ;
  ph.nSlots = nSlotsF;
  ph.cogRoot = cogRoot;
  int off = align16((int) sizeof(ph));
  auto place = [&](int& field, size_t bytes) {
    field = off;
    off = align16(off + (int) (bytes > 0 ? bytes : 4));
  };
  place(ph.offTask, task.size() * 4);
  place(ph.offBodySlotStart, slotStart.size() * 4);
  place(ph.offBodySlotList, slotList.size() * 4);
  place(ph.offSlotMask, slotMask.size() * 8);
  place(ph.offBodyInCog, inCog.size());
  place(ph.offColJoint, colJoint.size() * 4);
  place(ph.offInvWq, invWq.size() * 8);
  place(ph.offJlGain, jlGain.size() * 8);
  place(ph.offQ0, q0c.size() * 8);
  ph.totalBytes = off;
  std::vector<char> blob((size_t) ph.totalBytes, 0);
  memcpy(blob.data(), &ph, sizeof(ph));
  memcpy(blob.data() + ph.offTask, task.data(), task.size() * 4);
  memcpy(blob.data() + ph.offBodySlotStart, slotStart.data(), slotStart.size() * 4);
  memcpy(blob.data() + ph.offBodySlotList, slotList.data(), slotList.size() * 4);
  memcpy(blob.data() + ph.offSlotMask, slotMask.data(), slotMask.size() * 8);
  memcpy(blob.data() + ph.offBodyInCog, inCog.data(), inCog.size());
  memcpy(blob.data() + ph.offColJoint, colJoint.data(), colJoint.size() * 4);
  memcpy(blob.data() + ph.offInvWq, invWq.data(), invWq.size() * 8);
  memcpy(blob.data() + ph.offJlGain, jlGain.data(), jlGain.size() * 8);
  memcpy(blob.data() + ph.offQ0, q0c.data(), q0c.size() * 8);

  DevicePlan dp;
  RCSB_CUDA(cudaSetDevice(m->device));
  RCSB_CUDA(cudaMalloc(&dp.dev, blob.size()));
  RCSB_CUDA(cudaMemcpy(dp.dev, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  RCSB_CUDA(cudaDeviceSynchronize());
  dp.bytes = ph.totalBytes;
  dp.aux0 = nx;
  dp.aux1 = nSlotsF;
  auto ins = m->ikPlans.emplace(key, dp);
  *out = &ins.first->second;
  return RCSB_OK;
}

static int ikTasksImpl(const char* who, const RcsbModel* cm, int nTasks, const RcsbTaskSpec* tasks, long N,
                       double* q, const double* xDes, double lambda, double alpha, int iters, double* dx,
                       double* dq, double* det, bool constraint, double* nViol, unsigned flags, void* stream)
{
  RcsbModel* m = const_cast<RcsbModel*>(cm);
  if (!m || N < 0 || (N > 0 && (!q || !xDes)) || iters < 0 || nTasks <= 0 || !tasks)
  {
    rcsb_set_error("%s: invalid arguments", who);
    return RCSB_EINVAL;
  }
  const DevicePlan* plan = nullptr;
  int rc = buildIkTasksPlan(m, nTasks, tasks, &plan);
  if (rc != RCSB_OK || N == 0)
  {
    return rc;
  }
  /* rows the constraint solver may add: one per unconstrained joint that is not a Joint task */
  int nxCap = plan->aux0;
  if (constraint)
  {
    const int* jntJacobi = m->iarr(RCSB_A_JNT_JACOBI);
    for (int j = 0; j < m->hdr.dof; ++j)
    {
      bool own = false;
      for (int t = 0; t < nTasks; ++t)
      {
        own = own || (tasks[t].kind == RCSB_TASK_JOINT && tasks[t].effector == j);
      }
      nxCap += (jntJacobi[j] >= 0 && !own) ? 1 : 0;
    }
  }
  auto launch = [&](long n, double* q_, const double* x_, double* dx_, double* dq_, double* det_, double* nv_,
                    cudaStream_t st) -> int {
    rcsb::IkArgs a;
    memset(&a, 0, sizeof(a));
    a.model = m->dev;
    a.plan = plan->dev;
    a.modelBytes = m->hdr.totalBytes;
    a.planBytes = plan->bytes;
    a.N = n;
    a.q = q_;
    a.xDes = x_;
    a.lambda = lambda;
    a.alpha = alpha;
    a.iters = iters;
    a.dx = dx_;
    a.dq = dq_;
    a.det = det_;
    a.constraint = constraint ? 1 : 0;
    a.nViol = nv_;
    RCSB_CUDA(cudaSetDevice(m->device));
    cudaError_t e = rcsb_launch_ik_tasks(&a, m->hdr.nSlots, plan->aux0, nxCap, m->hdr.nJ, m->hdr.dof, plan->aux1,
                                         st);
    if (e == cudaErrorInvalidValue)
    {
      cudaGetLastError();
      rcsb_set_error("%s: problem too large (nx %d, nJ %d, dof %d; limits 24 / 40 / 72)", who, plan->aux0,
                     m->hdr.nJ, m->hdr.dof);
      return RCSB_EUNSUPPORTED;
    }
    RCSB_CUDA(e);
    g_launches.fetch_add(1);
    return RCSB_OK;
  };
  if (!(flags & RCSB_HOST_PTRS))
  {
    return launch(N, q, xDes, dx, dq, det, nViol, static_cast<cudaStream_t>(stream));
  }
  const size_t dof = (size_t) m->hdr.dof, nx = (size_t) plan->aux0;
  std::vector<StagedArray> arr;
  enum { I_QIN, I_X, I_QOUT, I_DX, I_DQ, I_DET, I_NV };
  arr.push_back({q, nullptr, dof, nullptr});
  arr.push_back({xDes, nullptr, nx, nullptr});
  arr.push_back({nullptr, q, dof, nullptr});
  arr.push_back({nullptr, dx, dx ? nx : 0, nullptr});
  arr.push_back({nullptr, dq, dq ? dof : 0, nullptr});
  arr.push_back({nullptr, det, (size_t) (det ? 1 : 0), nullptr});
  arr.push_back({nullptr, nViol, (size_t) (nViol ? 1 : 0), nullptr});
  return runStaged(m, N, arr, {-1, -1, I_QIN, -1, -1, -1, -1}, [&](long n, cudaStream_t st, int) {
    auto dv = [&](int i, const void* host) { return host ? arr[(size_t) i].dev : nullptr; };
    return launch(n, arr[I_QIN].dev, arr[I_X].dev, dv(I_DX, dx), dv(I_DQ, dq), dv(I_DET, det), dv(I_NV, nViol), st);
  });
}

extern "C" int rcsb_ik_tasks_rmr(const RcsbModel* cm, int nTasks, const RcsbTaskSpec* tasks, long N, double* q,
                                 const double* xDes, double lambda, double alpha, int iters, double* dx,
                                 double* dq, double* det, unsigned flags, void* stream)
{
  return ikTasksImpl("rcsb_ik_tasks_rmr", cm, nTasks, tasks, N, q, xDes, lambda, alpha, iters, dx, dq, det, false,
                     nullptr, flags, stream);
}

extern "C" int rcsb_ik_constraint_rmr(const RcsbModel* cm, int nTasks, const RcsbTaskSpec* tasks, long N,
                                      double* q, const double* xDes, double lambda, double alpha, int iters,
                                      double* dx, double* dq, double* det, double* nViolations, unsigned flags,
                                      void* stream)
{
  return ikTasksImpl("rcsb_ik_constraint_rmr", cm, nTasks, tasks, N, q, xDes, lambda, alpha, iters, dx, dq, det,
                     true, nViolations, flags, stream);
}

extern "C" int rcsb_ik_prio_rmr(const RcsbModel* cm, int nTasks, const RcsbTask* tasks, const int* priority,
                                const double* activation, long N, double* q, const double* xDes,
                                double lambda, double alpha, int iters, double* dq1, double* dq2,
                                double* dqNs, double* ok, unsigned flags, void* stream)
{
  RcsbModel* m = const_cast<RcsbModel*>(cm);
  if (!m || N < 0 || (N > 0 && (!q || !xDes)) || iters < 0 || nTasks <= 0 || !tasks || !priority)
  {
    rcsb_set_error("rcsb_ik_prio_rmr: invalid arguments");
    return RCSB_EINVAL;
  }
  /* tasks that take part (level 0 or 1, active), their x_des columns and levels per plan task */
  std::vector<RcsbTask> active;
  IkEx ex;
  int levels[4] = {-1, -1, -1, -1};
  int planTask = 0, col = 0;
  bool all = true;
  for (int t = 0; t < nTasks; ++t)
  {
    const int dim = taskDim(tasks[t].kind);
    const double act = activation ? activation[t] : 1.0;
    if (act > 0.0 && (priority[t] == 0 || priority[t] == 1))
    {
      active.push_back(tasks[t]);
      for (int k = 0; k < dim; ++k)
      {
        ex.xMap.push_back(col + k);
      }
      for (int k = 0; k < dim / 3; ++k, ++planTask)
      {
        if (planTask < 4)
        {
          levels[planTask] = priority[t];
        }
      }
    }
    else
    {
      all = false;
    }
    col += dim;
  }
  ex.nxAll = col;
  if (all)
  {
    ex.xMap.clear();
  }
  if (active.empty() || planTask > 4)
  {
    rcsb_set_error("rcsb_ik_prio_rmr: needs 1 to 4 three-dimensional tasks on the two priority levels (%d)",
                   planTask);
    return active.empty() ? RCSB_EINVAL : RCSB_EUNSUPPORTED;
  }
  const DevicePlan* plan = nullptr;
  int rc = buildIkPlan(m, (int) active.size(), active.data(), &plan);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  if (RcsGraph_countCoupledColumns(m) > 0)
  {
    rcsb_set_error("rcsb_ik_prio_rmr: kinematically coupled unconstrained joints are not supported");
    return RCSB_EUNSUPPORTED;
  }
  if (N == 0)
  {
    return RCSB_OK;
  }

  auto launch = [&](long n, double* dq_, const double* dx_, double* o1, double* o2, double* on, double* ook,
                    cudaStream_t st, int ws) -> int {
    rcsb::IkArgs a;
    memset(&a, 0, sizeof(a));
    RCSB_CUDA(cudaSetDevice(m->device));
    const double* xd = dx_;
    if (!ex.xMap.empty())
    {
      const int nxAct = (int) ex.xMap.size();
      const size_t need = (size_t) n * nxAct * sizeof(double) + 256;
      {
        std::lock_guard<std::recursive_mutex> lock(m->mtx);
        if (m->ikScratchBytes[ws] < need)
        {
          if (m->ikScratch[ws])
          {
            RCSB_CUDA(cudaFree(m->ikScratch[ws]));
            m->ikScratch[ws] = nullptr;
            m->ikScratchBytes[ws] = 0;
          }
          RCSB_CUDA(cudaMalloc(&m->ikScratch[ws], need));
          m->ikScratchBytes[ws] = need;
        }
      }
      int* dMap = reinterpret_cast<int*>(m->ikScratch[ws]);
      double* dX = reinterpret_cast<double*>(m->ikScratch[ws] + 256);
      RCSB_CUDA(cudaMemcpyAsync(dMap, ex.xMap.data(), sizeof(int) * (size_t) nxAct, cudaMemcpyHostToDevice, st));
      long blocks = (n * nxAct + 255) / 256;
      const long cap = (long) m->smCount * 8;
      blocks = blocks < cap ? blocks : cap;
      ikGatherKernel<<<(unsigned int) (blocks > 0 ? blocks : 1), 256, 0, st>>>(dx_, dX, n, ex.nxAll, nxAct, dMap);
      RCSB_CUDA(cudaGetLastError());
      g_launches.fetch_add(1);
      xd = dX;
    }
    a.model = m->dev;
    a.plan = plan->dev;
    a.modelBytes = m->hdr.totalBytes;
    a.planBytes = plan->bytes;
    a.N = n;
    a.q = dq_;
    a.xDes = xd;
    a.lambda = lambda;
    a.alpha = alpha;
    a.iters = iters;
    a.extended = 0;
    a.dq = o1;
    a.dqTs = o2;
    a.dqNs = on;
    a.det = ook;
    memcpy(a.taskLevel, levels, sizeof(levels));
    for (int t = 0; t < 4; ++t)
    {
      a.taskScale[t] = 1.0;
      a.taskAct[t] = 1.0;
    }
    cudaError_t e = rcsb_launch_ik_prio(&a, m->hdr.nSlots, plan->aux0, m->hdr.nJ, m->hdr.dof, st);
    if (e == cudaErrorInvalidValue)
    {
      cudaGetLastError();
      rcsb_set_error("rcsb_ik_prio_rmr: problem too large (nx %d, nJ %d, dof %d; limits 12 / 40 / 72)",
                     plan->aux0, m->hdr.nJ, m->hdr.dof);
      return RCSB_EUNSUPPORTED;
    }
    RCSB_CUDA(e);
    g_launches.fetch_add(1);
    return RCSB_OK;
  };

  if (!(flags & RCSB_HOST_PTRS))
  {
    return launch(N, q, xDes, dq1, dq2, dqNs, ok, static_cast<cudaStream_t>(stream), 0);
  }
  const size_t dof = (size_t) m->hdr.dof;
  std::vector<StagedArray> arr;
  enum { I_QIN, I_X, I_QOUT, I_D1, I_D2, I_DN, I_OK };
  arr.push_back({q, nullptr, dof, nullptr});
  arr.push_back({xDes, nullptr, (size_t) col, nullptr});
  arr.push_back({nullptr, q, dof, nullptr});
  arr.push_back({nullptr, dq1, dq1 ? dof : 0, nullptr});
  arr.push_back({nullptr, dq2, dq2 ? dof : 0, nullptr});
  arr.push_back({nullptr, dqNs, dqNs ? dof : 0, nullptr});
  arr.push_back({nullptr, ok, (size_t) (ok ? 1 : 0), nullptr});
  return runStaged(m, N, arr, {-1, -1, I_QIN, -1, -1, -1, -1}, [&](long n, cudaStream_t st, int slot) {
    auto dv = [&](int i, const void* host) { return host ? arr[(size_t) i].dev : nullptr; };
    return launch(n, arr[I_QIN].dev, arr[I_X].dev, dv(I_D1, dq1), dv(I_D2, dq2), dv(I_DN, dqNs), dv(I_OK, ok),
                  st, slot);
  });
}

extern "C" int rcsb_ik_rmr(const RcsbModel* cm, int nTasks, const RcsbTask* tasks, long N,
                           double* q, const double* xDes, double lambda, double alpha, int iters,
                           double* dx, double* dq, double* det, unsigned flags, void* stream)
{
  return rcsb_ik_rmr_ex(cm, nTasks, tasks, N, q, xDes, lambda, alpha, iters, dx, dq, det, nullptr, flags,
                        stream);
}
