/*
 * C ABI of rcs_b200 (include/rcsb.h), part 1: utilities, models, the forward-kinematics /
 * Jacobian entry points and the host-buffer staging pipe.
 *
 * This layer contains no kinematics: it validates shapes, flattens/uploads models, builds
 * the small per-query plans the kernels read, and enqueues the kernels of rcsb_fk.cu,
 * rcsb_kinetic.cu and rcsb_ik.cu. If there is no CUDA device every compute entry point
 * fails with RCSB_ENODEVICE — there is deliberately no host implementation to fall back to.
 */
#include "rcsb_priv.h"

#include <cctype>
#include <cstdlib>

extern "C" int rcsb_flatten(const RcsGraph* graph, void** blobOut, size_t* bytesOut);

namespace rcsbhost
{

std::atomic<unsigned long long> g_launches(0);

int checkModel(const RcsbModel* m, long N, int qdim, const double* q)
{
  if (!m)
  {
    rcsb_set_error("model is NULL");
    return RCSB_EINVAL;
  }
  if (N < 0 || (N > 0 && !q))
  {
    rcsb_set_error("invalid batch: N = %ld, q = %p", N, (const void*) q);
    return RCSB_EINVAL;
  }
  if (qdim != m->hdr.dof && qdim != m->hdr.nJ)
  {
    /* the reference aborts here: "q has wrong dimensions" (Rcs_graph.c:250) */
    rcsb_set_error("q has wrong dimensions: qdim = %d, dof is %d, nJ is %d", qdim, m->hdr.dof,
                   m->hdr.nJ);
    return RCSB_EINVAL;
  }
  return RCSB_OK;
}

/* ---- staging pipe for host-pointer calls ------------------------------------------------- */

static int pipeReserve(RcsbModel* m, int slot, size_t bytes)
{
  StagePipe& p = m->pipe;
  if (!p.stream[slot])
  {
    RCSB_CUDA(cudaStreamCreateWithFlags(&p.stream[slot], cudaStreamNonBlocking));
  }
  if (p.cap[slot] < bytes)
  {
    if (p.buf[slot])
    {
      RCSB_CUDA(cudaFree(p.buf[slot]));
      p.buf[slot] = nullptr;
      p.cap[slot] = 0;
    }
    RCSB_CUDA(cudaMalloc(&p.buf[slot], bytes));
    p.cap[slot] = bytes;
  }
  return RCSB_OK;
}

static size_t carveBytes(size_t doubles)
{
  return (doubles * sizeof(double) + 255) & ~(size_t) 255;
}

/* States per pipeline stage: at least 65,536 (smaller launches leave SMs idle), more for paths with
 * little traffic per state so that a stage moves about 48 MiB (the 4M-target IK workload ran 64
 * stages of 11 MB on two streams: 210 M solves/s end to end, 252 M with 262,144-state stages on
 * three), at most 1M. A multiple of 384 (block
 * sizes of the kernels). */
static const long kStageChunkMin = 1L << 16;
static long stageChunk(size_t bytesPerState)
{
  const char* env = getenv("RCSB_STAGE_CHUNK");   /* states per stage, for tests of the pipeline */
  if (env && atol(env) > 0)
  {
    return atol(env);
  }
  long c = (long) ((48u << 20) / (bytesPerState ? bytesPerState : 1));
  c = c < kStageChunkMin ? kStageChunkMin : (c > (1L << 20) ? (1L << 20) : c);
  return (c + 383) / 384 * 384;
}

static thread_local const StageTap* t_stageTap = nullptr;

void setStageTap(const StageTap* tap)
{
  t_stageTap = tap;
}

int runStaged(RcsbModel* m, long N, std::vector<StagedArray>& arrays,
              const std::vector<int>& aliasOutToIn,
              const std::function<int(long n, cudaStream_t st, int slot)>& launch)
{
  std::lock_guard<std::recursive_mutex> lock(m->mtx);
  RCSB_CUDA(cudaSetDevice(m->device));

  size_t bytesPerState = 0;
  for (size_t i = 0; i < arrays.size(); ++i)
  {
    bytesPerState += arrays[i].perState * sizeof(double);
  }
  const long want = stageChunk(bytesPerState);
  const long chunk = N < want ? N : want;
  std::vector<size_t> offs(arrays.size(), 0);
  size_t need = 0;
  for (size_t i = 0; i < arrays.size(); ++i)
  {
    const int al = i < aliasOutToIn.size() ? aliasOutToIn[i] : -1;
    if (al >= 0)
    {
      offs[i] = offs[(size_t) al];
    }
    else
    {
      offs[i] = need;
      need += carveBytes((size_t) chunk * arrays[i].perState);
    }
  }

  int nSlots = RCSB_PIPE_SLOTS;
  {
    const char* env = getenv("RCSB_STAGE_SLOTS");
    if (env && atoi(env) >= 1 && atoi(env) <= RCSB_PIPE_SLOTS)
    {
      nSlots = atoi(env);
    }
  }
  const StageTap* tap = t_stageTap;
  int tapArray = -1;
  if (tap && tap->hostOut)
  {
    for (size_t i = 0; i < arrays.size(); ++i)
    {
      tapArray = (arrays[i].hostOut == tap->hostOut && arrays[i].perState > 0) ? (int) i : tapArray;
    }
    if (tapArray >= 0 && tap->begin)
    {
      const int rcTap = tap->begin((int) ((N + chunk - 1) / chunk));
      if (rcTap != RCSB_OK)
      {
        return rcTap;
      }
    }
  }
  int slot = 0, stageIndex = 0;
  for (long s0 = 0; s0 < N; s0 += chunk, slot = (slot + 1) % nSlots, ++stageIndex)
  {
    const long n = (N - s0 < chunk) ? (N - s0) : chunk;
    int rc = pipeReserve(m, slot, need);
    if (rc != RCSB_OK)
    {
      return rc;
    }
    cudaStream_t st = m->pipe.stream[slot];
    for (size_t i = 0; i < arrays.size(); ++i)
    {
      StagedArray& A = arrays[i];
      A.dev = reinterpret_cast<double*>(m->pipe.buf[slot] + offs[i]);
      if (A.hostIn)
      {
        RCSB_CUDA(cudaMemcpyAsync(A.dev, A.hostIn + (size_t) s0 * A.perState,
                                  sizeof(double) * (size_t) n * A.perState,
                                  cudaMemcpyHostToDevice, st));
      }
    }
    rc = launch(n, st, slot);
    if (rc != RCSB_OK)
    {
      return rc;
    }
    if (tapArray >= 0 && tap->stage)
    {
      rc = tap->stage(arrays[(size_t) tapArray].dev, s0, n, stageIndex, st);
      if (rc != RCSB_OK)
      {
        return rc;
      }
    }
    for (size_t i = 0; i < arrays.size(); ++i)
    {
      StagedArray& A = arrays[i];
      if (A.hostOut)
      {
        RCSB_CUDA(cudaMemcpyAsync(A.hostOut + (size_t) s0 * A.perState, A.dev,
                                  sizeof(double) * (size_t) n * A.perState,
                                  cudaMemcpyDeviceToHost, st));
      }
    }
  }
  for (int i = 0; i < RCSB_PIPE_SLOTS; ++i)
  {
    if (m->pipe.stream[i])
    {
      RCSB_CUDA(cudaStreamSynchronize(m->pipe.stream[i]));
    }
  }
  return RCSB_OK;
}

}   // namespace rcsbhost

using namespace rcsbhost;

namespace
{

/* ---- FK plan ---------------------------------------------------------------------------- */

/* compact: the Jacobian records hold the columns of the effectors' own chains only,
 * [nEff][3][recCols] with recCols = the longest chain of the group, column ci = the chain's
 * ci-th Jacobian column in ascending order (rcsb_task_jacobians: its intermediates are read
 * back by combineKernel, structural zeros would be two thirds of the traffic) */
int buildFkPlan(RcsbModel* m, int nSel, const int* bodySel, int nEff, const int* effBody,
                const double* effPt, bool withMass, const DevicePlan** out, bool compact = false)
{
  const int nb = m->hdr.nb, dof = m->hdr.dof, nJ = m->hdr.nJ;

  if (withMass)
  {
    /* the fused path keeps 6 doubles per Jacobian column in registers */
    bool symmetric = true;
    const double* I = m->darr(RCSB_A_BODY_INERTIA);
    for (int b = 0; b < nb; ++b)
    {
      symmetric = symmetric && I[9 * b + 1] == I[9 * b + 3] && I[9 * b + 2] == I[9 * b + 6] &&
                  I[9 * b + 5] == I[9 * b + 7];
    }
    if (nJ < 1 || nJ > 8 || !symmetric)
    {
      rcsb_set_error("fused mass matrix: needs 1 <= nJ <= 8 (is %d) and symmetric inertia tensors", nJ);
      return RCSB_EUNSUPPORTED;
    }
  }

  if (nSel < 0 || nEff < 0 || (nSel > 0 && !bodySel) || (nEff > 0 && !effBody))
  {
    rcsb_set_error("invalid body / effector selection");
    return RCSB_EINVAL;
  }
  for (int i = 0; i < nSel; ++i)
  {
    if (bodySel[i] < 0 || bodySel[i] >= nb)
    {
      rcsb_set_error("bodySel[%d] = %d out of range [0, %d)", i, bodySel[i], nb);
      return RCSB_EINVAL;
    }
  }
  for (int i = 0; i < nEff; ++i)
  {
    if (effBody[i] < 0 || effBody[i] >= nb)
    {
      rcsb_set_error("effBody[%d] = %d out of range [0, %d)", i, effBody[i], nb);
      return RCSB_EINVAL;
    }
  }

  std::string key;
  key.push_back(withMass ? 'M' : (compact ? 'c' : 'k'));
  key.append(reinterpret_cast<const char*>(&nSel), sizeof(int));
  key.append(reinterpret_cast<const char*>(&nEff), sizeof(int));
  if (nSel > 0)
  {
    key.append(reinterpret_cast<const char*>(bodySel), sizeof(int) * nSel);
  }
  if (nEff > 0)
  {
    key.append(reinterpret_cast<const char*>(effBody), sizeof(int) * nEff);
    /* up to RCSB_FK_INLINE_PTS effector points are kernel arguments (FkArgs::effPtArg): the plan
     * depends on the topology of the query only. Larger point sets are part of the plan. */
    if (effPt && nEff > RCSB_FK_INLINE_PTS)
    {
      key.append(reinterpret_cast<const char*>(effPt), sizeof(double) * 3 * nEff);
    }
  }

  std::lock_guard<std::recursive_mutex> lock(m->mtx);
  auto it = m->fkPlans.find(key);
  if (it != m->fkPlans.end())
  {
    *out = &it->second;
    return RCSB_OK;
  }
  /* plans that carry their points (more than RCSB_FK_INLINE_PTS effectors with explicit points)
   * are bounded: beyond 64 of them the oldest ones are released once the device is idle */
  if (effPt && nEff > RCSB_FK_INLINE_PTS)
  {
    m->fkPointPlanKeys.push_back(key);
    if (m->fkPointPlanKeys.size() > 64)
    {
      RCSB_CUDA(cudaSetDevice(m->device));
      RCSB_CUDA(cudaDeviceSynchronize());
      while (m->fkPointPlanKeys.size() > 32)
      {
        auto old = m->fkPlans.find(m->fkPointPlanKeys.front());
        if (old != m->fkPlans.end())
        {
          cudaFree(old->second.dev);
          m->fkPlans.erase(old);
        }
        m->fkPointPlanKeys.erase(m->fkPointPlanKeys.begin());
      }
    }
  }

  const int* jntPrev = m->iarr(RCSB_A_JNT_PREV);
  const int* jntJacobi = m->iarr(RCSB_A_JNT_JACOBI);
  const int* jntFlags = m->iarr(RCSB_A_JNT_FLAGS);
  const int* bodyLastJnt = m->iarr(RCSB_A_BODY_LASTJNT);

  /* output slots: a body selected several times keeps its first slot only */
  const int nOut = (nSel == 0) ? nb : nSel;
  std::vector<int> bodyOut(nb, -1);
  if (nSel == 0)
  {
    for (int b = 0; b < nb; ++b)
    {
      bodyOut[b] = b;
    }
  }
  else
  {
    for (int i = 0; i < nSel; ++i)
    {
      if (bodyOut[bodySel[i]] >= 0)
      {
        rcsb_set_error("body %d selected twice", bodySel[i]);
        return RCSB_EINVAL;
      }
      bodyOut[bodySel[i]] = i;
    }
  }

  /* scratch slots for unconstrained joints on effector chains, ascending with the Jacobian
   * column so that neighbouring columns sit in neighbouring shared-memory rows */
  std::vector<int> jntScr(dof, -1), useCount(dof, 0);
  int nChain = 0;
  for (int e = 0; e < nEff; ++e)
  {
    for (int j = bodyLastJnt[effBody[e]]; j >= 0; j = jntPrev[j])
    {
      if (jntJacobi[j] >= 0)
      {
        useCount[j]++;
      }
    }
  }
  bool inPlace = true;
  for (int j = 0; j < dof; ++j)
  {
    if (useCount[j] > 0 || (withMass && jntJacobi[j] >= 0))
    {
      jntScr[j] = nChain++;   /* with the mass matrix: every column, slot == column */
      inPlace = inPlace && (useCount[j] <= 1);
    }
  }
  if (nChain > 0xffff || nEff > 2047)
  {
    rcsb_set_error("too many chain joints / effectors");
    return RCSB_EUNSUPPORTED;
  }

  int recCols = nJ;
  if (compact)
  {
    recCols = 1;
    for (int e = 0; e < nEff; ++e)
    {
      int nc = 0;
      for (int j = bodyLastJnt[effBody[e]]; j >= 0; j = jntPrev[j])
      {
        nc += jntJacobi[j] >= 0 ? 1 : 0;
      }
      recCols = nc > recCols ? nc : recCols;
    }
  }
  const int rec = nEff * 3 * recCols;
  std::vector<int> jtab(rec > 0 ? rec : 1, -1), jtabR(rec > 0 ? rec : 1, -1);
  std::vector<int> effChainStart(nEff + 1, 0), effChain;
  for (int e = 0; e < nEff; ++e)
  {
    int below = 0;   /* chain columns still to come on the way to the root */
    for (int j = bodyLastJnt[effBody[e]]; j >= 0; j = jntPrev[j])
    {
      below += jntJacobi[j] >= 0 ? 1 : 0;
    }
    for (int j = bodyLastJnt[effBody[e]]; j >= 0; j = jntPrev[j])
    {
      int col = jntJacobi[j];
      if (col < 0)
      {
        continue;
      }
      --below;
      if (compact)
      {
        col = below;   /* rank of the column among the chain's columns, ascending */
      }
      const bool isRot = (jntFlags[j] & RCSB_JF_ROT) != 0;
      effChain.push_back(jntScr[j] | (isRot ? (1 << 16) : 0));
      for (int r = 0; r < 3; ++r)
      {
        const int idx = (e * 3 + r) * recCols + col;
        if (inPlace)
        {
          jtab[idx] = (3 + r) * nChain + jntScr[j];        /* JT column (was: joint origin) */
          jtabR[idx] = isRot ? r * nChain + jntScr[j] : -1; /* JR column = axis */
        }
        else
        {
          jtab[idx] = jntScr[j] | (r << 16) | (isRot ? (1 << 18) : 0) | (e << 20);
        }
      }
    }
    effChainStart[e + 1] = (int) effChain.size();
  }
  if (effChain.empty())
  {
    effChain.push_back(0);
  }

  std::vector<int> effStart(nb + 1, 0), effList(nEff > 0 ? nEff : 1, 0);
  for (int e = 0; e < nEff; ++e)
  {
    effStart[effBody[e] + 1]++;
  }
  for (int b = 0; b < nb; ++b)
  {
    effStart[b + 1] += effStart[b];
  }
  {
    std::vector<int> fill(effStart.begin(), effStart.end() - 1);
    for (int e = 0; e < nEff; ++e)
    {
      effList[fill[effBody[e]]++] = e;
    }
  }

  RcsbFkPlanHeader ph;
  memset(&ph, 0, sizeof(ph));
  ph.nSel = nOut;
  ph.nEff = nEff;
  ph.nChain = nChain;
  ph.nScrRows = (nEff > 0 || withMass) ? 6 * nChain + 3 * nEff : 0;
  std::vector<int> pathMask((size_t) nb, 0);
  if (withMass)
  {
    ph.mass = 1;
    ph.nJ = nJ;
    /* the matrices are staged over the Jacobian scratch once the Jacobians are out */
    ph.mStageRow = 0;
    ph.nScrRows = ph.nScrRows > nJ * nJ ? ph.nScrRows : nJ * nJ;
    ph.invNJ2 = 1.0f / (float) (nJ * nJ);
    for (int j = 0; j < dof; ++j)
    {
      const int c = jntJacobi[j];
      if (c < 0)
      {
        continue;
      }
      if (jntFlags[j] & RCSB_JF_ROT)
      {
        ph.rotMask |= 1 << c;
      }
      for (int a = jntPrev[j]; a >= 0; a = jntPrev[a])
      {
        if (jntJacobi[a] >= 0)
        {
          ph.relMask[c] |= 1 << jntJacobi[a];
        }
      }
    }
    for (int b = 0; b < nb; ++b)
    {
      for (int j = bodyLastJnt[b]; j >= 0; j = jntPrev[j])
      {
        if (jntJacobi[j] >= 0)
        {
          pathMask[(size_t) b] |= 1 << jntJacobi[j];
        }
      }
    }
  }
  std::vector<int> massList;
  if (withMass)
  {
    const double* mass = m->darr(RCSB_A_BODY_MASS);
    bool chain = true;
    for (int k = 0; k < nJ; ++k)
    {
      chain = chain && ph.relMask[k] == (1 << k) - 1;
    }
    for (int cols = nJ; cols >= 1 && chain; --cols)
    {
      for (int b = 0; b < nb; ++b)
      {
        if (mass[b] > 0.0 && pathMask[(size_t) b] == (1 << cols) - 1)
        {
          chain = chain && bodyOut[b] >= 0;
          massList.insert(massList.end(), {b, bodyOut[b], cols, 0});
        }
      }
    }
    ph.massChain = chain ? 1 : 0;
    ph.nMass = chain ? (int) massList.size() / 4 : 0;
  }
  if (massList.empty())
  {
    massList.assign(4, 0);
  }
  if (rec > 0)
  {
    ph.zeroRow = ph.nScrRows++;
  }

  /* pruned walk: selected bodies, effector bodies and their ancestors */
  std::vector<int> visit, jntNext((size_t) (dof > 0 ? dof : 1), -1);
  ph.firstJnt = -1;
  if (nSel > 0 && !withMass)
  {
    const int* parent = m->iarr(RCSB_A_BODY_PARENT);
    const int* jStart = m->iarr(RCSB_A_BODY_JNT_START);
    const int* jCount = m->iarr(RCSB_A_BODY_JNT_COUNT);
    std::vector<char> need((size_t) nb, 0);
    auto mark = [&](int b) {
      for (; b >= 0 && !need[(size_t) b]; b = parent[b])
      {
        need[(size_t) b] = 1;
      }
    };
    for (int i = 0; i < nSel; ++i)
    {
      mark(bodySel[i]);
    }
    for (int e = 0; e < nEff; ++e)
    {
      mark(effBody[e]);
    }
    int nNeed = 0;
    for (int b = 0; b < nb; ++b)
    {
      nNeed += need[(size_t) b];
    }
    if (nNeed < nb)
    {
      /* the same simulation of the depth-first walk as in the flattener (rcsb_model.c),
       * restricted to the sub-tree */
      std::vector<int> kids((size_t) nb, 0), needSlot((size_t) nb, 0), loadKind((size_t) nb, 0),
        slotOf((size_t) nb, -1), depth((size_t) nb, 0);
      for (int b = 0; b < nb; ++b)
      {
        if (need[(size_t) b] && parent[b] >= 0)
        {
          kids[(size_t) parent[b]]++;
        }
      }
      int cur = -1;
      for (int b = 0; b < nb; ++b)
      {
        if (!need[(size_t) b])
        {
          continue;
        }
        const int pb = parent[b];
        if (pb < 0)
        {
          loadKind[(size_t) b] = RCSB_LOAD_IDENT;
        }
        else if (cur == pb)
        {
          loadKind[(size_t) b] = RCSB_LOAD_CONT;
        }
        else
        {
          loadKind[(size_t) b] = 0;
          needSlot[(size_t) pb] = 1;
        }
        cur = (kids[(size_t) b] == 0 && pb >= 0) ? pb : b;
      }
      int slots = 0, lastJ = -1;
      for (int b = 0; b < nb; ++b)
      {
        if (!need[(size_t) b])
        {
          continue;
        }
        const int pb = parent[b];
        const int inherited = pb >= 0 ? depth[(size_t) pb] : 0;
        depth[(size_t) b] = inherited;
        if (needSlot[(size_t) b])
        {
          slotOf[(size_t) b] = inherited;
          depth[(size_t) b] = inherited + 1;
          slots = depth[(size_t) b] > slots ? depth[(size_t) b] : slots;
        }
        visit.insert(visit.end(), {b, loadKind[(size_t) b] == 0 ? slotOf[(size_t) pb] : loadKind[(size_t) b],
                                   slotOf[(size_t) b], (kids[(size_t) b] == 0 && pb >= 0) ? 1 : 0});
        for (int j = jStart[b]; j < jStart[b] + jCount[b]; ++j)
        {
          if (lastJ >= 0)
          {
            jntNext[(size_t) lastJ] = j;
          }
          else
          {
            ph.firstJnt = j;
          }
          lastJ = j;
        }
      }
      if (slots <= RCSB_MAX_FRAME_SLOTS)
      {
        ph.nVisit = (int) visit.size() / 4;
        ph.prunedSlots = slots;
      }
    }
  }
  if (visit.empty())
  {
    visit.assign(4, 0);
  }

  /* ---- serial chain of rotations: tables of the straight-line walk (RcsbFkPlanHeader::chainNc) ---- */
  std::vector<double> chainG, chainX, chainMass;
  std::vector<int> chainOutStart, chainOut, chainJoint;
  if (!compact && nJ >= 1 && nJ <= 8 && dof == nJ && m->hdr.nCpl == 0 && (inPlace || nEff == 0))
  {
    const int* parent = m->iarr(RCSB_A_BODY_PARENT);
    const int* jStart = m->iarr(RCSB_A_BODY_JNT_START);
    const int* jCount = m->iarr(RCSB_A_BODY_JNT_COUNT);
    const int* bodyFlags = m->iarr(RCSB_A_BODY_FLAGS);
    const int* jntType = m->iarr(RCSB_A_JNT_TYPE);
    const double* ajp = m->darr(RCSB_A_JNT_AJP);
    const double* abp = m->darr(RCSB_A_BODY_ABP);
    const double* mass = m->darr(RCSB_A_BODY_MASS);
    const double* com = m->darr(RCSB_A_BODY_COM);
    const double* inertia = m->darr(RCSB_A_BODY_INERTIA);
    struct Xf
    {
      double v[12];
    };
    const Xf ident = {{0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1}};
    /* applying A, then B: {A.org + A.rot^T B.org, B.rot A.rot} */
    auto compose = [](const Xf& A, const double* B) {
      Xf Y;
      for (int i = 0; i < 3; ++i)
      {
        Y.v[i] = A.v[i] + (A.v[3 + i] * B[0] + A.v[6 + i] * B[1] + A.v[9 + i] * B[2]);
        for (int k = 0; k < 3; ++k)
        {
          Y.v[3 + 3 * i + k] = B[3 + 3 * i] * A.v[3 + k] + B[3 + 3 * i + 1] * A.v[6 + k] + B[3 + 3 * i + 2] * A.v[9 + k];
        }
      }
      return Y;
    };
    /* rows / columns permuted for bases kept with shifted rows: in = shift of the base the transform
     * starts from, out = shift of the frame it leads to */
    auto permute = [](const Xf& A, int in, int out) {
      Xf Y;
      for (int i = 0; i < 3; ++i)
      {
        Y.v[i] = A.v[(i + in) % 3];
        for (int k = 0; k < 3; ++k)
        {
          Y.v[3 + 3 * i + k] = A.v[3 + 3 * ((i + out) % 3) + (k + in) % 3];
        }
      }
      return Y;
    };
    std::vector<int> lvl((size_t) nb, 0), shiftOf((size_t) nJ + 1, 0);   /* shiftOf[L]: shift of the base of level L */
    std::vector<Xf> X((size_t) nb, ident), G((size_t) nJ, ident);
    bool ok = true;
    chainJoint.assign((size_t) nJ, -1);
    for (int b = 0; b < nb && ok; ++b)
    {
      int L = parent[b] >= 0 ? lvl[(size_t) parent[b]] : 0;
      Xf T = parent[b] >= 0 ? X[(size_t) parent[b]] : ident;
      for (int j = jStart[b]; j < jStart[b] + jCount[b] && ok; ++j)
      {
        if (jntFlags[j] & RCSB_JF_HAS_AJP)
        {
          T = compose(T, ajp + 12 * j);
        }
        ok = (jntFlags[j] & RCSB_JF_ROT) && jntJacobi[j] == L && L < nJ;
        if (ok)
        {
          G[(size_t) L] = T;
          chainJoint[(size_t) L] = j;
          shiftOf[(size_t) L + 1] = (jntType[j] % 3 + 1) % 3;   /* axis -> row 2 */
          ++L;
          T = ident;
        }
      }
      if (bodyFlags[b] & RCSB_BF_HAS_ABP)
      {
        T = compose(T, abp + 12 * b);
      }
      lvl[(size_t) b] = L;
      X[(size_t) b] = T;
    }
    for (int c = 0; c < nJ; ++c)
    {
      ok = ok && chainJoint[(size_t) c] >= 0;
    }
    if (ok)
    {
      for (int c = 0; c < nJ; ++c)
      {
        const Xf Gp = permute(G[(size_t) c], shiftOf[(size_t) c], shiftOf[(size_t) c + 1]);
        chainG.insert(chainG.end(), Gp.v, Gp.v + 12);
      }
      chainMass.assign((size_t) (nJ + 1) * 10, 0.0);
      chainOutStart.assign(1, 0);
      for (int L = 0; L <= nJ; ++L)
      {
        const int sh = shiftOf[(size_t) L];
        double* cm = chainMass.data() + 10 * (size_t) L;
        for (int b = 0; b < nb; ++b)
        {
          if (lvl[(size_t) b] != L)
          {
            continue;
          }
          const Xf Xp = permute(X[(size_t) b], sh, 0);
          bool rotIdent = true, noTrans = true;
          for (int i = 0; i < 3; ++i)
          {
            noTrans = noTrans && Xp.v[i] == 0.0;
            for (int k = 0; k < 3; ++k)
            {
              rotIdent = rotIdent && Xp.v[3 + 3 * i + k] == (i == k ? 1.0 : 0.0);
            }
          }
          const bool needed = bodyOut[(size_t) b] >= 0 || effStart[(size_t) b + 1] > effStart[(size_t) b];
          const int ent[4] = {b, bodyOut[(size_t) b],
                              (rotIdent ? RCSB_REL_ROT_IDENT : 0) | (noTrans ? RCSB_REL_NO_TRANS : 0) | (needed ? 1 : 0), 0};
          chainOut.insert(chainOut.end(), ent, ent + 4);
          chainX.insert(chainX.end(), Xp.v, Xp.v + 12);
          if (mass[b] > 0.0 && L > 0)
          {
            /* COM and inertia in the (shifted) coordinates of the level's base, about its origin */
            const double* Xr = Xp.v + 3;
            double c[3], Ic[9], t[9];
            for (int i = 0; i < 3; ++i)
            {
              c[i] = Xp.v[i] + (Xr[i] * com[3 * b] + Xr[3 + i] * com[3 * b + 1] + Xr[6 + i] * com[3 * b + 2]);
            }
            for (int i = 0; i < 3; ++i)
            {
              for (int k = 0; k < 3; ++k)
              {
                t[3 * i + k] = inertia[9 * b + 3 * i] * Xr[k] + inertia[9 * b + 3 * i + 1] * Xr[3 + k] +
                               inertia[9 * b + 3 * i + 2] * Xr[6 + k];
              }
            }
            for (int i = 0; i < 3; ++i)
            {
              for (int k = 0; k < 3; ++k)
              {
                Ic[3 * i + k] = Xr[i] * t[k] + Xr[3 + i] * t[3 + k] + Xr[6 + i] * t[6 + k];
              }
            }
            const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
            cm[0] += mass[b];
            cm[1] += mass[b] * c[0];
            cm[2] += mass[b] * c[1];
            cm[3] += mass[b] * c[2];
            cm[4] += Ic[0] + mass[b] * (cc - c[0] * c[0]);
            cm[5] += Ic[1] - mass[b] * c[0] * c[1];
            cm[6] += Ic[2] - mass[b] * c[0] * c[2];
            cm[7] += Ic[4] + mass[b] * (cc - c[1] * c[1]);
            cm[8] += Ic[5] - mass[b] * c[1] * c[2];
            cm[9] += Ic[8] + mass[b] * (cc - c[2] * c[2]);
          }
        }
        chainOutStart.push_back((int) chainOut.size() / 4);
      }
      ph.chainNc = nJ;
    }
  }
  ph.jacRecord = rec;
  ph.inPlace = inPlace ? 1 : 0;
  {
    /* output slots ascending along the walk? */
    int last = -1;
    bool mono = true;
    for (int b = 0; b < nb; ++b)
    {
      if (bodyOut[b] >= 0)
      {
        mono = mono && bodyOut[b] == last + 1;
        last = bodyOut[b];
      }
    }
    ph.velStream = mono ? 1 : 0;
  }
  ph.invJacRecord = rec > 0 ? 1.0f / (float) rec : 0.0f;
  int off = align16((int) sizeof(ph));
  ph.offBodyOut = off;
  off = align16(off + nb * 4);
  ph.offJntScr = off;
  off = align16(off + dof * 4);
  ph.offBodyEffStart = off;
  off = align16(off + (nb + 1) * 4);
  ph.offBodyEffList = off;
  off = align16(off + (int) effList.size() * 4);
  ph.offEffPt = off;
  off = align16(off + (nEff > 0 ? nEff : 1) * 24);
  ph.offJTab = off;
  off = align16(off + (int) jtab.size() * 4);
  ph.offJTabR = off;
  off = align16(off + (int) jtabR.size() * 4);
  ph.offEffChainStart = off;
  off = align16(off + (nEff + 1) * 4);
  ph.offEffChain = off;
  off = align16(off + (int) effChain.size() * 4);
  ph.offBodyPathMask = off;
  off = align16(off + nb * 4);
  ph.offMassList = off;
  off = align16(off + (int) massList.size() * 4);
  ph.offVisit = off;
  off = align16(off + (int) visit.size() * 4);
  ph.offJntNext = off;
  off = align16(off + (int) jntNext.size() * 4);
  if (ph.chainNc > 0)
  {
    ph.offChainG = off;
    off = align16(off + (int) chainG.size() * 8);
    ph.offChainX = off;
    off = align16(off + (int) chainX.size() * 8);
    ph.offChainOutStart = off;
    off = align16(off + (int) chainOutStart.size() * 4);
    ph.offChainOut = off;
    off = align16(off + (int) chainOut.size() * 4);
    ph.offChainMass = off;
    off = align16(off + (int) chainMass.size() * 8);
    ph.offChainJoint = off;
    off = align16(off + (int) chainJoint.size() * 4);
  }
  ph.totalBytes = off;

  std::vector<char> blob(ph.totalBytes, 0);
  memcpy(blob.data(), &ph, sizeof(ph));
  memcpy(blob.data() + ph.offBodyOut, bodyOut.data(), nb * 4);
  memcpy(blob.data() + ph.offJntScr, jntScr.data(), dof * 4);
  memcpy(blob.data() + ph.offBodyEffStart, effStart.data(), (nb + 1) * 4);
  memcpy(blob.data() + ph.offBodyEffList, effList.data(), effList.size() * 4);
  if (nEff > RCSB_FK_INLINE_PTS && effPt)
  {
    memcpy(blob.data() + ph.offEffPt, effPt, sizeof(double) * 3 * nEff);
  }
  memcpy(blob.data() + ph.offJTab, jtab.data(), jtab.size() * 4);
  memcpy(blob.data() + ph.offJTabR, jtabR.data(), jtabR.size() * 4);
  memcpy(blob.data() + ph.offEffChainStart, effChainStart.data(), (nEff + 1) * 4);
  memcpy(blob.data() + ph.offEffChain, effChain.data(), effChain.size() * 4);
  memcpy(blob.data() + ph.offBodyPathMask, pathMask.data(), (size_t) nb * 4);
  memcpy(blob.data() + ph.offMassList, massList.data(), massList.size() * 4);
  memcpy(blob.data() + ph.offVisit, visit.data(), visit.size() * 4);
  memcpy(blob.data() + ph.offJntNext, jntNext.data(), jntNext.size() * 4);
  if (ph.chainNc > 0)
  {
    memcpy(blob.data() + ph.offChainG, chainG.data(), chainG.size() * 8);
    memcpy(blob.data() + ph.offChainX, chainX.data(), chainX.size() * 8);
    memcpy(blob.data() + ph.offChainOutStart, chainOutStart.data(), chainOutStart.size() * 4);
    memcpy(blob.data() + ph.offChainOut, chainOut.data(), chainOut.size() * 4);
    memcpy(blob.data() + ph.offChainMass, chainMass.data(), chainMass.size() * 8);
    memcpy(blob.data() + ph.offChainJoint, chainJoint.data(), chainJoint.size() * 4);
  }

  DevicePlan dp;
  RCSB_CUDA(cudaSetDevice(m->device));
  RCSB_CUDA(cudaMalloc(&dp.dev, blob.size()));
  RCSB_CUDA(cudaMemcpy(dp.dev, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  /* pageable source: the call may return before the DMA has landed, and the first launch can
   * follow on a non-blocking stream that is not ordered against it */
  RCSB_CUDA(cudaDeviceSynchronize());
  dp.bytes = ph.totalBytes;
  dp.nScrRows = ph.nScrRows;
  dp.nSel = nOut;
  dp.nEff = nEff;
  dp.aux0 = compact ? recCols : ph.massChain;   /* compact plans never carry the mass tables */
  dp.aux1 = ph.nVisit > 0 ? ph.prunedSlots + 1 : 0;   /* 0: no pruned walk */
  dp.aux2 = ph.chainNc;                                /* > 0: the straight-line chain walk applies */

  auto ins = m->fkPlans.emplace(key, dp);
  *out = &ins.first->second;
  return RCSB_OK;
}

/* FK (+ Jacobians) on device pointers */
int fkDevice(RcsbModel* m, const DevicePlan* plan, long N, int qdim, const double* q,
             const double* qd, double* A_BI, double* A_JI, double* xdot, double* omega,
             double* qOut, double* JT, double* JR, double* qdOut, bool noCoupling, double* M,
             cudaStream_t stream, const double* effPt = nullptr)
{
  rcsb::FkArgs a;
  a.effPtInline = 0;
  if (plan->nEff > 0 && plan->nEff <= RCSB_FK_INLINE_PTS)
  {
    a.effPtInline = 1;
    memset(a.effPtArg, 0, sizeof(a.effPtArg));
    if (effPt)
    {
      memcpy(a.effPtArg, effPt, sizeof(double) * 3 * (size_t) plan->nEff);
    }
  }
  a.model = m->dev;
  a.plan = plan->dev;
  a.modelBytes = m->hdr.totalBytes;
  a.planBytes = plan->bytes;
  a.N = N;
  a.qdim = qdim;
  a.q = q;
  a.qd = qd;
  a.A_BI = A_BI;
  a.A_JI = A_JI;
  a.xdot = xdot;
  a.omega = omega;
  a.qOut = qOut;
  a.JT = JT;
  a.JR = JR;
  a.qdOut = qdOut;
  a.M = M;
  a.noCoupling = noCoupling ? 1 : 0;
  /* composite (tip-to-root) mass matrix when the plan allows it and the frames are written */
  a.massMode = M ? ((plan->aux0 && A_BI) ? 2 : 1) : 0;   /* (plans with M are never compact) */
  /* per-joint outputs and velocities need every body: the pruned walk serves frames and
   * Jacobians of a body selection only */
  a.pruned = (plan->aux1 > 0 && !A_JI && !qOut && !qd && !qdOut && !M) ? 1 : 0;
  /* serial chains of rotations: the straight-line kernel serves frames, Jacobians and M */
  {
    const char* noChain = getenv("RCSB_FK_GENERAL");   /* force the interpreting kernel (tests) */
    const bool chainOk = plan->aux2 > 0 && !(noChain && noChain[0] == '1') && !A_JI && !qOut && !qdOut &&
                         qdim == m->hdr.dof;
    /* with velocities: frames, x_dot, omega only (rcsb_fk); without: frames, Jacobians, M */
    a.chainNc = chainOk && (qd ? (plan->nEff == 0 && !M && !JT && !JR) : true) ? plan->aux2 : 0;
    a.velRecDoubles = 3 * plan->nSel;
  }
  a.scrDoubles = plan->nScrRows * (RCSB_FK_BLOCK(M != nullptr) + 1);
  RCSB_CUDA(cudaSetDevice(m->device));
  if (rcsbhost::fkSharedBytes(m, plan->bytes, plan->nScrRows, M != nullptr, qd != nullptr) > RCSB_SMEM_LIMIT)
  {
    rcsb_set_error("forward kinematics: the effectors' joint chains need %zu bytes of shared memory "
                   "(limit %d); ask for fewer effectors per call",
                   rcsbhost::fkSharedBytes(m, plan->bytes, plan->nScrRows, M != nullptr, qd != nullptr),
                   RCSB_SMEM_LIMIT);
    return RCSB_EUNSUPPORTED;
  }
  RCSB_CUDA(rcsb_launch_fk(&a, a.pruned ? plan->aux1 - 1 : m->hdr.nSlots, plan->nScrRows, stream));
  if (N > 0)
  {
    g_launches.fetch_add(1);
  }
  return RCSB_OK;
}

}   // namespace

/* ------------------------------------------------------------------------------------------ */
/* utilities                                                                                   */
/* ------------------------------------------------------------------------------------------ */

extern "C" const char* rcsb_last_error(void)
{
  return rcsb_error_string();
}

extern "C" int rcsb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return n;
}

/* PCI address of a device as sysfs spells it (rcsb_hostbind.c) */
extern "C" int rcsb_device_pci_bus_id(int device, char* busId, int len)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n ||
      cudaDeviceGetPCIBusId(busId, len, device) != cudaSuccess)
  {
    cudaGetLastError();
    return -1;
  }
  for (char* c = busId; *c; ++c)
  {
    *c = (char) tolower(*c);
  }
  return 0;
}

extern "C" void* rcsb_host_alloc(size_t bytes)
{
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess)
  {
    cudaGetLastError();
    rcsb_set_error("cudaHostAlloc of %zu bytes failed", bytes);
    return nullptr;
  }
  return p;
}

extern "C" void rcsb_host_free(void* p)
{
  if (p)
  {
    cudaFreeHost(p);
  }
}

extern "C" unsigned long long rcsb_kernel_launch_count(void)
{
  return g_launches.load();
}

/* ------------------------------------------------------------------------------------------ */
/* model                                                                                       */
/* ------------------------------------------------------------------------------------------ */

extern "C" int rcsb_model_create(const RcsGraph* graph, int device, RcsbModel** out)
{
  if (!graph || !out)
  {
    rcsb_set_error("rcsb_model_create: NULL argument");
    return RCSB_EINVAL;
  }
  *out = nullptr;

  int nDev = 0;
  cudaError_t e = cudaGetDeviceCount(&nDev);
  if (e != cudaSuccess || nDev == 0)
  {
    cudaGetLastError();
    rcsb_set_error("no CUDA device available (%s); rcs_b200 has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return RCSB_ENODEVICE;
  }
  if (device < 0 || device >= nDev)
  {
    rcsb_set_error("device %d out of range [0, %d)", device, nDev);
    return RCSB_EINVAL;
  }

  void* blob = nullptr;
  size_t bytes = 0;
  if (rcsb_flatten(graph, &blob, &bytes) != 0)
  {
    return RCSB_EINVAL;
  }

  RcsbModel* m = new RcsbModel;
  m->device = device;
  m->host.assign(static_cast<char*>(blob), static_cast<char*>(blob) + bytes);
  free(blob);
  memcpy(&m->hdr, m->host.data(), sizeof(RcsbFlatHeader));

  if (m->hdr.nSlots > RCSB_MAX_FRAME_SLOTS)
  {
    rcsb_set_error("tree needs %d parent-frame slots, at most %d supported", m->hdr.nSlots,
                   RCSB_MAX_FRAME_SLOTS);
    delete m;
    return RCSB_EUNSUPPORTED;
  }

  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess)
  {
    ce = cudaMalloc(&m->dev, bytes);
  }
  if (ce == cudaSuccess)
  {
    ce = cudaMemcpy(m->dev, m->host.data(), bytes, cudaMemcpyHostToDevice);
  }
  if (ce == cudaSuccess)
  {
    ce = cudaDeviceSynchronize();   /* pageable source: see buildFkPlan */
  }
  if (ce == cudaSuccess)
  {
    ce = cudaDeviceGetAttribute(&m->smCount, cudaDevAttrMultiProcessorCount, device);
  }
  if (ce != cudaSuccess)
  {
    delete m;
    return cudaFail(ce, "model upload");
  }

  *out = m;
  return RCSB_OK;
}

extern "C" void rcsb_model_destroy(RcsbModel* m)
{
  if (!m)
  {
    return;
  }
  cudaSetDevice(m->device);
  for (auto& kv : m->fkPlans)
  {
    cudaFree(kv.second.dev);
  }
  for (auto& kv : m->ikPlans)
  {
    cudaFree(kv.second.dev);
  }
  cudaFree(m->ktPlan[0].dev);
  cudaFree(m->ktPlan[1].dev);
  for (auto& kv : m->wbPlans)
  {
    cudaFree(kv.second.dev);
  }
  for (int i = 0; i < RCSB_PIPE_SLOTS; ++i)
  {
    cudaFree(m->ktScratch[i]);
    cudaFree(m->ikScratch[i]);
  }
  for (int i = 0; i < RCSB_PIPE_SLOTS; ++i)
  {
    if (m->pipe.buf[i])
    {
      cudaFree(m->pipe.buf[i]);
    }
    if (m->pipe.stream[i])
    {
      cudaStreamDestroy(m->pipe.stream[i]);
    }
  }
  cudaFree(m->dev);
  delete m;
}

extern "C" int rcsb_model_dof(const RcsbModel* m)
{
  return m ? m->hdr.dof : -1;
}

extern "C" int rcsb_model_nj(const RcsbModel* m)
{
  return m ? m->hdr.nJ : -1;
}

extern "C" int rcsb_model_nbodies(const RcsbModel* m)
{
  return m ? m->hdr.nb : -1;
}

extern "C" int rcsb_model_device(const RcsbModel* m)
{
  return m ? m->device : -1;
}

/* ------------------------------------------------------------------------------------------ */
/* forward kinematics + Jacobians                                                              */
/* ------------------------------------------------------------------------------------------ */

int rcsbhost::fkCommon(const RcsbModel* cm, long N, int qdim, const double* q, const double* qd,
                       int nSel, const int* bodySel, double* A_BI, double* A_JI, double* xdot,
                       double* omega, double* qOut, int nEff, const int* effBody,
                       const double* effPt, double* JT, double* JR, double* qdOut,
                       bool noCoupling, unsigned flags, void* stream, double* M)
{
  RcsbModel* m = const_cast<RcsbModel*>(cm);
  int rc = checkModel(m, N, qdim, q);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  if ((xdot || omega || qdOut) && !qd)
  {
    rcsb_set_error("xdot / omega / q_dot output requested without q_dot");
    return RCSB_EINVAL;
  }
  if ((JT || JR) && nEff <= 0)
  {
    rcsb_set_error("Jacobians requested without effectors");
    return RCSB_EINVAL;
  }

  const DevicePlan* plan = nullptr;
  if (M && qd)
  {
    rcsb_set_error("the fused mass matrix is computed without velocities");
    return RCSB_EINVAL;
  }
  rc = buildFkPlan(m, nSel, bodySel, (JT || JR) ? nEff : 0, effBody, effPt, M != nullptr, &plan);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  if (N == 0)
  {
    return RCSB_OK;
  }

  if (!(flags & RCSB_HOST_PTRS))
  {
    return fkDevice(m, plan, N, qdim, q, qd, A_BI, A_JI, xdot, omega, qOut, JT, JR, qdOut,
                    noCoupling, M, static_cast<cudaStream_t>(stream), effPt);
  }

  const size_t dof = (size_t) m->hdr.dof, nOut = (size_t) plan->nSel;
  const size_t recJ = (size_t) plan->nEff * 3 * m->hdr.nJ;
  std::vector<StagedArray> arr;
  enum { I_Q, I_QD, I_A, I_AJ, I_XD, I_OM, I_QO, I_JT, I_JR, I_QDO, I_M };
  arr.push_back({q, nullptr, (size_t) qdim, nullptr});
  arr.push_back({qd, nullptr, qd ? (size_t) qdim : 0, nullptr});
  arr.push_back({nullptr, A_BI, A_BI ? nOut * 12 : 0, nullptr});
  arr.push_back({nullptr, A_JI, A_JI ? dof * 12 : 0, nullptr});
  arr.push_back({nullptr, xdot, xdot ? nOut * 3 : 0, nullptr});
  arr.push_back({nullptr, omega, omega ? nOut * 3 : 0, nullptr});
  arr.push_back({nullptr, qOut, qOut ? dof : 0, nullptr});
  arr.push_back({nullptr, JT, JT ? recJ : 0, nullptr});
  arr.push_back({nullptr, JR, JR ? recJ : 0, nullptr});
  arr.push_back({nullptr, qdOut, qdOut ? dof : 0, nullptr});
  arr.push_back({nullptr, M, M ? (size_t) m->hdr.nJ * m->hdr.nJ : 0, nullptr});

  return runStaged(m, N, arr, {}, [&](long n, cudaStream_t st, int) {
    auto dv = [&](int i, const void* host) { return host ? arr[(size_t) i].dev : nullptr; };
    return fkDevice(m, plan, n, qdim, dv(I_Q, q), dv(I_QD, qd), dv(I_A, A_BI), dv(I_AJ, A_JI),
                    dv(I_XD, xdot), dv(I_OM, omega), dv(I_QO, qOut), dv(I_JT, JT), dv(I_JR, JR),
                    dv(I_QDO, qdOut), noCoupling, dv(I_M, M), st, effPt);
  });
}

/* Frames + compact world Jacobians of a few bodies, device pointers (rcsb_task_jacobians). */
int rcsbhost::fkCompactDevice(RcsbModel* m, long n, int qdim, const double* q, int G, const int* bodies,
                              double* A, double* JT, double* JR, cudaStream_t stream, int* recCols)
{
  const DevicePlan* plan = nullptr;
  const int rc = buildFkPlan(m, G, bodies, G, bodies, nullptr, false, &plan, true);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  *recCols = plan->aux0;
  if (n == 0)
  {
    return RCSB_OK;
  }
  return fkDevice(m, plan, n, qdim, q, nullptr, A, nullptr, nullptr, nullptr, nullptr, JT, JR, nullptr,
                  false, nullptr, stream);
}

extern "C" int rcsb_fk(const RcsbModel* model, long N, int qdim, const double* q,
                       const double* qd, int nSel, const int* bodySel, double* A_BI,
                       double* A_JI, double* xdot, double* omega, double* qOut, unsigned flags,
                       void* stream)
{
  return fkCommon(model, N, qdim, q, qd, nSel, bodySel, A_BI, A_JI, xdot, omega, qOut, 0,
                  nullptr, nullptr, nullptr, nullptr, nullptr, false, flags, stream);
}

extern "C" int rcsb_fk_jacobians(const RcsbModel* model, long N, int qdim, const double* q,
                                 int nSel, const int* bodySel, double* A_BI, int nEff,
                                 const int* effBody, const double* effPt, double* JT,
                                 double* JR, unsigned flags, void* stream)
{
  return fkCommon(model, N, qdim, q, nullptr, nSel, bodySel, A_BI, nullptr, nullptr, nullptr,
                  nullptr, nEff, effBody, effPt, JT, JR, nullptr, false, flags, stream);
}

/* FK + Jacobians + mass matrix in one launch (models with nJ <= 8 and symmetric inertia
 * tensors); other models take rcsb_fk_jacobians + rcsb_kinetic_terms. */
extern "C" int rcsb_fk_jacobians_mass(const RcsbModel* model, long N, int qdim, const double* q,
                                      int nSel, const int* bodySel, double* A_BI, int nEff,
                                      const int* effBody, const double* effPt, double* JT,
                                      double* JR, double* M, unsigned flags, void* stream)
{
  if (!M)
  {
    return rcsb_fk_jacobians(model, N, qdim, q, nSel, bodySel, A_BI, nEff, effBody, effPt, JT, JR,
                             flags, stream);
  }
  int rc = fkCommon(model, N, qdim, q, nullptr, nSel, bodySel, A_BI, nullptr, nullptr, nullptr,
                    nullptr, nEff, effBody, effPt, JT, JR, nullptr, false, flags, stream, M);
  if (rc != RCSB_EUNSUPPORTED)
  {
    return rc;
  }
  rc = rcsb_fk_jacobians(model, N, qdim, q, nSel, bodySel, A_BI, nEff, effBody, effPt, JT, JR, flags,
                         stream);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  return rcsb_kinetic_terms(model, N, qdim, q, nullptr, M, nullptr, nullptr, nullptr, flags, stream);
}
