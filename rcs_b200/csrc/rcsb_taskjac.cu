/*
 * Task Jacobians relative to a (possibly articulated) reference body / frame, batched:
 *   RcsGraph_3dPosJacobian    src/RcsCore/Rcs_kinematics.c:566-615
 *       world:      J = JT_2, rotated by A_3I if a refFrame is given
 *       relative:   J = A_xI (JT_2 - JT_1 + r_12 x JR_3),  x = refFrame if given and != refBdy,
 *                   else refBdy
 *   RcsGraph_3dOmegaJacobian  src/RcsCore/Rcs_kinematics.c:789-822
 *       world:      J = JR_2
 *       fixed ref:  J = A_1I JR_2          (refFrame == refBdy, not articulated)
 *       relative:   J = A_3I (JR_2 - JR_1)
 * (2 = effector, 1 = refBdy, 3 = refFrame.) The frames and world Jacobians of the bodies
 * involved come from fkKernel (rcsb_fk.cu) in passes of 131,072 states (compact records of
 * the bodies' own chain columns); combineKernel turns them into the task rows, one thread per
 * (state, task, column), consecutive threads on consecutive columns.
 */
#include <cstdlib>
#include "rcsb_priv.h"

namespace
{

struct CombineTask
{
  int kind;       /* RCSB_JAC_POS / RCSB_JAC_OMEGA */
  int mode;       /* 0 world (optionally rotated), 1 relative */
  int u2, u1, u3; /* slots of effector / refBdy / rotation-Jacobian body in the pass buffers, -1 */
  int uA;         /* slot of the body whose rotation is applied, -1 = none */
};

/* where the pass buffers hold body u: fkKernel runs once per group of bodies (as many as its
 * shared-memory scratch takes), each group state-major */
struct SlotInfo
{
  long offA, strideA;   /* doubles: frame of state s at tmp[offA + s strideA] */
  long offJ, strideJ;   /* JT at tmp[offJ + s strideJ], JR at tmp[offJ + jrShift + s strideJ] */
  long jrShift;
  int rowStride;        /* doubles between the rows of a 3 x rowStride Jacobian record: the records
                           are compact, they hold the columns of the body's own chain only */
  int pad;
};

struct CombineArgs
{
  const double* tmp;
  const SlotInfo* slots;
  double* out;        /* n x nTasks x 3 x nJ */
  const CombineTask* tasks;
  const int* colMap;  /* [bodies][nJ]: index of Jacobian column k in the body's compact record, -1 = zero */
  long n;
  int nTasks, nJ;
};

__global__ void __launch_bounds__(256) combineKernel(const CombineArgs a)
{
  const long total = a.n * a.nTasks * a.nJ;
  const long perState = (long) a.nTasks * a.nJ;
  for (long e = (long) blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long) gridDim.x * blockDim.x)
  {
    const long s = e / perState;
    const int rem = (int) (e - s * perState);
    const int t = rem / a.nJ, k = rem - t * a.nJ;
    const CombineTask tk = a.tasks[t];
    const long rec = 3L * a.nJ;
    auto frame = [&](int u) { return a.tmp + a.slots[u].offA + s * a.slots[u].strideA; };
    auto jt = [&](int u) { return a.tmp + a.slots[u].offJ + s * a.slots[u].strideJ; };
    auto jr = [&](int u) { return a.tmp + a.slots[u].offJ + a.slots[u].jrShift + s * a.slots[u].strideJ; };

    /* column k of body u's translation (rot = false) / rotation Jacobian; zero off its chain */
    auto column = [&](int u, bool rot, double (&c)[3]) {
      const int ci = a.colMap[u * a.nJ + k];
      c[0] = c[1] = c[2] = 0.0;
      if (ci >= 0)
      {
        const int rs = a.slots[u].rowStride;
        const double* j = (rot ? jr(u) : jt(u)) + ci;
        c[0] = j[0];
        c[1] = j[rs];
        c[2] = j[2 * rs];
      }
    };

    double v[3] = {0.0, 0.0, 0.0};
    if (tk.kind == RCSB_JAC_POS)
    {
      if (tk.u2 >= 0)
      {
        column(tk.u2, false, v);
      }
      if (tk.mode == 1)
      {
        double j1[3];
        column(tk.u1, false, j1);
        v[0] -= j1[0]; v[1] -= j1[1]; v[2] -= j1[2];
        if (tk.u3 >= 0)
        {
          const double* o1 = frame(tk.u1);
          double r[3] = {-o1[0], -o1[1], -o1[2]};
          if (tk.u2 >= 0)
          {
            const double* o2 = frame(tk.u2);
            r[0] += o2[0]; r[1] += o2[1]; r[2] += o2[2];
          }
          double j3[3];
          column(tk.u3, true, j3);
          const double w0 = j3[0], w1 = j3[1], w2 = j3[2];
          v[0] += r[1] * w2 - r[2] * w1;
          v[1] += -r[0] * w2 + r[2] * w0;
          v[2] += r[0] * w1 - r[1] * w0;
        }
      }
    }
    else
    {
      if (tk.u2 >= 0)
      {
        column(tk.u2, true, v);
      }
      if (tk.mode == 1 && tk.u1 >= 0)
      {
        double j1[3];
        column(tk.u1, true, j1);
        v[0] -= j1[0]; v[1] -= j1[1]; v[2] -= j1[2];
      }
    }
    if (tk.uA >= 0)
    {
      const double* R = frame(tk.uA) + 3;
      const double x = v[0], y = v[1], z = v[2];
      v[0] = R[0] * x + R[1] * y + R[2] * z;
      v[1] = R[3] * x + R[4] * y + R[5] * z;
      v[2] = R[6] * x + R[7] * y + R[8] * z;
    }
    double* o = a.out + (s * a.nTasks + t) * rec + k;
    o[0] = v[0];
    o[a.nJ] = v[1];
    o[2 * a.nJ] = v[2];
  }
}

}   // namespace

using namespace rcsbhost;

extern "C" int rcsb_task_jacobians(const RcsbModel* cm, long N, int qdim, const double* q, int nJac,
                                   const RcsbRelJac* jac, double* J, unsigned flags, void* stream)
{
  RcsbModel* m = const_cast<RcsbModel*>(cm);
  int rc = checkModel(m, N, qdim, q);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  if (nJac <= 0 || !jac || (N > 0 && !J))
  {
    rcsb_set_error("rcsb_task_jacobians: no tasks / no output");
    return RCSB_EINVAL;
  }
  const int nb = m->hdr.nb, nJ = m->hdr.nJ;
  const int* jntPrev = m->iarr(RCSB_A_JNT_PREV);
  const int* jntJacobi = m->iarr(RCSB_A_JNT_JACOBI);
  const int* bodyLastJnt = m->iarr(RCSB_A_BODY_LASTJNT);
  auto articulated = [&](int b) {   /* RcsBody_isArticulated, Rcs_body.c:688 */
    for (int j = bodyLastJnt[b]; j >= 0; j = jntPrev[j])
    {
      if (jntJacobi[j] >= 0)
      {
        return true;
      }
    }
    return false;
  };

  /* bodies involved -> slots of the pass buffers */
  std::vector<int> slotOf((size_t) nb, -1), bodies;
  auto slot = [&](int b) {
    if (b < 0)
    {
      return -1;
    }
    if (slotOf[(size_t) b] < 0)
    {
      slotOf[(size_t) b] = (int) bodies.size();
      bodies.push_back(b);
    }
    return slotOf[(size_t) b];
  };
  std::vector<CombineTask> tasks((size_t) nJac);
  for (int t = 0; t < nJac; ++t)
  {
    const RcsbRelJac& r = jac[t];
    if ((r.kind != RCSB_JAC_POS && r.kind != RCSB_JAC_OMEGA) || r.effector >= nb || r.refBdy >= nb ||
        r.refFrame >= nb || r.effector < -1 || r.refBdy < -1 || r.refFrame < -1)
    {
      rcsb_set_error("rcsb_task_jacobians: task %d is invalid", t);
      return RCSB_EINVAL;
    }
    CombineTask& c = tasks[(size_t) t];
    c.kind = r.kind;
    c.u2 = slot(r.effector);
    c.u1 = c.u3 = c.uA = -1;
    c.mode = 0;
    if (r.kind == RCSB_JAC_POS)
    {
      if (r.refBdy < 0)
      {
        c.uA = slot(r.refFrame);
      }
      else
      {
        c.mode = 1;
        c.u1 = slot(r.refBdy);
        c.u3 = slot(r.refFrame);
        c.uA = (r.refFrame >= 0 && r.refFrame != r.refBdy) ? slot(r.refFrame) : c.u1;
      }
    }
    else
    {
      if (r.refBdy < 0 && r.refFrame < 0)
      {
        /* world */
      }
      else if (r.refBdy >= 0 && r.refFrame == r.refBdy && !articulated(r.refBdy))
      {
        c.uA = slot(r.refBdy);
      }
      else
      {
        c.mode = 1;
        c.u1 = slot(r.refBdy);
        c.uA = slot(r.refFrame);
      }
    }
  }
  if (N == 0)
  {
    return RCSB_OK;
  }
  if (bodies.empty())
  {
    bodies.push_back(0);   /* only NULL bodies: all rows are zero, keep the pass well-formed */
  }
  const int nU = (int) bodies.size();
  std::vector<char> needT((size_t) nU, 0), needR((size_t) nU, 0);
  for (const CombineTask& c : tasks)
  {
    auto mark = [&](std::vector<char>& v, int u) {
      if (u >= 0)
      {
        v[(size_t) u] = 1;
      }
    };
    if (c.kind == RCSB_JAC_POS)
    {
      mark(needT, c.u2);
      if (c.mode == 1)
      {
        mark(needT, c.u1);
        mark(needR, c.u3);
      }
    }
    else
    {
      mark(needR, c.u2);
      if (c.mode == 1)
      {
        mark(needR, c.u1);
      }
    }
  }

  /* groups of bodies whose joint chains fit the FK kernel's shared-memory scratch together */
  std::vector<std::vector<int>> groups;
  {
    std::vector<char> onChain((size_t) m->hdr.dof, 0);
    int nChain = 0;
    for (int u = 0; u < nU; ++u)
    {
      int add = 0;
      for (int j = bodyLastJnt[bodies[(size_t) u]]; j >= 0; j = jntPrev[j])
      {
        add += (jntJacobi[j] >= 0 && !onChain[(size_t) j]) ? 1 : 0;
      }
      const int g = groups.empty() ? 0 : (int) groups.back().size();
      const int rows = 6 * (nChain + add) + 3 * (g + 1);
      /* plan blobs are small next to the scratch; 8 KB covers them */
      /* one body per launch by default: with the pruned walk a launch visits that body's
       * ancestors only and its scratch leaves room for two blocks per SM (measured on the
       * humanoid's cAction task set: 10.2 ms per 1M states with full groups, 8.0 ms with 1) */
      static const int maxGroup = getenv("RCSB_TASKJAC_GROUP") ? atoi(getenv("RCSB_TASKJAC_GROUP")) : 1;
      if (groups.empty() || g >= maxGroup ||
          fkSharedBytes(m, 8192 + 8 * 3 * nJ * (g + 1), rows) > RCSB_SMEM_LIMIT)
      {
        groups.emplace_back();
        std::fill(onChain.begin(), onChain.end(), 0);
        nChain = 0;
      }
      for (int j = bodyLastJnt[bodies[(size_t) u]]; j >= 0; j = jntPrev[j])
      {
        if (jntJacobi[j] >= 0 && !onChain[(size_t) j])
        {
          onChain[(size_t) j] = 1;
          ++nChain;
        }
      }
      groups.back().push_back(u);
    }
  }

  const bool host = (flags & RCSB_HOST_PTRS) != 0;
  cudaStream_t st = host ? nullptr : static_cast<cudaStream_t>(stream);
  std::lock_guard<std::recursive_mutex> lock(m->mtx);
  RCSB_CUDA(cudaSetDevice(m->device));

  /* states per pass: large enough to fill the GPU with the per-body launches (measured on the
   * humanoid's cAction task set, ms per 1M states: 8192: 12.9, 32768: 5.8, 131072: 5.5,
   * 262144: 5.2), bounded for the size of the intermediate buffer */
  static const long passStates = getenv("RCSB_TASKJAC_PASS") ? atol(getenv("RCSB_TASKJAC_PASS")) : 131072;
  const long chunk = N < passStates ? N : passStates;
  /* compact record width of every group and the column maps of the bodies */
  std::vector<int> colMap((size_t) nU * nJ, -1), groupCols(groups.size(), 1);
  size_t perState = 0;
  for (size_t gi = 0; gi < groups.size(); ++gi)
  {
    for (int u : groups[gi])
    {
      std::vector<int> cols;
      for (int j = bodyLastJnt[bodies[(size_t) u]]; j >= 0; j = jntPrev[j])
      {
        if (jntJacobi[j] >= 0)
        {
          cols.push_back(jntJacobi[j]);
        }
      }
      for (size_t i = 0; i < cols.size(); ++i)
      {
        colMap[(size_t) u * nJ + cols[i]] = (int) (cols.size() - 1 - i);   /* ascending rank */
      }
      groupCols[gi] = (int) cols.size() > groupCols[gi] ? (int) cols.size() : groupCols[gi];
    }
    perState += groups[gi].size() * (12 + 6 * (size_t) groupCols[gi]);
  }
  const size_t tmpDoubles = (size_t) chunk * perState;
  const size_t outDoubles = host ? (size_t) chunk * nJac * 3 * nJ : 0;
  const size_t qDoubles = host ? (size_t) chunk * qdim : 0;
  double* tmp = nullptr;
  CombineTask* dTasks = nullptr;
  SlotInfo* dSlots = nullptr;
  RCSB_CUDA(cudaMalloc(&tmp, (tmpDoubles + outDoubles + qDoubles) * sizeof(double)));
  int* dColMap = nullptr;
  cudaError_t e = cudaMalloc(&dTasks, tasks.size() * sizeof(CombineTask) + (size_t) nU * sizeof(SlotInfo) +
                                      colMap.size() * sizeof(int));
  if (e == cudaSuccess)
  {
    dSlots = reinterpret_cast<SlotInfo*>(dTasks + tasks.size());
    dColMap = reinterpret_cast<int*>(dSlots + nU);
    e = cudaMemcpyAsync(dTasks, tasks.data(), tasks.size() * sizeof(CombineTask),
                        cudaMemcpyHostToDevice, st);
  }
  if (e == cudaSuccess)
  {
    e = cudaMemcpyAsync(dColMap, colMap.data(), colMap.size() * sizeof(int), cudaMemcpyHostToDevice, st);
  }
  int result = RCSB_OK;
  std::vector<SlotInfo> slots((size_t) nU);
  long lastN = -1;
  for (long s0 = 0; s0 < N && e == cudaSuccess && result == RCSB_OK; s0 += chunk)
  {
    const long n = N - s0 < chunk ? N - s0 : chunk;
    double* dOut = host ? tmp + tmpDoubles : J + (size_t) s0 * nJac * 3 * nJ;
    const double* dQ = q + (size_t) s0 * qdim;
    if (host)
    {
      double* dq = tmp + tmpDoubles + outDoubles;
      e = cudaMemcpyAsync(dq, dQ, sizeof(double) * (size_t) n * qdim, cudaMemcpyHostToDevice, st);
      dQ = dq;
      if (e != cudaSuccess)
      {
        break;
      }
    }
    long base = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi)
    {
      const std::vector<int>& grp = groups[gi];
      const int G = (int) grp.size();
      const long nc = groupCols[gi];
      std::vector<int> gb((size_t) G);
      for (int i = 0; i < G; ++i)
      {
        gb[(size_t) i] = bodies[(size_t) grp[(size_t) i]];
      }
      double* A = tmp + base;
      double* JT = A + (size_t) n * G * 12;
      double* JR = JT + (size_t) n * G * 3 * nc;
      for (int i = 0; i < G; ++i)
      {
        SlotInfo& si = slots[(size_t) grp[(size_t) i]];
        si.offA = base + 12L * i;
        si.strideA = 12L * G;
        si.offJ = (JT - tmp) + 3L * nc * i;
        si.strideJ = 3L * nc * G;
        si.jrShift = JR - JT;
        si.rowStride = (int) nc;
        si.pad = 0;
      }
      /* only the Jacobians some task reads (a world position task needs no rotation Jacobian) */
      bool wantT = false, wantR = false;
      for (int i = 0; i < G; ++i)
      {
        wantT = wantT || needT[(size_t) grp[(size_t) i]];
        wantR = wantR || needR[(size_t) grp[(size_t) i]];
      }
      int recCols = 0;
      result = fkCompactDevice(m, n, qdim, dQ, G, gb.data(), A, wantT ? JT : nullptr,
                               wantR ? JR : nullptr, st, &recCols);
      if (result != RCSB_OK)
      {
        break;
      }
      if (recCols != nc)
      {
        rcsb_set_error("rcsb_task_jacobians: internal record width mismatch (%d, %ld)", recCols, nc);
        result = RCSB_EINVAL;
        break;
      }
      base += (long) n * G * (12 + 6L * nc);
    }
    if (result != RCSB_OK)
    {
      break;
    }
    if (n != lastN)
    {
      /* pageable source: the copy has left the host buffer when the call returns */
      e = cudaMemcpyAsync(dSlots, slots.data(), slots.size() * sizeof(SlotInfo), cudaMemcpyHostToDevice, st);
      lastN = n;
      if (e != cudaSuccess)
      {
        break;
      }
    }
    CombineArgs ca;
    ca.tmp = tmp;
    ca.slots = dSlots;
    ca.out = dOut;
    ca.tasks = dTasks;
    ca.colMap = dColMap;
    ca.n = n;
    ca.nTasks = nJac;
    ca.nJ = nJ;
    const long total = n * nJac * nJ;
    long blocks = (total + 255) / 256;
    const long cap = (long) m->smCount * 8;
    blocks = blocks < cap ? blocks : cap;
    combineKernel<<<(unsigned int) blocks, 256, 0, st>>>(ca);
    e = cudaGetLastError();
    g_launches.fetch_add(1);
    if (e == cudaSuccess && host)
    {
      e = cudaMemcpyAsync(J + (size_t) s0 * nJac * 3 * nJ, dOut,
                          sizeof(double) * (size_t) n * nJac * 3 * nJ, cudaMemcpyDeviceToHost, st);
    }
  }
  /* the pass buffers are released once the stream has consumed them */
  cudaError_t e2 = host ? cudaStreamSynchronize(st) : cudaSuccess;
  if (!host)
  {
    cudaFreeAsync(tmp, st);
    cudaFreeAsync(dTasks, st);
  }
  else
  {
    cudaFree(tmp);
    cudaFree(dTasks);
  }
  if (result != RCSB_OK)
  {
    return result;
  }
  RCSB_CUDA(e);
  RCSB_CUDA(e2);
  return RCSB_OK;
}
