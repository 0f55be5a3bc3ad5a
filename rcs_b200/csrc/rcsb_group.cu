/*
 * Multi-GPU entry points of the C ABI (include/rcsb.h, "device groups"): one process drives the
 * GPUs of a node. States are independent and the model is tiny and replicated, so a batch is
 * cut into contiguous slices (rcsb_group_slice, the rule of rcs_b200/shard.py), one per device;
 * each device gets its own host thread, stream, model replica and slice buffers, and runs the
 * same kernels as the single-device entry points on its slice. There is no data-path
 * collective. The one exchange of the path is an all-gather of per-rank summary statistics of a
 * per-state result (5 doubles per rank: count, sum, sum of squares, min, max), done with NCCL
 * between the devices of the group (single process, one communicator per device) before the
 * results go back to the host.
 *
 * NCCL is looked up at run time: the copy already in the process if there is one (PyTorch loads its
 * own libnccl.so.2), else $RCSB_NCCL_LIBRARY, else the system's libnccl.so.2. librcsb.so itself
 * does not link against it. A group of more than one device cannot be created without it.
 */
#include "rcsb_priv.h"

#include <dlfcn.h>
#include <nccl.h>
#include <thread>

using namespace rcsbhost;

namespace
{

struct NcclApi
{
  void* handle = nullptr;
  decltype(&ncclCommInitAll) commInitAll = nullptr;
  decltype(&ncclCommDestroy) commDestroy = nullptr;
  decltype(&ncclAllGather) allGather = nullptr;
  decltype(&ncclGetErrorString) errorString = nullptr;
  bool ok = false;
};

NcclApi& nccl()
{
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    /* 1. a copy that is already in the process (PyTorch brings its own libnccl.so.2: loading another
     *    one under the same soname first would make a later `import torch` bind to the wrong version),
     * 2. RCSB_NCCL_LIBRARY, 3. the system library */
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char* env = getenv("RCSB_NCCL_LIBRARY");
    if (!api.handle && env && *env)
    {
      api.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    }
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
    {
      if (!api.handle)
      {
        api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      }
    }
    if (!api.handle)
    {
      return;
    }
    api.commInitAll = reinterpret_cast<decltype(&ncclCommInitAll)>(dlsym(api.handle, "ncclCommInitAll"));
    api.commDestroy = reinterpret_cast<decltype(&ncclCommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
    api.allGather = reinterpret_cast<decltype(&ncclAllGather)>(dlsym(api.handle, "ncclAllGather"));
    api.errorString = reinterpret_cast<decltype(&ncclGetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
    api.ok = api.commInitAll && api.commDestroy && api.allGather && api.errorString;
  });
  return api;
}

struct Member
{
  int device = 0;
  RcsbModel* model = nullptr;
  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;
  char* buf = nullptr;         /* slice buffers (inputs, results), grown on demand */
  size_t cap = 0;
  double* stats = nullptr;     /* 5 doubles of this rank + nDev x 5 gathered */
  std::string error;
  int rc = RCSB_OK;
};

}   // namespace

struct RcsbGroup
{
  std::vector<Member> members;
  std::mutex mtx;              /* one call at a time per group */
};

namespace
{

int reserve(Member& mb, size_t bytes)
{
  if (mb.cap >= bytes)
  {
    return RCSB_OK;
  }
  if (mb.buf)
  {
    RCSB_CUDA(cudaFree(mb.buf));
    mb.buf = nullptr;
    mb.cap = 0;
  }
  RCSB_CUDA(cudaMalloc(&mb.buf, bytes));
  mb.cap = bytes;
  return RCSB_OK;
}

/*
 * Runs one call on every member, each on its own host thread: the member's slice goes through the
 * host-pointer entry point of the single-device API (H2D, kernels and D2H of successive stages
 * overlap on three streams, rcsb_api.cu), the summary statistics of the per-state value `statHost`
 * names are reduced on the device stage by stage while the value is still there (StageTap), their
 * partial results are merged, and the members all-gather the 5 doubles over NCCL.
 *   call(mb, first, n)  the host-pointer call on the slice [first, first + n)
 *   statHost            host address (whole batch) of the output the statistics are taken of, or NULL
 *   statPerState        doubles per state of that output, statOffset the entry within a state's record
 */
typedef std::function<int(Member& mb, long first, long n)> GroupCall;

int runGroup(RcsbGroup* g, long N, const GroupCall& call, const double* statHost, size_t statPerState,
             size_t statOffset, double* statsOut)
{
  std::lock_guard<std::mutex> lock(g->mtx);
  const int R = (int) g->members.size();
  std::vector<std::thread> threads;
  for (int r = 0; r < R; ++r)
  {
    threads.emplace_back([&, r]() {
      Member& mb = g->members[(size_t) r];
      mb.rc = [&]() -> int {
        long first = 0, n = 0;
        rcsb_group_slice(g, N, r, &first, &n);
        RCSB_CUDA(cudaSetDevice(mb.device));
        int nStages = 0;
        StageTap tap;
        if (statHost && n > 0)
        {
          tap.hostOut = statHost + (size_t) first * statPerState;
          tap.begin = [&](int stages) -> int {
            nStages = stages;
            return reserve(mb, (size_t) stages * 5 * sizeof(double));
          };
          tap.stage = [&](const double* dev, long, long cnt, int stageIndex, cudaStream_t st) -> int {
            return rcsb_summary_stats(dev + statOffset, cnt, (long) statPerState,
                                      reinterpret_cast<double*>(mb.buf) + 5 * stageIndex, st);
          };
          setStageTap(&tap);
        }
        int rc = n > 0 ? call(mb, first, n) : RCSB_OK;
        setStageTap(nullptr);
        if (rc != RCSB_OK)
        {
          return rc;
        }
        /* merge the stages' statistics (the call has synchronised its streams) */
        double mine[5] = {0.0, 0.0, 0.0, 1.0 / 0.0, -1.0 / 0.0};
        if (nStages > 0)
        {
          std::vector<double> part((size_t) nStages * 5);
          RCSB_CUDA(cudaMemcpy(part.data(), mb.buf, part.size() * sizeof(double), cudaMemcpyDeviceToHost));
          for (int k = 0; k < nStages; ++k)
          {
            const double* p5 = part.data() + 5 * k;
            mine[0] += p5[0];
            mine[1] += p5[1];
            mine[2] += p5[2];
            mine[3] = p5[3] < mine[3] ? p5[3] : mine[3];
            mine[4] = p5[4] > mine[4] ? p5[4] : mine[4];
          }
        }
        /* every rank takes part in the all-gather, also one with an empty slice (count 0) */
        RCSB_CUDA(cudaMemcpyAsync(mb.stats, mine, sizeof(mine), cudaMemcpyHostToDevice, mb.stream));
        if (R > 1)
        {
          const ncclResult_t e = nccl().allGather(mb.stats, mb.stats + 5, 5, ncclDouble, mb.comm, mb.stream);
          if (e != ncclSuccess)
          {
            rcsb_set_error("ncclAllGather: %s", nccl().errorString(e));
            return RCSB_ECUDA;
          }
        }
        else
        {
          RCSB_CUDA(cudaMemcpyAsync(mb.stats + 5, mb.stats, 5 * sizeof(double), cudaMemcpyDeviceToDevice,
                                    mb.stream));
        }
        if (r == 0 && statsOut)
        {
          RCSB_CUDA(cudaMemcpyAsync(statsOut, mb.stats + 5, sizeof(double) * 5 * (size_t) R,
                                    cudaMemcpyDeviceToHost, mb.stream));
        }
        RCSB_CUDA(cudaStreamSynchronize(mb.stream));
        return RCSB_OK;
      }();
      if (mb.rc != RCSB_OK)
      {
        mb.error = rcsb_last_error();   /* the message is thread-local: carry it over */
      }
    });
  }
  for (std::thread& t : threads)
  {
    t.join();
  }
  for (Member& mb : g->members)
  {
    if (mb.rc != RCSB_OK)
    {
      rcsb_set_error("device %d: %s", mb.device, mb.error.c_str());
      return mb.rc;
    }
  }
  return RCSB_OK;
}

/* host address of a member's slice of an array of the whole batch (NULL stays NULL) */
template <typename T>
T* sliceOf(T* base, long first, size_t perState)
{
  return base ? base + (size_t) first * perState : nullptr;
}

}   // namespace

extern "C" void rcsb_group_slice(const RcsbGroup* g, long N, int rank, long* first, long* count)
{
  const long R = g ? (long) g->members.size() : 1;
  const long base = N / R, extra = N % R;
  if (first)
  {
    *first = rank * base + (rank < extra ? rank : extra);
  }
  if (count)
  {
    *count = base + (rank < extra ? 1 : 0);
  }
}

extern "C" int rcsb_group_size(const RcsbGroup* g)
{
  return g ? (int) g->members.size() : 0;
}

extern "C" void rcsb_group_destroy(RcsbGroup* g)
{
  if (!g)
  {
    return;
  }
  for (Member& mb : g->members)
  {
    cudaSetDevice(mb.device);
    if (mb.comm)
    {
      nccl().commDestroy(mb.comm);
    }
    if (mb.stream)
    {
      cudaStreamDestroy(mb.stream);
    }
    cudaFree(mb.buf);
    cudaFree(mb.stats);
    rcsb_model_destroy(mb.model);
  }
  delete g;
}

extern "C" int rcsb_group_create(const RcsGraph* graph, int nDevices, const int* devices, RcsbGroup** out)
{
  if (!graph || !out || nDevices < 1)
  {
    rcsb_set_error("rcsb_group_create: invalid arguments");
    return RCSB_EINVAL;
  }
  *out = nullptr;
  const int avail = rcsb_device_count();
  if (avail < 1)
  {
    rcsb_set_error("no CUDA device available; rcs_b200 has no CPU fallback");
    return RCSB_ENODEVICE;
  }
  std::vector<int> devs((size_t) nDevices);
  for (int i = 0; i < nDevices; ++i)
  {
    devs[(size_t) i] = devices ? devices[i] : i;
    if (devs[(size_t) i] < 0 || devs[(size_t) i] >= avail)
    {
      rcsb_set_error("rcsb_group_create: device %d out of range [0, %d)", devs[(size_t) i], avail);
      return RCSB_EINVAL;
    }
    for (int k = 0; k < i; ++k)
    {
      if (devs[(size_t) k] == devs[(size_t) i])
      {
        rcsb_set_error("rcsb_group_create: device %d listed twice", devs[(size_t) i]);
        return RCSB_EINVAL;
      }
    }
  }
  if (nDevices > 1 && !nccl().ok)
  {
    rcsb_set_error("rcsb_group_create: libnccl.so.2 could not be loaded (%s); a group of %d devices needs it",
                   dlerror() ? dlerror() : "missing symbols", nDevices);
    return RCSB_EUNSUPPORTED;
  }

  RcsbGroup* g = new RcsbGroup;
  g->members.resize((size_t) nDevices);
  int rc = RCSB_OK;
  for (int i = 0; i < nDevices && rc == RCSB_OK; ++i)
  {
    Member& mb = g->members[(size_t) i];
    mb.device = devs[(size_t) i];
    rc = rcsb_model_create(graph, mb.device, &mb.model);
    if (rc == RCSB_OK)
    {
      cudaError_t e = cudaSetDevice(mb.device);
      if (e == cudaSuccess)
      {
        e = cudaStreamCreateWithFlags(&mb.stream, cudaStreamNonBlocking);
      }
      if (e == cudaSuccess)
      {
        e = cudaMalloc(&mb.stats, sizeof(double) * 5 * (size_t) (nDevices + 1));
      }
      if (e != cudaSuccess)
      {
        rc = cudaFail(e, "rcsb_group_create");
      }
    }
  }
  if (rc == RCSB_OK && nDevices > 1)
  {
    std::vector<ncclComm_t> comms((size_t) nDevices);
    const ncclResult_t e = nccl().commInitAll(comms.data(), nDevices, devs.data());
    if (e != ncclSuccess)
    {
      rcsb_set_error("ncclCommInitAll: %s", nccl().errorString(e));
      rc = RCSB_ECUDA;
    }
    else
    {
      for (int i = 0; i < nDevices; ++i)
      {
        g->members[(size_t) i].comm = comms[(size_t) i];
      }
    }
  }
  if (rc != RCSB_OK)
  {
    const std::string msg = rcsb_last_error();
    rcsb_group_destroy(g);
    rcsb_set_error("%s", msg.c_str());
    return rc;
  }
  *out = g;
  return RCSB_OK;
}

/* ---- calls ---------------------------------------------------------------------------------- */

extern "C" int rcsb_group_fk_jacobians_mass(RcsbGroup* g, long N, int qdim, const double* q, int nSel,
                                            const int* bodySel, double* A_BI, int nEff, const int* effBody,
                                            const double* effPt, double* JT, double* JR, double* M,
                                            double* stats)
{
  if (!g || N < 0 || (N > 0 && !q))
  {
    rcsb_set_error("rcsb_group_fk_jacobians_mass: invalid arguments");
    return RCSB_EINVAL;
  }
  const RcsbModel* m0 = g->members[0].model;
  const size_t nJ = (size_t) m0->hdr.nJ, nOut = nSel > 0 ? (size_t) nSel : (size_t) m0->hdr.nb;
  /* statistics: height (world z) of the first selected frame */
  return runGroup(g, N, [&](Member& mb, long first, long n) {
    return rcsb_fk_jacobians_mass(mb.model, n, qdim, sliceOf(q, first, (size_t) qdim), nSel, bodySel,
                                  sliceOf(A_BI, first, nOut * 12), nEff, effBody, effPt,
                                  sliceOf(JT, first, (size_t) nEff * 3 * nJ), sliceOf(JR, first, (size_t) nEff * 3 * nJ),
                                  sliceOf(M, first, nJ * nJ), RCSB_HOST_PTRS, nullptr);
  }, A_BI, nOut * 12, 2, stats);
}

extern "C" int rcsb_group_kinetic_terms(RcsbGroup* g, long N, int qdim, const double* q, const double* qd,
                                        double* M, double* h, double* Fg, double* E, double* stats)
{
  if (!g || N < 0 || (N > 0 && !q))
  {
    rcsb_set_error("rcsb_group_kinetic_terms: invalid arguments");
    return RCSB_EINVAL;
  }
  const size_t nJ = (size_t) g->members[0].model->hdr.nJ;
  /* statistics: total energy */
  return runGroup(g, N, [&](Member& mb, long first, long n) {
    return rcsb_kinetic_terms(mb.model, n, qdim, sliceOf(q, first, (size_t) qdim), sliceOf(qd, first, (size_t) qdim),
                              sliceOf(M, first, nJ * nJ), sliceOf(h, first, nJ), sliceOf(Fg, first, nJ),
                              sliceOf(E, first, 1), RCSB_HOST_PTRS, nullptr);
  }, E, 1, 0, stats);
}

extern "C" int rcsb_group_ik_rmr(RcsbGroup* g, int nTasks, const RcsbTask* tasks, long N, double* q,
                                 const double* xDes, double lambda, double alpha, int iters, double* dx,
                                 double* dq, double* det, double* stats)
{
  if (!g || N < 0 || (N > 0 && (!q || !xDes)) || nTasks <= 0 || !tasks)
  {
    rcsb_set_error("rcsb_group_ik_rmr: invalid arguments");
    return RCSB_EINVAL;
  }
  const size_t dof = (size_t) g->members[0].model->hdr.dof;
  size_t nx = 0;
  for (int t = 0; t < nTasks; ++t)
  {
    nx += tasks[t].kind == RCSB_TASK_XYZABC ? 6 : 3;
  }
  /* statistics: det of the last iteration; always computed, into a scratch array if the caller
   * does not ask for it */
  std::vector<double> detScratch;
  if (!det && N > 0)
  {
    detScratch.resize((size_t) N);
    det = detScratch.data();
  }
  return runGroup(g, N, [&](Member& mb, long first, long n) {
    return rcsb_ik_rmr(mb.model, nTasks, tasks, n, sliceOf(q, first, dof), sliceOf(xDes, first, nx), lambda, alpha,
                       iters, sliceOf(dx, first, nx), sliceOf(dq, first, dof), sliceOf(det, first, 1),
                       RCSB_HOST_PTRS, nullptr);
  }, det, 1, 0, stats);
}

extern "C" int rcsb_group_wholebody(RcsbGroup* g, long N, int qdim, const double* q, int nSel,
                                    const int* bodySel, double* A_BI, int nJac, const RcsbRelJac* jac,
                                    double* J, double* M, double* Jcog, double* stats)
{
  if (!g || N < 0 || (N > 0 && !q))
  {
    rcsb_set_error("rcsb_group_wholebody: invalid arguments");
    return RCSB_EINVAL;
  }
  const RcsbModel* m0 = g->members[0].model;
  const size_t nJ = (size_t) m0->hdr.nJ, nOut = nSel > 0 ? (size_t) nSel : (size_t) m0->hdr.nb;
  /* statistics: M[0][0] */
  return runGroup(g, N, [&](Member& mb, long first, long n) {
    return rcsb_wholebody(mb.model, n, qdim, sliceOf(q, first, (size_t) qdim), nSel, bodySel,
                          sliceOf(A_BI, first, nOut * 12), nJac, jac, sliceOf(J, first, (size_t) nJac * 3 * nJ),
                          sliceOf(M, first, nJ * nJ), sliceOf(Jcog, first, 3 * nJ), RCSB_HOST_PTRS, nullptr);
  }, M, nJ * nJ, 0, stats);
}
