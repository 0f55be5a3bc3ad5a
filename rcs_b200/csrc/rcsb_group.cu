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

/* one array of a group call: `perState` doubles per state, host pointer of the WHOLE batch */
struct GArray
{
  const double* in;
  double* out;
  size_t perState;
  double* dev;                 /* device address of this member's slice (set per member) */
};

size_t carve(size_t doubles)
{
  return (doubles * sizeof(double) + 255) & ~(size_t) 255;
}

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
 * Runs one call on every member: copy the member's slice of the inputs to its device, launch
 * (device-pointer entry point on the member's stream), reduce the statistics source, all-gather
 * the statistics over NCCL, copy the results back. `launch` returns the device address and
 * stride of the per-state value the statistics are taken of (or nullptr).
 */
typedef std::function<int(Member& mb, long n, std::vector<GArray>& arr, const double** statSrc,
                          long* statStride)> GroupLaunch;

int runGroup(RcsbGroup* g, long N, const std::vector<GArray>& arrays, const GroupLaunch& launch,
             double* statsOut)
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
        std::vector<GArray> arr = arrays;
        size_t need = 0;
        for (GArray& a : arr)
        {
          need += carve((size_t) n * a.perState);
        }
        int rc = reserve(mb, need > 0 ? need : 256);
        if (rc != RCSB_OK)
        {
          return rc;
        }
        size_t off = 0;
        for (GArray& a : arr)
        {
          a.dev = a.perState ? reinterpret_cast<double*>(mb.buf + off) : nullptr;
          off += carve((size_t) n * a.perState);
          if (a.in && a.perState && n > 0)
          {
            RCSB_CUDA(cudaMemcpyAsync(a.dev, a.in + (size_t) first * a.perState,
                                      sizeof(double) * (size_t) n * a.perState, cudaMemcpyHostToDevice,
                                      mb.stream));
          }
        }
        const double* statSrc = nullptr;
        long statStride = 1;
        if (n > 0)
        {
          rc = launch(mb, n, arr, &statSrc, &statStride);
          if (rc != RCSB_OK)
          {
            return rc;
          }
        }
        /* summary statistics of this rank, then the all-gather: every rank takes part, also one
         * with an empty slice (count 0) */
        if (statSrc && n > 0)
        {
          rc = rcsb_summary_stats(statSrc, n, statStride, mb.stats, mb.stream);
          if (rc != RCSB_OK)
          {
            return rc;
          }
        }
        else
        {
          const double empty[5] = {0.0, 0.0, 0.0, 1.0 / 0.0, -1.0 / 0.0};
          RCSB_CUDA(cudaMemcpyAsync(mb.stats, empty, sizeof(empty), cudaMemcpyHostToDevice, mb.stream));
        }
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
        for (GArray& a : arr)
        {
          if (a.out && a.perState && n > 0)
          {
            RCSB_CUDA(cudaMemcpyAsync(a.out + (size_t) first * a.perState, a.dev,
                                      sizeof(double) * (size_t) n * a.perState, cudaMemcpyDeviceToHost,
                                      mb.stream));
          }
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
  std::vector<GArray> arr;
  enum { I_Q, I_A, I_JT, I_JR, I_M };
  arr.push_back({q, nullptr, (size_t) qdim, nullptr});
  arr.push_back({nullptr, A_BI, A_BI ? nOut * 12 : 0, nullptr});
  arr.push_back({nullptr, JT, JT ? (size_t) nEff * 3 * nJ : 0, nullptr});
  arr.push_back({nullptr, JR, JR ? (size_t) nEff * 3 * nJ : 0, nullptr});
  arr.push_back({nullptr, M, M ? nJ * nJ : 0, nullptr});
  return runGroup(g, N, arr, [&](Member& mb, long n, std::vector<GArray>& a, const double** src, long* stride) {
    /* statistics: height (world z) of the first selected frame */
    if (a[I_A].dev)
    {
      *src = a[I_A].dev + 2;
      *stride = (long) nOut * 12;
    }
    return rcsb_fk_jacobians_mass(mb.model, n, qdim, a[I_Q].dev, nSel, bodySel, a[I_A].dev, nEff, effBody,
                                  effPt, a[I_JT].dev, a[I_JR].dev, a[I_M].dev, 0, mb.stream);
  }, stats);
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
  std::vector<GArray> arr;
  enum { I_Q, I_QD, I_M, I_H, I_G, I_E };
  arr.push_back({q, nullptr, (size_t) qdim, nullptr});
  arr.push_back({qd, nullptr, qd ? (size_t) qdim : 0, nullptr});
  arr.push_back({nullptr, M, M ? nJ * nJ : 0, nullptr});
  arr.push_back({nullptr, h, h ? nJ : 0, nullptr});
  arr.push_back({nullptr, Fg, Fg ? nJ : 0, nullptr});
  arr.push_back({nullptr, E, (size_t) (E ? 1 : 0), nullptr});
  return runGroup(g, N, arr, [&](Member& mb, long n, std::vector<GArray>& a, const double** src, long* stride) {
    if (a[I_E].dev)
    {
      *src = a[I_E].dev;     /* statistics: total energy */
      *stride = 1;
    }
    return rcsb_kinetic_terms(mb.model, n, qdim, a[I_Q].dev, a[I_QD].dev, a[I_M].dev, a[I_H].dev, a[I_G].dev,
                              a[I_E].dev, 0, mb.stream);
  }, stats);
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
  std::vector<GArray> arr;
  enum { I_Q, I_X, I_DX, I_DQ, I_DET };
  arr.push_back({q, q, dof, nullptr});          /* in-out */
  arr.push_back({xDes, nullptr, nx, nullptr});
  arr.push_back({nullptr, dx, dx ? nx : 0, nullptr});
  arr.push_back({nullptr, dq, dq ? dof : 0, nullptr});
  arr.push_back({nullptr, det, (size_t) 1, nullptr});   /* always computed: source of the statistics */
  return runGroup(g, N, arr, [&](Member& mb, long n, std::vector<GArray>& a, const double** src, long* stride) {
    *src = a[I_DET].dev;
    *stride = 1;
    return rcsb_ik_rmr(mb.model, nTasks, tasks, n, a[I_Q].dev, a[I_X].dev, lambda, alpha, iters, a[I_DX].dev,
                       a[I_DQ].dev, a[I_DET].dev, 0, mb.stream);
  }, stats);
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
  std::vector<GArray> arr;
  enum { I_Q, I_A, I_J, I_M, I_JC };
  arr.push_back({q, nullptr, (size_t) qdim, nullptr});
  arr.push_back({nullptr, A_BI, A_BI ? nOut * 12 : 0, nullptr});
  arr.push_back({nullptr, J, J ? (size_t) nJac * 3 * nJ : 0, nullptr});
  arr.push_back({nullptr, M, M ? nJ * nJ : 0, nullptr});
  arr.push_back({nullptr, Jcog, Jcog ? 3 * nJ : 0, nullptr});
  return runGroup(g, N, arr, [&](Member& mb, long n, std::vector<GArray>& a, const double** src, long* stride) {
    if (a[I_M].dev)
    {
      *src = a[I_M].dev;     /* statistics: M[0][0] */
      *stride = (long) (nJ * nJ);
    }
    return rcsb_wholebody(mb.model, n, qdim, a[I_Q].dev, nSel, bodySel, a[I_A].dev, nJac, jac, a[I_J].dev,
                          a[I_M].dev, a[I_JC].dev, 0, mb.stream);
  }, stats);
}
