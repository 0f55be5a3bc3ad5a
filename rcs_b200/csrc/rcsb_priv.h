/*
 * Private host-side declarations shared by the translation units that implement the C ABI
 * (rcsb_api.cu, rcsb_api_dyn.cu, rcsb_single.cu).
 */
#ifndef RCSB_PRIV_H
#define RCSB_PRIV_H

#include "rcsb.h"
#include "rcsb_internal.h"
#include "rcsb_kernels.h"
#include "rcsb_model.h"
#include "rcsb_plan.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace rcsbhost
{

extern std::atomic<unsigned long long> g_launches;

struct DevicePlan
{
  char* dev = nullptr;
  int bytes = 0;
  int nScrRows = 0;
  int nSel = 0;
  int nEff = 0;
  int aux0 = 0;
  int aux1 = 0;
  int aux2 = 0;
};

/* double-buffered device workspace of the RCSB_HOST_PTRS path */
/* stages of the host-buffer pipeline (and of every per-stage workspace of a model): three, so that
 * the copy-in of one chunk, the kernels of the next and the copy-out of a third overlap */
#define RCSB_PIPE_SLOTS 3
struct StagePipe
{
  cudaStream_t stream[RCSB_PIPE_SLOTS] = {};
  char* buf[RCSB_PIPE_SLOTS] = {};
  size_t cap[RCSB_PIPE_SLOTS] = {};
};

}   // namespace rcsbhost

struct RcsbModel
{
  int device = 0;
  int smCount = 0;
  std::vector<char> host;        /* flat model blob */
  char* dev = nullptr;
  RcsbFlatHeader hdr;
  std::recursive_mutex mtx;               /* guards the plan caches, workspaces and the staging pipe */
  std::map<std::string, rcsbhost::DevicePlan> fkPlans;
  std::vector<std::string> fkPointPlanKeys;   /* FK plans that carry effector points, oldest first (bounded) */
  std::map<std::string, rcsbhost::DevicePlan> ikPlans;
  rcsbhost::DevicePlan ktPlan[2];              /* kinetic-terms plans without / with velocities */
  std::map<std::string, rcsbhost::DevicePlan> wbPlans;   /* whole-body variants of the kinetic-terms plan */
  bool ktPlanBuilt[2] = {false, false};
  double* ktScratch[RCSB_PIPE_SLOTS] = {};
  size_t ktScratchDoubles[RCSB_PIPE_SLOTS] = {};
  char* ikScratch[RCSB_PIPE_SLOTS] = {};      /* gathered x_des of masked IK calls, per staging slot */
  size_t ikScratchBytes[RCSB_PIPE_SLOTS] = {};
  rcsbhost::StagePipe pipe;

  const int* iarr(int a) const
  {
    return reinterpret_cast<const int*>(host.data() + hdr.off[a]);
  }
  const double* darr(int a) const
  {
    return reinterpret_cast<const double*>(host.data() + hdr.off[a]);
  }
};

namespace rcsbhost
{

inline int cudaFail(cudaError_t e, const char* what)
{
  rcsb_set_error("%s: %s", what, cudaGetErrorString(e));
  cudaGetLastError();
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
  {
    return RCSB_ENODEVICE;
  }
  return RCSB_ECUDA;
}

#define RCSB_CUDA(call)                        \
  do                                           \
  {                                            \
    cudaError_t e_ = (call);                   \
    if (e_ != cudaSuccess)                     \
    {                                          \
      return rcsbhost::cudaFail(e_, #call);    \
    }                                          \
  } while (0)

inline int align16(int x)
{
  return (x + 15) & ~15;
}

int checkModel(const RcsbModel* m, long N, int qdim, const double* q);

/* dynamic shared memory of one fkKernel block: model + plan + Jacobian scratch */
#define RCSB_SMEM_LIMIT (227 * 1024)
inline size_t fkSharedBytes(const RcsbModel* m, int planBytes, int nScrRows, bool withMass = false,
                            bool withVelocity = false)
{
  const size_t threads = (size_t) RCSB_FK_BLOCK(withMass);
  return (size_t) m->hdr.totalBytes + (size_t) planBytes + (size_t) nScrRows * (threads + 1) * sizeof(double) +
         (withVelocity ? (threads / 32) * 2 * 32 * 33 * sizeof(double) : 0) + 16 /* staging mbarrier */;
}

/* One array of a staged (host-pointer) call: `perState` doubles per state; exactly one of
 * hostIn / hostOut is set; dev receives the device address of the current chunk. */
struct StagedArray
{
  const double* hostIn;
  double* hostOut;
  size_t perState;
  double* dev;
};

/*
 * Runs `launch(n, stream)` over stages of the batch: inputs are copied host -> device,
 * the kernel is enqueued, outputs are copied device -> host, rotating over RCSB_PIPE_SLOTS
 * streams so that the copies of one stage overlap the kernels of the others. In-out arrays
 * are listed twice (once as input, once as output, same perState) and share a buffer when
 * `aliasOutToIn` names the input's index. Returns after all results are in host memory.
 */
int runStaged(RcsbModel* m, long N, std::vector<StagedArray>& arrays,
              const std::vector<int>& aliasOutToIn,
              const std::function<int(long n, cudaStream_t st, int slot)>& launch);

/*
 * Per-stage hook of runStaged for the calling thread (device groups: the summary statistics of a
 * per-state result are reduced on the device, stage by stage, while the result is still there).
 * `hostOut` names the watched output array by the host address the caller passed for it; after the
 * kernels of a stage `stage(dev, s0, n, stageIndex, stream)` is called with the device address of
 * that stage's part of the array; `begin(nStages)` once before the first stage.
 */
struct StageTap
{
  const double* hostOut = nullptr;
  std::function<int(int nStages)> begin;
  std::function<int(const double* dev, long s0, long n, int stageIndex, cudaStream_t st)> stage;
};
void setStageTap(const StageTap* tap);   /* thread-local; nullptr clears it */

/* Forward kinematics (+ Jacobians) behind rcsb_fk / rcsb_fk_jacobians and the single-state
 * RcsGraph_* wrappers; qdOut and noCoupling are only reachable from the latter. */
int fkCommon(const RcsbModel* cm, long N, int qdim, const double* q, const double* qd, int nSel,
             const int* bodySel, double* A_BI, double* A_JI, double* xdot, double* omega,
             double* qOut, int nEff, const int* effBody, const double* effPt, double* JT,
             double* JR, double* qdOut, bool noCoupling, unsigned flags, void* stream,
             double* M = nullptr);

int fkCompactDevice(RcsbModel* m, long n, int qdim, const double* q, int G, const int* bodies,
                    double* A, double* JT, double* JR, cudaStream_t stream, int* recCols);

/* body selection and task list of a whole-body variant of the kinetic-terms plan */
struct WbSpec
{
  int nSel = 0;                    /* 0: all bodies */
  const int* bodySel = nullptr;
  std::vector<int> taskBodies;     /* body of slot u */
  std::vector<int> jacTasks;       /* 8 ints per task Jacobian: kind, mode, u2, u1, u3, uA, 0, 0 */
  std::string key;
};

/* optional inputs / outputs of one pass of the kinetic-terms kernels */
struct KtExtra
{
  double* cog = nullptr;
  double* Jcog = nullptr;
  const double* Fext = nullptr;
  const double* Fjnt = nullptr;
  double* qdd = nullptr;
  /* whole-body plans */
  const DevicePlan* plan = nullptr;   /* nullptr: the model's kinetic-terms plan */
  double* A_BI = nullptr;
  double* Jt = nullptr;
};

int buildKtPlan(RcsbModel* m, const WbSpec* wb, bool vel, DevicePlan** out);
int ktDevice(RcsbModel* m, long N, int qdim, const double* q, const double* qd, double* M,
             double* h, double* Fg, double* E, const KtExtra& x, cudaStream_t stream, int ws);

}   // namespace rcsbhost

#endif
