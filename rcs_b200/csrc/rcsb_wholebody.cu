/*
 * rcsb_wholebody (include/rcsb.h): frames of the selected bodies + task Jacobians + mass matrix +
 * COG Jacobian of the same states (BASELINE.json configs[4]). Reference per state:
 * RcsGraph_setState (src/RcsCore/Rcs_graph.c:231), ControllerBase::computeJ over XYZ / XYZABC /
 * COG tasks (src/RcsCore/ControllerBase.cpp:776; RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian,
 * Rcs_kinematics.c:566, :789), RcsGraph_computeMassMatrix (src/RcsCore/Rcs_dynamics.c:610).
 *
 * One pass of the kinetic-terms kernels (rcsb_kinetic.cu) does all of it: the walk kernel is the
 * one tree walk per state — it stores the frames (A_BI) and leaves, next to the joint columns and
 * mass runs of its record, the frames of the bodies the tasks refer to; the assembly kernel
 * turns the record into M, the COG Jacobian and the task rows. Two launches per pass, no other
 * intermediate in HBM (round 1: one pruned fkKernel per task body + combineKernel passes + the
 * kinetic pair, 57 launches per 1M states).
 */
#include "rcsb_priv.h"

using namespace rcsbhost;

/* task list -> slots of the bodies involved and the combination recipe of every task (the rules
 * of RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian, as in rcsb_task_jacobians) */
static int resolveTasks(const RcsbModel* m, int nJac, const RcsbRelJac* jac, WbSpec& spec)
{
  const int nb = m->hdr.nb;
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
  std::vector<int> slotOf((size_t) nb, -1);
  auto slot = [&](int b) {
    if (b < 0)
    {
      return -1;
    }
    if (slotOf[(size_t) b] < 0)
    {
      slotOf[(size_t) b] = (int) spec.taskBodies.size();
      spec.taskBodies.push_back(b);
    }
    return slotOf[(size_t) b];
  };
  for (int t = 0; t < nJac; ++t)
  {
    const RcsbRelJac& r = jac[t];
    if ((r.kind != RCSB_JAC_POS && r.kind != RCSB_JAC_OMEGA) || r.effector >= nb || r.refBdy >= nb ||
        r.refFrame >= nb || r.effector < -1 || r.refBdy < -1 || r.refFrame < -1)
    {
      rcsb_set_error("rcsb_wholebody: task %d is invalid", t);
      return RCSB_EINVAL;
    }
    int mode = 0, u2 = slot(r.effector), u1 = -1, u3 = -1, uA = -1;
    if (r.kind == RCSB_JAC_POS)
    {
      if (r.refBdy < 0)
      {
        uA = slot(r.refFrame);
      }
      else
      {
        mode = 1;
        u1 = slot(r.refBdy);
        u3 = slot(r.refFrame);
        uA = (r.refFrame >= 0 && r.refFrame != r.refBdy) ? slot(r.refFrame) : u1;
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
        uA = slot(r.refBdy);
      }
      else
      {
        mode = 1;
        u1 = slot(r.refBdy);
        uA = slot(r.refFrame);
      }
    }
    const int rec[8] = {r.kind, mode, u2, u1, u3, uA, 0, 0};
    spec.jacTasks.insert(spec.jacTasks.end(), rec, rec + 8);
  }
  return RCSB_OK;
}

extern "C" int rcsb_wholebody(const RcsbModel* cm, long N, int qdim, const double* q, int nSel,
                              const int* bodySel, double* A_BI, int nJac, const RcsbRelJac* jac,
                              double* J, double* M, double* Jcog, unsigned flags, void* stream)
{
  RcsbModel* m = const_cast<RcsbModel*>(cm);
  int rc = checkModel(m, N, qdim, q);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  if ((J != nullptr) != (nJac > 0 && jac != nullptr))
  {
    rcsb_set_error("rcsb_wholebody: task list and J must be given together");
    return RCSB_EINVAL;
  }
  if (nSel < 0 || (nSel > 0 && !bodySel))
  {
    rcsb_set_error("rcsb_wholebody: invalid body selection");
    return RCSB_EINVAL;
  }

  WbSpec spec;
  spec.nSel = nSel;
  spec.bodySel = bodySel;
  if (J)
  {
    rc = resolveTasks(m, nJac, jac, spec);
    if (rc != RCSB_OK)
    {
      return rc;
    }
  }
  spec.key.append(reinterpret_cast<const char*>(&nSel), sizeof(int));
  if (nSel > 0)
  {
    spec.key.append(reinterpret_cast<const char*>(bodySel), sizeof(int) * (size_t) nSel);
  }
  if (J)
  {
    spec.key.append(reinterpret_cast<const char*>(jac), sizeof(RcsbRelJac) * (size_t) nJac);
  }

  DevicePlan* plan = nullptr;
  rc = buildKtPlan(m, &spec, false, &plan);
  if (rc != RCSB_OK || N == 0)
  {
    return rc;
  }

  KtExtra x;
  x.plan = plan;
  if (!(flags & RCSB_HOST_PTRS))
  {
    x.A_BI = A_BI;
    x.Jt = J;
    x.Jcog = Jcog;
    return ktDevice(m, N, qdim, q, nullptr, M, nullptr, nullptr, nullptr, x, static_cast<cudaStream_t>(stream), 0);
  }

  const size_t nJ = (size_t) m->hdr.nJ;
  std::vector<StagedArray> arr;
  enum { I_Q, I_A, I_J, I_M, I_JC };
  arr.push_back({q, nullptr, (size_t) qdim, nullptr});
  arr.push_back({nullptr, A_BI, A_BI ? (size_t) plan->nSel * 12 : 0, nullptr});
  arr.push_back({nullptr, J, J ? (size_t) nJac * 3 * nJ : 0, nullptr});
  arr.push_back({nullptr, M, M ? nJ * nJ : 0, nullptr});
  arr.push_back({nullptr, Jcog, Jcog ? 3 * nJ : 0, nullptr});
  return runStaged(m, N, arr, {}, [&](long n, cudaStream_t st, int slot) {
    auto dv = [&](int i, const void* host) { return host ? arr[(size_t) i].dev : nullptr; };
    KtExtra d;
    d.plan = plan;
    d.A_BI = dv(I_A, A_BI);
    d.Jt = dv(I_J, J);
    d.Jcog = dv(I_JC, Jcog);
    return ktDevice(m, n, qdim, dv(I_Q, q), nullptr, dv(I_M, M), nullptr, nullptr, nullptr, d, st, slot);
  });
}
