/*
 * rcsb_wholebody (include/rcsb.h): frames of all bodies + task Jacobians + mass matrix + COG
 * Jacobian of the same states in one call (BASELINE.json configs[4]). Reference per state:
 * RcsGraph_setState (src/RcsCore/Rcs_graph.c:231), ControllerBase::computeJ
 * (src/RcsCore/ControllerBase.cpp:776), RcsGraph_computeMassMatrix (src/RcsCore/Rcs_dynamics.c:610).
 */
#include "rcsb_priv.h"

using namespace rcsbhost;

extern "C" int rcsb_wholebody(const RcsbModel* model, long N, int qdim, const double* q, int nSel,
                              const int* bodySel, double* A_BI, int nJac, const RcsbRelJac* jac,
                              double* J, double* M, double* Jcog, unsigned flags, void* stream)
{
  int rc = checkModel(model, N, qdim, q);
  if (rc != RCSB_OK)
  {
    return rc;
  }
  if ((J != nullptr) != (nJac > 0 && jac != nullptr))
  {
    rcsb_set_error("rcsb_wholebody: task list and J must be given together");
    return RCSB_EINVAL;
  }
  if (A_BI)
  {
    rc = rcsb_fk(model, N, qdim, q, nullptr, nSel, bodySel, A_BI, nullptr, nullptr, nullptr, nullptr,
                 flags, stream);
    if (rc != RCSB_OK)
    {
      return rc;
    }
  }
  if (J)
  {
    rc = rcsb_task_jacobians(model, N, qdim, q, nJac, jac, J, flags, stream);
    if (rc != RCSB_OK)
    {
      return rc;
    }
  }
  if (M || Jcog)
  {
    rc = rcsb_kinetic_terms_cog(model, N, qdim, q, nullptr, M, nullptr, nullptr, nullptr, nullptr, Jcog,
                                flags, stream);
  }
  return rc;
}
