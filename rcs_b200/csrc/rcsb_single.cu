/*
 * Single-state reference API (include/rcsb.h, bottom): the reference's own entry points
 *   RcsGraph_setState                  src/RcsCore/Rcs_graph.c:231
 *   RcsGraph_computeForwardKinematics  src/RcsCore/Rcs_graph.c:301
 *   RcsGraph_bodyPointJacobian         src/RcsCore/Rcs_kinematics.c:200
 *   RcsGraph_rotationJacobian          src/RcsCore/Rcs_kinematics.c:416
 *   RcsGraph_computeKineticTerms       src/RcsCore/Rcs_dynamics.c:439
 *   Rcs_directDynamics                 src/RcsCore/Rcs_dynamics.c:673
 *   RcsGraph_computeMassMatrix         src/RcsCore/Rcs_dynamics.c:610
 *   RcsGraph_computeGravityTorque      src/RcsCore/Rcs_dynamics.c:576
 *   RcsGraph_COG / RcsGraph_COGJacobian   src/RcsCore/Rcs_kinematics.c:904, :973
 *   RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian   src/RcsCore/Rcs_kinematics.c:566, :789
 * as N = 1 calls into the batched kernels, so that code (and tests) written against the
 * reference run unchanged on this library. Nothing is evaluated on the host: the wrappers
 * only move the state in and the results back into the RcsGraph / MatNd structures.
 *
 * The reference keeps "the state of the last forward kinematics" in the graph (A_BI, A_JI,
 * BODY->omega) and its Jacobian / dynamics functions read it. Here the graph's device cache
 * remembers the joint state of the last forward-kinematics call instead, and the Jacobian /
 * dynamics wrappers evaluate at that state.
 *
 * The flattened device model is created on first use (current CUDA device). Every call
 * re-flattens the graph on the host and rebuilds the device model if the mechanism (transforms,
 * mass properties, limits, constraints, couplings) has been edited since, so the wrappers read
 * the graph on every call like the reference does. The cache is released by RcsGraph_destroy()
 * or rcsb_graph_release_device_cache() (graphs created by another library).
 */
#include "rcsb_priv.h"

using namespace rcsbhost;

extern "C" int rcsb_flatten(const RcsGraph* graph, void** blobOut, size_t* bytesOut);

namespace
{

struct GraphCache
{
  RcsbModel* model = nullptr;
  std::vector<double> q, qd;      /* state of the last forward kinematics (dof each) */
  bool haveVel = false;
  bool coupled = true;            /* last FK applied the coupling rules */
  std::vector<double> frames, jframes, xdot, omega, qOut, qdOut;
};

/* graph -> device cache. A side table, not RcsGraph::userData: graphs created by the reference
 * library (same struct layout) may carry their own user data. */
std::mutex g_cacheMutex;
std::map<const RcsGraph*, GraphCache*> g_caches;

/* true if two flat models describe the same mechanism: everything but the reference state
 * (RCSB_A_JNT_QREF, the joint values the graph held when it was flattened; the single-state
 * wrappers always pass full state vectors and never read it) */
bool sameMechanism(const std::vector<char>& a, const char* b, size_t bytes)
{
  if (a.size() != bytes)
  {
    return false;
  }
  const RcsbFlatHeader* h = reinterpret_cast<const RcsbFlatHeader*>(a.data());
  const size_t q0 = (size_t) h->off[RCSB_A_JNT_QREF], q1 = q0 + sizeof(double) * (size_t) h->dof;
  return memcmp(a.data(), b, q0) == 0 && memcmp(a.data() + q1, b + q1, bytes - q1) == 0;
}

GraphCache* cacheOf(RcsGraph* self)
{
  GraphCache* c = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_cacheMutex);
    GraphCache*& slot = g_caches[self];
    if (!slot)
    {
      slot = new GraphCache;
    }
    c = slot;
  }
  if (c->model)
  {
    /* the reference reads the graph on every call: a caller may have edited A_BP, masses, joint
     * limits, constraints ... since the device model was made. Re-flatten (a few microseconds
     * of host work) and rebuild the device model when the mechanism has changed. */
    void* blob = nullptr;
    size_t bytes = 0;
    if (rcsb_flatten(self, &blob, &bytes) != 0)
    {
      return nullptr;
    }
    const bool same = sameMechanism(c->model->host, static_cast<const char*>(blob), bytes);
    free(blob);
    if (!same)
    {
      rcsb_model_destroy(c->model);
      c->model = nullptr;
    }
  }
  if (!c->model)
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
    {
      cudaGetLastError();
      dev = 0;
    }
    if (rcsb_model_create(self, dev, &c->model) != RCSB_OK)
    {
      c->model = nullptr;
      return nullptr;
    }
    if (c->q.size() != self->dof)
    {
      c->q.assign(self->q->ele, self->q->ele + self->dof);
      c->qd.assign(self->dof, 0.0);
    }
  }
  return c;
}

/* One forward-kinematics pass at (q, qd) [dof each], results into the graph. */
bool runFk(RcsGraph* self, GraphCache* c, const double* q, const double* qd, bool noCoupling,
           bool writeBack, const RcsBody* only = nullptr, bool subTree = false)
{
  const unsigned int dof = self->dof, nb = RcsGraph_numBodies(self);
  c->frames.resize((size_t) nb * 12);
  c->jframes.resize((size_t) (dof > 0 ? dof : 1) * 12);
  c->qOut.resize(dof > 0 ? dof : 1);
  if (qd)
  {
    c->xdot.resize((size_t) nb * 3);
    c->omega.resize((size_t) nb * 3);
    c->qdOut.resize(dof > 0 ? dof : 1);
  }
  const int rc = fkCommon(c->model, 1, (int) dof, q, qd, 0, nullptr, c->frames.data(),
                          dof > 0 ? c->jframes.data() : nullptr, qd ? c->xdot.data() : nullptr,
                          qd ? c->omega.data() : nullptr, c->qOut.data(), 0, nullptr, nullptr,
                          nullptr, nullptr, qd ? c->qdOut.data() : nullptr, noCoupling,
                          RCSB_HOST_PTRS, nullptr);
  if (rc != RCSB_OK)
  {
    RCSB_FATAL("forward kinematics failed: %s", rcsb_last_error());
    return false;
  }

  if (only)
  {
    /* RcsGraph_computeBodyKinematics: only this body (or its subtree) takes the new frames;
     * Jacobian columns are not renumbered (Rcs_graph.c:327-352 passes nJ = NULL) and the
     * state the Jacobian / dynamics wrappers evaluate at stays that of the last full pass */
    unsigned int idx = 0;
    RCSGRAPH_TRAVERSE_BODIES(self)
    {
      bool take = (BODY == only);
      for (const RcsBody* a = BODY->parent; a && subTree && !take; a = a->parent)
      {
        take = (a == only);
      }
      if (take)
      {
        memcpy(BODY->A_BI, &c->frames[(size_t) idx * 12], sizeof(HTr));
        if (qd)
        {
          memcpy(BODY->x_dot, &c->xdot[(size_t) idx * 3], 3 * sizeof(double));
          memcpy(BODY->omega, &c->omega[(size_t) idx * 3], 3 * sizeof(double));
        }
        RCSBODY_TRAVERSE_JOINTS(BODY)
        {
          memcpy(&JNT->A_JI, &c->jframes[(size_t) JNT->jointIndex * 12], sizeof(HTr));
        }
      }
      ++idx;
    }
    return true;
  }

  unsigned int b = 0, nJ = 0;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    memcpy(BODY->A_BI, &c->frames[(size_t) b * 12], sizeof(HTr));
    if (qd)
    {
      memcpy(BODY->x_dot, &c->xdot[(size_t) b * 3], 3 * sizeof(double));
      memcpy(BODY->omega, &c->omega[(size_t) b * 3], 3 * sizeof(double));
    }
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      memcpy(&JNT->A_JI, &c->jframes[(size_t) JNT->jointIndex * 12], sizeof(HTr));
      /* RcsGraph_internalBodyKinematics renumbers the Jacobian columns while it walks
       * (Rcs_graph.c:131-140) */
      JNT->jacobiIndex = JNT->constrained ? -1 : (int) nJ++;
    }
    ++b;
  }
  if (self->nJ != nJ)
  {
    self->nJ = nJ;
  }

  c->q.assign(c->qOut.begin(), c->qOut.begin() + dof);
  c->haveVel = (qd != nullptr);
  c->coupled = !noCoupling;
  if (qd)
  {
    c->qd.assign(c->qdOut.begin(), c->qdOut.begin() + dof);
  }
  if (writeBack)
  {
    memcpy(self->q->ele, c->q.data(), dof * sizeof(double));
    if (qd)
    {
      memcpy(self->q_dot->ele, c->qd.data(), dof * sizeof(double));
    }
  }
  return true;
}

}   // namespace

extern "C" void rcsb_graph_release_device_cache(RcsGraph* self)
{
  GraphCache* c = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_cacheMutex);
    auto it = g_caches.find(self);
    if (it != g_caches.end())
    {
      c = it->second;
      g_caches.erase(it);
    }
  }
  if (c)
  {
    rcsb_model_destroy(c->model);
    delete c;
  }
}

extern "C" bool RcsGraph_setState(RcsGraph* self, const MatNd* q_, const MatNd* q_dot_)
{
  if (!self)
  {
    RCSB_FATAL("RcsGraph_setState: graph is NULL");
    return false;
  }
  const unsigned int dof = self->dof;

  if (q_)
  {
    if (q_->m == self->nJ && q_->m != dof)
    {
      RcsGraph_stateVectorFromIK(self, q_, self->q);
    }
    else if (q_->m == dof)
    {
      if (self->q != q_)
      {
        memcpy(self->q->ele, q_->ele, dof * sizeof(double));
      }
    }
    else
    {
      RCSB_FATAL("q has wrong dimensions: [%d x %d], dof is %d, nJ is %d", q_->m, q_->n, dof,
                 self->nJ);
      return false;
    }
  }
  if (q_dot_)
  {
    if (q_dot_->m == self->nJ && q_dot_->m != dof)
    {
      RcsGraph_stateVectorFromIK(self, q_dot_, self->q_dot);
    }
    else if (q_dot_->m == dof)
    {
      if (self->q_dot != q_dot_)
      {
        memcpy(self->q_dot->ele, q_dot_->ele, dof * sizeof(double));
      }
    }
    else
    {
      RCSB_FATAL("q_dot has wrong dimensions: [%d x %d], dof is %d, nJ is %d", q_dot_->m,
                 q_dot_->n, dof, self->nJ);
      return false;
    }
  }

  GraphCache* c = cacheOf(self);
  if (!c)
  {
    RCSB_FATAL("RcsGraph_setState: %s", rcsb_last_error());
    return false;
  }

  std::vector<double> qIn(self->q->ele, self->q->ele + dof);
  if (!runFk(self, c, qIn.data(), q_dot_ ? self->q_dot->ele : nullptr, false, true))
  {
    return false;
  }

  /* "changed" as in RcsGraph_updateSlaveJoints (Rcs_graph.c:2756): a slave joint had to be
   * moved by more than 1e-8 */
  bool changed = false;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (JNT->coupledTo)
      {
        const double d = c->q[(size_t) JNT->jointIndex] - qIn[(size_t) JNT->jointIndex];
        changed = changed || (d > 1.0e-8) || (d < -1.0e-8);
      }
    }
  }
  return changed;
}

extern "C" void RcsGraph_computeForwardKinematics(RcsGraph* self, const MatNd* q,
                                                  const MatNd* q_dot)
{
  if (!self)
  {
    RCSB_FATAL("RcsGraph_computeForwardKinematics: graph is NULL");
    return;
  }
  if ((q && q->m != self->dof) || (q_dot && q_dot->m != self->dof))
  {
    RCSB_FATAL("RcsGraph_computeForwardKinematics: q / q_dot must have dof = %d rows", self->dof);
    return;
  }
  GraphCache* c = cacheOf(self);
  if (!c)
  {
    RCSB_FATAL("RcsGraph_computeForwardKinematics: %s", rcsb_last_error());
    return;
  }
  /* no slave update, always with velocities, graph->q untouched (Rcs_graph.c:301-322) */
  runFk(self, c, q ? q->ele : self->q->ele, q_dot ? q_dot->ele : self->q_dot->ele, true, false);
}

extern "C" void RcsGraph_computeBodyKinematics(RcsGraph* self, RcsBody* bdy, const MatNd* q,
                                               const MatNd* q_dot, bool subTree)
{
  if (!self || !bdy)
  {
    RCSB_FATAL("RcsGraph_computeBodyKinematics: NULL argument");
    return;
  }
  if ((q && q->m != self->dof) || (q_dot && q_dot->m != self->dof))
  {
    RCSB_FATAL("RcsGraph_computeBodyKinematics: q / q_dot must have dof = %d rows", self->dof);
    return;
  }
  GraphCache* c = cacheOf(self);
  if (!c)
  {
    RCSB_FATAL("RcsGraph_computeBodyKinematics: %s", rcsb_last_error());
    return;
  }
  runFk(self, c, q ? q->ele : self->q->ele, q_dot ? q_dot->ele : self->q_dot->ele, true, false, bdy,
        subTree);
}

namespace
{

void jacobianCommon(const RcsGraph* cself, const RcsBody* body, const double* B_pt,
                    double A_BI[3][3], MatNd* J, bool rotation, const char* who)
{
  RcsGraph* self = const_cast<RcsGraph*>(cself);
  if (!self || !J)
  {
    RCSB_FATAL("%s: NULL argument", who);
    return;
  }
  if (J->size < 3 * self->nJ)
  {
    RCSB_FATAL("%s: Jacobian has room for %u doubles, 3 x %u needed", who, J->size, self->nJ);
    return;
  }
  MatNd_reshapeAndSetZero(J, 3, self->nJ);
  if (!body || self->nJ == 0)
  {
    return;   /* NULL body = world frame: zero Jacobian (Rcs_kinematics.c:188-191, :424-427) */
  }
  GraphCache* c = cacheOf(self);
  if (!c)
  {
    RCSB_FATAL("%s: %s", who, rcsb_last_error());
    return;
  }
  const int eff = RcsGraph_bodyIndex(self, body);
  if (eff < 0)
  {
    RCSB_FATAL("%s: body \"%s\" does not belong to the graph", who, body->name ? body->name : "?");
    return;
  }
  const int bodySel[1] = {eff};
  const int rc = fkCommon(c->model, 1, (int) self->dof, c->q.data(), nullptr, 1, bodySel, nullptr,
                          nullptr, nullptr, nullptr, nullptr, 1, bodySel, B_pt,
                          rotation ? nullptr : J->ele, rotation ? J->ele : nullptr, nullptr,
                          !c->coupled, RCSB_HOST_PTRS, nullptr);
  if (rc != RCSB_OK)
  {
    RCSB_FATAL("%s: %s", who, rcsb_last_error());
    return;
  }
  if (A_BI)
  {
    /* J <- A_BI J: the columns expressed in the frame with rotation matrix A_BI
     * (MatNd_rotateSelf, Rcs_kinematics.c:160-163, :453-456) */
    const unsigned int n = self->nJ;
    for (unsigned int k = 0; k < n; ++k)
    {
      const double v[3] = {J->ele[k], J->ele[n + k], J->ele[2 * n + k]};
      for (int r = 0; r < 3; ++r)
      {
        J->ele[r * n + k] = A_BI[r][0] * v[0] + A_BI[r][1] * v[1] + A_BI[r][2] * v[2];
      }
    }
  }
}

}   // namespace

/* Jacobian of a point given in world coordinates that moves with the body
 * (Rcs_kinematics.c:178): the point is taken into the body frame of the last forward
 * kinematics; NULL point = body origin. */
extern "C" void RcsGraph_worldPointJacobian(const RcsGraph* self, const RcsBody* body,
                                            const double I_pt[3], double A_BI[3][3], MatNd* J)
{
  if (!body || !I_pt)
  {
    jacobianCommon(self, body, nullptr, A_BI, J, false, "RcsGraph_worldPointJacobian");
    return;
  }
  const HTr* A = body->A_BI;
  const double d[3] = {I_pt[0] - A->org[0], I_pt[1] - A->org[1], I_pt[2] - A->org[2]};
  const double B_pt[3] = {A->rot[0][0] * d[0] + A->rot[0][1] * d[1] + A->rot[0][2] * d[2],
                          A->rot[1][0] * d[0] + A->rot[1][1] * d[1] + A->rot[1][2] * d[2],
                          A->rot[2][0] * d[0] + A->rot[2][1] * d[1] + A->rot[2][2] * d[2]};
  jacobianCommon(self, body, B_pt, A_BI, J, false, "RcsGraph_worldPointJacobian");
}

extern "C" void RcsGraph_bodyPointJacobian(const RcsGraph* self, const RcsBody* body,
                                           const double B_pt[3], double A_BI[3][3], MatNd* J)
{
  jacobianCommon(self, body, B_pt, A_BI, J, false, "RcsGraph_bodyPointJacobian");
}

extern "C" void RcsGraph_rotationJacobian(const RcsGraph* self, const RcsBody* body,
                                          double A_BI[3][3], MatNd* J)
{
  jacobianCommon(self, body, nullptr, A_BI, J, true, "RcsGraph_rotationJacobian");
}

extern "C" double RcsGraph_computeKineticTerms(const RcsGraph* cself, MatNd* M, MatNd* h,
                                               MatNd* F_gravity)
{
  RcsGraph* self = const_cast<RcsGraph*>(cself);
  if (!self)
  {
    RCSB_FATAL("RcsGraph_computeKineticTerms: graph is NULL");
    return 0.0;
  }
  const unsigned int n = self->nJ;
  if ((M && M->size < n * n) || (h && h->size < n) || (F_gravity && F_gravity->size < n))
  {
    RCSB_FATAL("RcsGraph_computeKineticTerms: an output matrix is too small for nJ = %u", n);
    return 0.0;
  }
  GraphCache* c = cacheOf(self);
  if (!c)
  {
    RCSB_FATAL("RcsGraph_computeKineticTerms: %s", rcsb_last_error());
    return 0.0;
  }
  if (M)
  {
    MatNd_reshapeAndSetZero(M, n, n);
  }
  if (h)
  {
    MatNd_reshapeAndSetZero(h, n, 1);
  }
  if (F_gravity)
  {
    MatNd_reshapeAndSetZero(F_gravity, n, 1);
  }
  double E = 0.0;
  /* the reference reads graph->q_dot and the body velocities of the last FK (Rcs_dynamics.c:
   * 449-451, :346-377); both follow from the cached joint state */
  const int rc = rcsb_kinetic_terms(c->model, 1, (int) self->dof, c->q.data(),
                                    c->haveVel ? c->qd.data() : self->q_dot->ele,
                                    M ? M->ele : nullptr, h ? h->ele : nullptr,
                                    F_gravity ? F_gravity->ele : nullptr, &E, RCSB_HOST_PTRS,
                                    nullptr);
  if (rc != RCSB_OK)
  {
    RCSB_FATAL("RcsGraph_computeKineticTerms: %s", rcsb_last_error());
    return 0.0;
  }
  return E;
}

extern "C" void RcsGraph_computeMassMatrix(const RcsGraph* cself, MatNd* M)
{
  RcsGraph* self = const_cast<RcsGraph*>(cself);
  GraphCache* c = (self && M) ? cacheOf(self) : nullptr;
  if (!c || M->size < self->nJ * self->nJ)
  {
    RCSB_FATAL("RcsGraph_computeMassMatrix: %s", c ? "M is too small" : rcsb_last_error());
    return;
  }
  MatNd_reshapeAndSetZero(M, self->nJ, self->nJ);
  if (rcsb_kinetic_terms(c->model, 1, (int) self->dof, c->q.data(), nullptr, M->ele, nullptr,
                         nullptr, nullptr, RCSB_HOST_PTRS, nullptr) != RCSB_OK)
  {
    RCSB_FATAL("RcsGraph_computeMassMatrix: %s", rcsb_last_error());
  }
}

extern "C" void RcsGraph_computeGravityTorque(const RcsGraph* cself, MatNd* T_gravity)
{
  RcsGraph* self = const_cast<RcsGraph*>(cself);
  GraphCache* c = (self && T_gravity) ? cacheOf(self) : nullptr;
  if (!c || T_gravity->size < self->nJ)
  {
    RCSB_FATAL("RcsGraph_computeGravityTorque: %s", c ? "T_gravity is too small" : rcsb_last_error());
    return;
  }
  MatNd_reshapeAndSetZero(T_gravity, self->nJ, 1);
  /* J_cog^T (0, 0, -m g) = sum_b m_b JT_b^T (0, 0, -g): the F_gravity of the kinetic terms */
  if (rcsb_kinetic_terms(c->model, 1, (int) self->dof, c->q.data(), nullptr, nullptr, nullptr,
                         T_gravity->ele, nullptr, RCSB_HOST_PTRS, nullptr) != RCSB_OK)
  {
    RCSB_FATAL("RcsGraph_computeGravityTorque: %s", rcsb_last_error());
  }
}

extern "C" double RcsGraph_COG(const RcsGraph* cself, double r_cog[3])
{
  RcsGraph* self = const_cast<RcsGraph*>(cself);
  GraphCache* c = (self && r_cog) ? cacheOf(self) : nullptr;
  if (!c)
  {
    RCSB_FATAL("RcsGraph_COG: %s", rcsb_last_error());
    return 0.0;
  }
  r_cog[0] = r_cog[1] = r_cog[2] = 0.0;
  if (rcsb_cog(c->model, 1, (int) self->dof, c->q.data(), r_cog, nullptr, RCSB_HOST_PTRS,
               nullptr) != RCSB_OK)
  {
    RCSB_FATAL("RcsGraph_COG: %s", rcsb_last_error());
    return 0.0;
  }
  double m = 0.0;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    m += BODY->m > 0.0 ? BODY->m : 0.0;
  }
  return m;
}

extern "C" void RcsGraph_COGJacobian(const RcsGraph* cself, MatNd* J_cog)
{
  RcsGraph* self = const_cast<RcsGraph*>(cself);
  GraphCache* c = (self && J_cog) ? cacheOf(self) : nullptr;
  if (!c || J_cog->size < 3 * self->nJ)
  {
    RCSB_FATAL("RcsGraph_COGJacobian: %s", c ? "J_cog is too small" : rcsb_last_error());
    return;
  }
  MatNd_reshapeAndSetZero(J_cog, 3, self->nJ);
  if (self->nJ > 0 && rcsb_cog(c->model, 1, (int) self->dof, c->q.data(), nullptr, J_cog->ele,
                               RCSB_HOST_PTRS, nullptr) != RCSB_OK)
  {
    RCSB_FATAL("RcsGraph_COGJacobian: %s", rcsb_last_error());
  }
}

namespace
{

void relJacobian(const RcsGraph* cself, int kind, const RcsBody* effector, const RcsBody* refBdy,
                 const RcsBody* refFrame, MatNd* J, const char* who)
{
  RcsGraph* self = const_cast<RcsGraph*>(cself);
  GraphCache* c = (self && J) ? cacheOf(self) : nullptr;
  if (!c || J->size < 3 * self->nJ)
  {
    RCSB_FATAL("%s: %s", who, c ? "J is too small" : rcsb_last_error());
    return;
  }
  MatNd_reshapeAndSetZero(J, 3, self->nJ);
  if (self->nJ == 0)
  {
    return;
  }
  RcsbRelJac r;
  r.kind = kind;
  r.effector = effector ? RcsGraph_bodyIndex(self, effector) : -1;
  r.refBdy = refBdy ? RcsGraph_bodyIndex(self, refBdy) : -1;
  r.refFrame = refFrame ? RcsGraph_bodyIndex(self, refFrame) : -1;
  if (rcsb_task_jacobians(c->model, 1, (int) self->dof, c->q.data(), 1, &r, J->ele, RCSB_HOST_PTRS,
                          nullptr) != RCSB_OK)
  {
    RCSB_FATAL("%s: %s", who, rcsb_last_error());
  }
}

}   // namespace

extern "C" void RcsGraph_3dPosJacobian(const RcsGraph* self, const RcsBody* effector,
                                       const RcsBody* refBdy, const RcsBody* refFrame, MatNd* J)
{
  relJacobian(self, RCSB_JAC_POS, effector, refBdy, refFrame, J, "RcsGraph_3dPosJacobian");
}

extern "C" void RcsGraph_3dOmegaJacobian(const RcsGraph* self, const RcsBody* effector,
                                         const RcsBody* refBdy, const RcsBody* refFrame, MatNd* J)
{
  relJacobian(self, RCSB_JAC_OMEGA, effector, refBdy, refFrame, J, "RcsGraph_3dOmegaJacobian");
}

extern "C" double Rcs_directDynamics(const RcsGraph* cgraph, const MatNd* F_ext, const MatNd* F_jnt,
                                     MatNd* q_ddot)
{
  RcsGraph* self = const_cast<RcsGraph*>(cgraph);
  GraphCache* c = (self && q_ddot) ? cacheOf(self) : nullptr;
  const unsigned int n = self ? self->nJ : 0;
  if (!c || q_ddot->size < n || (F_ext && (F_ext->m != n || F_ext->n != 1)) ||
      (F_jnt && (F_jnt->m != n || F_jnt->n != 1)))
  {
    RCSB_FATAL("Rcs_directDynamics: %s", c ? "F_ext / F_jnt / q_ddot must be nJ x 1" : rcsb_last_error());
    return 0.0;
  }
  MatNd_reshapeAndSetZero(q_ddot, n, 1);
  double E = 0.0;
  if (rcsb_direct_dynamics(c->model, 1, (int) self->dof, c->q.data(),
                           c->haveVel ? c->qd.data() : self->q_dot->ele, F_ext ? F_ext->ele : nullptr,
                           F_jnt ? F_jnt->ele : nullptr, q_ddot->ele, &E, RCSB_HOST_PTRS,
                           nullptr) != RCSB_OK)
  {
    RCSB_FATAL("Rcs_directDynamics: %s", rcsb_last_error());
    return 0.0;
  }
  return E;
}
