/*
 * ref_api — C entry points that drive the UNMODIFIED reference (HRI-EU/Rcs RcsCore,
 * compiled from /root/reference by oracle/ref_build/Makefile) over batches of states.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity anchor ("oracle/_ref") and the
 * "reference" CPU baseline of bench.py. It contains no algorithm of its own: every
 * number it returns is produced by a reference function, one state at a time, exactly as
 * the reference's own apps call them:
 *   RcsGraph_create / RcsGraph_clone         src/RcsCore/Rcs_graph.c:398, :2193
 *   RcsGraph_setState                        src/RcsCore/Rcs_graph.c:231
 *   RcsGraph_bodyPointJacobian               src/RcsCore/Rcs_kinematics.c:200
 *   RcsGraph_rotationJacobian                src/RcsCore/Rcs_kinematics.c:416
 *   RcsGraph_COG / RcsGraph_COGJacobian      src/RcsCore/Rcs_kinematics.c:904, :973
 *   RcsGraph_computeKineticTerms             src/RcsCore/Rcs_dynamics.c:439
 *   RcsGraph_computeMassMatrix               src/RcsCore/Rcs_dynamics.c:610
 *   ControllerBase::computeDX/computeJ/...   src/RcsCore/ControllerBase.cpp:776-1332
 *   IkSolverRMR::solveRightInverse           src/RcsCore/IkSolverRMR.cpp:403
 * The loop of ref_ik_rmr mirrors examples/ExampleIK.cpp:154-177.
 *
 * Threading: batches are cut into contiguous slices, one std::thread per slice, one
 * RcsGraph_clone (or ControllerBase copy + IkSolverRMR) per thread.
 */
#include <Rcs_typedef.h>
#include <Rcs_graph.h>
#include <Rcs_body.h>
#include <Rcs_joint.h>
#include <Rcs_kinematics.h>
#include <Rcs_dynamics.h>
#include <Rcs_macros.h>
#include <Rcs_MatNd.h>
#include <Rcs_Mat3d.h>
#include <Rcs_Vec3d.h>
#include <Rcs_resourcePath.h>
#include <ControllerBase.h>
#include <IkSolverRMR.h>
#include <IkSolverPrioRMR.h>
#include <IkSolverConstraintRMR.h>
#include <Rcs_URDFParser.h>
#include <strings.h>

#include <unistd.h>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace
{

int g_threads = 1;

/*
 * Timing of a batch call. Every slice runs on its own thread with its own RcsGraph_clone (or
 * controller copy). A slice calls sliceReady() once its clone and work matrices exist: the
 * call returns when every slice has got there, and the clock of the batch starts at that
 * moment (the reference's apps clone once per thread, outside their loops). sliceDone() stops
 * the slice's clock before it frees its clone. The batch time is first start -> last stop.
 * Slices that never call sliceReady() are timed over their whole lifetime, clone included.
 */
struct SliceClock
{
  std::mutex mtx;
  std::condition_variable cv;
  int expected = 0;
  int arrived = 0;
  bool started = false;
  std::chrono::steady_clock::time_point t0, t1;
};

SliceClock* g_clock = NULL;   /* one batch call at a time (the Python wrapper is serial) */

void sliceReady()
{
  SliceClock* c = g_clock;
  std::unique_lock<std::mutex> lk(c->mtx);
  if (++c->arrived == c->expected)
  {
    c->t0 = std::chrono::steady_clock::now();
    c->started = true;
    c->cv.notify_all();
  }
  else
  {
    c->cv.wait(lk, [c]() { return c->started; });
  }
}

void sliceDone()
{
  SliceClock* c = g_clock;
  auto now = std::chrono::steady_clock::now();
  std::lock_guard<std::mutex> lk(c->mtx);
  if (now > c->t1)
  {
    c->t1 = now;
  }
}

template <class F>
double parallelSlices(long N, F body)
{
  int nt = g_threads < 1 ? 1 : g_threads;
  if ((long) nt > N)
  {
    nt = N > 0 ? (int) N : 1;
  }
  SliceClock clock;
  clock.expected = nt;
  g_clock = &clock;
  auto t0 = std::chrono::steady_clock::now();
  clock.t1 = t0;
  if (nt == 1)
  {
    body(0, 0L, N);
  }
  else
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
    {
      long lo = N * t / nt, hi = N * (t + 1) / nt;
      th.emplace_back([=]() { body(t, lo, hi); });
    }
    for (auto& x : th)
    {
      x.join();
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  g_clock = NULL;
  if (clock.started)
  {
    return std::chrono::duration<double>(clock.t1 - clock.t0).count();
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

std::vector<RcsBody*> dfsBodies(const RcsGraph* g)
{
  std::vector<RcsBody*> v;
  RCSGRAPH_TRAVERSE_BODIES(g)
  {
    v.push_back(BODY);
  }
  return v;
}

int bodyIndex(const std::vector<RcsBody*>& v, const RcsBody* b)
{
  if (!b)
  {
    return -1;
  }
  for (size_t i = 0; i < v.size(); ++i)
  {
    if (v[i] == b)
    {
      return (int) i;
    }
  }
  return -1;
}

void putDouble(std::ostringstream& o, double x)
{
  char buf[40];
  snprintf(buf, sizeof(buf), "%.17g", x);
  // JSON has no inf/nan; DBL_MAX prints fine.
  o << buf;
}

void putHTr(std::ostringstream& o, const HTr* A)
{
  if (!A)
  {
    o << "null";
    return;
  }
  o << "[";
  for (int i = 0; i < 3; ++i)
  {
    putDouble(o, A->org[i]);
    o << ",";
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
    {
      putDouble(o, A->rot[i][j]);
      if (i * 3 + j < 8)
      {
        o << ",";
      }
    }
  o << "]";
}

void setQ(RcsGraph* g, MatNd* qv, MatNd* qdv, int qdim, const double* q, const double* qd)
{
  MatNd_reshape(qv, qdim, 1);
  memcpy(qv->ele, q, sizeof(double) * qdim);
  if (qd)
  {
    MatNd_reshape(qdv, qdim, 1);
    memcpy(qdv->ele, qd, sizeof(double) * qdim);
  }
  RcsGraph_setState(g, qv, qd ? qdv : NULL);
}

struct IkHandle
{
  Rcs::ControllerBase* controller;
  std::string xml;
};

}   // namespace

extern "C" {

void ref_set_threads(int n)
{
  g_threads = n;
}

void ref_set_log_level(int l)
{
  RcsLogLevel = l;
}

void ref_add_resource_path(const char* p)
{
  Rcs_addResourcePath(p);
}

void* ref_graph_create(const char* xmlFile)
{
  /* URDF robots go through the reference's own URDF loader (Rcs_URDFParser.c:1152) */
  const size_t len = strlen(xmlFile);
  if (len > 5 && strcasecmp(xmlFile + len - 5, ".urdf") == 0)
  {
    return RcsGraph_fromURDFFile(xmlFile);
  }
  return RcsGraph_create(xmlFile);
}

/* RcsGraph_clone (Rcs_graph.c:2193): the copy every worker thread of this driver evaluates on */
void* ref_graph_clone(void* g)
{
  return RcsGraph_clone((RcsGraph*) g);
}

void ref_graph_destroy(void* g)
{
  RcsGraph_destroy((RcsGraph*) g);
}

int ref_graph_dof(void* g)
{
  return (int) ((RcsGraph*) g)->dof;
}

int ref_graph_nj(void* g)
{
  return (int) ((RcsGraph*) g)->nJ;
}

int ref_graph_nbodies(void* g)
{
  return (int) dfsBodies((RcsGraph*) g).size();
}

int ref_graph_check(void* g, int* nErrors, int* nWarnings)
{
  return RcsGraph_check((RcsGraph*) g, nErrors, nWarnings) ? 1 : 0;
}

int ref_body_index(void* g, const char* name)
{
  RcsGraph* graph = (RcsGraph*) g;
  return bodyIndex(dfsBodies(graph), RcsGraph_getBodyByName(graph, name));
}

void ref_free(void* p)
{
  free(p);
}

/* Canonical JSON dump of everything the hot path reads from the loaded graph. */
char* ref_graph_layout_json(void* g)
{
  RcsGraph* graph = (RcsGraph*) g;
  std::vector<RcsBody*> bodies = dfsBodies(graph);
  std::ostringstream o;
  o << "{\"dof\":" << graph->dof << ",\"nJ\":" << graph->nJ
    << ",\"nBodies\":" << bodies.size() << ",\"bodies\":[";

  for (size_t i = 0; i < bodies.size(); ++i)
  {
    RcsBody* b = bodies[i];
    o << (i ? "," : "") << "{\"name\":\"" << b->name << "\",\"parent\":"
      << bodyIndex(bodies, b->parent) << ",\"firstChild\":" << bodyIndex(bodies, b->firstChild)
      << ",\"lastChild\":" << bodyIndex(bodies, b->lastChild)
      << ",\"next\":" << bodyIndex(bodies, b->next) << ",\"prev\":" << bodyIndex(bodies, b->prev)
      << ",\"m\":";
    putDouble(o, b->m);
    o << ",\"rigid_body_joints\":" << (b->rigid_body_joints ? "true" : "false")
      << ",\"physicsSim\":" << b->physicsSim << ",\"A_BP\":";
    putHTr(o, b->A_BP);
    o << ",\"Inertia\":";
    putHTr(o, b->Inertia);
    o << ",\"joints\":[";
    bool first = true;
    for (RcsJoint* j = b->jnt; j; j = j->next)
    {
      o << (first ? "" : ",") << j->jointIndex;
      first = false;
    }
    int nShapes = 0;
    if (b->shape)
    {
      for (RcsShape** s = b->shape; *s; ++s)
      {
        nShapes++;
      }
    }
    o << "],\"nShapes\":" << nShapes << "}";
  }
  o << "],\"joints\":[";

  bool firstJ = true;
  RCSGRAPH_TRAVERSE_JOINTS(graph)
  {
    RcsJoint* j = JNT;
    int bIdx = -1;
    for (size_t i = 0; i < bodies.size() && bIdx < 0; ++i)
      for (RcsJoint* k = bodies[i]->jnt; k; k = k->next)
        if (k == j)
        {
          bIdx = (int) i;
        }
    o << (firstJ ? "" : ",") << "{\"name\":\"" << j->name << "\",\"body\":" << bIdx
      << ",\"type\":" << j->type << ",\"dirIdx\":" << j->dirIdx
      << ",\"jointIndex\":" << j->jointIndex << ",\"jacobiIndex\":" << j->jacobiIndex
      << ",\"constrained\":" << (j->constrained ? "true" : "false")
      << ",\"ctrlType\":" << j->ctrlType
      << ",\"prev\":" << (j->prev ? j->prev->jointIndex : -1)
      << ",\"coupledTo\":" << (j->coupledTo ? j->coupledTo->jointIndex : -1);
    const char* names[] = {"q0", "q_init", "q_min", "q_max", "weightJL", "weightCA",
                           "weightMetric", "maxTorque", "speedLimit", "accLimit", "decLimit"};
    const double vals[] = {j->q0, j->q_init, j->q_min, j->q_max, j->weightJL, j->weightCA,
                           j->weightMetric, j->maxTorque, j->speedLimit, j->accLimit, j->decLimit};
    for (int k = 0; k < 11; ++k)
    {
      o << ",\"" << names[k] << "\":";
      putDouble(o, vals[k]);
    }
    o << ",\"couplingFactors\":[";
    if (j->couplingFactors)
    {
      for (unsigned int k = 0; k < j->couplingFactors->m * j->couplingFactors->n; ++k)
      {
        o << (k ? "," : "");
        putDouble(o, j->couplingFactors->ele[k]);
      }
    }
    o << "],\"A_JP\":";
    putHTr(o, j->A_JP);
    o << "}";
    firstJ = false;
  }
  o << "],\"q\":[";
  for (unsigned int i = 0; i < graph->dof; ++i)
  {
    o << (i ? "," : "");
    putDouble(o, graph->q->ele[i]);
  }
  o << "]}";
  return strdup(o.str().c_str());
}

/* FK over a batch. q: N x qdim (qdim == dof or nJ), qd: NULL or N x qdim. Outputs may be
 * NULL. A_BI: N x nb x 12, A_JI: N x dof x 12, xdot/omega: N x nb x 3, qOut: N x dof. */
double ref_fk(void* g, long N, int qdim, const double* q, const double* qd, double* A_BI,
              double* A_JI, double* xdot, double* omega, double* qOut)
{
  RcsGraph* master = (RcsGraph*) g;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    RcsGraph* graph = RcsGraph_clone(master);
    std::vector<RcsBody*> bodies = dfsBodies(graph);
    const size_t nb = bodies.size();
    const unsigned int dof = graph->dof;
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* qdv = MatNd_create(dof, 1);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, qdv, qdim, q + i * qdim, qd ? qd + i * qdim : NULL);
      if (A_BI)
      {
        for (size_t b = 0; b < nb; ++b)
        {
          memcpy(A_BI + (i * nb + b) * 12, bodies[b]->A_BI, 12 * sizeof(double));
        }
      }
      if (xdot)
      {
        for (size_t b = 0; b < nb; ++b)
        {
          memcpy(xdot + (i * nb + b) * 3, bodies[b]->x_dot, 3 * sizeof(double));
        }
      }
      if (omega)
      {
        for (size_t b = 0; b < nb; ++b)
        {
          memcpy(omega + (i * nb + b) * 3, bodies[b]->omega, 3 * sizeof(double));
        }
      }
      if (A_JI)
      {
        RCSGRAPH_TRAVERSE_JOINTS(graph)
        {
          memcpy(A_JI + (i * dof + JNT->jointIndex) * 12, &JNT->A_JI, 12 * sizeof(double));
        }
      }
      if (qOut)
      {
        memcpy(qOut + i * dof, graph->q->ele, dof * sizeof(double));
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(qdv);
    RcsGraph_destroy(graph);
  });
}

/* FK + per-effector translation / rotation Jacobians. effPt: nEff x 3 body-fixed points
 * or NULL (body origin). JT, JR: N x nEff x 3 x nJ (either may be NULL). A_BI as ref_fk. */
static double fkJacobiansImpl(void* g, long N, int qdim, const double* q, int nEff, const int* effBody,
                              const double* effPt, double* A_BI, double* JT, double* JR, double* M);

double ref_fk_jacobians(void* g, long N, int qdim, const double* q, int nEff, const int* effBody,
                        const double* effPt, double* A_BI, double* JT, double* JR)
{
  return fkJacobiansImpl(g, N, qdim, q, nEff, effBody, effPt, A_BI, JT, JR, NULL);
}

/* The same pass plus RcsGraph_computeMassMatrix (Rcs_dynamics.c:610) on the same graph state:
 * ONE RcsGraph_setState per state for "FK + Jacobians + M(q)". M: N x nJ x nJ. */
double ref_fk_jacobians_mass(void* g, long N, int qdim, const double* q, int nEff, const int* effBody,
                             const double* effPt, double* A_BI, double* JT, double* JR, double* M)
{
  return fkJacobiansImpl(g, N, qdim, q, nEff, effBody, effPt, A_BI, JT, JR, M);
}

static double fkJacobiansImpl(void* g, long N, int qdim, const double* q, int nEff, const int* effBody,
                              const double* effPt, double* A_BI, double* JT, double* JR, double* M)
{
  RcsGraph* master = (RcsGraph*) g;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    RcsGraph* graph = RcsGraph_clone(master);
    std::vector<RcsBody*> bodies = dfsBodies(graph);
    const size_t nb = bodies.size();
    const unsigned int dof = graph->dof, nJ = graph->nJ;
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* J = MatNd_create(3, nJ);
    MatNd* Mm = MatNd_create(nJ, nJ);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, qdim, q + i * qdim, NULL);
      if (A_BI)
      {
        for (size_t b = 0; b < nb; ++b)
        {
          memcpy(A_BI + (i * nb + b) * 12, bodies[b]->A_BI, 12 * sizeof(double));
        }
      }
      for (int e = 0; e < nEff; ++e)
      {
        const RcsBody* eb = bodies[effBody[e]];
        if (JT)
        {
          RcsGraph_bodyPointJacobian(graph, eb, effPt ? effPt + 3 * e : NULL, NULL, J);
          memcpy(JT + ((size_t) i * nEff + e) * 3 * nJ, J->ele, 3 * nJ * sizeof(double));
        }
        if (JR)
        {
          RcsGraph_rotationJacobian(graph, eb, NULL, J);
          memcpy(JR + ((size_t) i * nEff + e) * 3 * nJ, J->ele, 3 * nJ * sizeof(double));
        }
      }
      if (M)
      {
        RcsGraph_computeMassMatrix(graph, Mm);
        memcpy(M + (size_t) i * nJ * nJ, Mm->ele, sizeof(double) * nJ * nJ);
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(J);
    MatNd_destroy(Mm);
    RcsGraph_destroy(graph);
  });
}

/* Center of gravity and its Jacobian. cog: N x 3, Jcog: N x 3 x nJ. */
double ref_cog(void* g, long N, int qdim, const double* q, double* cog, double* Jcog)
{
  RcsGraph* master = (RcsGraph*) g;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    RcsGraph* graph = RcsGraph_clone(master);
    const unsigned int dof = graph->dof, nJ = graph->nJ;
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* J = MatNd_create(3, nJ);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, qdim, q + i * qdim, NULL);
      if (cog)
      {
        RcsGraph_COG(graph, cog + 3 * i);
      }
      if (Jcog)
      {
        RcsGraph_COGJacobian(graph, J);
        memcpy(Jcog + (size_t) i * 3 * nJ, J->ele, 3 * nJ * sizeof(double));
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(J);
    RcsGraph_destroy(graph);
  });
}

/* Task Jacobians relative to a (possibly articulated) reference body / frame:
 * RcsGraph_3dPosJacobian (kind 1) / RcsGraph_3dOmegaJacobian (kind 2), Rcs_kinematics.c:566, :789.
 * spec: nJac x {kind, effector, refBdy, refFrame} (depth-first body indices, -1 = NULL).
 * J: N x nJac x 3 x nJ. */
double ref_task_jacobians(void* g, long N, int qdim, const double* q, int nJac, const int* spec,
                          double* Jout)
{
  RcsGraph* master = (RcsGraph*) g;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    RcsGraph* graph = RcsGraph_clone(master);
    std::vector<RcsBody*> bodies = dfsBodies(graph);
    const unsigned int dof = graph->dof, nJ = graph->nJ;
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* J = MatNd_create(3, nJ);
    auto body = [&](int idx) -> const RcsBody* { return idx < 0 ? NULL : bodies[idx]; };
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, qdim, q + i * qdim, NULL);
      for (int t = 0; t < nJac; ++t)
      {
        const int* sp = spec + 4 * t;
        if (sp[0] == 1)
        {
          RcsGraph_3dPosJacobian(graph, body(sp[1]), body(sp[2]), body(sp[3]), J);
        }
        else
        {
          RcsGraph_3dOmegaJacobian(graph, body(sp[1]), body(sp[2]), body(sp[3]), J);
        }
        memcpy(Jout + ((size_t) i * nJac + t) * 3 * nJ, J->ele, 3 * nJ * sizeof(double));
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(J);
    RcsGraph_destroy(graph);
  });
}

/* Kinetic terms. M: N x nJ x nJ, h: N x nJ, Fg: N x nJ, E: N (any may be NULL). */
double ref_kinetic_terms(void* g, long N, int qdim, const double* q, const double* qd,
                         double* M, double* h, double* Fg, double* E)
{
  RcsGraph* master = (RcsGraph*) g;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    RcsGraph* graph = RcsGraph_clone(master);
    const unsigned int dof = graph->dof, nJ = graph->nJ;
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* qdv = MatNd_create(dof, 1);
    MatNd* Mm = MatNd_create(nJ, nJ);
    MatNd* hm = MatNd_create(nJ, 1);
    MatNd* gm = MatNd_create(nJ, 1);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, qdv, qdim, q + i * qdim, qd ? qd + i * qdim : NULL);
      double e = RcsGraph_computeKineticTerms(graph, M ? Mm : NULL, h ? hm : NULL, Fg ? gm : NULL);
      if (M)
      {
        memcpy(M + (size_t) i * nJ * nJ, Mm->ele, sizeof(double) * nJ * nJ);
      }
      if (h)
      {
        memcpy(h + (size_t) i * nJ, hm->ele, sizeof(double) * nJ);
      }
      if (Fg)
      {
        memcpy(Fg + (size_t) i * nJ, gm->ele, sizeof(double) * nJ);
      }
      if (E)
      {
        E[i] = e;
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(qdv);
    MatNd_destroy(Mm);
    MatNd_destroy(hm);
    MatNd_destroy(gm);
    RcsGraph_destroy(graph);
  });
}

/* Joint accelerations through Rcs_directDynamics (Rcs_dynamics.c:673): M qdd = F_g + F_ext -
 * F_jnt + h. Fext, Fjnt: N x nJ or NULL; qdd: N x nJ; E: N or NULL. */
double ref_direct_dynamics(void* g, long N, int qdim, const double* q, const double* qd,
                           const double* Fext, const double* Fjnt, double* qdd, double* E)
{
  RcsGraph* master = (RcsGraph*) g;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    RcsGraph* graph = RcsGraph_clone(master);
    const unsigned int dof = graph->dof, nJ = graph->nJ;
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* qdv = MatNd_create(dof, 1);
    MatNd* fe = MatNd_create(nJ, 1);
    MatNd* fj = MatNd_create(nJ, 1);
    MatNd* acc = MatNd_create(nJ, 1);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, qdv, qdim, q + i * qdim, qd ? qd + i * qdim : NULL);
      if (Fext)
      {
        memcpy(fe->ele, Fext + (size_t) i * nJ, sizeof(double) * nJ);
      }
      if (Fjnt)
      {
        memcpy(fj->ele, Fjnt + (size_t) i * nJ, sizeof(double) * nJ);
      }
      double e = Rcs_directDynamics(graph, Fext ? fe : NULL, Fjnt ? fj : NULL, acc);
      memcpy(qdd + (size_t) i * nJ, acc->ele, sizeof(double) * nJ);
      if (E)
      {
        E[i] = e;
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(qdv);
    MatNd_destroy(fe);
    MatNd_destroy(fj);
    MatNd_destroy(acc);
    RcsGraph_destroy(graph);
  });
}

/* Mass matrix through RcsGraph_computeMassMatrix. */
double ref_mass_matrix(void* g, long N, int qdim, const double* q, double* M)
{
  RcsGraph* master = (RcsGraph*) g;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    RcsGraph* graph = RcsGraph_clone(master);
    const unsigned int dof = graph->dof, nJ = graph->nJ;
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* Mm = MatNd_create(nJ, nJ);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, qdim, q + i * qdim, NULL);
      RcsGraph_computeMassMatrix(graph, Mm);
      memcpy(M + (size_t) i * nJ * nJ, Mm->ele, sizeof(double) * nJ * nJ);
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(Mm);
    RcsGraph_destroy(graph);
  });
}

/* Joint metric and joint limit gradient (IK state space). invWq: nJ, dH: N x nJ, cost: N. */
void ref_inv_wq(void* g, double* invWq)
{
  RcsGraph* graph = (RcsGraph*) g;
  MatNd* w = MatNd_create(graph->nJ, 1);
  RcsGraph_getInvWq(graph, w, RcsStateIK);
  memcpy(invWq, w->ele, sizeof(double) * graph->nJ);
  MatNd_destroy(w);
}

double ref_joint_limit_gradient(void* g, long N, int qdim, const double* q, double* dH,
                                double* cost)
{
  RcsGraph* master = (RcsGraph*) g;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    RcsGraph* graph = RcsGraph_clone(master);
    const unsigned int dof = graph->dof, nJ = graph->nJ;
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* d = MatNd_create(1, nJ);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, qdim, q + i * qdim, NULL);
      double c = RcsGraph_jointLimitGradient(graph, d, RcsStateIK);
      memcpy(dH + (size_t) i * nJ, d->ele, sizeof(double) * nJ);
      if (cost)
      {
        cost[i] = c;
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(d);
    RcsGraph_destroy(graph);
  });
}

/* ---- Controller + IkSolverRMR ------------------------------------------------------- */

/* tasks: newline-free list "XYZ:BallTip;ABC:BallTip" → <Task controlVariable= effector=> */
void* ref_ik_create(const char* graphFile, const char* taskSpec)
{
  std::ostringstream x;
  x << "<Controller graph=\"" << graphFile << "\">";
  std::string spec(taskSpec);
  size_t pos = 0;
  int k = 0;
  while (pos < spec.size())
  {
    size_t end = spec.find(';', pos);
    if (end == std::string::npos)
    {
      end = spec.size();
    }
    std::string item = spec.substr(pos, end - pos);
    /* kind:effector[:refBdy[:refFrame]] */
    std::vector<std::string> f;
    size_t a = 0;
    while (a <= item.size())
    {
      size_t b = item.find(':', a);
      if (b == std::string::npos)
      {
        b = item.size();
      }
      f.push_back(item.substr(a, b - a));
      a = b + 1;
    }
    if (f.size() >= 2)
    {
      /* POLARX / POLARY / POLARZ: TaskPolar2D with that axisDirection */
      std::string polarAxis;
      if (f[0].size() == 6 && f[0].compare(0, 5, "POLAR") == 0)
      {
        polarAxis = f[0].substr(5);
        f[0] = "POLAR";
      }
      x << "<Task name=\"t" << k++ << "\" controlVariable=\"" << f[0] << "\"";
      if (!polarAxis.empty())
      {
        x << " axisDirection=\"" << polarAxis << "\"";
      }
      if (f[0] == "Joint")
      {
        x << " jnt=\"" << f[1] << "\"";       /* TaskJoint: the second field is the joint's name */
        if (f.size() >= 3 && !f[2].empty())
        {
          x << " refJnt=\"" << f[2] << "\"";   /* third: reference joint, fourth: refGain */
          if (f.size() >= 4 && !f[3].empty())
          {
            x << " refGain=\"" << f[3] << "\"";
          }
        }
        f.resize(2);
      }
      else if (!f[1].empty())
      {
        x << " effector=\"" << f[1] << "\"";
      }
      if (f.size() >= 3 && !f[2].empty())
      {
        x << " refBdy=\"" << f[2] << "\"";
      }
      if (f.size() >= 4 && !f[3].empty())
      {
        x << " refFrame=\"" << f[3] << "\"";
      }
      x << " active=\"true\"/>";
    }
    pos = end + 1;
  }
  x << "</Controller>";
  IkHandle* h = new IkHandle;
  h->xml = x.str();
  h->controller = new Rcs::ControllerBase(h->xml, true);
  return h;
}

void ref_ik_destroy(void* ik)
{
  IkHandle* h = (IkHandle*) ik;
  delete h->controller;
  delete h;
}

int ref_ik_task_dim(void* ik)
{
  return (int) ((IkHandle*) ik)->controller->getTaskDim();
}

void* ref_ik_graph(void* ik)
{
  return ((IkHandle*) ik)->controller->getGraph();
}

/* Task vector x (N x nx) and stacked task Jacobian (N x nx x nJ) at states q. */
double ref_ik_task_kinematics(void* ik, long N, int qdim, const double* q, double* x, double* J)
{
  IkHandle* h = (IkHandle*) ik;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    Rcs::ControllerBase c(*h->controller);
    RcsGraph* graph = c.getGraph();
    const unsigned int dof = graph->dof, nJ = graph->nJ, nx = c.getTaskDim();
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* xm = MatNd_create(nx, 1);
    MatNd* Jm = MatNd_create(nx, nJ);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, qdim, q + i * qdim, NULL);
      if (x)
      {
        c.computeX(xm);
        memcpy(x + (size_t) i * nx, xm->ele, sizeof(double) * nx);
      }
      if (J)
      {
        c.computeJ(Jm);
        memcpy(J + (size_t) i * nx * nJ, Jm->ele, sizeof(double) * nx * nJ);
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(xm);
    MatNd_destroy(Jm);
  });
}

/* Whole-body evaluation of one state in ONE pass on the controller's graph: RcsGraph_setState
 * (frames of every body, A_BI: N x nb x 12), ControllerBase::computeJ (stacked task Jacobian,
 * J: N x nx x nJ) and RcsGraph_computeMassMatrix (M: N x nJ x nJ). Any output may be NULL. */
double ref_ik_wholebody(void* ik, long N, int qdim, const double* q, double* A_BI, double* J, double* M)
{
  IkHandle* h = (IkHandle*) ik;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    Rcs::ControllerBase c(*h->controller);
    RcsGraph* graph = c.getGraph();
    std::vector<RcsBody*> bodies = dfsBodies(graph);
    const size_t nb = bodies.size();
    const unsigned int dof = graph->dof, nJ = graph->nJ, nx = c.getTaskDim();
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* Jm = MatNd_create(nx, nJ);
    MatNd* Mm = MatNd_create(nJ, nJ);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, qdim, q + i * qdim, NULL);
      if (A_BI)
      {
        for (size_t b = 0; b < nb; ++b)
        {
          memcpy(A_BI + (i * nb + b) * 12, bodies[b]->A_BI, 12 * sizeof(double));
        }
      }
      if (J)
      {
        c.computeJ(Jm);
        memcpy(J + (size_t) i * nx * nJ, Jm->ele, sizeof(double) * nx * nJ);
      }
      if (M)
      {
        RcsGraph_computeMassMatrix(graph, Mm);
        memcpy(M + (size_t) i * nJ * nJ, Mm->ele, sizeof(double) * nJ * nJ);
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(Jm);
    MatNd_destroy(Mm);
  });
}

/*
 * RMR iterations as in examples/ExampleIK.cpp:154-177.
 *   q     : N x dof, in: start state, out: state after `iters` steps
 *   xDes  : N x nx task-space targets
 *   dxOut : N x nx task error computed in the LAST iteration (before its step), or NULL
 *   dqOut : N x dof  joint displacement of the LAST iteration, or NULL
 *   det   : N determinant of (J W J^T + lambda I) in the LAST iteration, or NULL
 */
static double ikRmrImpl(void* ik, long N, double* q, const double* xDes, double lambda, double alpha,
                        int iters, double* dxOut, double* dqOut, double* det, bool leftInverse);

double ref_ik_rmr(void* ik, long N, double* q, const double* xDes, double lambda, double alpha,
                  int iters, double* dxOut, double* dqOut, double* det)
{
  return ikRmrImpl(ik, N, q, xDes, lambda, alpha, iters, dxOut, dqOut, det, false);
}

/*
 * The same loop with IkSolverConstraintRMR::solveRightInverse (IkSolverConstraintRMR.cpp:156): the step
 * of the user's tasks, then, if it leaves a joint closer than 2 degrees to a limit, once more with a
 * joint constraint per such joint. No collision model, all user tasks active.
 */
double ref_ik_constraint_rmr(void* ik, long N, double* q, const double* xDes, double lambda, double alpha,
                             int iters, double* dxOut, double* dqOut, double* det)
{
  IkHandle* h = (IkHandle*) ik;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    Rcs::ControllerBase c(*h->controller);
    Rcs::IkSolverConstraintRMR solver(&c);
    RcsGraph* graph = c.getGraph();
    const unsigned int dof = graph->dof, nJ = graph->nJ, nx = c.getTaskDim();
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* dq = MatNd_create(dof, 1);
    MatNd* a = MatNd_create(c.getNumberOfTasks(), 1);
    MatNd* xd = MatNd_create(nx, 1);
    MatNd* dx = MatNd_create(nx, 1);
    MatNd* dH = MatNd_create(1, nJ);
    MatNd_setElementsTo(a, 1.0);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, dof, q + i * dof, NULL);
      memcpy(xd->ele, xDes + (size_t) i * nx, sizeof(double) * nx);
      for (int it = 0; it < iters; ++it)
      {
        c.computeDX(dx, xd);
        c.computeJointlimitGradient(dH);
        MatNd_constMulSelf(dH, alpha);
        solver.solveRightInverse(dq, dx, dH, a, lambda);
        MatNd_addSelf(graph->q, dq);
        RcsGraph_setState(graph, NULL, NULL);
      }
      memcpy(q + i * dof, graph->q->ele, sizeof(double) * dof);
      if (dxOut)
      {
        memcpy(dxOut + (size_t) i * nx, dx->ele, sizeof(double) * nx);
      }
      if (dqOut)
      {
        memcpy(dqOut + (size_t) i * dof, dq->ele, sizeof(double) * dof);
      }
      if (det)
      {
        det[i] = solver.getDeterminant();
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(dq);
    MatNd_destroy(a);
    MatNd_destroy(xd);
    MatNd_destroy(dx);
    MatNd_destroy(dH);
  });
}

/* the same loop with IkSolverRMR::solveLeftInverse (IkSolverRMR.cpp:205) */
double ref_ik_rmr_left(void* ik, long N, double* q, const double* xDes, double lambda, double alpha,
                       int iters, double* dxOut, double* dqOut, double* det)
{
  return ikRmrImpl(ik, N, q, xDes, lambda, alpha, iters, dxOut, dqOut, det, true);
}

/*
 * The same loop with the remaining arguments of the reference's solver interface:
 *   activation [nTasks]      passed to computeDX (if scaleDx) and to the solver
 *   lambdaRows [active rows] or NULL: the m x 1 lambda of solveRightInverse(..., const MatNd* lambda)
 *   blending                 IkSolverRMR::setActivationBlending: 0 Binary, 1 Linear, 2 Approximate,
 *                            3 IndependentTasks, 4 DependentTasks (left inverse)
 *   dxOut  N x nx: dx of the last iteration, ACTIVE rows first (compressToActive), rest untouched
 *   dqTs, dqNs N x dof: solveRightInverse / solveLeftInverse (dq_ts, dq_ns, ...) of the last iteration
 */
double ref_ik_rmr_ex(void* ik, long N, double* q, const double* xDes, const double* activation,
                     int scaleDx, const double* lambdaRows, int nLambdaRows, double lambda, double alpha,
                     int iters, int leftInverse, int blending, double* dxOut, double* dqTs, double* dqNs,
                     double* det)
{
  IkHandle* h = (IkHandle*) ik;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    Rcs::ControllerBase c(*h->controller);
    Rcs::IkSolverRMR solver(&c);
    static const char* modes[] = {"Binary", "Linear", "Approximate", "IndependentTasks", "DependentTasks"};
    solver.setActivationBlending(modes[blending]);
    RcsGraph* graph = c.getGraph();
    const unsigned int dof = graph->dof, nJ = graph->nJ, nx = c.getTaskDim();
    const unsigned int nTasks = c.getNumberOfTasks();
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* ts = MatNd_create(dof, 1);
    MatNd* ns = MatNd_create(dof, 1);
    MatNd* a = MatNd_create(nTasks, 1);
    MatNd* xd = MatNd_create(nx, 1);
    MatNd* dx = MatNd_create(nx, 1);
    MatNd* dxc = MatNd_create(nx, 1);
    MatNd* dH = MatNd_create(1, nJ);
    MatNd* lam = MatNd_create(nLambdaRows > 0 ? nLambdaRows : 1, 1);
    for (unsigned int t = 0; t < nTasks; ++t)
    {
      a->ele[t] = activation ? activation[t] : 1.0;
    }
    for (int r = 0; r < nLambdaRows; ++r)
    {
      lam->ele[r] = lambdaRows[r];
    }
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, dof, q + i * dof, NULL);
      memcpy(xd->ele, xDes + (size_t) i * nx, sizeof(double) * nx);
      for (int it = 0; it < iters; ++it)
      {
        MatNd_reshape(dx, nx, 1);
        if (scaleDx)
        {
          c.computeDX(dx, xd, a);      /* active rows only, scaled by the activations */
        }
        else
        {
          c.computeDX(dx, xd);         /* all rows; the solver compresses them */
        }
        c.computeJointlimitGradient(dH);
        MatNd_constMulSelf(dH, alpha);
        if (leftInverse)
        {
          solver.solveLeftInverse(ts, ns, dx, dH, a, lambda);
        }
        else if (nLambdaRows > 0)
        {
          solver.solveRightInverse(ts, ns, dx, dH, a, lam);
        }
        else
        {
          solver.solveRightInverse(ts, ns, dx, dH, a, lambda);
        }
        MatNd_addSelf(graph->q, ts);
        MatNd_addSelf(graph->q, ns);
        RcsGraph_setState(graph, NULL, NULL);
      }
      memcpy(q + i * dof, graph->q->ele, sizeof(double) * dof);
      if (dxOut)
      {
        if (dx->m == nx && !scaleDx)
        {
          c.compressToActive(dxc, dx, a);
          memcpy(dxOut + (size_t) i * nx, dxc->ele, sizeof(double) * dxc->m);
        }
        else
        {
          memcpy(dxOut + (size_t) i * nx, dx->ele, sizeof(double) * dx->m);
        }
      }
      if (dqTs)
      {
        memcpy(dqTs + (size_t) i * dof, ts->ele, sizeof(double) * dof);
      }
      if (dqNs)
      {
        memcpy(dqNs + (size_t) i * dof, ns->ele, sizeof(double) * dof);
      }
      if (det)
      {
        det[i] = solver.getDeterminant();
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(ts);
    MatNd_destroy(ns);
    MatNd_destroy(a);
    MatNd_destroy(xd);
    MatNd_destroy(dx);
    MatNd_destroy(dxc);
    MatNd_destroy(dH);
    MatNd_destroy(lam);
  });
}

/*
 * Prioritised steps: IkSolverPrioRMR::solve(dq1, dq2, dq_ns, dx, dH, activation, lambda)
 * (src/RcsCore/IkSolverPrioRMR.cpp:178), looped like examples/ExampleIK.cpp. priority [nTasks]:
 * level of every task (IkSolverPrioRMR::setTaskPriority). Outputs of the last iteration.
 */
double ref_ik_prio_rmr(void* ik, long N, double* q, const double* xDes, const int* priority,
                       const double* activation, double lambda, double alpha, int iters, double* dq1Out,
                       double* dq2Out, double* dqNsOut, double* okOut)
{
  IkHandle* h = (IkHandle*) ik;
  /* IkSolverPrioRMR's constructor re-reads the controller's xml FILE for "prio" attributes
   * (IkSolverPrioRMR.cpp:108-140): give it a controller that was created from a file */
  char tmpName[] = "/tmp/rcsb_ref_controller_XXXXXX.xml";
  const int fd = mkstemps(tmpName, 4);
  if (fd < 0)
  {
    return -1.0;
  }
  FILE* tf = fdopen(fd, "w");
  fputs(h->xml.c_str(), tf);
  fclose(tf);
  Rcs::ControllerBase* fileController = new Rcs::ControllerBase(std::string(tmpName), true);
  const double secs = parallelSlices(N, [=](int, long lo, long hi)
  {
    Rcs::ControllerBase c(*fileController);
    Rcs::IkSolverPrioRMR solver(&c);
    RcsGraph* graph = c.getGraph();
    const unsigned int dof = graph->dof, nJ = graph->nJ, nx = c.getTaskDim();
    const unsigned int nTasks = c.getNumberOfTasks();
    for (unsigned int t = 0; t < nTasks; ++t)
    {
      solver.setTaskPriority(t, priority[t]);
    }
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* d1 = MatNd_create(dof, 1);
    MatNd* d2 = MatNd_create(dof, 1);
    MatNd* dn = MatNd_create(dof, 1);
    MatNd* a = MatNd_create(nTasks, 1);
    MatNd* xd = MatNd_create(nx, 1);
    MatNd* dx = MatNd_create(nx, 1);
    MatNd* dH = MatNd_create(1, nJ);
    for (unsigned int t = 0; t < nTasks; ++t)
    {
      a->ele[t] = activation ? activation[t] : 1.0;
    }
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, dof, q + i * dof, NULL);
      memcpy(xd->ele, xDes + (size_t) i * nx, sizeof(double) * nx);
      bool ok = true;
      for (int it = 0; it < iters; ++it)
      {
        MatNd_reshape(dx, nx, 1);
        c.computeDX(dx, xd);
        c.computeJointlimitGradient(dH);
        MatNd_constMulSelf(dH, alpha);
        MatNd_reshape(d1, dof, 1);
        MatNd_reshape(d2, dof, 1);
        MatNd_reshape(dn, dof, 1);
        ok = solver.solve(d1, d2, dn, dx, dH, a, lambda);
        MatNd_addSelf(graph->q, d1);
        MatNd_addSelf(graph->q, d2);
        MatNd_addSelf(graph->q, dn);
        RcsGraph_setState(graph, NULL, NULL);
      }
      memcpy(q + i * dof, graph->q->ele, sizeof(double) * dof);
      if (dq1Out)
      {
        memcpy(dq1Out + (size_t) i * dof, d1->ele, sizeof(double) * dof);
      }
      if (dq2Out)
      {
        memcpy(dq2Out + (size_t) i * dof, d2->ele, sizeof(double) * dof);
      }
      if (dqNsOut)
      {
        memcpy(dqNsOut + (size_t) i * dof, dn->ele, sizeof(double) * dof);
      }
      if (okOut)
      {
        okOut[i] = ok ? 1.0 : 0.0;
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(d1);
    MatNd_destroy(d2);
    MatNd_destroy(dn);
    MatNd_destroy(a);
    MatNd_destroy(xd);
    MatNd_destroy(dx);
    MatNd_destroy(dH);
  });
  delete fileController;
  unlink(tmpName);
  return secs;
}

}   // extern "C"

static double ikRmrImpl(void* ik, long N, double* q, const double* xDes, double lambda, double alpha,
                        int iters, double* dxOut, double* dqOut, double* det, bool leftInverse)
{
  IkHandle* h = (IkHandle*) ik;
  return parallelSlices(N, [=](int, long lo, long hi)
  {
    Rcs::ControllerBase c(*h->controller);
    Rcs::IkSolverRMR solver(&c);
    RcsGraph* graph = c.getGraph();
    const unsigned int dof = graph->dof, nJ = graph->nJ, nx = c.getTaskDim();
    MatNd* qv = MatNd_create(dof, 1);
    MatNd* dq = MatNd_create(dof, 1);
    MatNd* a = MatNd_create(c.getNumberOfTasks(), 1);
    MatNd* xd = MatNd_create(nx, 1);
    MatNd* dx = MatNd_create(nx, 1);
    MatNd* dH = MatNd_create(1, nJ);
    MatNd_setElementsTo(a, 1.0);
    sliceReady();
    for (long i = lo; i < hi; ++i)
    {
      setQ(graph, qv, NULL, dof, q + i * dof, NULL);
      memcpy(xd->ele, xDes + (size_t) i * nx, sizeof(double) * nx);
      for (int it = 0; it < iters; ++it)
      {
        c.computeDX(dx, xd);
        c.computeJointlimitGradient(dH);
        MatNd_constMulSelf(dH, alpha);
        if (leftInverse)
        {
          solver.solveLeftInverse(dq, dx, dH, a, lambda);
        }
        else
        {
          solver.solveRightInverse(dq, dx, dH, a, lambda);
        }
        MatNd_addSelf(graph->q, dq);
        RcsGraph_setState(graph, NULL, NULL);
      }
      memcpy(q + i * dof, graph->q->ele, sizeof(double) * dof);
      if (dxOut)
      {
        memcpy(dxOut + (size_t) i * nx, dx->ele, sizeof(double) * nx);
      }
      if (dqOut)
      {
        memcpy(dqOut + (size_t) i * dof, dq->ele, sizeof(double) * dof);
      }
      if (det)
      {
        det[i] = solver.getDeterminant();
      }
    }
    sliceDone();
    MatNd_destroy(qv);
    MatNd_destroy(dq);
    MatNd_destroy(a);
    MatNd_destroy(xd);
    MatNd_destroy(dx);
    MatNd_destroy(dH);
  });
}
