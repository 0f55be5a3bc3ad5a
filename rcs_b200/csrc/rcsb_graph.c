/*
 * Host-side kinematic tree of rcs_b200: construction, traversal, indexing, coupling
 * rules, cloning, validation and the canonical layout dump.
 *
 * Everything here is bookkeeping that happens once per graph. It reproduces the
 * reference's integer layout exactly (body order, jointIndex, jacobiIndex, joint chain
 * links), because the device model (rcsb_model.c) is flattened from these structures.
 * Citations are to /root/reference/src/RcsCore.
 */
#include "rcsb_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------- */
/* errors                                                                               */
/* ---------------------------------------------------------------------------------- */

int RcsLogLevel = 0;
static int g_abortOnFatal = 1;
static __thread char g_lastError[1024] = "";

void Rcs_setFatalMode(int abortOnFatal)
{
  g_abortOnFatal = abortOnFatal;
}

void rcsb_set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_lastError, sizeof(g_lastError), fmt, ap);
  va_end(ap);
}

const char* rcsb_error_string(void)
{
  return g_lastError;
}

void rcsb_fatal(const char* file, int line, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_lastError, sizeof(g_lastError), fmt, ap);
  va_end(ap);

  if (g_abortOnFatal)
  {
    fprintf(stderr, "[rcsb FATAL %s:%d] %s\n", file, line, g_lastError);
    abort();
  }
}

/* ---------------------------------------------------------------------------------- */
/* MatNd                                                                                */
/* ---------------------------------------------------------------------------------- */

MatNd* MatNd_create(unsigned int m, unsigned int n)
{
  MatNd* self = (MatNd*) calloc(1, sizeof(MatNd));
  self->m = m;
  self->n = n;
  self->size = m * n;
  self->ele = (double*) calloc(self->size > 0 ? self->size : 1, sizeof(double));
  return self;
}

void MatNd_destroy(MatNd* self)
{
  if (!self)
  {
    return;
  }
  if (!self->stackMem)
  {
    free(self->ele);
  }
  free(self);
}

/* Reshapes within the allocated size; growing beyond it is a contract violation in the
 * reference (src/RcsCore/Rcs_MatNd.c MatNd_reshape RCHECK). Heap matrices are grown here
 * instead of aborting. */
void MatNd_reshape(MatNd* self, unsigned int m, unsigned int n)
{
  if (m * n > self->size)
  {
    if (self->stackMem)
    {
      RCSB_FATAL("MatNd_reshape to %u x %u exceeds stack matrix size %u", m, n, self->size);
      return;
    }
    self->ele = (double*) realloc(self->ele, sizeof(double) * m * n);
    memset(self->ele + self->size, 0, sizeof(double) * (m * n - self->size));
    self->size = m * n;
  }
  self->m = m;
  self->n = n;
}

void MatNd_reshapeAndSetZero(MatNd* self, unsigned int m, unsigned int n)
{
  MatNd_reshape(self, m, n);
  memset(self->ele, 0, sizeof(double) * m * n);
}

/* ---------------------------------------------------------------------------------- */
/* traversal                                                                            */
/* ---------------------------------------------------------------------------------- */

/* src/RcsCore/Rcs_body.c:573 */
RcsBody* RcsBody_depthFirstTraversalGetNext(const RcsBody* body)
{
  if (!body)
  {
    return NULL;
  }
  if (body->firstChild)
  {
    return body->firstChild;
  }
  if (body->next)
  {
    return body->next;
  }
  for (const RcsBody* up = body->parent; up; up = up->parent)
  {
    if (up->next)
    {
      return up->next;
    }
  }
  return NULL;
}

/* src/RcsCore/Rcs_body.c:1706 */
RcsJoint* RcsBody_lastJointBeforeBody(const RcsBody* body)
{
  if (!body)
  {
    return NULL;
  }
  RcsJoint* jnt = body->jnt;
  while (!jnt)
  {
    if (!body->parent)
    {
      return NULL;
    }
    body = body->parent;
    jnt = body->jnt;
  }
  while (jnt->next)
  {
    jnt = jnt->next;
  }
  return jnt;
}

bool RcsJoint_isRotation(const RcsJoint* j)
{
  return j->type == RCSJOINT_ROT_X || j->type == RCSJOINT_ROT_Y || j->type == RCSJOINT_ROT_Z;
}

/* First depth-first match (src/RcsCore/Rcs_graph.c:1116). Generic bodies are not
 * supported by this loader. */
RcsBody* RcsGraph_getBodyByName(const RcsGraph* self, const char* name)
{
  if (!self || !name)
  {
    return NULL;
  }
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    if (strcmp(name, BODY->name) == 0)
    {
      return BODY;
    }
  }
  return NULL;
}

RcsJoint* RcsGraph_getJointByName(const RcsGraph* self, const char* name)
{
  if (!self || !name)
  {
    return NULL;
  }
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (strcmp(name, JNT->name) == 0)
      {
        return JNT;
      }
    }
  }
  return NULL;
}

unsigned int RcsGraph_numBodies(const RcsGraph* self)
{
  unsigned int n = 0;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    n++;
  }
  return n;
}

RcsBody* RcsGraph_getBodyByIndex(const RcsGraph* self, unsigned int idx)
{
  unsigned int i = 0;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    if (i++ == idx)
    {
      return BODY;
    }
  }
  return NULL;
}

int RcsGraph_bodyIndex(const RcsGraph* self, const RcsBody* body)
{
  int i = 0;
  if (!body)
  {
    return -1;
  }
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    if (BODY == body)
    {
      return i;
    }
    i++;
  }
  return -1;
}

RcsJoint* RcsGraph_getJointByIndex(const RcsGraph* self, unsigned int jointIndex)
{
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if ((unsigned int) JNT->jointIndex == jointIndex)
      {
        return JNT;
      }
    }
  }
  return NULL;
}

double RcsGraph_mass(const RcsGraph* self)
{
  double m = 0.0;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    m += BODY->m;
  }
  return m;
}

/* ---------------------------------------------------------------------------------- */
/* construction                                                                         */
/* ---------------------------------------------------------------------------------- */

/* A body always becomes the LAST child of its parent (or the last top-level sibling),
 * so depth-first order differs from file order when "prev" points backwards
 * (src/RcsCore/Rcs_graph.c:2535-2598). */
void RcsGraph_insertBody(RcsGraph* graph, RcsBody* parent, RcsBody* body)
{
  if (!body)
  {
    return;
  }
  if (body->jnt || body->prev || body->next || body->firstChild || body->lastChild)
  {
    RCSB_FATAL("RcsGraph_insertBody: body \"%s\" must be unlinked and without joints",
               body->name ? body->name : "?");
    return;
  }

  body->parent = parent;

  if (parent)
  {
    if (parent->lastChild)
    {
      parent->lastChild->next = body;
      body->prev = parent->lastChild;
    }
    else
    {
      parent->firstChild = body;
    }
    parent->lastChild = body;
  }
  else if (graph->root)
  {
    RcsBody* b = graph->root;
    while (b->next)
    {
      b = b->next;
    }
    b->next = body;
    body->prev = b;
  }
  else
  {
    graph->root = body;
  }
}

/* Appends the joint to the body's joint list; jointIndex is the creation count and is
 * renumbered later; the first joint of a body is chained to the last joint on the way
 * to the root (src/RcsCore/Rcs_graph.c:2604-2655). */
void RcsGraph_insertJoint(RcsGraph* graph, RcsBody* body, RcsJoint* jnt)
{
  graph->dof++;

  if (!graph->q)
  {
    graph->q = MatNd_create(graph->dof, 1);
  }
  else if (graph->q->m < graph->dof)
  {
    MatNd_reshape(graph->q, graph->dof, 1);
  }

  jnt->jointIndex = (int) graph->dof - 1;

  if (!body->jnt)
  {
    body->jnt = jnt;
    jnt->prev = body->parent ? RcsBody_lastJointBeforeBody(body->parent) : NULL;
  }
  else
  {
    RcsJoint* last = body->jnt;
    while (last->next)
    {
      last = last->next;
    }
    last->next = jnt;
    jnt->prev = last;
  }
}

/* ---------------------------------------------------------------------------------- */
/* joint coupling (src/RcsCore/Rcs_joint.c:318-462)                                    */
/* ---------------------------------------------------------------------------------- */

static double couplingPoly(double x, const MatNd* c)
{
  const int orderM1 = (int) c->m - 1;
  double y = 0.0;
  for (int i = 0; i <= orderM1; ++i)
  {
    y += c->ele[i] * pow(x, orderM1 - i);
  }
  return y;
}

static double couplingPolyDerivative(double x, const MatNd* c)
{
  const unsigned int orderM1 = c->m - 1;
  double dy = 0.0;
  for (unsigned int i = 0; i < orderM1; ++i)
  {
    dy += (orderM1 - i) * c->ele[i] * pow(x, (int) (orderM1 - 1 - i));
  }
  return dy;
}

double RcsJoint_computeSlaveJointAngle(const RcsJoint* slave, double q_master)
{
  const RcsJoint* master = slave->coupledTo;
  if (!master)
  {
    return 0.0;
  }

  if (slave->couplingFactors->size == 1)
  {
    return slave->q_init + slave->couplingFactors->ele[0] * (q_master - master->q_init);
  }

  /* polynomial coupling: tangent continuation outside of the master's range */
  if (q_master < master->q_min)
  {
    const double sens = couplingPolyDerivative(master->q_min - master->q_init,
                                               slave->couplingFactors);
    return slave->q_min - sens * (master->q_min - q_master);
  }
  if (q_master > master->q_max)
  {
    const double sens = couplingPolyDerivative(master->q_max - master->q_init,
                                               slave->couplingFactors);
    return slave->q_max + sens * (q_master - master->q_max);
  }
  return slave->q_init + couplingPoly(q_master - master->q_init, slave->couplingFactors);
}

double RcsJoint_computeSlaveJointVelocity(const RcsJoint* slave, double q_master,
                                          double q_dot_master)
{
  const RcsJoint* master = slave->coupledTo;
  if (!master)
  {
    return 0.0;
  }
  if (slave->couplingFactors->size == 1)
  {
    return slave->couplingFactors->ele[0] * q_dot_master;
  }
  /* the reference evaluates the derivative at the clipped master angle itself */
  double q_ltd = q_master < master->q_min ? master->q_min
                                          : (q_master > master->q_max ? master->q_max : q_master);
  return couplingPolyDerivative(q_ltd, slave->couplingFactors) * q_dot_master;
}

/* ---------------------------------------------------------------------------------- */
/* indexing                                                                             */
/* ---------------------------------------------------------------------------------- */

/* Renumber jointIndex in depth-first body order / in-body joint order and assign
 * jacobiIndex as the running count of unconstrained joints (-1 if constrained);
 * q and q_dot follow their joints (src/RcsCore/Rcs_graph.c:2932-2997). */
static void recomputeJointIndices(RcsGraph* self)
{
  unsigned int nq = 0, nj = 0;
  const unsigned int oldDof = self->dof;
  double* qOld = (double*) calloc(oldDof + 1, sizeof(double));
  double* qdOld = (double*) calloc(oldDof + 1, sizeof(double));
  if (self->q)
  {
    memcpy(qOld, self->q->ele, sizeof(double) * (self->q->m < oldDof ? self->q->m : oldDof));
  }
  if (self->q_dot)
  {
    memcpy(qdOld, self->q_dot->ele,
           sizeof(double) * (self->q_dot->m < oldDof ? self->q_dot->m : oldDof));
  }

  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if ((unsigned int) JNT->jointIndex != nq)
      {
        self->q->ele[nq] = qOld[JNT->jointIndex];
        self->q_dot->ele[nq] = qdOld[JNT->jointIndex];
        JNT->jointIndex = (int) nq;
      }
      nq++;
      if (!JNT->constrained)
      {
        JNT->jacobiIndex = (int) nj++;
      }
      else
      {
        JNT->jacobiIndex = -1;
      }
    }
  }

  self->q->m = nq;
  self->q_dot->m = nq;
  self->dof = nq;
  self->nJ = nj;
  free(qOld);
  free(qdOld);
}

/* src/RcsCore/Rcs_graph.c:2999-3067 */
void RcsGraph_makeJointsConsistent(RcsGraph* self)
{
  if (!self->q)
  {
    self->q = MatNd_create(self->dof, 1);
  }
  if (!self->q_dot)
  {
    self->q_dot = MatNd_create(self->dof, 1);
  }
  else if (self->q_dot->m < self->dof)
  {
    MatNd_reshape(self->q_dot, self->dof, 1);
  }

  recomputeJointIndices(self);

  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      self->q->ele[JNT->jointIndex] = JNT->q0;

      if (!JNT->coupledJointName)
      {
        continue;
      }

      RcsJoint* master = RcsGraph_getJointByName(self, JNT->coupledJointName);
      if (!master)
      {
        RCSB_FATAL("No coupled joint \"%s\"", JNT->coupledJointName);
        continue;
      }
      JNT->coupledTo = master;

      /* the slave's range and centre are those of the master mapped by the coupling */
      double q_min = RcsJoint_computeSlaveJointAngle(JNT, master->q_min);
      double q_max = RcsJoint_computeSlaveJointAngle(JNT, master->q_max);
      double q0 = RcsJoint_computeSlaveJointAngle(JNT, master->q0);
      double q_init = RcsJoint_computeSlaveJointAngle(JNT, master->q_init);

      if (q_min > q_max)
      {
        double t = q_min;
        q_min = q_max;
        q_max = t;
      }

      JNT->q_min = q_min;
      JNT->q_max = q_max;
      JNT->q0 = q0;
      JNT->q_init = q_init;

      self->q->ele[JNT->jointIndex] =
        RcsJoint_computeSlaveJointAngle(JNT, self->q->ele[master->jointIndex]);
    }
  }
}

/* ---------------------------------------------------------------------------------- */
/* state vector helpers                                                                 */
/* ---------------------------------------------------------------------------------- */

void RcsGraph_stateVectorFromIK(const RcsGraph* self, const MatNd* q_ik, MatNd* q_full)
{
  if (q_ik->m != self->nJ || q_full->m != self->dof)
  {
    RCSB_FATAL("stateVectorFromIK: q_ik->m=%u (nJ=%u), q_full->m=%u (dof=%u)", q_ik->m,
               self->nJ, q_full->m, self->dof);
    return;
  }
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (JNT->jacobiIndex != -1)
      {
        q_full->ele[JNT->jointIndex] = q_ik->ele[JNT->jacobiIndex];
      }
    }
  }
}

void RcsGraph_stateVectorToIK(const RcsGraph* self, const MatNd* q_full, MatNd* q_ik)
{
  if (q_full->m != self->dof || q_ik->size < self->nJ)
  {
    RCSB_FATAL("stateVectorToIK: q_full->m=%u (dof=%u), q_ik->size=%u (nJ=%u)", q_full->m,
               self->dof, q_ik->size, self->nJ);
    return;
  }
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (JNT->jacobiIndex != -1)
      {
        q_ik->ele[JNT->jacobiIndex] = q_full->ele[JNT->jointIndex];
      }
    }
  }
  q_ik->m = self->nJ;
  q_ik->n = 1;
}

void RcsGraph_getInvWq(const RcsGraph* self, MatNd* invWq, RcsStateType type)
{
  MatNd_reshape(invWq, type == RcsStateFull ? self->dof : self->nJ, 1);
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      const double w = JNT->weightMetric * (JNT->q_max - JNT->q_min);
      if (type == RcsStateFull)
      {
        invWq->ele[JNT->jointIndex] = w;
      }
      else if (JNT->jacobiIndex != -1)
      {
        invWq->ele[JNT->jacobiIndex] = w;
      }
    }
  }
}

int RcsGraph_countCoupledJoints(const RcsGraph* self)
{
  int n = 0;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (JNT->coupledTo && !JNT->constrained)
      {
        n++;
      }
    }
  }
  return n;
}

/* Applies the coupling rules to an explicit state (src/RcsCore/Rcs_graph.c:2756): slave angle
 * and rate from the master's; reports whether an angle moved by more than 1e-8. */
bool RcsGraph_updateSlaveJoints(const RcsGraph* self, MatNd* q, MatNd* q_dot)
{
  if (!self || !q || q->m != self->dof || q->n != 1 ||
      (q_dot && (q_dot->m != self->dof || q_dot->n != 1)))
  {
    RCSB_FATAL("RcsGraph_updateSlaveJoints: q / q_dot must be %u x 1", self ? self->dof : 0);
    return false;
  }
  bool moved = false;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      const RcsJoint* master = JNT->coupledTo;
      if (!master)
      {
        continue;
      }
      const double qm = q->ele[master->jointIndex];
      const double qs = RcsJoint_computeSlaveJointAngle(JNT, qm);
      moved = moved || fabs(qs - q->ele[JNT->jointIndex]) > 1.0e-8;
      q->ele[JNT->jointIndex] = qs;
      if (q_dot)
      {
        q_dot->ele[JNT->jointIndex] =
          RcsJoint_computeSlaveJointVelocity(JNT, qm, q_dot->ele[master->jointIndex]);
      }
    }
  }
  return moved;
}

/* Coupling matrices of the reduced joint space (src/RcsCore/Rcs_graph.c:2823): column c of A
 * (nJ x nqr) belongs to an uncoupled joint and holds 1 in its own row and the sensitivity
 * d q_slave / d q_master in the rows of its slaves, the column scaled by 1 / |column sum|;
 * invA = (A^T A)^-1 A^T (A^T A is diagonal). Returns the number of coupled joints. */
int RcsGraph_coupledJointMatrix(const RcsGraph* self, MatNd* A, MatNd* invA)
{
  const unsigned int nq = self->nJ;
  const int nCoupled = RcsGraph_countCoupledJoints(self);
  const unsigned int nqr = nq - (unsigned int) nCoupled;
  if (!A || !invA || A->size < nq * nqr || invA->size < nq * nqr)
  {
    RCSB_FATAL("RcsGraph_coupledJointMatrix: A and invA need room for %u x %u doubles", nq, nqr);
    return nCoupled;
  }
  MatNd_reshapeAndSetZero(A, nq, nqr);
  MatNd_reshapeAndSetZero(invA, nqr, nq);
  /* reduced column of every master (unconstrained, uncoupled joint), in jacobiIndex order */
  int* redCol = (int*) malloc(sizeof(int) * (nq + 1));
  double* sum = (double*) calloc(nq + 1, sizeof(double));
  double* sq = (double*) calloc(nq + 1, sizeof(double));
  for (unsigned int i = 0; i < nq; ++i)
  {
    redCol[i] = -1;
  }
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (!JNT->constrained && !JNT->coupledTo)
      {
        redCol[JNT->jacobiIndex] = 0;   /* marks a master; numbered below */
      }
    }
  }
  int next = 0;
  for (unsigned int i = 0; i < nq; ++i)
  {
    if (redCol[i] == 0)
    {
      redCol[i] = next++;
    }
  }
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (JNT->constrained)
      {
        continue;
      }
      const RcsJoint* master = JNT->coupledTo;
      double entry = 1.0;
      int col = redCol[JNT->jacobiIndex];
      if (master)
      {
        if (master->coupledTo || master->jacobiIndex < 0)
        {
          RCSB_FATAL("joint \"%s\" is coupled to \"%s\", which is itself coupled or constrained",
                     JNT->name, master->name);
          continue;
        }
        entry = RcsJoint_computeSlaveJointVelocity(JNT, self->q->ele[master->jointIndex], 1.0);
        col = redCol[master->jacobiIndex];
      }
      A->ele[(unsigned int) JNT->jacobiIndex * nqr + (unsigned int) col] = entry;
      sum[col] += entry;
    }
  }
  for (unsigned int c = 0; c < nqr; ++c)
  {
    const double scale = fabs(sum[c]);
    for (unsigned int r = 0; r < nq && scale > 0.0; ++r)
    {
      const double a = A->ele[r * nqr + c] / scale;
      A->ele[r * nqr + c] = a;
      invA->ele[c * nq + r] = a;
      sq[c] += a * a;
    }
  }
  for (unsigned int c = 0; c < nqr; ++c)
  {
    const double inv = 1.0 / sq[c];
    for (unsigned int r = 0; r < nq; ++r)
    {
      invA->ele[c * nq + r] *= inv;
    }
  }
  free(redCol);
  free(sum);
  free(sq);
  return nCoupled;
}

/* Gradient of the joint-limit cost 1/2 sum w ((q - q0) / range)^2 over the unconstrained joints
 * with a non-empty range, at the graph's state (src/RcsCore/Rcs_kinematics.c:1496); dH is
 * 1 x dof or 1 x nJ. Returns the cost. */
double RcsGraph_jointLimitGradient(const RcsGraph* self, MatNd* dH, RcsStateType type)
{
  const unsigned int n = (type == RcsStateFull) ? self->dof : self->nJ;
  if (!dH || dH->size < n)
  {
    RCSB_FATAL("RcsGraph_jointLimitGradient: dH needs room for %u doubles", n);
    return 0.0;
  }
  MatNd_reshapeAndSetZero(dH, 1, n);
  double cost = 0.0;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      const double range = JNT->q_max - JNT->q_min;
      if (JNT->jacobiIndex < 0 || !(range > 0.0))
      {
        continue;
      }
      const double d = self->q->ele[JNT->jointIndex] - JNT->q0;
      dH->ele[type == RcsStateIK ? JNT->jacobiIndex : JNT->jointIndex] = JNT->weightJL * d / (range * range);
      cost += JNT->weightJL * (d / range) * (d / range);
    }
  }
  return 0.5 * cost;
}

/* World coordinates of a body-fixed point from the frames of the last forward kinematics
 * (src/RcsCore/Rcs_kinematics.c:47); NULL body = world frame, NULL point = body origin. */
void RcsGraph_bodyPoint(const RcsBody* body, const double B_pt[3], double I_pt[3])
{
  if (!body)
  {
    for (int i = 0; i < 3; ++i)
    {
      I_pt[i] = B_pt ? B_pt[i] : 0.0;
    }
    return;
  }
  const HTr* A = body->A_BI;
  for (int i = 0; i < 3; ++i)
  {
    I_pt[i] = A->org[i];
    if (B_pt)
    {
      I_pt[i] += A->rot[0][i] * B_pt[0] + A->rot[1][i] * B_pt[1] + A->rot[2][i] * B_pt[2];
    }
  }
}

void RcsGraph_getDefaultState(const RcsGraph* self, MatNd* q0)
{
  MatNd_reshape(q0, self->dof, 1);
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      q0->ele[JNT->jointIndex] = JNT->q0;
    }
  }
}

/* ---------------------------------------------------------------------------------- */
/* destroy / clone                                                                      */
/* ---------------------------------------------------------------------------------- */

static void destroyJoint(RcsJoint* j)
{
  free(j->name);
  free(j->A_JP);
  free(j->coupledJointName);
  MatNd_destroy(j->couplingFactors);
  free(j);
}

static void destroyBody(RcsBody* b)
{
  for (RcsJoint* j = b->jnt; j;)
  {
    RcsJoint* n = j->next;
    destroyJoint(j);
    j = n;
  }
  if (b->shape)
  {
    for (RcsShape** s = b->shape; *s; ++s)
    {
      free((*s)->meshFile);
      free((*s)->textureFile);
      free((*s)->color);
      free((*s)->material);
      free(*s);
    }
    free(b->shape);
  }
  free(b->name);
  free(b->xmlName);
  free(b->suffix);
  free(b->A_BP);
  free(b->A_BI);
  free(b->Inertia);
  free(b);
}

void rcsb_graph_release_device_cache(RcsGraph* self);   /* rcsb_api.cu (weak here) */
__attribute__((weak)) void rcsb_graph_release_device_cache(RcsGraph* self)
{
  (void) self;
}

void RcsGraph_destroy(RcsGraph* self)
{
  if (!self)
  {
    return;
  }
  rcsb_graph_release_device_cache(self);
  RcsBody* b = self->root;
  while (b)
  {
    RcsBody* n = RcsBody_depthFirstTraversalGetNext(b);
    destroyBody(b);
    b = n;
  }
  MatNd_destroy(self->q);
  MatNd_destroy(self->q_dot);
  free(self->xmlFile);
  free(self);
}

static char* cloneString(const char* s)
{
  if (!s)
  {
    return NULL;
  }
  char* d = (char*) malloc(strlen(s) + 1);
  strcpy(d, s);
  return d;
}

/* Deep copy preserving order and indices (src/RcsCore/Rcs_graph.c:2193). */
RcsGraph* RcsGraph_clone(const RcsGraph* src)
{
  if (!src)
  {
    return NULL;
  }

  const unsigned int nb = RcsGraph_numBodies(src);
  const RcsBody** srcBodies = (const RcsBody**) calloc(nb + 1, sizeof(RcsBody*));
  RcsBody** dstBodies = (RcsBody**) calloc(nb + 1, sizeof(RcsBody*));
  RcsJoint** dstJoints = (RcsJoint**) calloc(src->dof + 1, sizeof(RcsJoint*));

  RcsGraph* dst = (RcsGraph*) calloc(1, sizeof(RcsGraph));
  dst->xmlFile = cloneString(src->xmlFile);
  dst->q = MatNd_create(0, 1);

  unsigned int i = 0;
  RCSGRAPH_TRAVERSE_BODIES(src)
  {
    srcBodies[i++] = BODY;
  }

  for (i = 0; i < nb; ++i)
  {
    const RcsBody* s = srcBodies[i];
    RcsBody* d = (RcsBody*) calloc(1, sizeof(RcsBody));
    dstBodies[i] = d;
    d->m = s->m;
    d->rigid_body_joints = s->rigid_body_joints;
    d->physicsSim = s->physicsSim;
    memcpy(d->x_dot, s->x_dot, sizeof(d->x_dot));
    memcpy(d->omega, s->omega, sizeof(d->omega));
    d->confidence = s->confidence;
    d->name = cloneString(s->name);
    d->xmlName = cloneString(s->xmlName);
    d->suffix = cloneString(s->suffix);
    d->A_BP = s->A_BP ? HTr_clone(s->A_BP) : NULL;
    d->A_BI = HTr_clone(s->A_BI);
    d->Inertia = HTr_clone(s->Inertia);

    unsigned int ns = 0;
    if (s->shape)
    {
      while (s->shape[ns])
      {
        ns++;
      }
    }
    d->shape = (RcsShape**) calloc(ns + 1, sizeof(RcsShape*));
    for (unsigned int k = 0; k < ns; ++k)
    {
      d->shape[k] = (RcsShape*) malloc(sizeof(RcsShape));
      memcpy(d->shape[k], s->shape[k], sizeof(RcsShape));
      d->shape[k]->meshFile = cloneString(s->shape[k]->meshFile);
      d->shape[k]->textureFile = cloneString(s->shape[k]->textureFile);
      d->shape[k]->color = cloneString(s->shape[k]->color);
      d->shape[k]->material = cloneString(s->shape[k]->material);
      d->shape[k]->userData = NULL;
    }

    /* DFS order guarantees that the parent has been cloned already */
    RcsBody* parent = NULL;
    for (unsigned int k = 0; k < i && !parent; ++k)
    {
      if (srcBodies[k] == s->parent)
      {
        parent = dstBodies[k];
      }
    }
    RcsGraph_insertBody(dst, parent, d);

    for (const RcsJoint* sj = s->jnt; sj; sj = sj->next)
    {
      RcsJoint* dj = (RcsJoint*) calloc(1, sizeof(RcsJoint));
      RcsGraph_insertJoint(dst, d, dj);
      RcsJoint* prevLink = dj->prev;
      RcsJoint* nextLink = dj->next;
      *dj = *sj;
      dj->prev = prevLink;
      dj->next = nextLink;
      dj->coupledTo = NULL;
      dj->name = cloneString(sj->name);
      dj->A_JP = sj->A_JP ? HTr_clone(sj->A_JP) : NULL;
      dj->coupledJointName = cloneString(sj->coupledJointName);
      dj->couplingFactors = NULL;
      if (sj->couplingFactors)
      {
        dj->couplingFactors = MatNd_create(sj->couplingFactors->m, sj->couplingFactors->n);
        memcpy(dj->couplingFactors->ele, sj->couplingFactors->ele,
               sizeof(double) * sj->couplingFactors->m * sj->couplingFactors->n);
      }
      dstJoints[sj->jointIndex] = dj;
    }
  }

  /* indices were copied verbatim; relink the coupling pointers */
  for (i = 0; i < nb; ++i)
  {
    for (const RcsJoint* sj = srcBodies[i]->jnt; sj; sj = sj->next)
    {
      if (sj->coupledTo)
      {
        dstJoints[sj->jointIndex]->coupledTo = dstJoints[sj->coupledTo->jointIndex];
      }
    }
  }

  dst->dof = src->dof;
  dst->nJ = src->nJ;
  MatNd_reshape(dst->q, src->dof, 1);
  memcpy(dst->q->ele, src->q->ele, sizeof(double) * src->dof);
  dst->q_dot = MatNd_create(src->dof, 1);
  memcpy(dst->q_dot->ele, src->q_dot->ele, sizeof(double) * src->dof);

  free(srcBodies);
  free(dstBodies);
  free(dstJoints);
  return dst;
}

/* ---------------------------------------------------------------------------------- */
/* validation (contract: src/RcsCore/Rcs_graph.h:452-482)                               */
/* ---------------------------------------------------------------------------------- */

bool RcsGraph_check(const RcsGraph* self, int* nErrorsOut, int* nWarningsOut)
{
  int nErrors = 0, nWarnings = 0;

  if (!self)
  {
    if (nErrorsOut)
    {
      *nErrorsOut = 1;
    }
    if (nWarningsOut)
    {
      *nWarningsOut = 0;
    }
    return false;
  }

  unsigned int nJoints = 0, nFree = 0;
  static const int rbjOrder[6] = {RCSJOINT_TRANS_X, RCSJOINT_TRANS_Y, RCSJOINT_TRANS_Z,
                                  RCSJOINT_ROT_X, RCSJOINT_ROT_Y, RCSJOINT_ROT_Z};

  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    /* connectivity */
    if (BODY->firstChild && BODY->firstChild->parent != BODY)
    {
      nErrors++;
    }
    if (BODY->next && BODY->next->prev != BODY)
    {
      nErrors++;
    }
    if (BODY->parent && !BODY->prev && BODY->parent->firstChild != BODY)
    {
      nErrors++;
    }
    if (BODY->m < 0.0)
    {
      nErrors++;
      RCSB_LOG(1, "Body \"%s\" has negative mass", BODY->name);
    }
    if (BODY->m == 0.0 && BODY->Inertia)
    {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          if (BODY->Inertia->rot[i][j] != 0.0)
          {
            nWarnings++;
            i = 3;
            break;
          }
    }

    /* duplicate names */
    for (RcsBody* other = RcsBody_depthFirstTraversalGetNext(BODY); other;
         other = RcsBody_depthFirstTraversalGetNext(other))
    {
      if (strcmp(other->name, BODY->name) == 0)
      {
        nErrors++;
        RCSB_LOG(1, "Duplicate body name \"%s\"", BODY->name);
      }
    }

    if (BODY->rigid_body_joints)
    {
      int k = 0;
      RCSBODY_TRAVERSE_JOINTS(BODY)
      {
        if (k < 6 && JNT->type != rbjOrder[k])
        {
          nErrors++;
        }
        if (k > 0 && JNT->A_JP && !HTr_isIdentity(JNT->A_JP))
        {
          nErrors++;
        }
        k++;
      }
      if (k != 6)
      {
        nErrors++;
      }
    }

    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      nJoints++;
      if (JNT->jointIndex < 0 || JNT->jointIndex >= (int) self->dof)
      {
        nErrors++;
      }
      if (JNT->jacobiIndex < -1 || JNT->jacobiIndex >= (int) self->nJ)
      {
        nErrors++;
      }
      if (JNT->constrained != (JNT->jacobiIndex == -1))
      {
        nErrors++;
      }
      if (!JNT->constrained)
      {
        nFree++;
      }
      if (JNT->dirIdx < 0 || JNT->dirIdx > 2 || JNT->type < 0 || JNT->type > 5 ||
          JNT->dirIdx != JNT->type % 3)
      {
        nErrors++;
      }
      /* the reference's check counts polynomial couplings (5 / 9 coefficients) as errors although
       * its loader and RcsGraph_setState accept them (src/RcsCore/Rcs_graph.c:2047-2063) */
      if (JNT->couplingFactors && JNT->couplingFactors->size != 1)
      {
        nErrors++;
        RCSB_LOG(1, "Incorrect number of coupling factors in joint \"%s\": %u", JNT->name,
                 JNT->couplingFactors->size);
      }
      if ((JNT->coupledJointName != NULL) != (JNT->coupledTo != NULL))
      {
        nErrors++;
      }
      if (JNT->coupledTo && (JNT->coupledTo->coupledTo || !JNT->couplingFactors))
      {
        nErrors++;
      }
      if (JNT->q0 < JNT->q_min || JNT->q0 > JNT->q_max)
      {
        nWarnings++;
      }
      if (JNT->jointIndex >= 0 && JNT->jointIndex < (int) self->dof)
      {
        const double qi = self->q->ele[JNT->jointIndex];
        if (!isfinite(qi) || !isfinite(self->q_dot->ele[JNT->jointIndex]))
        {
          nErrors++;
        }
        else if (qi < JNT->q_min || qi > JNT->q_max)
        {
          nWarnings++;
        }
      }
    }
  }

  if (nJoints != self->dof || self->q->m != self->dof || nFree != self->nJ)
  {
    nErrors++;
  }

  if (nErrorsOut)
  {
    *nErrorsOut = nErrors;
  }
  if (nWarningsOut)
  {
    *nWarningsOut = nWarnings;
  }
  return nErrors == 0 && nWarnings == 0;
}

/* ---------------------------------------------------------------------------------- */
/* canonical layout dump                                                                */
/* ---------------------------------------------------------------------------------- */

typedef struct
{
  char* buf;
  size_t len, cap;
} StrBuf;

static void sbPut(StrBuf* s, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  int n = vsnprintf(NULL, 0, fmt, ap);
  va_end(ap);
  if (n < 0)
  {
    return;
  }
  if (s->len + (size_t) n + 1 > s->cap)
  {
    s->cap = (s->cap + (size_t) n + 1) * 2;
    s->buf = (char*) realloc(s->buf, s->cap);
  }
  va_start(ap, fmt);
  vsnprintf(s->buf + s->len, (size_t) n + 1, fmt, ap);
  va_end(ap);
  s->len += (size_t) n;
}

static void sbHTr(StrBuf* s, const HTr* A)
{
  if (!A)
  {
    sbPut(s, "null");
    return;
  }
  const double* v = &A->org[0];
  sbPut(s, "[");
  for (int i = 0; i < 12; ++i)
  {
    sbPut(s, "%s%.17g", i ? "," : "", i < 3 ? v[i] : A->rot[(i - 3) / 3][(i - 3) % 3]);
  }
  sbPut(s, "]");
}

char* RcsGraph_layoutJSON(const RcsGraph* self)
{
  StrBuf s = {NULL, 0, 0};
  sbPut(&s, "{\"dof\":%u,\"nJ\":%u,\"nBodies\":%u,\"bodies\":[", self->dof, self->nJ,
        RcsGraph_numBodies(self));

  bool first = true;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    sbPut(&s, "%s{\"name\":\"%s\",\"parent\":%d,\"firstChild\":%d,\"lastChild\":%d,"
          "\"next\":%d,\"prev\":%d,\"m\":%.17g,\"rigid_body_joints\":%s,\"physicsSim\":%d,"
          "\"A_BP\":", first ? "" : ",", BODY->name, RcsGraph_bodyIndex(self, BODY->parent),
          RcsGraph_bodyIndex(self, BODY->firstChild), RcsGraph_bodyIndex(self, BODY->lastChild),
          RcsGraph_bodyIndex(self, BODY->next), RcsGraph_bodyIndex(self, BODY->prev), BODY->m,
          BODY->rigid_body_joints ? "true" : "false", BODY->physicsSim);
    sbHTr(&s, BODY->A_BP);
    sbPut(&s, ",\"Inertia\":");
    sbHTr(&s, BODY->Inertia);
    sbPut(&s, ",\"joints\":[");
    bool fj = true;
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      sbPut(&s, "%s%d", fj ? "" : ",", JNT->jointIndex);
      fj = false;
    }
    int nShapes = 0;
    if (BODY->shape)
    {
      while (BODY->shape[nShapes])
      {
        nShapes++;
      }
    }
    sbPut(&s, "],\"nShapes\":%d}", nShapes);
    first = false;
  }

  sbPut(&s, "],\"joints\":[");
  first = true;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      sbPut(&s, "%s{\"name\":\"%s\",\"body\":%d,\"type\":%d,\"dirIdx\":%d,\"jointIndex\":%d,"
            "\"jacobiIndex\":%d,\"constrained\":%s,\"ctrlType\":%d,\"prev\":%d,\"coupledTo\":%d",
            first ? "" : ",", JNT->name, RcsGraph_bodyIndex(self, BODY), JNT->type, JNT->dirIdx,
            JNT->jointIndex, JNT->jacobiIndex, JNT->constrained ? "true" : "false", JNT->ctrlType,
            JNT->prev ? JNT->prev->jointIndex : -1,
            JNT->coupledTo ? JNT->coupledTo->jointIndex : -1);
      const char* names[] = {"q0", "q_init", "q_min", "q_max", "weightJL", "weightCA",
                             "weightMetric", "maxTorque", "speedLimit", "accLimit", "decLimit"};
      const double vals[] = {JNT->q0, JNT->q_init, JNT->q_min, JNT->q_max, JNT->weightJL,
                             JNT->weightCA, JNT->weightMetric, JNT->maxTorque, JNT->speedLimit,
                             JNT->accLimit, JNT->decLimit};
      for (int k = 0; k < 11; ++k)
      {
        sbPut(&s, ",\"%s\":%.17g", names[k], vals[k]);
      }
      sbPut(&s, ",\"couplingFactors\":[");
      if (JNT->couplingFactors)
      {
        for (unsigned int k = 0; k < JNT->couplingFactors->m * JNT->couplingFactors->n; ++k)
        {
          sbPut(&s, "%s%.17g", k ? "," : "", JNT->couplingFactors->ele[k]);
        }
      }
      sbPut(&s, "],\"A_JP\":");
      sbHTr(&s, JNT->A_JP);
      sbPut(&s, "}");
      first = false;
    }
  }

  sbPut(&s, "],\"q\":[");
  for (unsigned int i = 0; i < self->dof; ++i)
  {
    sbPut(&s, "%s%.17g", i ? "," : "", self->q->ele[i]);
  }
  sbPut(&s, "]}");
  return s.buf;
}
