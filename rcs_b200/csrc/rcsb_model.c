/*
 * Flattener: RcsGraph -> flat model blob (rcsb_model.h).
 *
 * Besides copying parameters into depth-first arrays it precomputes what the kernels would
 * otherwise have to chase pointers for:
 *  - where each body finds its parent frame during a depth-first walk (registers if the
 *    parent was visited just before, else a slot of a small per-state frame stack) and
 *    which bodies must park their frame in such a slot (bodies with more than one child);
 *  - the backward joint chain of every body (RcsBody_lastJointBeforeBody + jnt->prev,
 *    src/RcsCore/Rcs_body.c:1706, Rcs_graph.c:2604-2655) as integer links.
 */
#include "rcsb_internal.h"
#include "rcsb_model.h"

#include <stdlib.h>
#include <string.h>

static int alignUp16(int x)
{
  return (x + 15) & ~15;
}

/* Exactly-identity rotation / exactly-zero translation of a relative transform: the kernels
 * skip the corresponding products (the result is bit-identical). */
static int relFlags(const HTr* A)
{
  int f = 0;
  bool ident = true;
  for (int i = 0; i < 3; ++i)
  {
    for (int j = 0; j < 3; ++j)
    {
      ident = ident && (A->rot[i][j] == (i == j ? 1.0 : 0.0));
    }
  }
  if (ident)
  {
    f |= RCSB_REL_ROT_IDENT;
  }
  if (A->org[0] == 0.0 && A->org[1] == 0.0 && A->org[2] == 0.0)
  {
    f |= RCSB_REL_NO_TRANS;
  }
  return f;
}

int rcsb_flatten(const RcsGraph* graph, void** blobOut, size_t* bytesOut)
{
  if (!graph || !blobOut || !bytesOut)
  {
    rcsb_set_error("rcsb_flatten: NULL argument");
    return -1;
  }

  const int nb = (int) RcsGraph_numBodies(graph);
  const int dof = (int) graph->dof;

  RcsBody** bodies = (RcsBody**) calloc((size_t) nb + 1, sizeof(RcsBody*));
  int nCpl = 0, k = 0;
  RCSGRAPH_TRAVERSE_BODIES(graph)
  {
    bodies[k++] = BODY;
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (JNT->coupledTo)
      {
        nCpl++;
      }
    }
  }

  /* layout */
  RcsbFlatHeader hdr;
  memset(&hdr, 0, sizeof(hdr));
  hdr.nb = nb;
  hdr.dof = dof;
  hdr.nJ = (int) graph->nJ;
  hdr.nCpl = nCpl;

  int sizes[RCSB_A_COUNT];
  for (int a = 0; a < RCSB_A_COUNT; ++a)
  {
    if (a <= RCSB_A_BODY_LASTJNT)
    {
      sizes[a] = nb * (int) sizeof(int);
    }
    else if (a <= RCSB_A_JNT_BODY)
    {
      sizes[a] = dof * (int) sizeof(int);
    }
    else
    {
      sizes[a] = 0;
    }
  }
  sizes[RCSB_A_BODY_ABP] = nb * 12 * (int) sizeof(double);
  sizes[RCSB_A_BODY_MASS] = nb * (int) sizeof(double);
  sizes[RCSB_A_BODY_COM] = nb * 3 * (int) sizeof(double);
  sizes[RCSB_A_BODY_INERTIA] = nb * 9 * (int) sizeof(double);
  sizes[RCSB_A_JNT_AJP] = dof * 12 * (int) sizeof(double);
  for (int a = RCSB_A_JNT_QREF; a <= RCSB_A_JNT_WMETRIC; ++a)
  {
    sizes[a] = dof * (int) sizeof(double);
  }
  sizes[RCSB_A_CPL] = nCpl * (int) sizeof(RcsbCoupling);

  int off = alignUp16((int) sizeof(RcsbFlatHeader));
  for (int a = 0; a < RCSB_A_COUNT; ++a)
  {
    hdr.off[a] = off;
    off = alignUp16(off + sizes[a]);
  }
  hdr.totalBytes = alignUp16(off + 8);   /* multiple of 16 for vector copies */
  hdr.totalBytes = (hdr.totalBytes + 15) & ~15;

  char* blob = (char*) calloc(1, (size_t) hdr.totalBytes);
#define IARR(a) ((int*) (blob + hdr.off[a]))
#define DARR(a) ((double*) (blob + hdr.off[a]))

  int* nChildren = (int*) calloc((size_t) nb + 1, sizeof(int));
  int* slotOf = (int*) calloc((size_t) nb + 1, sizeof(int));
  int* branchDepth = (int*) calloc((size_t) nb + 1, sizeof(int));
  int nSlots = 0;

  for (int b = 0; b < nb; ++b)
  {
    const int p = RcsGraph_bodyIndex(graph, bodies[b]->parent);
    IARR(RCSB_A_BODY_PARENT)[b] = p;
    if (p >= 0)
    {
      nChildren[p]++;
    }
  }

  /* Where does each body find its parent frame? Simulate the depth-first walk. A leaf body
   * is evaluated on a register copy, so after it the registers hold its parent's frame
   * again; after any other body they hold that body's frame. A parent whose frame is not
   * in registers when one of its children comes up has to park it in a frame-stack slot. */
  int* needSlot = (int*) calloc((size_t) nb + 1, sizeof(int));
  int* loadKind = (int*) calloc((size_t) nb + 1, sizeof(int));
  {
    int cur = -1;   /* body whose frame the registers hold */
    for (int b = 0; b < nb; ++b)
    {
      const int p = IARR(RCSB_A_BODY_PARENT)[b];
      if (p < 0)
      {
        loadKind[b] = RCSB_LOAD_IDENT;
      }
      else if (cur == p)
      {
        loadKind[b] = RCSB_LOAD_CONT;
      }
      else
      {
        loadKind[b] = 0;   /* from the parent's slot */
        needSlot[p] = 1;
      }
      cur = (nChildren[b] == 0 && p >= 0) ? p : b;
    }
  }

  /* slots are used like a stack: depth = number of slot-holding ancestors */
  for (int b = 0; b < nb; ++b)
  {
    const int p = IARR(RCSB_A_BODY_PARENT)[b];
    const int inherited = p >= 0 ? branchDepth[p] : 0;
    if (needSlot[b])
    {
      slotOf[b] = inherited;
      branchDepth[b] = inherited + 1;
      if (branchDepth[b] > nSlots)
      {
        nSlots = branchDepth[b];
      }
    }
    else
    {
      slotOf[b] = -1;
      branchDepth[b] = inherited;
    }
  }
  hdr.nSlots = nSlots;

  int ci = 0;
  for (int b = 0; b < nb; ++b)
  {
    const RcsBody* B = bodies[b];
    const int p = IARR(RCSB_A_BODY_PARENT)[b];
    int flags = 0;

    IARR(RCSB_A_BODY_SAVE)[b] = slotOf[b];
    IARR(RCSB_A_BODY_LOAD)[b] = (loadKind[b] == 0) ? slotOf[p] : loadKind[b];
    if (nChildren[b] == 0 && p >= 0)
    {
      flags |= RCSB_BF_LEAF;
    }

    double* abp = DARR(RCSB_A_BODY_ABP) + 12 * b;
    if (B->A_BP)
    {
      flags |= RCSB_BF_HAS_ABP | relFlags(B->A_BP);
      memcpy(abp, B->A_BP, sizeof(HTr));
    }
    else
    {
      abp[3] = abp[7] = abp[11] = 1.0;
    }

    if (B->m != 0.0)
    {
      flags |= RCSB_BF_HAS_MASS;
      hdr.nMassive++;
    }
    DARR(RCSB_A_BODY_MASS)[b] = B->m;
    if (B->Inertia)
    {
      memcpy(DARR(RCSB_A_BODY_COM) + 3 * b, B->Inertia->org, 3 * sizeof(double));
      memcpy(DARR(RCSB_A_BODY_INERTIA) + 9 * b, B->Inertia->rot, 9 * sizeof(double));
    }
    IARR(RCSB_A_BODY_FLAGS)[b] = flags;

    const RcsJoint* last = RcsBody_lastJointBeforeBody(B);
    IARR(RCSB_A_BODY_LASTJNT)[b] = last ? last->jointIndex : -1;

    IARR(RCSB_A_BODY_JNT_START)[b] = B->jnt ? B->jnt->jointIndex : 0;
    int cnt = 0;
    for (const RcsJoint* J = B->jnt; J; J = J->next)
    {
      const int j = J->jointIndex;
      if (j != IARR(RCSB_A_BODY_JNT_START)[b] + cnt || j < 0 || j >= dof)
      {
        rcsb_set_error("rcsb_flatten: joints of body \"%s\" are not contiguous in jointIndex "
                       "order (call RcsGraph_makeJointsConsistent first)", B->name);
        free(blob);
        free(bodies);
        free(nChildren);
        free(slotOf);
        free(branchDepth);
        free(needSlot);
        free(loadKind);
        return -1;
      }
      cnt++;

      int jf = 0;
      IARR(RCSB_A_JNT_TYPE)[j] = J->type;
      if (J->type <= RCSJOINT_ROT_Z)
      {
        jf |= RCSB_JF_ROT;
      }
      if (J->constrained)
      {
        jf |= RCSB_JF_CONSTRAINED;
      }
      double* ajp = DARR(RCSB_A_JNT_AJP) + 12 * j;
      if (J->A_JP)
      {
        jf |= RCSB_JF_HAS_AJP | relFlags(J->A_JP);
        memcpy(ajp, J->A_JP, sizeof(HTr));
      }
      else
      {
        ajp[3] = ajp[7] = ajp[11] = 1.0;
      }
      IARR(RCSB_A_JNT_FLAGS)[j] = jf;
      IARR(RCSB_A_JNT_JACOBI)[j] = J->jacobiIndex;
      IARR(RCSB_A_JNT_PREV)[j] = J->prev ? J->prev->jointIndex : -1;
      IARR(RCSB_A_JNT_BODY)[j] = b;
      IARR(RCSB_A_JNT_CPL)[j] = -1;

      DARR(RCSB_A_JNT_QREF)[j] = graph->q->ele[j];
      DARR(RCSB_A_JNT_Q0)[j] = J->q0;
      DARR(RCSB_A_JNT_QMIN)[j] = J->q_min;
      DARR(RCSB_A_JNT_QMAX)[j] = J->q_max;
      DARR(RCSB_A_JNT_WJL)[j] = J->weightJL;
      DARR(RCSB_A_JNT_WMETRIC)[j] = J->weightMetric;

      if (J->coupledTo)
      {
        RcsbCoupling* c = ((RcsbCoupling*) (blob + hdr.off[RCSB_A_CPL])) + ci;
        IARR(RCSB_A_JNT_CPL)[j] = ci++;
        c->slave = j;
        c->master = J->coupledTo->jointIndex;
        c->nCoef = (int) J->couplingFactors->size;
        for (int i = 0; i < c->nCoef && i < 9; ++i)
        {
          c->coef[i] = J->couplingFactors->ele[i];
        }
        c->sQInit = J->q_init;
        c->sQMin = J->q_min;
        c->sQMax = J->q_max;
        c->mQInit = J->coupledTo->q_init;
        c->mQMin = J->coupledTo->q_min;
        c->mQMax = J->coupledTo->q_max;
      }
    }
    IARR(RCSB_A_BODY_JNT_COUNT)[b] = cnt;
  }
#undef IARR
#undef DARR

  memcpy(blob, &hdr, sizeof(hdr));
  *blobOut = blob;
  *bytesOut = (size_t) hdr.totalBytes;

  free(bodies);
  free(nChildren);
  free(slotOf);
  free(branchDepth);
  free(needSlot);
  free(loadKind);
  return 0;
}
