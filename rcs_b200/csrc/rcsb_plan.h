/*
 * Per-query plans: small device-resident blobs that tell a kernel which results the caller
 * asked for (selected bodies, effectors, tasks), built on the host once per distinct query
 * and cached in the model.
 */
#ifndef RCSB_PLAN_H
#define RCSB_PLAN_H

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Jacobian scratch (shared memory, one column per state, structure-of-arrays over rows):
 *   row k * nChain + s        k = 0..2   axis of chain joint s            (-> JR column)
 *   row (3 + k) * nChain + s  k = 0..2   origin of chain joint s          (-> JT column *)
 *   row 6 * nChain + 3 e + k             world point of effector e
 * (*) "in place" plans: every chain joint serves exactly one effector, the owning thread
 *     replaces the joint origin by axis x (point - origin) (or by the axis for a
 *     translational joint) once the effector point is known, and the copy-out is a plain
 *     gather: jtabT / jtabR give the scratch row of every element of a state's record, -1
 *     for structural zeros.
 *     Otherwise (chains share joints) the copy-out evaluates the cross product itself from
 *     jtab entries  slot | row << 16 | isRot << 18 | effector << 20  (-1 = zero).
 */
#define RCSB_JTAB_SLOT(t) ((t) & 0xffff)
#define RCSB_JTAB_ROW(t) (((t) >> 16) & 3)
#define RCSB_JTAB_ISROT(t) (((t) >> 18) & 1)
#define RCSB_JTAB_EFF(t) ((t) >> 20)

typedef struct
{
  int nSel;        /* frames written per state */
  int nEff;        /* effectors */
  int nChain;      /* joints lying on some effector's chain = scratch slots */
  int nScrRows;    /* shared-memory scratch rows per state: 6 nChain + 3 nEff */
  int jacRecord;   /* doubles per state in JT (and in JR): nEff * 3 * nJ */
  int inPlace;     /* 1: in-place plan (see above) */
  int totalBytes;  /* multiple of 16 */
  int offBodyOut;  /* int[nb]: output slot of the body's frame, -1 = not selected */
  int offJntScr;   /* int[dof]: scratch slot of the joint, -1 = not on a chain */
  int offBodyEffStart;  /* int[nb + 1]: CSR over effectors attached to each body */
  int offBodyEffList;   /* int[nEff] */
  int offEffPt;    /* double[nEff * 3] body-fixed points */
  int offJTab;     /* int[jacRecord]: jtab, or jtabT for in-place plans */
  int offJTabR;    /* int[jacRecord]: jtabR (in-place plans) */
  int offEffChainStart; /* int[nEff + 1]: CSR over (slot | isRot << 16) per effector */
  int offEffChain;      /* int[sum of chain lengths] */
  float invJacRecord;
  int pad;
} RcsbFkPlanHeader;

#ifdef __cplusplus
}
#endif

#endif
