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

/* jtab entry of the Jacobian copy-out (one per element of a state's 3 x nJ x nEff record):
 * -1 = structural zero, else  slot | row << 16 | isRot << 18 | effector << 20            */
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
  int totalBytes;  /* multiple of 16 */
  int offBodyOut;  /* int[nb]: output slot of the body's frame, -1 = not selected */
  int offJntScr;   /* int[dof]: scratch slot of the joint, -1 = not on a chain */
  int offBodyEffStart;  /* int[nb + 1]: CSR over effectors attached to each body */
  int offBodyEffList;   /* int[nEff] */
  int offEffPt;    /* double[nEff * 3] body-fixed points */
  int offJTab;     /* int[jacRecord] */
  float invJacRecord;
  int pad;
} RcsbFkPlanHeader;

#ifdef __cplusplus
}
#endif

#endif
