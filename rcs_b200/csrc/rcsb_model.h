/*
 * Flattened, structure-of-arrays description of a kinematic tree ("flat model"): what the
 * CUDA kernels read instead of the pointer-linked RcsGraph. One contiguous blob, built on
 * the host by rcsb_model.c, uploaded once per model and staged into shared memory once per
 * thread block.
 *
 * Bodies are in depth-first order (index = reference body order), joints in jointIndex
 * order; the joints of a body are contiguous. Every array starts 8-byte aligned at the byte
 * offset recorded in the header.
 */
#ifndef RCSB_MODEL_H
#define RCSB_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

/* body flags */
#define RCSB_BF_HAS_ABP 1      /* A_BP is not the identity */
#define RCSB_BF_HAS_MASS 2     /* m != 0: contributes to the kinetic terms */
#define RCSB_BF_LEAF 4         /* no children (and not a root): evaluated on a register copy */
#define RCSB_BF_ABP_ROT_IDENT 8   /* rotation part of A_BP is exactly the identity */
#define RCSB_BF_ABP_NO_TRANS 16   /* translation part of A_BP is exactly zero */
/* joint flags */
#define RCSB_JF_HAS_AJP 1
#define RCSB_JF_CONSTRAINED 2
#define RCSB_JF_ROT 4          /* rotational (else translational) */
#define RCSB_JF_AJP_ROT_IDENT 8   /* rotation part of A_JP is exactly the identity */
#define RCSB_JF_AJP_NO_TRANS 16   /* translation part of A_JP is exactly zero */
/* the *_ROT_IDENT / *_NO_TRANS bits share their values between body and joint flags */
#define RCSB_REL_ROT_IDENT 8
#define RCSB_REL_NO_TRANS 16

/* frame sources of a body (bodyLoad) */
#define RCSB_LOAD_CONT (-2)    /* parent is the previously visited body: keep registers */
#define RCSB_LOAD_IDENT (-1)   /* root body: start from the world frame */

enum
{
  /* int arrays, length nb */
  RCSB_A_BODY_PARENT = 0,
  RCSB_A_BODY_JNT_START,
  RCSB_A_BODY_JNT_COUNT,
  RCSB_A_BODY_FLAGS,
  RCSB_A_BODY_LOAD,       /* RCSB_LOAD_* or frame-stack slot holding the parent frame */
  RCSB_A_BODY_SAVE,       /* frame-stack slot to keep this body's frame in, or -1 */
  RCSB_A_BODY_LASTJNT,    /* jointIndex of RcsBody_lastJointBeforeBody(body), -1 if none */
  /* int arrays, length dof */
  RCSB_A_JNT_TYPE,        /* RCSJOINT_TYPE 0..5; dirIdx = type % 3 */
  RCSB_A_JNT_FLAGS,
  RCSB_A_JNT_JACOBI,      /* jacobiIndex, -1 if constrained */
  RCSB_A_JNT_PREV,        /* jointIndex of jnt->prev (chain towards the root), -1 at the end */
  RCSB_A_JNT_CPL,         /* index into the coupling array if this joint is a slave, else -1 */
  RCSB_A_JNT_BODY,        /* body index the joint belongs to */
  /* double arrays */
  RCSB_A_BODY_ABP,        /* nb x 12 (org, row-major rot) */
  RCSB_A_BODY_MASS,       /* nb */
  RCSB_A_BODY_COM,        /* nb x 3, body frame */
  RCSB_A_BODY_INERTIA,    /* nb x 9, about the COM, body frame */
  RCSB_A_JNT_AJP,         /* dof x 12 */
  RCSB_A_JNT_QREF,        /* dof: state of the graph at model creation */
  RCSB_A_JNT_Q0,          /* dof */
  RCSB_A_JNT_QMIN,        /* dof */
  RCSB_A_JNT_QMAX,        /* dof */
  RCSB_A_JNT_WJL,         /* dof */
  RCSB_A_JNT_WMETRIC,     /* dof */
  RCSB_A_CPL,             /* nCpl x RcsbCoupling */
  RCSB_A_COUNT
};

/* kinematic coupling of a slave joint (src/RcsCore/Rcs_joint.c:375-462) */
typedef struct
{
  int slave;          /* jointIndex */
  int master;         /* jointIndex */
  int nCoef;          /* 1 (linear), 5 or 9 (polynomial) */
  int pad;
  double coef[9];
  double sQInit, sQMin, sQMax;   /* slave */
  double mQInit, mQMin, mQMax;   /* master */
} RcsbCoupling;

typedef struct
{
  int nb;             /* bodies */
  int dof;            /* joints */
  int nJ;             /* unconstrained joints */
  int nCpl;           /* slave joints */
  int nSlots;         /* depth of the parent-frame stack a depth-first walk needs */
  int nMassive;       /* bodies with m != 0 */
  int totalBytes;     /* header + arrays */
  int pad;
  int off[RCSB_A_COUNT];
} RcsbFlatHeader;

#ifdef __cplusplus
}
#endif

#endif
