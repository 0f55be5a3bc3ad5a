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
  /* fused mass matrix (fkKernel<..., MASS>: models with nJ <= 8 and symmetric inertia tensors):
   * every Jacobian column has a scratch slot (slot == column), the bodies' path masks say
   * which columns move them, M is staged in nJ x nJ extra scratch rows from mStageRow on */
  int velStream;       /* 1: the selected bodies come out in depth-first order, so x_dot / omega
                          can stream through 32-record windows (fkKernel<VEL>) */
  int mass;            /* 1: plan carries the tables below */
  int nJ;
  int rotMask;         /* bit k: column k is a rotational joint */
  int mStageRow;       /* first scratch row of the M staging area */
  int offBodyPathMask; /* int[nb]: bit k set if column k moves the body */
  int relMask[8];      /* bit i of relMask[k]: column i is a proper ancestor of column k */
  float invNJ2;
  /* serial-chain variant of the fused mass matrix (fkKernel<..., MASS = 2>): every column is an
   * ancestor of the next one (relMask[k] == (1 << k) - 1) and the frame of every massive body
   * is among the outputs. The massive bodies are then listed by descending number of columns
   * that move them, and M is built after the walk from composite inertias, tip to root. */
  int massChain;
  int nMass;           /* entries of the list */
  int offMassList;     /* int4[nMass]: {body, output slot, number of columns moving it, 0} */
  int zeroRow;         /* scratch row kept at zero (last row): structural zeros of JT / JR */
  /* pruned walk (fkKernel<..., PRUNED>): when only some bodies are asked for, the walk visits
   * them, the effector bodies and their ancestors only; the frame-stack program is rebuilt for
   * that sub-tree */
  int nVisit;          /* 0: the plan has no pruned walk (every body is needed) */
  int offVisit;        /* int4[nVisit]: {body, parent-frame source (RCSB_LOAD_*, or slot), slot to
                          save the frame in or -1, 1 if the body is a leaf of the sub-tree} */
  int offJntNext;      /* int[dof]: joint visited after joint j, -1 = none */
  int firstJnt;        /* first joint of the pruned walk, -1 = none */
  int prunedSlots;     /* frame-stack depth of the pruned walk */
  /* Serial chains of rotations (fkKernel<..., CHAIN = nJ>): no constrained or coupled joint, every
   * joint a rotation, every column an ancestor of the next one, nJ <= 8 - an arm with any number of
   * joint-less side bodies. "Level" L of a body = number of columns that move it; its frame is a
   * CONSTANT transform X_b away from the base of its level (the frame right after the rotation of
   * column L - 1; the world frame for L = 0), and the base of column c is a constant transform G_c
   * away from the previous base. The walk is then straight-line code over the columns (compile-time
   * count) with the bodies of a level as side outputs; nothing of the flat model is interpreted.
   * Bases are kept with their rows cyclically shifted so that the column's axis is row 2 (as in the
   * IK chain kernels, RcsbIkPlanHeader::offProg): the shifts are folded into G_c / X_b / the mass
   * constants on the host as exact permutations. */
  int chainNc;          /* 0: not eligible */
  int offChainG;        /* double[chainNc][12] */
  int offChainX;        /* double[nb][12], in the order of offChainOut */
  int offChainOutStart; /* int[chainNc + 2]: CSR over the levels 0 .. chainNc */
  int offChainOut;      /* int4[nb]: {body, output slot or -1, RCSB_REL_* flags of X | 1 if the frame is
                           needed (selected, or an effector is attached), 0} */
  int offChainMass;     /* double[chainNc + 1][10]: mass, first moment (3) and inertia about the base
                           origin (xx xy xz yy yz zz) of the massive bodies of level L, base coordinates */
  int offChainJoint;    /* int[chainNc]: jointIndex of the column */
  int pad[2];
} RcsbFkPlanHeader;

/*
 * Kinetic-terms plan. "Links" are the bodies that own at least one unconstrained joint;
 * every massive body is lumped into its nearest ancestor-or-self link, so that subtree
 * (composite) quantities only have to be kept per link.
 *
 * The computation runs as two kernels joined by a per-state record in global memory:
 *   ktWalkKernel      one thread per state, depth-first walk; the mass properties of the
 *                     bodies lumped into the current link are summed in one accumulator, which
 *                     is emitted as a "run" whenever the walk moves on to another link
 *   ktAssembleKernel  a group of lanes per state; sums runs into links and links into their
 *                     ancestors (composites), then builds M, h, F_gravity and E
 * Record of one state (recDoubles doubles, offsets in the order the walk produces them):
 *   column k   colRec[k] + {0..2 joint axis, 3..5 joint origin, 6 joint rate (IK space; plans with
 *                           velocities only)}
 *   run r      runRec[r] + {0 m, 1..3 m c, 4..12 inertia about the world origin (row-major,
 *                           not symmetrised: the reference uses the tensor of the file as
 *                           given), 13..15 bias force, 16..18 bias moment about the world
 *                           origin (plans with velocities only)} of the bodies of the run,
 *                           belonging to link runLink[r]
 *   tail at recV: potential energy, then m and m c (3) of the massive bodies that no
 *   unconstrained joint moves (they count for the centre of gravity only)
 */
/* doubles per column / run record: with velocities the joint rate and the bias wrench ride along */
#define RCSB_KT_COL_REC_N(vel) ((vel) ? 7 : 6)
#define RCSB_KT_LINK_REC_N(vel) ((vel) ? 19 : 13)
/* assembly kernel, per column in shared memory: S = {w[3], v[3], rate} (spatial axis of the joint
 * about the world origin: rotation w = a, v = o x a; translation w = 0, v = a), stride 7;
 * F = {p[3], L[3]} (+ L with the transposed inertia for unsymmetric models), stride 7 / 9 */
#define RCSB_KT_S_STRIDE 7
#define RCSB_KT_F_STRIDE(sym) ((sym) ? 7 : 9)

typedef struct
{
  int nLinks;
  int nJ;
  int totalBytes;
  int recDoubles;       /* doubles per state record (multiple of 4) */
  int nRuns;
  int recV;             /* offset of the 5-double tail: V, m_free, (m c)_free */
  int nPairs;           /* structurally non-zero entries of the upper triangle of M */
  int symmetric;        /* 1: every inertia tensor of the model is symmetric (M = M^T) */
  int offBodyLink;      /* int[nb]: link the body is lumped into, -1 = none (no joint moves it) */
  int offBodyEmit;      /* int[nb + 1]: 1 if the accumulator is emitted (and cleared) before the
                           body is visited ([nb]: after the last body) */
  int offLinkParent;    /* int[nLinks]: nearest proper ancestor link, -1 = none */
  int offColJoint;      /* int[nJ]: jointIndex of Jacobian column */
  int offColLink;       /* int[nJ]: link of the column's joint */
  int offColIsRot;      /* int[nJ] */
  int offColRec;        /* int[nJ] */
  int offRunLink;       /* int[nRuns] */
  int offRunRec;        /* int[nRuns] */
  /* assembly kernel: per-state shared-memory layout (doubles), see rcsbKtAsmLayout():
   *   [0, region0)   the record as the walk wrote it; the first run of a link becomes the
   *                  link's composite in place. With matrix staging (direct dynamics: the
   *                  Cholesky solve works on it) region0 also takes the nJ x nJ matrix once the
   *                  record is consumed; without staging M goes straight to global memory.
   *   S + 7 k        column k: axis, origin, rate (copied out of the record)
   *   F + 9 k        column k: p, L, L with the transposed inertia
   *   tail           copy of the record's 5-double tail (8 slots)
   *   rhs            nJ doubles: right-hand side / solution of the direct dynamics */
  int nExtraRuns;       /* runs that are not the first of their link */
  int offLinkHome;      /* int[nLinks]: offset of the link's composite (= its first run) */
  int offExtraRuns;     /* int2[nExtraRuns]: {offset of the run, offset of the link's composite} */
  int offColHome;       /* int[nJ]: composite of the column's link */
  int offPairs;         /* int4[nPairs], structurally non-zero upper-triangle entries (ancestor
                           column i, column k): {S offset of i, F offset of k,
                           (i nJ + k) | (k nJ + i) << 16, (i == k) << 1 | (S offset of k) << 8} */
  int pad0;
  int walkBytes;        /* leading part of the plan the walk kernel needs (header, bodyLink,
                           bodyEmit, bodyOut, bodyFrameRec); multiple of 16 */
  /* whole-body plans (rcsb_wholebody): the walk also writes the frames of the selected bodies
   * (A_BI, like fkKernel) and puts the frames of the bodies the task Jacobians refer to into the
   * record; the assembly kernel builds the task rows from them and the column records. */
  int wholeBody;        /* 1: the tables below are present */
  int nSel;             /* frames written per state */
  int nTaskBodies;      /* bodies referred to by the tasks (slots u) */
  int nJac;             /* task Jacobians (3 rows each) */
  int offBodyOut;       /* int[nb]: A_BI output slot of the body, -1 = not selected        (walk part) */
  int offBodyFrameRec;  /* int[nb]: 1 if the walk emits the body's frame into the record  (walk part) */
  int offTaskBodyRec;   /* int[nTaskBodies]: record offset of the frame of task body u */
  int offJacTasks;      /* int[nJac][8]: {kind, mode, u2, u1, u3, uA, 0, 0}, see rcsb_taskjac.cu */
  int offTaskColOn;     /* unsigned char[nTaskBodies * nJ]: 1 if column k moves task body u */
  int vel;              /* 1: records carry joint rates and bias wrenches (RCSB_KT_*_REC_N) */
  int nRoots;           /* links without a parent link */
  int offRootLinks;     /* int[nRoots]: record offset of their composites */
  int nJacEntries;      /* structurally non-zero (task Jacobian, column) pairs */
  int offJacEntries;    /* int[nJacEntries]: column | task << 8 | on2 << 16 | on1 << 17 | on3 << 18
                           (onX: the column moves the task's effector / refBdy / refFrame body) */
  int offLinkSteps;     /* int2[nLinks], links in descending (reverse depth-first) order: {offset of
                           the link's composite, where it goes next: -1 = into the next entry (its parent
                           is the preceding link: the sum stays in a register), -2 = nowhere (no parent
                           link), else the offset of the parent's composite} */
  int pad1[1];
} RcsbKtPlanHeader;

typedef struct
{
  int S, F, tail, rhs, doubles;
} RcsbKtAsmLayout;

#ifdef __CUDACC__
#define RCSB_HD __host__ __device__
#else
#define RCSB_HD
#endif
static inline RCSB_HD RcsbKtAsmLayout rcsbKtAsmLayout(int recDoubles, int nJ, int stageM, int symmetric)
{
  RcsbKtAsmLayout l;
  int region0 = (stageM && nJ * nJ > recDoubles) ? nJ * nJ : recDoubles;
  region0 = (region0 + 3) & ~3;
  l.S = region0;
  l.F = l.S + ((RCSB_KT_S_STRIDE * nJ + 3) & ~3);
  l.tail = l.F + ((RCSB_KT_F_STRIDE(symmetric) * nJ + 3) & ~3);
  l.rhs = l.tail + 8;
  l.doubles = l.rhs + nJ;
  return l;
}

/*
 * IK plan: tasks (all active), their effectors' joint chains and the per-column constants
 * of the resolved-motion-rate step.
 */
typedef struct
{
  int nTasks;
  int nx;               /* 3 per task */
  int nJ;
  int totalBytes;
  int offTaskKind;      /* int[nTasks]: RCSB_TASK_XYZ / RCSB_TASK_ABC */
  int offTaskBody;      /* int[nTasks] */
  int offBodyTaskStart; /* int[nb + 1]: CSR over tasks attached to each body */
  int offBodyTaskList;  /* int[nTasks] */
  int offJntNeeded;     /* int[dof]: 1 if the joint lies on some task's chain */
  int offChainStart;    /* int[nTasks + 1]: CSR over (column | isRot << 16) per task */
  int offChain;         /* int[sum of chain lengths] */
  int offColJoint;      /* int[nJ] jointIndex of each Jacobian column */
  int offInvWq;         /* double[nJ]: weightMetric (q_max - q_min)   (Rcs_graph.c:1087) */
  int offJlGain;        /* double[nJ]: weightJL / (q_max - q_min)^2, 0 if the range is not
                           positive (Rcs_kinematics.c:1496) */
  int offQ0;            /* double[nJ] */
  /*
   * Single-chain fast path (rcsb_ik.cu, ikChainKernel): all tasks sit on one effector whose
   * backward joint chain has chainNc <= 32 unconstrained joints (<= 8: ikChainKernel, registers;
   * 9..32: ikChainLoopKernel, shared memory). The walk from the root to the
   * effector is a straight-line program: per chain column c the steps progStart[c] ..
   * progStart[c + 1] (constant transforms and constrained joints) precede the column's own
   * joint motion; the steps progStart[chainNc] .. progStart[chainNc + 1] lead on to the
   * effector body. chainNc == 0: not eligible, the general kernel runs.
   */
  int nqr;              /* columns of the reduced problem: nJ minus the unconstrained slave joints */
  int offColRed;        /* int[nJ]: column of the coupling matrix A the joint belongs to (its own
                           reduced index, or its master's) */
  int offColCpl;        /* int[nJ]: coupling record of a slave column, -1 otherwise */
  int offChainXf;       /* double[nXf][12]: the chain's relative transforms, rows / columns
                           permuted for the frame representation of the chain kernels (below) */
  int chainNc;
  int nOther;           /* unconstrained joints that are not on the chain */
  int offProgStart;     /* int[chainNc + 2] */
  int offProg;          /* int4[nSteps]: {op, a, b, 0}; op 0: relative transform a of chainXf, b =
                           its RCSB_REL_* flags; op 1: joint of type a, jointIndex b, value from q.
                           The chain kernels keep the rows of the running rotation matrix cyclically
                           shifted so that the axis of the NEXT joint is row 2 and the rows it
                           mixes are rows 0 and 1: every joint then runs the same code (no
                           three-way selection by axis). The shifts are static, so they are folded
                           into the transforms on the host: rot' = P_out rot P_in^T, org' = P_in org
                           (exact permutations); a shift change with no transform to carry it
                           is op 2 (rows shifted by a, registers only); the last step restores
                           shift 0. A translational joint only reads its axis: the row that holds it
                           is op 1's fourth entry / chainType >> 4. */
  int offChainType;     /* int[chainNc]: RCSJOINT_TYPE of the column's joint | axis row << 4 */
  int offChainJoint;    /* int[chainNc]: jointIndex */
  int offChainW;        /* double[chainNc]: invWq */
  int offChainGain;     /* double[chainNc]: joint-limit gain */
  int offChainQ0;       /* double[chainNc] */
  int offOtherJoint;    /* int[nOther] jointIndex */
  int offOtherW;        /* double[nOther] */
  int offOtherGain;     /* double[nOther] */
  int offOtherQ0;       /* double[nOther] */
  /*
   * Several effectors without coupled joints (ikMultiKernel): the union of the tasks' chain
   * joints gets multiNu slots (ascending Jacobian column) described by the chain* arrays above
   * (type, jointIndex, invWq, joint-limit gain, q0); the other* arrays list the remaining
   * unconstrained joints. multiNu == 0: not eligible.
   */
  int multiNu;
  int nEffs;            /* distinct effector bodies */
  int offJntSlot;       /* int[dof]: slot of the joint, -1 = not on any task's chain */
  int offSlotTaskMask;  /* int[multiNu]: bit t set if the slot lies on the chain of task t */
  int offTaskEff;       /* int[nTasks]: effector index of the task */
  int offBodyEffIdx;    /* int[nb]: effector index of the body, -1 = none */
  /* pruned walk of ikMultiKernel: only the bodies above some effector, depth first */
  int nWalk;
  int walkSlots;        /* frame-stack slots the pruned walk needs */
  int offWalk;          /* int4[nWalk]: {body, load (RCSB_LOAD_* or slot), save slot or -1, 0} */
  int offSegXf;         /* double[chainNc + 1][12], 0 = none: the steps of every program segment folded
                           into ONE relative transform (row shifts included as exact permutations,
                           empty segments as the identity). Present when the program has no joint steps
                           (op 1) and every column is a rotation: the specialised ikChainKernel instances
                           walk the chain without interpreting the program. */
} RcsbIkPlanHeader;

/*
 * Plan of the general task-stack solver (rcsb_ik_tasks_rmr, ikTasksKernel): any mix of the task kinds
 * of rcsb.h with reference bodies / frames. Bodies the tasks refer to get frame slots; slot 0 is
 * reserved for "no body" (world: identity frame, no columns).
 */
typedef struct
{
  int nTasks;
  int nx;               /* rows of the stacked Jacobian */
  int nJ;
  int totalBytes;
  int nSlots;           /* frame slots incl. slot 0 */
  int cogRoot;          /* body whose subtree the COG tasks refer to, -1 = the whole graph, -2 = no COG task */
  int offTask;          /* int[nTasks][8]: {kind, row, dim, slotE, slotR, slotF, aux (joint column / COG
                           component), jointIndex} */
  int offBodySlotStart; /* int[nb + 1]: CSR over the frame slots of each body */
  int offBodySlotList;  /* int[number of slots - 1] */
  int offSlotMask;      /* unsigned long long[nSlots]: bit k set if column k moves the slot's body */
  int offBodyInCog;     /* unsigned char[nb]: 1 if the body counts for the COG tasks */
  int offColJoint;      /* int[nJ] */
  int offInvWq;         /* double[nJ] */
  int offJlGain;        /* double[nJ] */
  int offQ0;            /* double[nJ] */
  int pad;
} RcsbIkTasksPlanHeader;

#ifdef __cplusplus
}
#endif

#endif
