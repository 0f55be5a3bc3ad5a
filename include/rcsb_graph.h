/*
 * rcsb_graph.h — host-side data model and graph loader of rcs_b200 (plain C).
 *
 * This is the API surface the reference exposes for the hot path: a kinematic tree of
 * RcsBody / RcsJoint nodes owned by an RcsGraph, loaded from the reference's XML graph
 * files. Struct and field names follow the reference so that code written against
 * HRI-EU/Rcs reads the same here:
 *   HTr       src/RcsCore/Rcs_HTr.h:62-66      {org[3], rot[3][3]}, rows of rot = frame axes
 *   MatNd     src/RcsCore/Rcs_MatNd.h:92-99    row-major, caller-allocated, callee reshapes
 *   RcsJoint  src/RcsCore/Rcs_typedef.h:66-94
 *   RcsShape  src/RcsCore/Rcs_typedef.h:133-149
 *   RcsBody   src/RcsCore/Rcs_typedef.h:187-214
 *   RcsGraph  src/RcsCore/Rcs_typedef.h:247-259
 * The structs are binary compatible with the reference's (same fields, same order, same types;
 * tests/test_abi_cpu.py compares sizes and offsets with the reference's own headers, and
 * tests/test_abi_gpu.py hands a graph created by libRcsCore's RcsGraph_create to
 * rcsb_model_create): a graph built by either library can be read by the other. A graph is
 * destroyed by the library that created it.
 *
 * Only the loader and bookkeeping run on the host. Everything that evaluates a state
 * (forward kinematics, Jacobians, kinetic terms, the IK step) runs in the CUDA kernels
 * behind rcsb.h; the single-state RcsGraph_* functions at the bottom of this header are
 * N=1 calls into that same path and fail loudly without a GPU. There is no CPU fallback.
 */
#ifndef RCSB_GRAPH_H
#define RCSB_GRAPH_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCS_GRAVITY (9.81)

typedef struct
{
  double org[3];
  double rot[3][3];
} HTr;

typedef struct _MatNd
{
  double* ele;
  unsigned int m;
  unsigned int n;
  unsigned int size;
  bool stackMem;
} MatNd;

/* src/RcsCore/Rcs_typedef.h:48-57 */
typedef enum
{
  RCSJOINT_ROT_X = 0,
  RCSJOINT_ROT_Y = 1,
  RCSJOINT_ROT_Z = 2,
  RCSJOINT_TRANS_X = 3,
  RCSJOINT_TRANS_Y = 4,
  RCSJOINT_TRANS_Z = 5
} RCSJOINT_TYPE;

typedef enum
{
  RCSJOINT_CTRL_POSITION = 0,
  RCSJOINT_CTRL_VELOCITY,
  RCSJOINT_CTRL_TORQUE
} RCSJOINT_CTRL_TYPE;

/* src/RcsCore/Rcs_typedef.h:98-116 */
typedef enum
{
  RCSSHAPE_NONE = 0,
  RCSSHAPE_SSL = 1,
  RCSSHAPE_SSR = 2,
  RCSSHAPE_MESH = 3,
  RCSSHAPE_BOX = 4,
  RCSSHAPE_CYLINDER = 5,
  RCSSHAPE_REFFRAME = 6,
  RCSSHAPE_SPHERE = 7,
  RCSSHAPE_CONE = 8,
  RCSSHAPE_GPISF = 9,
  RCSSHAPE_TORUS = 10,
  RCSSHAPE_OCTREE = 11,
  RCSSHAPE_POINT = 12,
  RCSSHAPE_MARKER = 13
} RCSSHAPE_TYPE;

typedef enum
{
  RCSBODY_PHYSICS_NONE = 0,
  RCSBODY_PHYSICS_KINEMATIC = 1,
  RCSBODY_PHYSICS_DYNAMIC = 2,
  RCSBODY_PHYSICS_FIXED = 3
} RCSBODY_PHYSICS_SIMULATION_TYPE;

typedef enum
{
  RcsStateFull = 0,
  RcsStateIK
} RcsStateType;

typedef struct _RcsJoint RcsJoint;
typedef struct _RcsShape RcsShape;
typedef struct _RcsBody RcsBody;
typedef struct _RcsGraph RcsGraph;

struct _RcsJoint
{
  char* name;
  double q0;
  double q_init;
  double q_min;
  double q_max;
  double weightJL;
  double weightCA;
  double weightMetric;
  bool constrained;
  int type;          /* RCSJOINT_TYPE */
  int dirIdx;        /* 0,1,2 for x,y,z */
  int jointIndex;    /* index into q (dof) */
  int jacobiIndex;   /* Jacobian column, -1 if constrained */
  HTr* A_JP;         /* NULL = identity */
  HTr A_JI;          /* world frame after the joint's motion (filled by RcsGraph_setState) */
  double maxTorque;
  double speedLimit;
  double accLimit;
  double decLimit;
  int ctrlType;
  char* coupledJointName;
  MatNd* couplingFactors;
  RcsJoint* prev;    /* previous joint on the chain towards the root */
  RcsJoint* next;    /* next joint of the same body */
  RcsJoint* coupledTo;
};

struct _RcsShape
{
  int type;           /* RCSSHAPE_TYPE */
  HTr A_CB;
  double extents[3];
  double scale;
  bool resizeable;
  char computeType;
  char* meshFile;     /* the four strings are carried (clone, destroy), never interpreted here */
  char* textureFile;
  char* color;
  char* material;
  void* userData;
};

struct _RcsBody
{
  double m;
  bool rigid_body_joints;
  int physicsSim;
  double x_dot[3];
  double omega[3];
  double confidence;
  char* name;
  char* xmlName;
  char* suffix;
  HTr* A_BP;          /* NULL = identity */
  HTr* A_BI;          /* world frame (filled by RcsGraph_setState) */
  HTr* Inertia;       /* org = COM in body frame, rot = inertia about COM, body frame */
  RcsBody* parent;
  RcsBody* firstChild;
  RcsBody* lastChild;
  RcsBody* next;
  RcsBody* prev;
  RcsShape** shape;   /* NULL-terminated */
  RcsJoint* jnt;      /* first joint of the body */
  void* extraInfo;    /* generic bodies: the body they are linked to */
};

typedef struct _RcsSensor RcsSensor;   /* src/RcsCore/Rcs_typedef.h:233-243; not on the hot path */

struct _RcsGraph
{
  RcsBody* root;
  RcsBody gBody[10];  /* generic bodies (placeholders the reference's tasks can be re-linked to);
                         kept for layout compatibility, this loader leaves them unlinked */
  unsigned int dof;
  unsigned int nJ;
  MatNd* q;
  MatNd* q_dot;
  char* xmlFile;
  RcsSensor* sensor;  /* sensors are not loaded here (always NULL in graphs of this loader) */
  void* userData;     /* the reference leaves it to extensions; the single-state wrappers of rcsb.h
                         keep their device model cache in a side table keyed by the graph instead */
};

/* ---- MatNd (src/RcsCore/Rcs_MatNd.h): the few calls the path needs ------------------- */
MatNd* MatNd_create(unsigned int m, unsigned int n);
void MatNd_destroy(MatNd* self);
void MatNd_reshape(MatNd* self, unsigned int m, unsigned int n);
void MatNd_reshapeAndSetZero(MatNd* self, unsigned int m, unsigned int n);

/* ---- Graph construction / loading --------------------------------------------------- */

/* src/RcsCore/Rcs_graph.c:398 — file name is searched as given, then relative to every
 * path added with Rcs_addResourcePath(). Returns NULL if the file cannot be loaded. */
RcsGraph* RcsGraph_create(const char* xmlFile);
/* src/RcsCore/Rcs_URDFParser.c:1152 — a URDF <robot> file (revolute, continuous, prismatic and fixed
 * joints, any axis, mimic joints, <model_state>; box / cylinder / sphere shapes, mesh names only).
 * URDF sub-trees inside a <Graph> file are loaded through its <URDF file= prev= suffix= transform=
 * rigid_body_joints=> nodes (src/RcsCore/Rcs_graphParser.c:1438). */
RcsGraph* RcsGraph_fromURDFFile(const char* urdfFile);
/* src/RcsCore/Rcs_graph.c:366 (buffer variant of RcsGraph_create) */
RcsGraph* RcsGraph_createFromBuffer(const char* buffer, unsigned int size);
void RcsGraph_destroy(RcsGraph* self);
/* src/RcsCore/Rcs_graph.c:2193 */
RcsGraph* RcsGraph_clone(const RcsGraph* src);
/* src/RcsCore/Rcs_graph.c:1887 — structural validation; returns true if no errors. */
bool RcsGraph_check(const RcsGraph* self, int* nErrors, int* nWarnings);
/* src/RcsCore/Rcs_resourcePath.h:77 */
bool Rcs_addResourcePath(const char* path);
void Rcs_clearResourcePath(void);

/* src/RcsCore/Rcs_graph.c:2535, :2604, :2999 — programmatic construction. */
void RcsGraph_insertBody(RcsGraph* graph, RcsBody* parent, RcsBody* body);
void RcsGraph_insertJoint(RcsGraph* graph, RcsBody* body, RcsJoint* jnt);
void RcsGraph_makeJointsConsistent(RcsGraph* self);

/* Traversal: depth-first body order (src/RcsCore/Rcs_body.c:573) defines body indices;
 * joints in body order / in-body order define jointIndex (Rcs_graph.c:2932). */
RcsBody* RcsBody_depthFirstTraversalGetNext(const RcsBody* body);
RcsJoint* RcsBody_lastJointBeforeBody(const RcsBody* body);   /* Rcs_body.c:1706 */
RcsBody* RcsGraph_getBodyByName(const RcsGraph* self, const char* name);
RcsJoint* RcsGraph_getJointByName(const RcsGraph* self, const char* name);
unsigned int RcsGraph_numBodies(const RcsGraph* self);
RcsBody* RcsGraph_getBodyByIndex(const RcsGraph* self, unsigned int dfsIndex);
int RcsGraph_bodyIndex(const RcsGraph* self, const RcsBody* body);
RcsJoint* RcsGraph_getJointByIndex(const RcsGraph* self, unsigned int jointIndex);
double RcsGraph_mass(const RcsGraph* self);

/* src/RcsCore/Rcs_graph.c:628, :670 */
void RcsGraph_stateVectorFromIK(const RcsGraph* self, const MatNd* q_ik, MatNd* q_full);
void RcsGraph_stateVectorToIK(const RcsGraph* self, const MatNd* q_full, MatNd* q_ik);
/* src/RcsCore/Rcs_graph.c:1087, :2804 */
void RcsGraph_getInvWq(const RcsGraph* self, MatNd* invWq, RcsStateType type);
int RcsGraph_countCoupledJoints(const RcsGraph* self);
/* src/RcsCore/Rcs_graph.c:2756, :2823; Rcs_kinematics.c:1496, :47 (host-side helpers of the
 * single-state API: O(dof) loops over the joint list, nothing to batch) */
bool RcsGraph_updateSlaveJoints(const RcsGraph* self, MatNd* q, MatNd* q_dot);
int RcsGraph_coupledJointMatrix(const RcsGraph* self, MatNd* A, MatNd* invA);
double RcsGraph_jointLimitGradient(const RcsGraph* self, MatNd* dH, RcsStateType type);
void RcsGraph_bodyPoint(const RcsBody* body, const double B_pt[3], double I_pt[3]);
void RcsGraph_getDefaultState(const RcsGraph* self, MatNd* q0);

/* Canonical JSON description of everything the hot path reads from the graph (tree
 * links, joint and body parameters, indices). Caller frees with free(). Used by the
 * parity tests to compare integer layout bit for bit with the reference loader. */
char* RcsGraph_layoutJSON(const RcsGraph* self);

/* Writes the graph as a self-contained <Graph> XML file (explicit mass / inertia /
 * cogVector, shapes dropped) that RcsGraph_create() loads back to the same layout.
 * Counterpart of RcsGraph_fprintXML, src/RcsCore/Rcs_graph.c:3072. Returns 0 on success. */
int RcsGraph_writeXML(const RcsGraph* self, const char* fileName);

/* ---- Error behaviour ------------------------------------------------------------------
 * The reference terminates the process on contract violations (RFATAL / RCHECK,
 * src/RcsCore/Rcs_macros.h:141-151, :312-355). Mode 1 (default) does the same: message to
 * stderr, then abort(). Mode 0 records the message (rcsb_last_error() in rcsb.h) and
 * returns an error value from the offending call instead. */
void Rcs_setFatalMode(int abortOnFatal);

#ifdef __cplusplus
}
#endif

#endif
