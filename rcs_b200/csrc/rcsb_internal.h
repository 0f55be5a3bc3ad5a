/*
 * Internal declarations shared by the host C sources of rcs_b200 (not installed).
 */
#ifndef RCSB_INTERNAL_H
#define RCSB_INTERNAL_H

#include "rcsb_graph.h"

#include <stdarg.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- errors / logging ------------------------------------------------------------- */
extern int RcsLogLevel;                       /* like the reference's global, default 0 */
void rcsb_set_error(const char* fmt, ...);    /* records the message for rcsb_last_error() */
void rcsb_fatal(const char* file, int line, const char* fmt, ...);   /* RFATAL analogue */
const char* rcsb_error_string(void);

#define RCSB_FATAL(...) rcsb_fatal(__FILE__, __LINE__, __VA_ARGS__)
#define RCSB_LOG(level, ...)                                              \
  do                                                                      \
  {                                                                       \
    if ((level) <= RcsLogLevel)                                           \
    {                                                                     \
      fprintf(stderr, "[rcsb %s:%d] ", __FILE__, __LINE__);               \
      fprintf(stderr, __VA_ARGS__);                                       \
      fprintf(stderr, "\n");                                              \
    }                                                                     \
  } while (0)

/* ---- small host math used at load time only --------------------------------------- */
void HTr_setIdentity(HTr* A);
bool HTr_isIdentity(const HTr* A);
HTr* HTr_clone(const HTr* A);
/* A_2I <- A_21 * A_1I (in place on A_2I, which enters as A_1I) */
void HTr_transformSelf(HTr* A_2I, const HTr* A_21);
void Mat3d_fromEulerAngles(double A_KI[3][3], const double ea[3]);
void Mat3d_toEulerAngles(double ea[3], double A_KI[3][3]);
void Mat3d_mul(double A_3I[3][3], double A_32[3][3], double A_2I[3][3]);
void Mat3d_similarityTransform(double B_I[3][3], double A_CB[3][3], double C_I[3][3]);
void Quat_toRotationMatrix(double A_BI[3][3], const double q[4]);
void HTr_from2Points(HTr* A_KI, const double p1[3], const double p2[3]);

/* ---- element-only XML DOM (expat) ------------------------------------------------- */
typedef struct RcsbXmlAttr
{
  char* name;
  char* value;
  struct RcsbXmlAttr* next;
} RcsbXmlAttr;

typedef struct RcsbXmlNode
{
  char* name;
  RcsbXmlAttr* attrs;
  struct RcsbXmlNode* children;
  struct RcsbXmlNode* last;
  struct RcsbXmlNode* parent;
  struct RcsbXmlNode* next;
} RcsbXmlNode;

RcsbXmlNode* rcsb_xml_parse_file(const char* fileName);
RcsbXmlNode* rcsb_xml_parse_buffer(const char* buf, size_t size);
void rcsb_xml_free(RcsbXmlNode* root);
const char* rcsb_xml_attr(const RcsbXmlNode* n, const char* name);   /* NULL if absent */

/* attribute helpers with the reference's parsing rules (src/RcsCore/Rcs_parser.c) */
unsigned int rcsb_xml_num_strings(const RcsbXmlNode* n, const char* tag);
bool rcsb_xml_vecn(const RcsbXmlNode* n, const char* tag, double* x, unsigned int cnt);
bool rcsb_xml_bool(const RcsbXmlNode* n, const char* tag, bool* x);
bool rcsb_xml_htr(const RcsbXmlNode* n, const char* tag, HTr* A);
unsigned int rcsb_xml_string(const RcsbXmlNode* n, const char* tag, char* str, unsigned int cap);

/* ---- graph internals --------------------------------------------------------------- */
bool Rcs_getAbsoluteFileName(const char* fileName, char* absFileName);   /* 512 bytes */
RcsGraph* RcsGraph_createFromXmlNode(const RcsbXmlNode* node);
/* rcsb_urdf.c: RcsGraph_rootBodyFromURDFFile (src/RcsCore/Rcs_URDFParser.c:924) */
RcsBody* rcsb_urdf_root_body(const char* filename, const char* suffix, const HTr* A_BP, unsigned int* dof);
void RcsBody_computeInertiaTensor(const RcsBody* self, HTr* I);
double RcsJoint_computeSlaveJointAngle(const RcsJoint* slave, double q_master);
double RcsJoint_computeSlaveJointVelocity(const RcsJoint* slave, double q_master,
                                          double q_dot_master);
bool RcsJoint_isRotation(const RcsJoint* j);

#define RCSGRAPH_TRAVERSE_BODIES(graph) \
  for (RcsBody* BODY = (graph)->root; BODY; BODY = RcsBody_depthFirstTraversalGetNext(BODY))

#define RCSBODY_TRAVERSE_JOINTS(body) for (RcsJoint* JNT = (body)->jnt; JNT; JNT = JNT->next)

#ifdef __cplusplus
}
#endif

#endif
