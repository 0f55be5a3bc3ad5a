/*
 * URDF input: <robot> files -> RcsGraph, and <URDF file=...> nodes inside a <Graph>.
 * Counterpart of src/RcsCore/Rcs_URDFParser.c (RcsGraph_fromURDFFile :1152,
 * RcsGraph_rootBodyFromURDFFile :924, parseBodyURDF :274, parseJointURDF :457, connectURDF :612,
 * parseShapeURDF :65, RcsGraph_parseModelState :885) and of the <URDF> branch of the graph parser
 * (src/RcsCore/Rcs_graphParser.c:1438-1600). The graph that comes out has the reference's body and
 * joint order, relative transforms, limits, couplings and mass properties, so everything after the
 * loader (flattening, kernels) is shared with the native format.
 *
 * origin rpy: URDF gives fixed-axis roll / pitch / yaw; the reference converts them to its own
 * relative x-y-z Euler angles with Ken Shoemake's converter (Graphics Gems IV, "Euler Angle
 * Conversion", vendored as external/EulerAngles.c: Eul_ToHMatrix with order XYZs, Eul_FromHMatrix
 * with order XYZr) and builds the rotation from those. The detour through the angles is kept because
 * it is not exact at gimbal lock (pitch = +-90 degrees is common in robot descriptions): below
 * cos(pitch) = 16 FLT_EPSILON the converter drops one angle, and the frames of the reference carry
 * that.
 *
 * Not loaded: mesh geometry (the file name is kept; a link without <inertial> gets its default
 * inertia from its box / cylinder / sphere shapes only), textures.
 */
#include "rcsb_internal.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

static char* cloneString(const char* s)
{
  char* c = (char*) malloc(strlen(s) + 1);
  strcpy(c, s);
  return c;
}

static char* cloneWithSuffix(const char* s, const char* suffix)
{
  const size_t n = strlen(s) + (suffix ? strlen(suffix) : 0) + 1;
  char* c = (char*) malloc(n);
  strcpy(c, s);
  if (suffix)
  {
    strcat(c, suffix);
  }
  return c;
}

static bool tagIs(const RcsbXmlNode* n, const char* name)
{
  return n && strcasecmp(n->name, name) == 0;
}

static const RcsbXmlNode* childByName(const RcsbXmlNode* n, const char* name)
{
  for (const RcsbXmlNode* c = n->children; c; c = c->next)
  {
    if (strcasecmp(c->name, name) == 0)
    {
      return c;
    }
  }
  return NULL;
}

static HTr* newHTr(void)
{
  HTr* A = (HTr*) calloc(1, sizeof(HTr));
  HTr_setIdentity(A);
  return A;
}

/* fixed-axis x-y-z angles -> the reference's relative x-y-z angles -> rotation (see the header) */
static void originToHTr(HTr* A, const double rpy[3], const double xyz[3])
{
  /* Eul_ToHMatrix, order XYZs: (i, j, k) = (x, y, z), even parity, static frame */
  const double ci = cos(rpy[0]), cj = cos(rpy[1]), ch = cos(rpy[2]);
  const double si = sin(rpy[0]), sj = sin(rpy[1]), sh = sin(rpy[2]);
  const double cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
  double M[3][3];
  M[0][0] = cj * ch;
  M[0][1] = sj * sc - cs;
  M[0][2] = sj * cc + ss;
  M[1][0] = cj * sh;
  M[1][1] = sj * ss + cc;
  M[1][2] = sj * cs - sc;
  M[2][0] = -sj;
  M[2][1] = cj * si;
  M[2][2] = cj * ci;

  /* Eul_FromHMatrix, order XYZr: (i, j, k) = (z, y, x), odd parity, rotating frame */
  const double cy = sqrt(M[2][2] * M[2][2] + M[1][2] * M[1][2]);
  double ex, ey, ez;
  if (cy > 16.0 * FLT_EPSILON)
  {
    ex = atan2(M[0][1], M[0][0]);
    ey = atan2(-M[0][2], cy);
    ez = atan2(M[1][2], M[2][2]);
  }
  else
  {
    ex = atan2(-M[1][0], M[1][1]);
    ey = atan2(-M[0][2], cy);
    ez = 0.0;
  }
  /* odd parity: all signs flip; rotating frame: first and last angle swap */
  const double ea[3] = {-ez, -ey, -ex};

  A->org[0] = xyz[0];
  A->org[1] = xyz[1];
  A->org[2] = xyz[2];
  Mat3d_fromEulerAngles(A->rot, ea);
}

static void readOrigin(const RcsbXmlNode* parent, HTr* A)
{
  double xyz[3] = {0.0, 0.0, 0.0}, rpy[3] = {0.0, 0.0, 0.0};
  const RcsbXmlNode* o = childByName(parent, "origin");
  if (o)
  {
    rcsb_xml_vecn(o, "xyz", xyz, 3);
    rcsb_xml_vecn(o, "rpy", rpy, 3);
  }
  originToHTr(A, rpy, xyz);
}

/* <visual> / <collision> (Rcs_URDFParser.c:65-272) */
static RcsShape* shapeFromURDF(const RcsbXmlNode* node)
{
  RcsShape* shape = (RcsShape*) calloc(1, sizeof(RcsShape));
  readOrigin(node, &shape->A_CB);
  shape->scale = 1.0;

  const RcsbXmlNode* geometry = childByName(node, "geometry");
  if (!geometry)
  {
    RCSB_FATAL("URDF: <%s> without <geometry>", node->name);
    free(shape);
    return NULL;
  }
  const RcsbXmlNode* g;
  if ((g = childByName(geometry, "box")))
  {
    rcsb_xml_vecn(g, "size", shape->extents, 3);
    shape->type = RCSSHAPE_BOX;
  }
  else if ((g = childByName(geometry, "cylinder")))
  {
    rcsb_xml_vecn(g, "radius", &shape->extents[0], 1);
    rcsb_xml_vecn(g, "length", &shape->extents[2], 1);
    shape->type = RCSSHAPE_CYLINDER;
  }
  else if ((g = childByName(geometry, "sphere")))
  {
    rcsb_xml_vecn(g, "radius", &shape->extents[0], 1);
    shape->type = RCSSHAPE_SPHERE;
  }
  else if ((g = childByName(geometry, "mesh")))
  {
    shape->type = RCSSHAPE_MESH;
    char meshFile[256] = "", full[512] = "";
    rcsb_xml_string(g, "filename", meshFile, 256);
    if (strncmp(meshFile, "package://", 10) == 0)
    {
      if (!Rcs_getAbsoluteFileName(meshFile + 10, full))
      {
        snprintf(full, sizeof(full), "%s", meshFile + 10);
      }
    }
    else
    {
      Rcs_getAbsoluteFileName(meshFile, full);
    }
    shape->meshFile = cloneString(full);
    double scale3[3] = {1.0, 1.0, 1.0};
    rcsb_xml_vecn(g, "scale", scale3, 3);
    if (!(scale3[0] == scale3[1] && scale3[0] == scale3[2]))
    {
      RCSB_FATAL("URDF: only uniform mesh scaling is supported");
    }
    shape->scale = scale3[0];
  }
  else
  {
    RCSB_LOG(1, "URDF: unsupported shape type below <geometry>");
  }

  const RcsbXmlNode* material = childByName(node, "material");
  const RcsbXmlNode* color = material ? childByName(material, "color") : NULL;
  if (color)
  {
    double rgba[4] = {1.0, 1.0, 1.0, 1.0};
    rcsb_xml_vecn(color, "rgba", rgba, 4);
    char buf[10];
    snprintf(buf, sizeof(buf), "#%02x%02x%02x%02x", (int) (rgba[0] * 255), (int) (rgba[1] * 255),
             (int) (rgba[2] * 255), (int) (rgba[3] * 255));
    shape->color = cloneString(buf);
  }
  else
  {
    shape->color = cloneString("DEFAULT");
  }
  shape->material = cloneString("default");
  /* RCSSHAPE_COMPUTE_GRAPHICS = 4 for <visual>, DISTANCE | PHYSICS = 1 | 2 for <collision>
   * (src/RcsCore/Rcs_typedef.h:120-126) */
  shape->computeType = tagIs(node, "visual") ? 4 : 3;
  return shape;
}

/* <link> (Rcs_URDFParser.c:274-417) */
static RcsBody* bodyFromURDF(const RcsbXmlNode* node)
{
  char name[256] = "";
  if (!tagIs(node, "link") || rcsb_xml_string(node, "name", name, 256) == 0)
  {
    return NULL;
  }
  if (strncmp(name, "GenericBody", 11) == 0)
  {
    RCSB_LOG(4, "The name \"GenericBody\" is reserved for internal use");
    return NULL;
  }
  RcsBody* body = (RcsBody*) calloc(1, sizeof(RcsBody));
  body->A_BI = newHTr();
  body->Inertia = newHTr();
  memset(body->Inertia->rot, 0, sizeof(body->Inertia->rot));
  body->xmlName = cloneString(name);
  body->name = cloneString(name);

  size_t nCollision = 0, nShapes = 0, nInertial = 0, k = 0;
  for (const RcsbXmlNode* c = node->children; c; c = c->next)
  {
    nCollision += strcmp(c->name, "collision") == 0;
    nShapes += strcmp(c->name, "collision") == 0 || strcmp(c->name, "visual") == 0;
  }
  body->shape = (RcsShape**) calloc(nShapes + 1, sizeof(RcsShape*));

  for (const RcsbXmlNode* c = node->children; c; c = c->next)
  {
    if (tagIs(c, "inertial"))
    {
      ++nInertial;
      const RcsbXmlNode* n;
      if ((n = childByName(c, "origin")))
      {
        rcsb_xml_vecn(n, "xyz", body->Inertia->org, 3);
      }
      if ((n = childByName(c, "mass")))
      {
        rcsb_xml_vecn(n, "value", &body->m, 1);
      }
      if ((n = childByName(c, "inertia")))
      {
        /* as the reference fills it: the lower triangle is MINUS the upper one */
        double (*I)[3] = body->Inertia->rot;
        rcsb_xml_vecn(n, "ixx", &I[0][0], 1);
        rcsb_xml_vecn(n, "ixy", &I[0][1], 1);
        rcsb_xml_vecn(n, "ixz", &I[0][2], 1);
        I[1][0] = -I[0][1];
        rcsb_xml_vecn(n, "iyy", &I[1][1], 1);
        rcsb_xml_vecn(n, "iyz", &I[1][2], 1);
        I[2][0] = -I[0][2];
        I[2][1] = -I[1][2];
        rcsb_xml_vecn(n, "izz", &I[2][2], 1);
      }
    }
    else if ((tagIs(c, "visual") || tagIs(c, "collision")) && k < nShapes)
    {
      RcsShape* s = shapeFromURDF(c);
      if (s)
      {
        body->shape[k++] = s;
      }
    }
  }
  if (nInertial == 0)
  {
    RcsBody_computeInertiaTensor(body, body->Inertia);
  }
  body->rigid_body_joints = false;
  if (body->m > 0.0)
  {
    body->physicsSim = nCollision == 0 ? RCSBODY_PHYSICS_NONE : RCSBODY_PHYSICS_DYNAMIC;
  }
  else
  {
    body->physicsSim = nCollision > 0 ? RCSBODY_PHYSICS_KINEMATIC : RCSBODY_PHYSICS_NONE;
  }
  return body;
}

/* <joint>, everything but the connectivity (Rcs_URDFParser.c:457-578) */
static RcsJoint* jointFromURDF(const RcsbXmlNode* node)
{
  char name[256] = "", type[256] = "";
  if (!tagIs(node, "joint"))
  {
    return NULL;
  }
  if (rcsb_xml_string(node, "name", name, 256) == 0)
  {
    RCSB_FATAL("Joint name not specified in URDF description");
    return NULL;
  }
  RcsJoint* jnt = (RcsJoint*) calloc(1, sizeof(RcsJoint));
  jnt->name = cloneString(name);
  jnt->weightJL = 1.0;
  jnt->weightCA = 1.0;
  jnt->weightMetric = 1.0;
  jnt->ctrlType = RCSJOINT_CTRL_POSITION;
  jnt->constrained = false;
  jnt->accLimit = DBL_MAX;
  jnt->decLimit = DBL_MAX;
  jnt->maxTorque = DBL_MAX;
  jnt->speedLimit = DBL_MAX;
  HTr_setIdentity(&jnt->A_JI);

  if (childByName(node, "origin"))
  {
    HTr A;
    readOrigin(node, &A);
    if (!HTr_isIdentity(&A))
    {
      jnt->A_JP = HTr_clone(&A);
    }
  }
  const RcsbXmlNode* limit = childByName(node, "limit");
  if (limit)
  {
    rcsb_xml_vecn(limit, "effort", &jnt->maxTorque, 1);
    rcsb_xml_vecn(limit, "lower", &jnt->q_min, 1);
    rcsb_xml_vecn(limit, "upper", &jnt->q_max, 1);
    rcsb_xml_vecn(limit, "velocity", &jnt->speedLimit, 1);
    jnt->accLimit = DBL_MAX;
  }
  if (rcsb_xml_string(node, "type", type, 256) == 0)
  {
    RCSB_FATAL("URDF joint \"%s\" has no type", name);
  }
  if (strcasecmp(type, "continuous") == 0)
  {
    jnt->q_min = -1.0;
    jnt->q_max = 1.0;
    jnt->constrained = true;
  }
  jnt->q0 = (jnt->q_max + jnt->q_min) / 2.0;
  jnt->q_init = jnt->q0;

  const RcsbXmlNode* mimic = childByName(node, "mimic");
  if (mimic)
  {
    char master[256] = "";
    if (rcsb_xml_string(mimic, "joint", master, 256) == 0)
    {
      RCSB_FATAL("Couldn't find tag \"joint\" in mimic joint %s", jnt->name);
    }
    jnt->coupledJointName = cloneString(master);
    jnt->constrained = true;   /* kinematically overwritten after the forward kinematics */
    jnt->couplingFactors = MatNd_create(1, 1);
    jnt->couplingFactors->ele[0] = 1.0;
    rcsb_xml_vecn(mimic, "multiplier", jnt->couplingFactors->ele, 1);
    /* the offset is parked in q_init until the master's q_init is known */
    jnt->q_init = 0.0;
    rcsb_xml_vecn(mimic, "offset", &jnt->q_init, 1);
  }
  return jnt;
}

static RcsBody* findBody(RcsBody** v, const char* name)
{
  for (; *v; ++v)
  {
    if (strcasecmp(name, (*v)->name) == 0)
    {
      return *v;
    }
  }
  return NULL;
}

static RcsJoint* findJoint(RcsJoint** v, const char* name)
{
  for (; *v; ++v)
  {
    if (strcasecmp(name, (*v)->name) == 0)
    {
      return *v;
    }
  }
  return NULL;
}

/* Vec3d_orthonormalVec (Rcs_Vec3d.c:893) and Mat3d_fromVec with dir = 2 (Rcs_Mat3d.c:1706) */
static void frameWithZ(double A[3][3], const double v[3])
{
  const double sqrLength = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (sqrLength == 0.0)
  {
    memset(A, 0, 9 * sizeof(double));
    A[0][0] = A[1][1] = A[2][2] = 1.0;
    return;
  }
  const double c = 1.0 / sqrt(sqrLength);
  for (int i = 0; i < 3; ++i)
  {
    A[2][i] = v[i] * c;
  }
  /* y: orthonormal to z */
  {
    const double* z = A[2];
    const double sq = z[0] * z[0] + z[1] * z[1] + z[2] * z[2];
    double nz[3], tmp[3] = {0.0, 0.0, 1.0};
    const double inv = 1.0 / sqrt(sq);
    nz[0] = z[0] * inv;
    nz[1] = z[1] * inv;
    nz[2] = z[2] * inv;
    const double cosPhi = nz[0] * tmp[0] + nz[1] * tmp[1] + nz[2] * tmp[2];
    if (cosPhi * cosPhi > 0.99)
    {
      tmp[0] = 1.0;
      tmp[1] = 0.0;
      tmp[2] = 0.0;
    }
    double* y = A[1];
    y[0] = nz[1] * tmp[2] - tmp[1] * nz[2];
    y[1] = -(nz[0] * tmp[2] - tmp[0] * nz[2]);
    y[2] = nz[0] * tmp[1] - tmp[0] * nz[1];
    const double len = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    if (len > 0.0)
    {
      const double s = 1.0 / len;
      y[0] *= s;
      y[1] *= s;
      y[2] *= s;
    }
  }
  /* x = y x z */
  A[0][0] = A[1][1] * A[2][2] - A[2][1] * A[1][2];
  A[0][1] = -(A[1][0] * A[2][2] - A[2][0] * A[1][2]);
  A[0][2] = A[1][0] * A[2][1] - A[2][0] * A[1][1];
}

static RcsJoint* lastJointBefore(const RcsBody* b)
{
  return b ? RcsBody_lastJointBeforeBody(b) : NULL;
}

static void appendChild(RcsBody* parent, RcsBody* child)
{
  child->parent = parent;
  if (!parent->firstChild)
  {
    parent->firstChild = child;
    parent->lastChild = child;
    child->prev = NULL;
    child->next = NULL;
  }
  else
  {
    RcsBody* last = parent->firstChild;
    while (last->next)
    {
      last = last->next;
    }
    last->next = child;
    child->prev = last;
    parent->lastChild = child;
  }
}

/* parent / child links of one <joint> (Rcs_URDFParser.c:612-850) */
static bool connectURDF(const RcsbXmlNode* node, RcsBody** bdyVec, RcsJoint** jntVec, const char* suffix)
{
  char jointName[256] = "", parentName[300] = "", childName[300] = "", type[256] = "";
  rcsb_xml_string(node, "name", jointName, 256);
  const RcsbXmlNode* pn = childByName(node, "parent");
  const RcsbXmlNode* cn = childByName(node, "child");
  if (!pn || !cn)
  {
    RCSB_FATAL("\"parent\" and \"child\" are required in joint definition of \"%s\"", jointName);
    return false;
  }
  rcsb_xml_string(pn, "link", parentName, 256);
  rcsb_xml_string(cn, "link", childName, 256);
  if (suffix)
  {
    strncat(parentName, suffix, sizeof(parentName) - strlen(parentName) - 1);
    strncat(childName, suffix, sizeof(childName) - strlen(childName) - 1);
  }
  RcsBody* parent = findBody(bdyVec, parentName);
  RcsBody* child = findBody(bdyVec, childName);
  if (!parent || !child)
  {
    RCSB_FATAL("URDF joint \"%s\": link \"%s\" not found", jointName, parent ? childName : parentName);
    return false;
  }
  rcsb_xml_string(node, "type", type, 256);
  const bool prismatic = strcasecmp(type, "prismatic") == 0;

  if (strcasecmp(type, "revolute") == 0 || strcasecmp(type, "continuous") == 0 || prismatic)
  {
    if (child->parent)
    {
      RCSB_FATAL("URDF link \"%s\" has two parents", child->name);
      return false;
    }
    appendChild(parent, child);
    char jn[300];
    snprintf(jn, sizeof(jn), "%s%s", jointName, suffix ? suffix : "");
    RcsJoint* jnt = findJoint(jntVec, jn);
    if (!jnt)
    {
      RCSB_FATAL("Joint \"%s\" not found", jn);
      return false;
    }
    jnt->dirIdx = 2;
    jnt->type = prismatic ? RCSJOINT_TRANS_Z : RCSJOINT_ROT_Z;
    const RcsbXmlNode* axisNode = childByName(node, "axis");
    if (axisNode)
    {
      double ax[3] = {1.0, 0.0, 0.0};
      rcsb_xml_vecn(axisNode, "xyz", ax, 3);
      if (ax[0] == 1.0 && ax[1] == 0.0 && ax[2] == 0.0)
      {
        jnt->dirIdx = 0;
        jnt->type = prismatic ? RCSJOINT_TRANS_X : RCSJOINT_ROT_X;
      }
      else if (ax[0] == 0.0 && ax[1] == 1.0 && ax[2] == 0.0)
      {
        jnt->dirIdx = 1;
        jnt->type = prismatic ? RCSJOINT_TRANS_Y : RCSJOINT_ROT_Y;
      }
      else if (ax[0] == 0.0 && ax[1] == 0.0 && ax[2] == 1.0)
      {
        jnt->dirIdx = 2;
        jnt->type = prismatic ? RCSJOINT_TRANS_Z : RCSJOINT_ROT_Z;
      }
      else
      {
        /* any other direction: the joint turns about z of a frame A_NJ whose z is the axis; the
         * frame goes into the joint's relative transform, its transpose into the child's */
        double A_NJ[3][3], T[3][3];
        frameWithZ(A_NJ, ax);
        if (jnt->A_JP)
        {
          memcpy(T, jnt->A_JP->rot, sizeof(T));
          Mat3d_mul(jnt->A_JP->rot, A_NJ, T);
        }
        else
        {
          jnt->A_JP = newHTr();
          memcpy(jnt->A_JP->rot, A_NJ, sizeof(A_NJ));
        }
        if (!child->A_BP)
        {
          child->A_BP = newHTr();
          for (int i = 0; i < 3; ++i)
          {
            for (int k = 0; k < 3; ++k)
            {
              child->A_BP->rot[i][k] = A_NJ[k][i];
            }
          }
        }
        else
        {
          double A_JN[3][3];
          for (int i = 0; i < 3; ++i)
          {
            for (int k = 0; k < 3; ++k)
            {
              A_JN[i][k] = A_NJ[k][i];
            }
          }
          memcpy(T, jnt->A_JP->rot, sizeof(T));
          Mat3d_mul(jnt->A_JP->rot, T, A_JN);
        }
      }
    }
    child->jnt = jnt;
    jnt->prev = lastJointBefore(parent);
    jnt->next = NULL;
  }
  else if (strcasecmp(type, "fixed") == 0)
  {
    if (childByName(node, "origin"))
    {
      HTr A;
      readOrigin(node, &A);
      if (!HTr_isIdentity(&A))
      {
        free(child->A_BP);
        child->A_BP = HTr_clone(&A);
      }
    }
    if (child->parent)
    {
      RCSB_FATAL("URDF link \"%s\" has two parents", child->name);
      return false;
    }
    appendChild(parent, child);
    if (child->physicsSim == RCSBODY_PHYSICS_DYNAMIC)
    {
      child->physicsSim = RCSBODY_PHYSICS_FIXED;
    }
  }
  else
  {
    RCSB_LOG(0, "URDF joint \"%s\": type \"%s\" is not supported (as in the reference)", jointName, type);
  }
  return true;
}

/* <model_state> of a URDF file: joint positions become q_init and q0 (Rcs_URDFParser.c:885) */
static void modelStateFromURDF(const RcsbXmlNode* robot, RcsJoint** jntVec, const char* suffix)
{
  for (const RcsbXmlNode* n = robot->children; n; n = n->next)
  {
    if (!tagIs(n, "model_state"))
    {
      continue;
    }
    for (const RcsbXmlNode* js = n->children; js; js = js->next)
    {
      if (!tagIs(js, "joint_state"))
      {
        continue;
      }
      char jn[300] = "";
      rcsb_xml_string(js, "joint", jn, 256);
      if (suffix)
      {
        strncat(jn, suffix, sizeof(jn) - strlen(jn) - 1);
      }
      RcsJoint* jnt = findJoint(jntVec, jn);
      if (!jnt)
      {
        RCSB_FATAL("Joint \"%s\" not found", jn);
        continue;
      }
      rcsb_xml_vecn(js, "position", &jnt->q_init, 1);
      jnt->q0 = jnt->q_init;
    }
  }
}

/* RcsGraph_rootBodyFromURDFFile (Rcs_URDFParser.c:924-1146) */
RcsBody* rcsb_urdf_root_body(const char* filename, const char* suffix, const HTr* A_BP, unsigned int* dof)
{
  RcsbXmlNode* robot = rcsb_xml_parse_file(filename);
  if (!robot || strcmp(robot->name, "robot") != 0)
  {
    rcsb_set_error("\"%s\" is not a URDF file (no <robot> root element)", filename);
    rcsb_xml_free(robot);
    return NULL;
  }
  unsigned int numLinks = 0, numJnts = 0;
  for (const RcsbXmlNode* c = robot->children; c; c = c->next)
  {
    numLinks += strcmp(c->name, "link") == 0;
    numJnts += strcmp(c->name, "joint") == 0;
  }
  RcsBody** bdyVec = (RcsBody**) calloc(numLinks + 1, sizeof(RcsBody*));
  RcsJoint** jntVec = (RcsJoint**) calloc(numJnts + 1, sizeof(RcsJoint*));
  unsigned int nb = 0, nj = 0;
  for (const RcsbXmlNode* c = robot->children; c; c = c->next)
  {
    RcsBody* b = nb < numLinks ? bodyFromURDF(c) : NULL;
    if (b)
    {
      if (suffix)
      {
        char* n = cloneWithSuffix(b->name, suffix);
        free(b->name);
        b->name = n;
      }
      bdyVec[nb++] = b;
    }
  }
  for (const RcsbXmlNode* c = robot->children; c; c = c->next)
  {
    RcsJoint* j = nj < numJnts ? jointFromURDF(c) : NULL;
    if (j)
    {
      if (suffix)
      {
        char* n = cloneWithSuffix(j->name, suffix);
        free(j->name);
        j->name = n;
        if (j->coupledJointName)
        {
          n = cloneWithSuffix(j->coupledJointName, suffix);
          free(j->coupledJointName);
          j->coupledJointName = n;
        }
      }
      jntVec[nj++] = j;
    }
  }
  if (dof)
  {
    *dof = numJnts;
  }
  for (const RcsbXmlNode* c = robot->children; c; c = c->next)
  {
    if (tagIs(c, "joint"))
    {
      connectURDF(c, bdyVec, jntVec, suffix);
    }
  }
  modelStateFromURDF(robot, jntVec, suffix);

  /* the first link without a parent is the root, further ones follow it as siblings */
  RcsBody* root = NULL;
  for (RcsBody** bv = bdyVec; *bv; ++bv)
  {
    if ((*bv)->parent)
    {
      continue;
    }
    if (!root)
    {
      root = *bv;
    }
    else
    {
      RcsBody* b = root;
      while (b->next)
      {
        b = b->next;
      }
      b->next = *bv;
      (*bv)->prev = b;
    }
    if (A_BP)
    {
      if (!(*bv)->A_BP)
      {
        (*bv)->A_BP = newHTr();
      }
      HTr_transformSelf((*bv)->A_BP, A_BP);
    }
  }
  if (!root)
  {
    rcsb_set_error("Couldn't find root link in URDF model \"%s\" (cyclic?)", filename);
    free(bdyVec);
    free(jntVec);
    rcsb_xml_free(robot);
    return NULL;
  }

  /* joints were connected in file order: a joint's predecessor may have been created later */
  for (RcsBody* b = root; b; b = RcsBody_depthFirstTraversalGetNext(b))
  {
    RcsJoint* prevJnt = lastJointBefore(b->parent);
    for (RcsJoint* j = b->jnt; j; j = j->next)
    {
      if (prevJnt != j->prev)
      {
        j->prev = prevJnt;
        j->next = NULL;
      }
    }
  }

  /* URDF: q = c q_master + o; here: q = q_init + c (q_master - q_master_init), so
   * q_init = c q_master_init + o (the offset was parked in q_init) */
  for (RcsBody* b = root; b; b = RcsBody_depthFirstTraversalGetNext(b))
  {
    for (RcsJoint* j = b->jnt; j; j = j->next)
    {
      if (!j->coupledJointName)
      {
        continue;
      }
      RcsJoint* master = NULL;
      for (unsigned int i = 0; i < nj && !master; ++i)
      {
        master = strcmp(jntVec[i]->name, j->coupledJointName) == 0 ? jntVec[i] : NULL;
      }
      if (!master)
      {
        RCSB_FATAL("URDF mimic joint \"%s\": master \"%s\" not found", j->name, j->coupledJointName);
        continue;
      }
      j->q_init = j->couplingFactors->ele[0] * master->q_init + j->q_init;
    }
  }
  free(bdyVec);
  free(jntVec);
  rcsb_xml_free(robot);
  return root;
}

/* RcsGraph_fromURDFFile (Rcs_URDFParser.c:1152-1224) */
RcsGraph* RcsGraph_fromURDFFile(const char* configFile)
{
  char filename[512] = "";
  if (!configFile || !Rcs_getAbsoluteFileName(configFile, filename))
  {
    rcsb_set_error("URDF file \"%s\" not found in resource path", configFile ? configFile : "NULL");
    return NULL;
  }
  RcsBody* root = rcsb_urdf_root_body(filename, NULL, NULL, NULL);
  if (!root)
  {
    return NULL;
  }
  RcsGraph* self = (RcsGraph*) calloc(1, sizeof(RcsGraph));
  self->root = root;
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      JNT->jointIndex = (int) self->dof++;
      JNT->jacobiIndex = JNT->constrained ? -1 : (int) self->nJ++;
    }
  }
  self->q = MatNd_create(self->dof, 1);
  self->q_dot = MatNd_create(self->dof, 1);
  self->xmlFile = cloneString(configFile);
  RcsGraph_makeJointsConsistent(self);
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (JNT->coupledTo)
      {
        self->q->ele[JNT->jointIndex] =
          RcsJoint_computeSlaveJointAngle(JNT, self->q->ele[JNT->coupledTo->jointIndex]);
      }
    }
  }
  int errs = 0, warnings = 0;
  if (!RcsGraph_check(self, &errs, &warnings))
  {
    rcsb_set_error("Check for URDF graph \"%s\" failed: %d errors %d warnings", configFile, errs, warnings);
    RcsGraph_destroy(self);
    return NULL;
  }
  return self;
}
