/*
 * XML graph loader of rcs_b200: <Graph>/<Group>/<Body>/<Joint>/<Shape> files of the
 * reference (config/xml/...) -> RcsGraph.
 *
 * Written against the behaviour of the reference loader, not its code:
 *   RcsGraph_create                src/RcsCore/Rcs_graph.c:398-506
 *   RcsGraph_createFromXmlNode     src/RcsCore/Rcs_graphParser.c:1707-1770
 *   RcsGraph_parseBodies           src/RcsCore/Rcs_graphParser.c:1203-1635
 *   RcsBody_createFromXML          src/RcsCore/Rcs_graphParser.c:854-1180
 *   RcsBody_initJoint              src/RcsCore/Rcs_graphParser.c:504-849
 *   RcsBody_initShape              src/RcsCore/Rcs_graphParser.c:145-498
 *   default inertia                src/RcsCore/Rcs_body.c:869-956, Rcs_shape.c:83-445
 *   model_state                    src/RcsCore/Rcs_graphParser.c:55-140
 * The rules that decide the integer layout are spelled out in SURVEY.md, appendix A.
 *
 * <URDF> nodes: rcsb_urdf.c. Not supported (not needed for the hot path; logged and skipped):
 * <OpenRave>, .bvh files, generic bodies, sensors, mesh volumes (MESH shapes count as zero volume,
 * which is also what the reference computes when the mesh file is missing).
 */
#include "rcsb_internal.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>

#define RCSGRAPH_MAX_GROUPDEPTH 32

/* ---------------------------------------------------------------------------------- */
/* resource path (src/RcsCore/Rcs_resourcePath.cpp:100-240)                             */
/* ---------------------------------------------------------------------------------- */

#define RCSB_MAX_PATHS 64
static char* g_paths[RCSB_MAX_PATHS];
static int g_nPaths = 0;

static bool fileExists(const char* f)
{
  struct stat st;
  return f && stat(f, &st) == 0 && S_ISREG(st.st_mode);
}

bool Rcs_addResourcePath(const char* path)
{
  if (!path || !*path || g_nPaths >= RCSB_MAX_PATHS)
  {
    return false;
  }
  size_t n = strlen(path);
  char* p = (char*) malloc(n + 2);
  strcpy(p, path);
  if (p[n - 1] != '/')
  {
    p[n] = '/';
    p[n + 1] = '\0';
  }
  for (int i = 0; i < g_nPaths; ++i)
  {
    if (strcmp(g_paths[i], p) == 0)
    {
      free(p);
      return false;
    }
  }
  g_paths[g_nPaths++] = p;
  return true;
}

void Rcs_clearResourcePath(void)
{
  for (int i = 0; i < g_nPaths; ++i)
  {
    free(g_paths[i]);
  }
  g_nPaths = 0;
}

bool Rcs_getAbsoluteFileName(const char* fileName, char* absFileName)
{
  if (!fileName)
  {
    return false;
  }
  if (fileExists(fileName))
  {
    if (absFileName)
    {
      snprintf(absFileName, 512, "%s", fileName);
    }
    return true;
  }
  for (int i = 0; i < g_nPaths; ++i)
  {
    char full[512];
    snprintf(full, sizeof(full), "%s%s", g_paths[i], fileName);
    if (fileExists(full))
    {
      if (absFileName)
      {
        strcpy(absFileName, full);
      }
      return true;
    }
  }
  return false;
}

/* ---------------------------------------------------------------------------------- */
/* shapes and default inertia                                                           */
/* ---------------------------------------------------------------------------------- */

static bool isTag(const RcsbXmlNode* n, const char* name)
{
  return n && strcmp(n->name, name) == 0;
}

static char* strClone(const char* s)
{
  char* d = (char*) malloc(strlen(s) + 1);
  strcpy(d, s);
  return d;
}

static RcsShape* shapeFromXML(const RcsbXmlNode* node, const RcsBody* body)
{
  RcsShape* sh = (RcsShape*) calloc(1, sizeof(RcsShape));
  HTr_setIdentity(&sh->A_CB);
  sh->scale = 1.0;

  static const struct
  {
    const char* name;
    int type;
  } types[] = {{"SSL", RCSSHAPE_SSL},       {"SSR", RCSSHAPE_SSR},       {"BOX", RCSSHAPE_BOX},
               {"CYLINDER", RCSSHAPE_CYLINDER}, {"MESH", RCSSHAPE_MESH},   {"FRAME", RCSSHAPE_REFFRAME},
               {"SPHERE", RCSSHAPE_SPHERE}, {"CONE", RCSSHAPE_CONE},     {"GPISF", RCSSHAPE_GPISF},
               {"TORUS", RCSSHAPE_TORUS},   {"OCTREE", RCSSHAPE_OCTREE}, {"POINT", RCSSHAPE_POINT},
               {"MARKER", RCSSHAPE_MARKER}};
  const char* t = rcsb_xml_attr(node, "type");
  for (size_t i = 0; t && i < sizeof(types) / sizeof(types[0]); ++i)
  {
    if (strcmp(t, types[i].name) == 0)
    {
      sh->type = types[i].type;
    }
  }

  if (sh->type == RCSSHAPE_REFFRAME)
  {
    sh->extents[0] = sh->extents[1] = sh->extents[2] = 0.9;
  }

  rcsb_xml_vecn(node, "extents", sh->extents, 3);
  rcsb_xml_vecn(node, "radius", &sh->extents[0], 1);
  rcsb_xml_vecn(node, "length", &sh->extents[2], 1);
  rcsb_xml_htr(node, "transform", &sh->A_CB);
  rcsb_xml_vecn(node, "scale", &sh->scale, 1);

  if (rcsb_xml_attr(node, "from2Points"))
  {
    double p[6];
    if (rcsb_xml_vecn(node, "from2Points", p, 6))
    {
      HTr A2p;
      HTr_from2Points(&A2p, &p[0], &p[3]);
      HTr_transformSelf(&sh->A_CB, &A2p);
      const double d[3] = {p[0] - p[3], p[1] - p[4], p[2] - p[5]};
      sh->extents[2] = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    }
  }

  if (rcsb_xml_attr(node, "quat"))
  {
    double q[4];
    if (rcsb_xml_vecn(node, "quat", q, 4))
    {
      Quat_toRotationMatrix(sh->A_CB.rot, q);
    }
    memset(sh->A_CB.org, 0, sizeof(sh->A_CB.org));
    rcsb_xml_vecn(node, "pos", sh->A_CB.org, 3);
  }

  /* compute-type flags (defaults, then explicit attributes) */
  bool distance = true, graphics = true, physics = true, softPhysics = false, depth = false;
  if (body->physicsSim == RCSBODY_PHYSICS_NONE)
  {
    physics = false;
  }
  if (sh->type == RCSSHAPE_MESH)
  {
    physics = false;
    distance = false;
  }
  else
  {
    graphics = false;
  }
  if (sh->type == RCSSHAPE_REFFRAME)
  {
    graphics = true;
    distance = false;
    physics = false;
  }
  if (sh->type == RCSSHAPE_MARKER)
  {
    graphics = distance = physics = false;
  }
  rcsb_xml_bool(node, "distance", &distance);
  rcsb_xml_bool(node, "physics", &physics);
  rcsb_xml_bool(node, "graphics", &graphics);
  rcsb_xml_bool(node, "softPhysics", &softPhysics);
  rcsb_xml_bool(node, "depth", &depth);
  sh->computeType = (char) ((distance ? 1 : 0) | (physics ? 2 : 0) | (graphics ? 4 : 0) |
                            (softPhysics ? 16 : 0) | (depth ? 32 : 0));
  return sh;
}

/* src/RcsCore/Rcs_shape.c:83-160 */
static double shapeVolume(const RcsShape* s)
{
  const double* e = s->extents;
  switch (s->type)
  {
    case RCSSHAPE_SSL:
      return M_PI * e[0] * e[0] * e[2] + 4.0 / 3.0 * M_PI * pow(e[0], 3);
    case RCSSHAPE_SSR:
    {
      const double r = 0.5 * e[2];
      const double g = M_PI * r * r;
      double v = e[0] * e[1] * e[2];
      v += g * e[0];
      v += g * e[1];
      v += 0.75 * M_PI * pow(r, 3);
      return v;
    }
    case RCSSHAPE_BOX:
      return e[0] * e[1] * e[2];
    case RCSSHAPE_CYLINDER:
      return M_PI * pow(e[0], 2) * e[2];
    case RCSSHAPE_SPHERE:
      return 4.0 / 3.0 * M_PI * pow(e[0], 3);
    case RCSSHAPE_CONE:
      return M_PI / 3.0 * pow(e[0], 2) * e[2];
    case RCSSHAPE_TORUS:
      return (M_PI * pow(e[2] / 2, 2)) * (2.0 * M_PI * e[0]);
    default:   /* NONE, FRAME, POINT, MARKER, MESH (no mesh data), GPISF, OCTREE */
      return 0.0;
  }
}

/* src/RcsCore/Rcs_shape.c:231-262: SSL and cone reference points are not their COM */
static void shapeLocalCOM(const RcsShape* s, double r_com[3])
{
  double offs = 0.0;
  if (s->type == RCSSHAPE_SSL)
  {
    offs = 0.5 * s->extents[2];
  }
  else if (s->type == RCSSHAPE_CONE)
  {
    offs = 0.25 * s->extents[2];
  }
  else
  {
    memcpy(r_com, s->A_CB.org, 3 * sizeof(double));
    return;
  }
  for (int i = 0; i < 3; ++i)
  {
    const double b = s->A_CB.rot[0][i] * 0.0 + s->A_CB.rot[1][i] * 0.0 + s->A_CB.rot[2][i] * offs;
    r_com[i] = s->A_CB.org[i] + b;
  }
}

static void cylinderInertia(double I[3], double m, double r, double h)
{
  I[0] = (1.0 / 12.0) * m * (3.0 * r * r + h * h);
  I[1] = I[0];
  I[2] = 0.5 * m * r * r;
}

/* diagonal inertia about the shape's COM in the shape frame
 * (src/RcsCore/Rcs_shape.c:266-445) */
static void shapeInertiaDiag(const RcsShape* s, double density, double I[3])
{
  const double m = density * shapeVolume(s);
  const double r = s->extents[0], h = s->extents[2];
  I[0] = I[1] = I[2] = 0.0;

  switch (s->type)
  {
    case RCSSHAPE_SSR:
    case RCSSHAPE_BOX:
    {
      const double xx = pow(s->extents[0], 2), yy = pow(s->extents[1], 2), zz = pow(s->extents[2], 2);
      I[0] = (m / 12.0) * (yy + zz);
      I[1] = (m / 12.0) * (xx + zz);
      I[2] = (m / 12.0) * (xx + yy);
      break;
    }
    case RCSSHAPE_SSL:
    {
      /* capsule = cylinder + two hemispheres shifted by h/2 + 3r/8 */
      double I_cyl[3], I_hem[3];
      const double m_cyl = M_PI * r * r * h * density;
      cylinderInertia(I_cyl, m_cyl, r, h);
      const double m_hem = (2.0 / 3.0) * M_PI * r * r * r * density;
      I_hem[0] = 83.0 / 320.0 * m_hem * r * r;
      I_hem[1] = I_hem[0];
      I_hem[2] = 0.4 * m_hem * r * r;
      const double off = 0.5 * h + 3.0 * r / 8.0;
      const double steiner = m_hem * off * off;
      I_hem[0] += steiner;
      I_hem[1] += steiner;
      for (int i = 0; i < 3; ++i)
      {
        I_hem[i] *= 2.0;
        I[i] = I_cyl[i] + I_hem[i];
      }
      break;
    }
    case RCSSHAPE_CYLINDER:
      cylinderInertia(I, m, r, h);
      break;
    case RCSSHAPE_SPHERE:
      I[0] = 0.4 * m * r * r;
      I[1] = I[0];
      I[2] = I[0];
      break;
    case RCSSHAPE_CONE:
      I[0] = (3.0 / 80.0) * m * (4 * r * r + h * h);
      I[1] = I[0];
      I[2] = (3.0 / 10.0) * m * r * r;
      break;
    case RCSSHAPE_TORUS:
      I[0] = m * (4.0 * r * r + 5.0 * (h / 2.0) * (h / 2.0)) / 8.0;
      I[1] = I[0];
      I[2] = m * (4.0 * r * r + 3.0 * (h / 2.0) * (h / 2.0)) / 4.0;
      break;
    default:
      break;
  }
}

static double bodyVolume(const RcsBody* b)
{
  double v = 0.0;
  for (RcsShape** s = b->shape; s && *s; ++s)
  {
    v += shapeVolume(*s);
  }
  return v;
}

/* COM = volume-weighted mean of the shape COMs; inertia = sum over shapes of the shape's
 * inertia rotated into the body frame plus the parallel-axis term to the body COM, with
 * a uniform density m / V over ALL shapes (src/RcsCore/Rcs_body.c:869-956). */
void RcsBody_computeInertiaTensor(const RcsBody* self, HTr* I)
{
  memset(I->rot, 0, sizeof(I->rot));
  memset(I->org, 0, sizeof(I->org));

  if (self->m == 0.0)
  {
    return;
  }
  const double v_bdy = bodyVolume(self);
  if (v_bdy == 0.0)
  {
    return;
  }

  for (RcsShape** s = self->shape; s && *s; ++s)
  {
    double r_sh[3];
    const double v_sh = shapeVolume(*s);
    shapeLocalCOM(*s, r_sh);
    for (int i = 0; i < 3; ++i)
    {
      r_sh[i] *= v_sh / v_bdy;
      I->org[i] += r_sh[i];
    }
  }

  const double density = self->m / v_bdy;

  for (RcsShape** s = self->shape; s && *s; ++s)
  {
    double diag[3], I_s[3][3] = {{0}}, r_sgc[3], d[3];
    shapeInertiaDiag(*s, density, diag);
    I_s[0][0] = diag[0];
    I_s[1][1] = diag[1];
    I_s[2][2] = diag[2];
    Mat3d_similarityTransform(I_s, (*s)->A_CB.rot, I_s);

    shapeLocalCOM(*s, r_sgc);
    for (int i = 0; i < 3; ++i)
    {
      d[i] = I->org[i] - r_sgc[i];
    }
    const double m_sh = density * shapeVolume(*s);

    /* parallel-axis (Steiner) term, src/RcsCore/Rcs_basicMath.c:443 */
    I_s[0][0] += m_sh * (d[1] * d[1] + d[2] * d[2]);
    I_s[1][1] += m_sh * (d[0] * d[0] + d[2] * d[2]);
    I_s[2][2] += m_sh * (d[0] * d[0] + d[1] * d[1]);
    I_s[0][1] -= m_sh * (d[0] * d[1]);
    I_s[1][0] -= m_sh * (d[0] * d[1]);
    I_s[0][2] -= m_sh * (d[0] * d[2]);
    I_s[2][0] -= m_sh * (d[0] * d[2]);
    I_s[1][2] -= m_sh * (d[1] * d[2]);
    I_s[2][1] -= m_sh * (d[1] * d[2]);

    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
      {
        I->rot[i][j] += I_s[i][j];
      }
  }
}

/* ---------------------------------------------------------------------------------- */
/* joints                                                                               */
/* ---------------------------------------------------------------------------------- */

static void relTransformFromXML(const RcsbXmlNode* node, HTr** dst)
{
  if (rcsb_xml_attr(node, "transform"))
  {
    HTr A;
    HTr_setIdentity(&A);
    rcsb_xml_htr(node, "transform", &A);
    if (!HTr_isIdentity(&A))
    {
      *dst = HTr_clone(&A);
    }
  }
  if (rcsb_xml_attr(node, "quat"))
  {
    if (rcsb_xml_attr(node, "transform"))
    {
      RCSB_FATAL("\"transform\" is not allowed if \"quat\" exists");
      return;
    }
    HTr A;
    double q[4];
    HTr_setIdentity(&A);
    if (rcsb_xml_vecn(node, "quat", q, 4))
    {
      Quat_toRotationMatrix(A.rot, q);
    }
    rcsb_xml_vecn(node, "pos", A.org, 3);
    if (!HTr_isIdentity(&A))
    {
      free(*dst);
      *dst = HTr_clone(&A);
    }
  }
}

static int g_unnamedJointId = 0;
static int g_unnamedBodyId = 0;

static RcsJoint* jointFromXML(RcsGraph* self, RcsBody* b, const RcsbXmlNode* node,
                              const char* suffix, const HTr* A_group)
{
  RcsJoint* jnt = (RcsJoint*) calloc(1, sizeof(RcsJoint));
  RcsGraph_insertJoint(self, b, jnt);
  HTr_setIdentity(&jnt->A_JI);

  const char* name = rcsb_xml_attr(node, "name");
  char buf[512];
  if (name)
  {
    snprintf(buf, sizeof(buf), "%s%s", name, suffix);
  }
  else
  {
    snprintf(buf, sizeof(buf), "unnamed joint%s %d", suffix, g_unnamedJointId++);
  }
  jnt->name = strClone(buf);

  relTransformFromXML(node, &jnt->A_JP);

  /* the group transform lands on the first joint of the first body of the group */
  if (b->jnt == jnt && !HTr_isIdentity(A_group))
  {
    if (!jnt->A_JP)
    {
      jnt->A_JP = HTr_clone(A_group);
    }
    else
    {
      HTr_transformSelf(jnt->A_JP, A_group);
    }
  }

  rcsb_xml_bool(node, "constraint", &jnt->constrained);

  const char* type = rcsb_xml_attr(node, "type");
  if (type)
  {
    static const char* names[6] = {"RotX", "RotY", "RotZ", "TransX", "TransY", "TransZ"};
    int found = -1;
    for (int i = 0; i < 6; ++i)
    {
      if (strcmp(type, names[i]) == 0)
      {
        found = i;
      }
    }
    if (found < 0)
    {
      RCSB_FATAL("Joint \"%s\": unknown joint type \"%s\"", jnt->name, type);
      return jnt;
    }
    jnt->type = found;
    jnt->dirIdx = found % 3;
  }

  /* range: "min q0 max", "min max" (q0 = centre) or "v" (+-v, q0 = 0); degrees for
   * rotational joints */
  const unsigned int rangeEle = rcsb_xml_num_strings(node, "range");
  if (rangeEle >= 1 && rangeEle <= 3)
  {
    double ka[3];
    if (!rcsb_xml_vecn(node, "range", ka, rangeEle))
    {
      RCSB_FATAL("Joint \"%s\": malformed range", jnt->name);
      return jnt;
    }
    jnt->q_min = ka[0];
    if (rangeEle == 3)
    {
      jnt->q0 = ka[1];
      jnt->q_max = ka[2];
    }
    else if (rangeEle == 2)
    {
      jnt->q0 = 0.5 * (ka[1] + ka[0]);
      jnt->q_max = ka[1];
    }
    else
    {
      jnt->q_min = -ka[0];
      jnt->q0 = 0.0;
      jnt->q_max = ka[0];
    }
    if (!(jnt->q_min <= jnt->q_max) || !(jnt->q0 >= jnt->q_min && jnt->q0 <= jnt->q_max))
    {
      RCSB_FATAL("Joint \"%s\": inconsistent range q_min=%f q0=%f q_max=%f", jnt->name,
                 jnt->q_min, jnt->q0, jnt->q_max);
      return jnt;
    }
    if (RcsJoint_isRotation(jnt))
    {
      jnt->q_min *= (M_PI / 180.0);
      jnt->q0 *= (M_PI / 180.0);
      jnt->q_max *= (M_PI / 180.0);
    }
    jnt->q_init = jnt->q0;
  }
  else
  {
    if (!rcsb_xml_attr(node, "coupledTo"))
    {
      RCSB_FATAL("Joint \"%s\" has no range and is not coupled to another joint", jnt->name);
      return jnt;
    }
    jnt->q_init = 0.0;
  }

  jnt->weightJL = 1.0;
  rcsb_xml_vecn(node, "weightJL", &jnt->weightJL, 1);
  jnt->weightCA = 1.0;
  rcsb_xml_vecn(node, "weightCA", &jnt->weightCA, 1);
  jnt->weightMetric = 1.0;
  rcsb_xml_vecn(node, "weightMetric", &jnt->weightMetric, 1);

  jnt->ctrlType = RCSJOINT_CTRL_POSITION;
  jnt->maxTorque = DBL_MAX;
  const char* ctrl = rcsb_xml_attr(node, "ctrlType");
  if (ctrl)
  {
    if (!strcmp(ctrl, "Position") || !strcmp(ctrl, "pos"))
    {
      jnt->ctrlType = RCSJOINT_CTRL_POSITION;
    }
    else if (!strcmp(ctrl, "Velocity") || !strcmp(ctrl, "vel"))
    {
      jnt->ctrlType = RCSJOINT_CTRL_VELOCITY;
    }
    else if (!strcmp(ctrl, "Torque") || !strcmp(ctrl, "tor"))
    {
      jnt->maxTorque = 1.0;
      jnt->ctrlType = RCSJOINT_CTRL_TORQUE;
    }
    else
    {
      RCSB_FATAL("Joint \"%s\": unknown ctrlType \"%s\"", jnt->name, ctrl);
      return jnt;
    }
  }

  rcsb_xml_vecn(node, "torqueLimit", &jnt->maxTorque, 1);
  jnt->speedLimit = DBL_MAX;
  rcsb_xml_vecn(node, "speedLimit", &jnt->speedLimit, 1);
  jnt->accLimit = DBL_MAX;
  rcsb_xml_vecn(node, "accelerationLimit", &jnt->accLimit, 1);
  jnt->decLimit = jnt->accLimit;
  rcsb_xml_vecn(node, "decelerationLimit", &jnt->decLimit, 1);

  double gearRatio = 1.0;
  if (rcsb_xml_vecn(node, "gearRatio", &gearRatio, 1))
  {
    const double rpm2rad = M_PI / 30.0;
    jnt->speedLimit *= rpm2rad / gearRatio;
  }
  if (gearRatio == 1.0 && RcsJoint_isRotation(jnt))
  {
    if (jnt->speedLimit != DBL_MAX)
    {
      jnt->speedLimit *= (M_PI / 180.0);
    }
    if (jnt->accLimit != DBL_MAX)
    {
      jnt->accLimit *= (M_PI / 180.0);
    }
    if (jnt->decLimit != DBL_MAX)
    {
      jnt->decLimit *= (M_PI / 180.0);
    }
  }

  const char* coupled = rcsb_xml_attr(node, "coupledTo");
  if (coupled && *coupled)
  {
    snprintf(buf, sizeof(buf), "%s%s", coupled, suffix);
    jnt->coupledJointName = strClone(buf);
    unsigned int polyGrad = rcsb_xml_num_strings(node, "couplingFactor");
    if (polyGrad == 0)
    {
      polyGrad = 1;
    }
    if (polyGrad != 1 && polyGrad != 5 && polyGrad != 9)
    {
      RCSB_FATAL("Joint \"%s\": %u coupling coefficients (1, 5 or 9 supported)", jnt->name,
                 polyGrad);
      return jnt;
    }
    jnt->couplingFactors = MatNd_create(polyGrad, 1);
    for (unsigned int i = 0; i < polyGrad; ++i)
    {
      jnt->couplingFactors->ele[i] = 1.0;
    }
    rcsb_xml_vecn(node, "couplingFactor", jnt->couplingFactors->ele, polyGrad);
  }

  return jnt;
}

/* six constrained joints TransX,Y,Z, RotX,Y,Z (src/RcsCore/Rcs_body.c:1741-1812) */
static RcsJoint* createRigidBodyJoints(RcsGraph* self, RcsBody* b, const double q_rbj[6])
{
  RcsJoint* first = NULL;
  static const int types[6] = {RCSJOINT_TRANS_X, RCSJOINT_TRANS_Y, RCSJOINT_TRANS_Z,
                               RCSJOINT_ROT_X, RCSJOINT_ROT_Y, RCSJOINT_ROT_Z};
  for (int i = 0; i < 6; ++i)
  {
    RcsJoint* jnt = (RcsJoint*) calloc(1, sizeof(RcsJoint));
    RcsGraph_insertJoint(self, b, jnt);
    HTr_setIdentity(&jnt->A_JI);
    char buf[512];
    snprintf(buf, sizeof(buf), "%s_rigidBodyJnt%d", b->name, i);
    jnt->name = strClone(buf);
    jnt->constrained = true;
    jnt->type = types[i];
    jnt->dirIdx = i % 3;
    jnt->q0 = q_rbj[i];
    jnt->q_max = q_rbj[i] + 2.0 * M_PI;
    jnt->q_min = q_rbj[i] - 2.0 * M_PI;
    jnt->q_init = q_rbj[i];
    jnt->weightMetric = 1.0;
    jnt->speedLimit = DBL_MAX;
    jnt->maxTorque = DBL_MAX;
    if (i == 0)
    {
      first = jnt;
    }
  }
  return first;
}

/* ---------------------------------------------------------------------------------- */
/* bodies                                                                               */
/* ---------------------------------------------------------------------------------- */

static RcsBody* bodyFromXML(RcsGraph* self, const RcsbXmlNode* node, const char* suffix,
                            HTr* A_group, bool firstInGroup, RcsBody* groupRoot)
{
  if (!isTag(node, "Body"))
  {
    return NULL;
  }

  RcsBody* b = (RcsBody*) calloc(1, sizeof(RcsBody));
  b->A_BI = (HTr*) malloc(sizeof(HTr));
  HTr_setIdentity(b->A_BI);

  char msg[512];
  const char* name = rcsb_xml_attr(node, "name");
  if (name)
  {
    snprintf(msg, 256, "%s", name);
  }
  else
  {
    snprintf(msg, 256, "unnamed body %d", g_unnamedBodyId++);
  }
  if (strncmp(msg, "GenericBody", 11) == 0)
  {
    RCSB_FATAL("The name \"GenericBody\" is reserved");
  }
  b->xmlName = strClone(msg);
  b->suffix = strClone(suffix);
  snprintf(msg, sizeof(msg), "%s%s", b->xmlName, suffix);
  b->name = strClone(msg);

  /* parent: group default, overridden by "prev" (with the group suffix first, except
   * for the first body of a group) */
  RcsBody* parent = groupRoot;
  const char* prev = rcsb_xml_attr(node, "prev");
  if (prev)
  {
    if (!firstInGroup)
    {
      snprintf(msg, sizeof(msg), "%s%s", prev, suffix);
      parent = RcsGraph_getBodyByName(self, msg);
    }
    if (!parent)
    {
      parent = RcsGraph_getBodyByName(self, prev);
    }
  }

  RcsGraph_insertBody(self, parent, b);

  relTransformFromXML(node, &b->A_BP);

  const char* physics = rcsb_xml_attr(node, "physics");
  if (!physics || !strcmp(physics, "none"))
  {
    b->physicsSim = RCSBODY_PHYSICS_NONE;
  }
  else if (!strcmp(physics, "kinematic"))
  {
    b->physicsSim = RCSBODY_PHYSICS_KINEMATIC;
  }
  else if (!strcmp(physics, "dynamic"))
  {
    b->physicsSim = RCSBODY_PHYSICS_DYNAMIC;
  }
  else if (!strcmp(physics, "fixed"))
  {
    b->physicsSim = RCSBODY_PHYSICS_FIXED;
  }
  else
  {
    RCSB_FATAL("Unknown physics simulation type \"%s\"", physics);
  }

  /* rigid body joints: "true" | 6 values (xyz + Euler deg) | 12 values (+6 weights) */
  int nJoints = 0;
  if (rcsb_xml_attr(node, "rigid_body_joints"))
  {
    double q_rbj[12] = {0};
    b->rigid_body_joints = true;
    nJoints = 6;
    const unsigned int nStr = rcsb_xml_num_strings(node, "rigid_body_joints");
    if (nStr == 1)
    {
      rcsb_xml_bool(node, "rigid_body_joints", &b->rigid_body_joints);
    }
    else if (nStr == 6 || nStr == 12)
    {
      rcsb_xml_vecn(node, "rigid_body_joints", q_rbj, nStr);
      for (int i = 3; i < 6; ++i)
      {
        q_rbj[i] *= M_PI / 180.0;
      }
    }
    else
    {
      RCSB_FATAL("rigid_body_joints of body \"%s\" has %u entries", b->name, nStr);
    }

    RcsJoint* rbj0 = createRigidBodyJoints(self, b, q_rbj);
    if (nStr == 12)
    {
      int k = 0;
      for (RcsJoint* j = rbj0; j; j = j->next)
      {
        j->weightMetric = q_rbj[6 + k++];
      }
    }
    if (!HTr_isIdentity(A_group))
    {
      rbj0->A_JP = HTr_clone(A_group);
    }
  }

  /* shapes first: the default inertia depends on them */
  unsigned int nShapes = 0;
  for (const RcsbXmlNode* c = node->children; c; c = c->next)
  {
    nShapes += isTag(c, "Shape") ? 1 : 0;
  }
  b->shape = (RcsShape**) calloc(nShapes + 1, sizeof(RcsShape*));
  nShapes = 0;
  for (const RcsbXmlNode* c = node->children; c; c = c->next)
  {
    if (isTag(c, "Shape"))
    {
      b->shape[nShapes++] = shapeFromXML(c, b);
    }
  }

  b->Inertia = (HTr*) calloc(1, sizeof(HTr));
  rcsb_xml_vecn(node, "mass", &b->m, 1);
  RcsBody_computeInertiaTensor(b, b->Inertia);
  rcsb_xml_vecn(node, "inertia", &b->Inertia->rot[0][0], 9);
  rcsb_xml_vecn(node, "cogVector", b->Inertia->org, 3);

  unsigned int xmlJntCount = 0;
  HTr identity;
  HTr_setIdentity(&identity);
  for (const RcsbXmlNode* c = node->children; c; c = c->next)
  {
    if (isTag(c, "Joint"))
    {
      if (b->rigid_body_joints)
      {
        RCSB_FATAL("Body \"%s\": no additional joints for rigid body joints", b->name);
        break;
      }
      jointFromXML(self, b, c, suffix, xmlJntCount == 0 ? A_group : &identity);
      nJoints++;
      xmlJntCount++;
    }
  }

  /* a jointless first body takes the group transform on its own A_BP */
  if (nJoints == 0 && !HTr_isIdentity(A_group))
  {
    if (!b->A_BP)
    {
      b->A_BP = HTr_clone(A_group);
    }
    else
    {
      HTr_transformSelf(b->A_BP, A_group);
    }
  }

  /* consumed: only the first body of a group gets the group transform */
  HTr_setIdentity(A_group);
  return b;
}

/* Siblings are processed by iteration, children of Graph / Group by recursion; the
 * firstInGroup flag is dropped once a body has been created or a nested <Graph> has
 * been left (src/RcsCore/Rcs_graphParser.c:1203-1635). */
static void rewireJointChains(RcsBody* b)
{
  if (b->jnt)
  {
    b->jnt->prev = b->parent ? RcsBody_lastJointBeforeBody(b->parent) : NULL;
  }
  for (RcsBody* c = b->firstChild; c; c = c->next)
  {
    rewireJointChains(c);
  }
}

/* <URDF file= prev= suffix= transform= rigid_body_joints=> (src/RcsCore/Rcs_graphParser.c:1438-1600):
 * the URDF tree becomes the last child of `prev` (or a further root), optionally on six rigid-body
 * joints */
static void urdfFromXML(const RcsbXmlNode* node, RcsGraph* self, const char* suffix)
{
  char tmp[256] = "", filename[512] = "";
  RcsBody* pB = NULL;
  if (rcsb_xml_string(node, "prev", tmp, 256) > 0)
  {
    pB = RcsGraph_getBodyByName(self, tmp);
    if (!pB)
    {
      RCSB_FATAL("Body \"%s\" not found, which was specified as prev for an URDF node", tmp);
      return;
    }
  }
  tmp[0] = '\0';
  rcsb_xml_string(node, "file", tmp, 256);
  if (!Rcs_getAbsoluteFileName(tmp, filename))
  {
    RCSB_FATAL("Couldn't open urdf file \"%s\"", tmp);
    return;
  }
  char urdfSuffix[132] = "", ndExt[300];
  rcsb_xml_string(node, "suffix", urdfSuffix, 132);
  snprintf(ndExt, sizeof(ndExt), "%s%s", suffix, urdfSuffix);

  HTr A_local;
  HTr_setIdentity(&A_local);
  rcsb_xml_htr(node, "transform", &A_local);
  unsigned int dof = 0;
  RcsBody* urdfRoot = rcsb_urdf_root_body(filename, ndExt, &A_local, &dof);
  if (!urdfRoot)
  {
    RCSB_FATAL("Couldn't get URDF root from file \"%s\"", filename);
    return;
  }
  self->dof += dof;
  MatNd* q = MatNd_create(self->dof, 1);
  if (self->q)
  {
    memcpy(q->ele, self->q->ele, sizeof(double) * self->q->m);
    MatNd_destroy(self->q);
  }
  self->q = q;

  double q_rbj[12];
  memset(q_rbj, 0, sizeof(q_rbj));
  unsigned int nStr = 0;
  if (rcsb_xml_attr(node, "rigid_body_joints"))
  {
    urdfRoot->rigid_body_joints = true;
    nStr = rcsb_xml_num_strings(node, "rigid_body_joints");
    if (nStr == 1)
    {
      rcsb_xml_bool(node, "rigid_body_joints", &urdfRoot->rigid_body_joints);
    }
    else if (nStr == 6 || nStr == 12)
    {
      rcsb_xml_vecn(node, "rigid_body_joints", q_rbj, nStr);
      for (int i = 3; i < 6; ++i)
      {
        q_rbj[i] *= M_PI / 180.0;
      }
    }
    else
    {
      RCSB_FATAL("rigid_body_joints of body \"%s\" has %u entries", urdfRoot->name, nStr);
    }
  }
  if (urdfRoot->rigid_body_joints)
  {
    /* the URDF root has no joints of its own: the six joints are its only ones */
    RcsJoint* rbj0 = createRigidBodyJoints(self, urdfRoot, q_rbj);
    if (nStr == 12)
    {
      int k = 0;
      for (RcsJoint* j = rbj0; j; j = j->next)
      {
        j->weightMetric = q_rbj[6 + k++];
      }
    }
    if (!HTr_isIdentity(&A_local))
    {
      rbj0->A_JP = HTr_clone(&A_local);
      HTr_setIdentity(urdfRoot->A_BP);
    }
  }
  else if (urdfRoot->physicsSim != RCSBODY_PHYSICS_NONE)
  {
    urdfRoot->physicsSim = (pB && (pB->physicsSim == RCSBODY_PHYSICS_DYNAMIC || pB->physicsSim == RCSBODY_PHYSICS_FIXED))
                             ? RCSBODY_PHYSICS_FIXED : RCSBODY_PHYSICS_KINEMATIC;
  }

  /* attached last, as in the reference */
  if (pB)
  {
    urdfRoot->parent = pB;
    if (pB->firstChild)
    {
      pB->lastChild->next = urdfRoot;
      urdfRoot->prev = pB->lastChild;
      pB->lastChild = urdfRoot;
    }
    else
    {
      pB->firstChild = urdfRoot;
      pB->lastChild = urdfRoot;
    }
  }
  else if (!self->root)
  {
    self->root = urdfRoot;
  }
  else
  {
    RcsBody* b = self->root;
    while (b->next)
    {
      b = b->next;
    }
    b->next = urdfRoot;
    urdfRoot->prev = b;
  }

  /* The reference wires the joints of the URDF tree (and the rigid-body joints) while the tree has
   * no parent, so in the graph it returns the backward joint chain (jnt->prev) of every URDF body
   * ends at the URDF root: below an articulated `prev` body Jacobians and mass matrix lack the
   * columns of the joints above (x_dot != J q_dot) — until the graph is cloned: RcsGraph_clone
   * re-inserts every joint and wires the complete chains (the reference's controllers and solvers
   * work on clones). Here the chains are complete from the start: the first joint of every body of
   * the tree is re-wired now that the parent is known (tests/test_urdf_cpu.py, test_urdf_gpu.py). */
  rewireJointChains(urdfRoot);
}

static void parseBodies(const RcsbXmlNode* node, RcsGraph* self, const char* suffix, HTr* A,
                        bool firstInGroup, int level, RcsBody* root[RCSGRAPH_MAX_GROUPDEPTH])
{
  for (; node; node = node->next)
  {
    if (level >= RCSGRAPH_MAX_GROUPDEPTH - 2)
    {
      RCSB_FATAL("Group level exceeds maximum: %d", level);
      return;
    }

    if (isTag(node, "Graph"))
    {
      const char* rp = rcsb_xml_attr(node, "resourcePath");
      if (rp)
      {
        char* tmp = strClone(rp);
        for (char* tok = strtok(tmp, " "); tok; tok = strtok(NULL, " "))
        {
          Rcs_addResourcePath(tok);
        }
        free(tmp);
      }
      parseBodies(node->children, self, suffix, A, firstInGroup, level, root);
      firstInGroup = false;
    }
    else if (isTag(node, "Group"))
    {
      HTr A_local, A_group = *A;
      HTr_setIdentity(&A_local);
      rcsb_xml_htr(node, "transform", &A_local);
      HTr_transformSelf(&A_group, &A_local);

      /* the reference keeps group suffixes in 16-byte buffers */
      char name[32] = "", ndExt[16];
      rcsb_xml_string(node, "name", name, 32);
      snprintf(ndExt, sizeof(ndExt), "%s%s", suffix, name);

      root[level + 1] = root[level];
      char prevName[32] = "";
      if (rcsb_xml_string(node, "prev", prevName, 32) > 0)
      {
        RcsBody* pB = RcsGraph_getBodyByName(self, prevName);
        if (pB)
        {
          root[level + 1] = pB;
        }
      }
      parseBodies(node->children, self, ndExt, &A_group, true, level + 1, root);
    }
    else if (isTag(node, "URDF"))
    {
      urdfFromXML(node, self, suffix);
    }
    else if (isTag(node, "OpenRave"))
    {
      RCSB_LOG(1, "<%s> nodes are not supported by this loader - skipped", node->name);
    }
    else
    {
      RcsBody* nb = bodyFromXML(self, node, suffix, A, firstInGroup, root[level]);
      if (nb)
      {
        firstInGroup = false;
      }
    }
  }
}

/* ---------------------------------------------------------------------------------- */
/* model state                                                                          */
/* ---------------------------------------------------------------------------------- */

static bool parseModelState(const RcsbXmlNode* graphNode, RcsGraph* self, const char* mdlName)
{
  bool success = false;
  for (const RcsbXmlNode* n = graphNode->children; n; n = n->next)
  {
    if (strcasecmp(n->name, "model_state") != 0)
    {
      continue;
    }
    const char* model = rcsb_xml_attr(n, "model");
    if (strcmp(mdlName, model ? model : "") != 0)
    {
      continue;
    }
    success = true;
    for (const RcsbXmlNode* js = n->children; js; js = js->next)
    {
      if (strcasecmp(js->name, "joint_state") != 0)
      {
        continue;
      }
      RcsJoint* jnt = RcsGraph_getJointByName(self, rcsb_xml_attr(js, "joint"));
      if (!jnt)
      {
        continue;
      }
      double q;
      if (rcsb_xml_vecn(js, "position", &q, 1))
      {
        q *= RcsJoint_isRotation(jnt) ? (M_PI / 180.0) : 1.0;
        jnt->q0 = q;
        self->q->ele[jnt->jointIndex] = q;
        RCSGRAPH_TRAVERSE_BODIES(self)
        {
          RCSBODY_TRAVERSE_JOINTS(BODY)
          {
            if (JNT->coupledTo == jnt)
            {
              const double qs = RcsJoint_computeSlaveJointAngle(JNT, q);
              JNT->q0 = qs;
              self->q->ele[JNT->jointIndex] = qs;
            }
          }
        }
      }
      if (rcsb_xml_vecn(js, "velocity", &q, 1))
      {
        q *= RcsJoint_isRotation(jnt) ? (M_PI / 180.0) : 1.0;
        self->q_dot->ele[jnt->jointIndex] = q;
      }
    }
  }
  return success;
}

/* ---------------------------------------------------------------------------------- */
/* entry points                                                                         */
/* ---------------------------------------------------------------------------------- */

RcsGraph* RcsGraph_createFromXmlNode(const RcsbXmlNode* node)
{
  if (!node)
  {
    return NULL;
  }

  RcsGraph* self = (RcsGraph*) calloc(1, sizeof(RcsGraph));
  self->xmlFile = strClone("Created_from_xml_node");
  self->q = MatNd_create(0, 1);

  HTr A_rel;
  HTr_setIdentity(&A_rel);
  RcsBody* root[RCSGRAPH_MAX_GROUPDEPTH];
  memset(root, 0, sizeof(root));

  /* siblings of the top node are not part of this graph */
  RcsbXmlNode top = *node;
  top.next = NULL;
  parseBodies(&top, self, "", &A_rel, false, 0, root);

  self->q_dot = MatNd_create(self->dof, 1);
  RcsGraph_makeJointsConsistent(self);

  const char* mdlName = rcsb_xml_attr(node, "name");
  if (mdlName && *mdlName)
  {
    parseModelState(node, self, mdlName);
  }

  /* Slave joints follow their masters (first step of RcsGraph_setState,
   * src/RcsCore/Rcs_graph.c:275). The forward kinematics itself is device work. */
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

  return self;
}

RcsGraph* RcsGraph_createFromBuffer(const char* buffer, unsigned int size)
{
  RcsbXmlNode* root = rcsb_xml_parse_buffer(buffer, size);
  if (!root)
  {
    return NULL;
  }
  RcsGraph* self = isTag(root, "Graph") ? RcsGraph_createFromXmlNode(root) : NULL;
  rcsb_xml_free(root);
  return self;
}

RcsGraph* RcsGraph_create(const char* cfgFile)
{
  char filename[512] = "";
  if (!cfgFile)
  {
    return NULL;
  }
  if (!Rcs_getAbsoluteFileName(cfgFile, filename))
  {
    RcsGraph* self = RcsGraph_createFromBuffer(cfgFile, (unsigned int) strlen(cfgFile) + 1);
    if (!self)
    {
      rcsb_set_error("RcsGraph config file \"%s\" not found in resource path", cfgFile);
    }
    return self;
  }

  RcsbXmlNode* root = rcsb_xml_parse_file(filename);
  if (!root)
  {
    return NULL;
  }

  RcsGraph* self = NULL;
  const RcsbXmlNode* graphNode = NULL;

  if (isTag(root, "Graph"))
  {
    graphNode = root;
  }
  else
  {
    /* e.g. a <Controller> file with an inlined <Graph> */
    for (const RcsbXmlNode* c = root->children; c && !graphNode; c = c->next)
    {
      if (strcasecmp(c->name, "Graph") == 0)
      {
        graphNode = c;
      }
    }
  }

  if (graphNode)
  {
    self = RcsGraph_createFromXmlNode(graphNode);
    if (self)
    {
      free(self->xmlFile);
      self->xmlFile = strClone(filename);
    }
  }
  else
  {
    /* <Controller graph="file.xml"> : load the graph it names */
    char graphFile[256] = "";
    if (rcsb_xml_string(root, "graph", graphFile, 256) > 0)
    {
      rcsb_xml_free(root);
      /* resolve relative to the referring file first */
      char sibling[768];
      const char* slash = strrchr(filename, '/');
      if (slash && graphFile[0] != '/')
      {
        snprintf(sibling, sizeof(sibling), "%.*s%s", (int) (slash - filename + 1), filename,
                 graphFile);
        if (fileExists(sibling))
        {
          return RcsGraph_create(sibling);
        }
      }
      return RcsGraph_create(graphFile);
    }
    rcsb_set_error("No <Graph> found in \"%s\"", filename);
  }

  rcsb_xml_free(root);
  return self;
}
