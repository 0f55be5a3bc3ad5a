/*
 * RcsGraph -> self-contained <Graph> XML (counterpart of RcsGraph_fprintXML,
 * src/RcsCore/Rcs_graph.c:3072, Rcs_body.c:1817, Rcs_joint.c:470).
 *
 * The file keeps everything the kinematics / dynamics path reads (tree, relative
 * transforms as 12 numbers, joint parameters, mass / cogVector / inertia) and drops what it
 * does not (shapes, sensors, colours, groups, includes). Numbers are written with 17
 * significant digits; angular quantities, which the loader multiplies by pi/180, are
 * written as the degree value that maps back to the stored radians bit for bit whenever
 * such a value exists within a few ulps. Loading the file with RcsGraph_create (or with the
 * reference's loader) reproduces the integer layout and, in practice, every parameter.
 */
#include "rcsb_internal.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

/* degree value d with d * (pi/180) == rad exactly, if one exists near rad * 180/pi */
static double exactDegrees(double rad)
{
  const double d0 = rad * (180.0 / M_PI);
  if (d0 * (M_PI / 180.0) == rad)
  {
    return d0;
  }
  double up = d0, dn = d0;
  for (int i = 0; i < 8; ++i)
  {
    up = nextafter(up, INFINITY);
    if (up * (M_PI / 180.0) == rad)
    {
      return up;
    }
    dn = nextafter(dn, -INFINITY);
    if (dn * (M_PI / 180.0) == rad)
    {
      return dn;
    }
  }
  return d0;
}

static void putHTrAttr(FILE* f, const char* attr, const HTr* A)
{
  fprintf(f, " %s=\"", attr);
  for (int i = 0; i < 3; ++i)
  {
    fprintf(f, "%.17g ", A->org[i]);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
    {
      fprintf(f, "%.17g%s", A->rot[i][j], (i == 2 && j == 2) ? "" : " ");
    }
  fprintf(f, "\"");
}

static void putJoint(FILE* f, const RcsJoint* j)
{
  static const char* typeNames[6] = {"RotX", "RotY", "RotZ", "TransX", "TransY", "TransZ"};
  const bool rot = RcsJoint_isRotation(j);

  fprintf(f, "    <Joint name=\"%s\" type=\"%s\"", j->name, typeNames[j->type]);

  if (j->coupledTo)
  {
    /* the loader derives a slave's range from its master; only q_init must survive */
    if (j->q_init != 0.0)
    {
      const double v = rot ? exactDegrees(j->q_init) : j->q_init;
      fprintf(f, " range=\"%.17g %.17g %.17g\"", v, v, v);
    }
  }
  else
  {
    /* q_init is the range centre of the file; q0 differs from it only through a
     * model_state, which is written at the end of the file */
    double r[3] = {j->q_min, j->q_init, j->q_max};
    if (rot)
    {
      for (int i = 0; i < 3; ++i)
      {
        r[i] = exactDegrees(r[i]);
      }
    }
    /* keep the loader's consistency check happy even if rounding moved a bound */
    if (r[1] < r[0])
    {
      r[0] = r[1];
    }
    if (r[1] > r[2])
    {
      r[2] = r[1];
    }
    fprintf(f, " range=\"%.17g %.17g %.17g\"", r[0], r[1], r[2]);
  }

  if (j->A_JP)
  {
    putHTrAttr(f, "transform", j->A_JP);
  }
  if (j->constrained)
  {
    fprintf(f, " constraint=\"true\"");
  }
  fprintf(f, " weightJL=\"%.17g\" weightCA=\"%.17g\" weightMetric=\"%.17g\"", j->weightJL,
          j->weightCA, j->weightMetric);

  static const char* ctrlNames[3] = {"pos", "vel", "tor"};
  fprintf(f, " ctrlType=\"%s\"", ctrlNames[j->ctrlType]);
  if (j->maxTorque != DBL_MAX)
  {
    fprintf(f, " torqueLimit=\"%.17g\"", j->maxTorque);
  }
  if (j->speedLimit != DBL_MAX)
  {
    fprintf(f, " speedLimit=\"%.17g\"", rot ? exactDegrees(j->speedLimit) : j->speedLimit);
  }
  if (j->accLimit != DBL_MAX)
  {
    fprintf(f, " accelerationLimit=\"%.17g\"", rot ? exactDegrees(j->accLimit) : j->accLimit);
  }
  if (j->decLimit != j->accLimit)
  {
    fprintf(f, " decelerationLimit=\"%.17g\"", rot ? exactDegrees(j->decLimit) : j->decLimit);
  }
  if (j->coupledTo)
  {
    fprintf(f, " coupledTo=\"%s\" couplingFactor=\"", j->coupledTo->name);
    const unsigned int n = j->couplingFactors->m * j->couplingFactors->n;
    for (unsigned int i = 0; i < n; ++i)
    {
      fprintf(f, "%.17g%s", j->couplingFactors->ele[i], i + 1 < n ? " " : "");
    }
    fprintf(f, "\"");
  }
  fprintf(f, " />\n");
}

int RcsGraph_writeXML(const RcsGraph* self, const char* fileName)
{
  if (!self || !fileName)
  {
    rcsb_set_error("RcsGraph_writeXML: NULL argument");
    return -1;
  }
  FILE* f = fopen(fileName, "w");
  if (!f)
  {
    rcsb_set_error("RcsGraph_writeXML: cannot open \"%s\"", fileName);
    return -1;
  }

  static const char* physNames[4] = {"none", "kinematic", "dynamic", "fixed"};
  fprintf(f, "<Graph name=\"rcsb_export\" >\n\n");

  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    /* rigid body joints whose first joint carries a (group) transform need a group again */
    const bool rbj = BODY->rigid_body_joints && BODY->jnt;
    const bool grouped = rbj && BODY->jnt->A_JP;
    if (grouped)
    {
      fprintf(f, "  <Group");
      putHTrAttr(f, "transform", BODY->jnt->A_JP);
      fprintf(f, " >\n");
    }

    fprintf(f, "  <Body name=\"%s\"", BODY->name);
    if (BODY->parent)
    {
      fprintf(f, " prev=\"%s\"", BODY->parent->name);
    }
    if (BODY->A_BP)
    {
      putHTrAttr(f, "transform", BODY->A_BP);
    }
    fprintf(f, " physics=\"%s\"", physNames[BODY->physicsSim]);
    fprintf(f, " mass=\"%.17g\"", BODY->m);
    fprintf(f, " cogVector=\"%.17g %.17g %.17g\"", BODY->Inertia->org[0], BODY->Inertia->org[1],
            BODY->Inertia->org[2]);
    fprintf(f, " inertia=\"");
    for (int i = 0; i < 9; ++i)
    {
      fprintf(f, "%.17g%s", BODY->Inertia->rot[i / 3][i % 3], i < 8 ? " " : "");
    }
    fprintf(f, "\"");

    if (rbj)
    {
      bool unitWeights = true;
      double q[6] = {0}, w[6] = {0};
      int k = 0;
      RCSBODY_TRAVERSE_JOINTS(BODY)
      {
        if (k < 6)
        {
          q[k] = k >= 3 ? exactDegrees(JNT->q_init) : JNT->q_init;
          w[k] = JNT->weightMetric;
          unitWeights = unitWeights && (w[k] == 1.0);
        }
        k++;
      }
      fprintf(f, " rigid_body_joints=\"%.17g %.17g %.17g %.17g %.17g %.17g", q[0], q[1], q[2],
              q[3], q[4], q[5]);
      if (!unitWeights)
      {
        fprintf(f, " %.17g %.17g %.17g %.17g %.17g %.17g", w[0], w[1], w[2], w[3], w[4], w[5]);
      }
      fprintf(f, "\" />\n");
    }
    else if (BODY->jnt)
    {
      fprintf(f, " >\n");
      RCSBODY_TRAVERSE_JOINTS(BODY)
      {
        putJoint(f, JNT);
      }
      fprintf(f, "  </Body>\n");
    }
    else
    {
      fprintf(f, " />\n");
    }

    if (grouped)
    {
      fprintf(f, "  </Group>\n");
    }
    fprintf(f, "\n");
  }

  /* joints whose centre was moved by a model_state (src/RcsCore/Rcs_graphParser.c:55-140):
   * q0 is restored the same way; slaves follow their master on load */
  fprintf(f, "  <model_state model=\"rcsb_export\" time_stamp=\"\" >\n");
  RCSGRAPH_TRAVERSE_BODIES(self)
  {
    RCSBODY_TRAVERSE_JOINTS(BODY)
    {
      if (!JNT->coupledTo && JNT->q0 != JNT->q_init)
      {
        fprintf(f, "    <joint_state joint=\"%s\" position=\"%.17g\" />\n", JNT->name,
                RcsJoint_isRotation(JNT) ? exactDegrees(JNT->q0) : JNT->q0);
      }
    }
  }
  fprintf(f, "  </model_state>\n\n");

  fprintf(f, "</Graph>\n");
  fclose(f);
  return 0;
}
