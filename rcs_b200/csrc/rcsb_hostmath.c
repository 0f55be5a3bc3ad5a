/*
 * Load-time host math of rcs_b200: the handful of 3x3 / rigid-transform helpers the XML
 * loader needs to turn attributes into relative transforms and default inertia. Nothing
 * here evaluates a joint state — that is the CUDA kernels' job (rcsb_kernels.cu).
 *
 * Conventions follow the reference (src/RcsCore/Rcs_HTr.h:55-66): rot is row-major, its
 * rows are the axes of the frame expressed in the parent, A_21 rotates frame 1 into 2.
 */
#include "rcsb_internal.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

void HTr_setIdentity(HTr* A)
{
  memset(A, 0, sizeof(HTr));
  A->rot[0][0] = A->rot[1][1] = A->rot[2][2] = 1.0;
}

/* Exact comparison on purpose: the reference drops a relative transform only if every
 * element is exactly the identity's (src/RcsCore/Rcs_HTr.c:197, Rcs_Mat3d.c:449). */
bool HTr_isIdentity(const HTr* A)
{
  for (int i = 0; i < 3; ++i)
  {
    if (A->org[i] != 0.0)
    {
      return false;
    }
    for (int j = 0; j < 3; ++j)
    {
      if (A->rot[i][j] != (i == j ? 1.0 : 0.0))
      {
        return false;
      }
    }
  }
  return true;
}

HTr* HTr_clone(const HTr* A)
{
  HTr* B = (HTr*) malloc(sizeof(HTr));
  memcpy(B, A, sizeof(HTr));
  return B;
}

/* A_3I = A_32 * A_2I (src/RcsCore/Rcs_Mat3d.c:375) */
void Mat3d_mul(double A_3I[3][3], double A_32[3][3], double A_2I[3][3])
{
  double t[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
    {
      t[i][j] = A_32[i][0] * A_2I[0][j] + A_32[i][1] * A_2I[1][j] + A_32[i][2] * A_2I[2][j];
    }
  memcpy(A_3I, t, sizeof(t));
}

/* A_2I.rot = A_21.rot * A_1I.rot ; A_2I.org = A_1I.org + A_1I.rot^T * A_21.org
 * (src/RcsCore/Rcs_HTr.c:291-305) */
void HTr_transformSelf(HTr* A_2I, const HTr* A_21)
{
  HTr A_1I = *A_2I;
  Mat3d_mul(A_2I->rot, (double(*)[3]) A_21->rot, A_1I.rot);
  for (int i = 0; i < 3; ++i)
  {
    A_2I->org[i] = A_1I.org[i] + (A_1I.rot[0][i] * A_21->org[0] + A_1I.rot[1][i] * A_21->org[1] +
                                  A_1I.rot[2][i] * A_21->org[2]);
  }
}

/* x-y-z Euler angles -> rotation matrix (src/RcsCore/Rcs_Mat3d.c:1152) */
void Mat3d_fromEulerAngles(double A_KI[3][3], const double ea[3])
{
  const double sa = sin(ea[0]), ca = cos(ea[0]);
  const double sb = sin(ea[1]), cb = cos(ea[1]);
  const double sg = sin(ea[2]), cg = cos(ea[2]);

  A_KI[0][0] = cb * cg;
  A_KI[0][1] = ca * sg + sa * sb * cg;
  A_KI[0][2] = sa * sg - ca * sb * cg;
  A_KI[1][0] = -cb * sg;
  A_KI[1][1] = ca * cg - sa * sb * sg;
  A_KI[1][2] = sa * cg + ca * sb * sg;
  A_KI[2][0] = sb;
  A_KI[2][1] = -sa * cb;
  A_KI[2][2] = ca * cb;
}

/* Inverse of the above with the gimbal branch (src/RcsCore/Rcs_Mat3d.c:956) */
void Mat3d_toEulerAngles(double ea[3], double A_KI[3][3])
{
  const double cy = sqrt(A_KI[0][0] * A_KI[0][0] + A_KI[1][0] * A_KI[1][0]);

  if (cy > 16.0 * FLT_EPSILON)
  {
    ea[0] = -atan2(A_KI[2][1], A_KI[2][2]);
    ea[1] = -atan2(-A_KI[2][0], cy);
    ea[2] = -atan2(A_KI[1][0], A_KI[0][0]);
  }
  else
  {
    ea[0] = -atan2(-A_KI[1][2], A_KI[1][1]);
    ea[1] = -atan2(-A_KI[2][0], cy);
    ea[2] = 0.0;
  }
}

/* B_I = A_CB^T * C_I * A_CB (src/RcsCore/Rcs_Mat3d.c:364) */
void Mat3d_similarityTransform(double B_I[3][3], double A_CB[3][3], double C_I[3][3])
{
  double tmp[3][3], out[3][3];
  Mat3d_mul(tmp, C_I, A_CB);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
    {
      out[i][j] = A_CB[0][i] * tmp[0][j] + A_CB[1][i] * tmp[1][j] + A_CB[2][i] * tmp[2][j];
    }
  memcpy(B_I, out, sizeof(out));
}

/* Quaternion (w x y z) -> rotation matrix, row-major "A_BI" convention
 * (src/RcsCore/Rcs_quaternion.c:133) */
void Quat_toRotationMatrix(double A_BI[3][3], const double q[4])
{
  const double qw = q[0], qx = q[1], qy = q[2], qz = q[3];

  A_BI[0][0] = 1.0 - 2.0 * qy * qy - 2.0 * qz * qz;
  A_BI[1][0] = 2.0 * qx * qy - 2.0 * qz * qw;
  A_BI[2][0] = 2.0 * qx * qz + 2.0 * qy * qw;
  A_BI[0][1] = 2.0 * qx * qy + 2.0 * qz * qw;
  A_BI[1][1] = 1.0 - 2.0 * qx * qx - 2.0 * qz * qz;
  A_BI[2][1] = 2.0 * qy * qz - 2.0 * qx * qw;
  A_BI[0][2] = 2.0 * qx * qz - 2.0 * qy * qw;
  A_BI[1][2] = 2.0 * qy * qz + 2.0 * qx * qw;
  A_BI[2][2] = 1.0 - 2.0 * qx * qx - 2.0 * qy * qy;
}

static double vec3Normalize(double v[3])
{
  const double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (len > 0.0)
  {
    v[0] /= len;
    v[1] /= len;
    v[2] /= len;
  }
  return len;
}

static void vec3Cross(double c[3], const double a[3], const double b[3])
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* Frame at p1 with z pointing to p2 (src/RcsCore/Rcs_HTr.c:415) */
void HTr_from2Points(HTr* A_KI, const double p1[3], const double p2[3])
{
  double* ex = A_KI->rot[0];
  double* ey = A_KI->rot[1];
  double* ez = A_KI->rot[2];
  double tmp[3] = {0.0, 0.0, 1.0};

  memcpy(A_KI->org, p1, 3 * sizeof(double));
  for (int i = 0; i < 3; ++i)
  {
    ez[i] = p2[i] - p1[i];
  }

  if (vec3Normalize(ez) == 0.0)
  {
    HTr_setIdentity(A_KI);
    memcpy(A_KI->org, p1, 3 * sizeof(double));
    return;
  }

  double c = ez[2];   /* cos of angle between ez and the z axis */
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double ang = acos(c);
  if ((fabs(ang) < 1.0e-5) || (fabs(ang - M_PI) < 1.0e-5))
  {
    tmp[0] = 1.0;
    tmp[2] = 0.0;
  }

  vec3Cross(ey, ez, tmp);
  vec3Normalize(ey);
  vec3Cross(ex, ey, ez);
  vec3Normalize(ex);
}
