// Lane-group forward kinematics experiment: is a walk with THREE
// lanes per state faster than fkKernel's one thread per state?
//
// The frame update of the reference (Rcs_graph.c:61-226) is column-separable: R' = A R acts on every
// column of R alone, o' = o + R^T a on every component of o alone, and a joint rotation mixes two ROWS
// of R, i.e. again every column alone. So lane j of a group can own column j of R and component j of
// o (4 doubles instead of 12) and the walk needs no exchange at all; sin / cos of the joints are
// computed once per group (lane j takes joints j, j+3, j+6) and passed round by shuffles. Only the
// store needs the other lanes: a frame is 96 contiguous bytes = three 32-byte sectors, lane j writes
// sector j, which holds two of its own values and two of its neighbours' (two shuffle rounds). The
// three sectors of a frame then leave from three neighbouring lanes of one store instruction — the
// pattern tools/store_pattern.cu measured at 6.0 TB/s against 5.0 TB/s for one sector per thread.
//
// Both kernels walk the same synthetic serial chain shaped like BASELINE.json configs[1] (12 bodies,
// 7 revolute joints about z, constant relative transforms in constant memory, all 12 frames
// written: 56 B in, 1152 B out per state) and must produce bit-identical frames.
//   "thread":  one thread per state, three 32-byte stores per frame (fkKernel's walk and store path)
//   "lanes3":  three lanes per state as described (10 states per warp, lanes 30 / 31 idle)
// Prints one JSON line per kernel: ms per launch, GB/s, registers, and whether the outputs agree.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/lanegroup_fk tools/lanegroup_fk.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define NB 12
#define NJ 7

struct Chain
{
  double A[NB][9];    /* rotation of the body relative to its parent (before the joint) */
  double a[NB][3];    /* translation, in parent coordinates */
  int joint[NB];      /* index of the revolute joint (about the body's z axis) or -1 */
};
__constant__ Chain c_chain;

__device__ __forceinline__ void store4(double* p, double a, double b, double c, double d)
{
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c),
               "d"(d)
               : "memory");
}

/* ---- one thread per state -------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) threadKernel(const double* __restrict__ q, double* __restrict__ out, long n)
{
  const long s = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
  {
    return;
  }
  double qs[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
  {
    qs[j] = q[s * NJ + j];
  }
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, o[3] = {0, 0, 0};
  double* rec = out + s * (NB * 12);
#pragma unroll
  for (int b = 0; b < NB; ++b)
  {
    /* o' = o + R^T a, R' = A R */
    const double* A = c_chain.A[b];
    const double* a = c_chain.a[b];
#pragma unroll
    for (int j = 0; j < 3; ++j)
    {
      o[j] += R[j] * a[0] + R[3 + j] * a[1] + R[6 + j] * a[2];
    }
    double T[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
#pragma unroll
      for (int j = 0; j < 3; ++j)
      {
        T[3 * i + j] = A[3 * i] * R[j] + A[3 * i + 1] * R[3 + j] + A[3 * i + 2] * R[6 + j];
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i)
    {
      R[i] = T[i];
    }
    if (c_chain.joint[b] >= 0)
    {
      double sn, cs;
      sincos(qs[c_chain.joint[b]], &sn, &cs);
#pragma unroll
      for (int j = 0; j < 3; ++j)
      {
        const double u = R[j], v = R[3 + j];
        R[j] = cs * u + sn * v;
        R[3 + j] = cs * v - sn * u;
      }
    }
    store4(rec + 12 * b, o[0], o[1], o[2], R[0]);
    store4(rec + 12 * b + 4, R[1], R[2], R[3], R[4]);
    store4(rec + 12 * b + 8, R[5], R[6], R[7], R[8]);
  }
}

/* ---- three lanes per state ------------------------------------------------------------------- */
__device__ __forceinline__ double shflD(double v, int src)
{
  return __shfl_sync(0xffffffffu, v, src);
}

__global__ void __launch_bounds__(128) lanes3Kernel(const double* __restrict__ q, double* __restrict__ out, long n)
{
  const int lane = threadIdx.x & 31;
  const long warp = ((long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int g = lane / 3, j = lane - 3 * g;          /* group in the warp, column owned */
  const bool idle = lane >= 30;
  long s = warp * 10 + g;
  const bool live = !idle && s < n;
  if (warp * 10 >= n)
  {
    return;
  }
  if (!live)
  {
    s = n - 1;   /* shadow a valid state; nothing is stored */
  }
  const int base = 3 * (idle ? 9 : g);               /* first lane of the group (idle lanes: any) */

  /* lane j: sin / cos of joints j, j + 3, j + 6 */
  double sn[3], cs[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    const int jj = j + 3 * k;
    sn[k] = 0.0;
    cs[k] = 1.0;
    if (jj < NJ)
    {
      sincos(q[s * NJ + jj], &sn[k], &cs[k]);
    }
  }
  /* column j of R (c0, c1, c2 = R[0][j], R[1][j], R[2][j]) and o[j] */
  double c0 = j == 0 ? 1.0 : 0.0, c1 = j == 1 ? 1.0 : 0.0, c2 = j == 2 ? 1.0 : 0.0, oj = 0.0;
  double* rec = out + s * (NB * 12) + 4 * j;          /* this lane's sector of every frame */
#pragma unroll
  for (int b = 0; b < NB; ++b)
  {
    const double* A = c_chain.A[b];
    const double* a = c_chain.a[b];
    oj += c0 * a[0] + c1 * a[1] + c2 * a[2];
    const double t0 = A[0] * c0 + A[1] * c1 + A[2] * c2;
    const double t1 = A[3] * c0 + A[4] * c1 + A[5] * c2;
    const double t2 = A[6] * c0 + A[7] * c1 + A[8] * c2;
    c0 = t0;
    c1 = t1;
    c2 = t2;
    const int jn = c_chain.joint[b];                  /* compile-time after unrolling? no: constant memory */
    if (jn >= 0)
    {
      /* the joint's sin / cos live in lane jn % 3 of the group, slot jn / 3 */
      const int k = jn / 3, src = base + jn - 3 * k;
      const double mySn = k == 0 ? sn[0] : (k == 1 ? sn[1] : sn[2]);
      const double myCs = k == 0 ? cs[0] : (k == 1 ? cs[1] : cs[2]);
      const double S = shflD(mySn, src), C = shflD(myCs, src);
      const double u = c0, v = c1;
      c0 = C * u + S * v;
      c1 = C * v - S * u;
    }
    /* sectors: 0 = {o0 o1 o2 R00}, 1 = {R01 R02 R10 R11}, 2 = {R12 R20 R21 R22}; lane j holds
     * oj, R0j, R1j, R2j. Two rounds, every lane sends one value and receives one:
     *   round 1: lane0 <- o1 (lane1), lane1 <- R02 (lane2), lane2 <- R20 (lane0)
     *   round 2: lane0 <- o2 (lane2), lane1 <- R10 (lane0), lane2 <- R21 (lane1) */
    const double send1 = j == 0 ? c2 : (j == 1 ? oj : c0);
    const double send2 = j == 0 ? c1 : (j == 1 ? c2 : oj);
    const int from1 = base + (j == 0 ? 1 : (j == 1 ? 2 : 0));
    const int from2 = base + (j == 0 ? 2 : (j == 1 ? 0 : 1));
    const double r1 = shflD(send1, from1);
    const double r2 = shflD(send2, from2);
    if (live)
    {
      if (j == 0)
      {
        store4(rec + 12 * b, oj, r1, r2, c0);
      }
      else if (j == 1)
      {
        store4(rec + 12 * b, c0, r1, r2, c1);
      }
      else
      {
        store4(rec + 12 * b, c1, r1, r2, c2);
      }
    }
  }
}

static void check(cudaError_t e, const char* what)
{
  if (e != cudaSuccess)
  {
    fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e));
    exit(1);
  }
}

int main(int argc, char** argv)
{
  const long n = argc > 1 ? atol(argv[1]) : 1048576;
  const int reps = argc > 2 ? atoi(argv[2]) : 20;
  const size_t pad = argc > 3 ? (size_t) atol(argv[3]) : 0;   /* dynamic shared memory: limits the resident blocks */
  Chain h;
  srand(7);
  int nj = 0;
  for (int b = 0; b < NB; ++b)
  {
    /* a proper rotation from three Euler angles, a translation of a few decimetres */
    const double x = 0.3 * b + 0.1, y = 1.1 - 0.2 * b, z = 0.05 * b * b;
    const double cx = cos(x), sx = sin(x), cy = cos(y), sy = sin(y), cz = cos(z), sz = sin(z);
    const double R[9] = {cy * cz, sx * sy * cz + cx * sz, -cx * sy * cz + sx * sz,
                         -cy * sz, -sx * sy * sz + cx * cz, cx * sy * sz + sx * cz,
                         sy, -sx * cy, cx * cy};
    memcpy(h.A[b], R, sizeof(R));
    h.a[b][0] = 0.1 * (b % 3);
    h.a[b][1] = 0.05 * b;
    h.a[b][2] = 0.3 - 0.01 * b;
    h.joint[b] = (b >= 2 && nj < NJ && b != 6 && b != 10) ? nj++ : -1;
  }
  check(cudaMemcpyToSymbol(c_chain, &h, sizeof(h)), "chain");
  std::vector<double> hq((size_t) n * NJ);
  for (size_t i = 0; i < hq.size(); ++i)
  {
    hq[i] = 3.0 * ((double) rand() / RAND_MAX) - 1.5;
  }
  double *q, *o1, *o2;
  const size_t outBytes = (size_t) n * NB * 12 * sizeof(double);
  check(cudaMalloc(&q, hq.size() * sizeof(double)), "q");
  check(cudaMalloc(&o1, outBytes), "o1");
  check(cudaMalloc(&o2, outBytes), "o2");
  check(cudaMemcpy(q, hq.data(), hq.size() * sizeof(double), cudaMemcpyHostToDevice), "h2d");
  check(cudaMemset(o1, 0, outBytes), "memset");
  check(cudaMemset(o2, 0xff, outBytes), "memset");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double bytes = (double) n * (NJ * 8 + NB * 96);
  for (int variant = 0; variant < 2; ++variant)
  {
    const long threads = variant == 0 ? n : (n + 9) / 10 * 32;
    const dim3 grid((unsigned) ((threads + 127) / 128));
    cudaFuncAttributes fa;
    const void* fn = variant == 0 ? (const void*) threadKernel : (const void*) lanes3Kernel;
    check(cudaFuncGetAttributes(&fa, fn), "attr");
    check(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pad), "smem");
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, 128, pad);
    float best = 1e9f, sum = 0.f;
    for (int r = 0; r < reps + 3; ++r)
    {
      cudaEventRecord(e0);
      if (variant == 0)
      {
        threadKernel<<<grid, 128, pad>>>(q, o1, n);
      }
      else
      {
        lanes3Kernel<<<grid, 128, pad>>>(q, o2, n);
      }
      cudaEventRecord(e1);
      check(cudaEventSynchronize(e1), "run");
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (r >= 3)
      {
        best = ms < best ? ms : best;
        sum += ms;
      }
    }
    printf("{\"kernel\": \"%s\", \"states\": %ld, \"ms_mean\": %.4f, \"ms_best\": %.4f, \"gbs_mean\": %.1f, "
           "\"registers\": %d, \"blocks_per_sm\": %d}\n",
           variant == 0 ? "thread" : "lanes3", n, sum / reps, best, bytes / (sum / reps) * 1e-6, fa.numRegs, perSm);
  }
  std::vector<double> a((size_t) 4096 * NB * 12), b(a.size());
  const size_t tail = (size_t) (n - 4096) * NB * 12;
  cudaMemcpy(a.data(), o1 + tail, a.size() * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), o2 + tail, b.size() * sizeof(double), cudaMemcpyDeviceToHost);
  size_t diff = 0;
  double maxAbs = 0.0;
  for (size_t i = 0; i < a.size(); ++i)
  {
    diff += memcmp(&a[i], &b[i], 8) != 0;
    const double d = a[i] > b[i] ? a[i] - b[i] : b[i] - a[i];
    maxAbs = d > maxAbs || d != d ? d : maxAbs;
  }
  cudaMemcpy(a.data(), o1, a.size() * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), o2, b.size() * sizeof(double), cudaMemcpyDeviceToHost);
  for (size_t i = 0; i < a.size(); ++i)
  {
    diff += memcmp(&a[i], &b[i], 8) != 0;
    const double d = a[i] > b[i] ? a[i] - b[i] : b[i] - a[i];
    maxAbs = d > maxAbs || d != d ? d : maxAbs;
  }
  printf("{\"check\": \"first and last 4096 states\", \"differing_doubles\": %zu, \"max_abs_diff\": %.3e}\n", diff,
         maxAbs);
  return maxAbs < 1e-13 ? 0 : 2;
}
