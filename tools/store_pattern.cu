// Write-stream microbenchmark for the roofline discussion of fkKernel: how fast
// can one thread per state write its record when the record is produced piecewise?
//   pattern "frame":  thread t writes FRAMES x 96 B (three 32-byte st.global.v4.f64, L1
//                     no-allocate) into its own record of RECORD bytes -> every 32-byte sector is
//                     its own L2 request, the 32 lanes of an instruction hit 32 different records
//   pattern "row":    the warp writes the 32 records of its states as contiguous rows (coalesced
//                     8-byte stores), the order fkKernel uses for the Jacobian / M copy-out
// Prints one JSON line per pattern with the achieved GB/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/store_pattern tools/store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void store4(double* p, double a, double b, double c, double d)
{
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c),
               "d"(d)
               : "memory");
}

template <int FRAMES>
__global__ void __launch_bounds__(128, 4) frameKernel(double* out, long n, int recDoubles, int spin)
{
  const long s = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
  {
    return;
  }
  double* rec = out + s * recDoubles;
  double x = (double) s;
  for (int f = 0; f < FRAMES; ++f)
  {
    for (int k = 0; k < spin; ++k)   /* some dependent arithmetic between the frames, like a walk */
    {
      x = fma(x, 1.0000001, 1e-9);
    }
    store4(rec + 12 * f, x, x + 1, x + 2, x + 3);
    store4(rec + 12 * f + 4, x + 4, x + 5, x + 6, x + 7);
    store4(rec + 12 * f + 8, x + 8, x + 9, x + 10, x + 11);
  }
}

/* "triple": same records and bytes as "frame", but the three sectors of a frame are written by
 * three neighbouring lanes of ONE instruction (lane l of store j writes sector (32 j + l) % 3 of
 * state (32 j + l) / 3 of the warp): 96 contiguous bytes per state and instruction */
template <int FRAMES>
__global__ void __launch_bounds__(128, 4) tripleKernel(double* out, long n, int recDoubles)
{
  const long w0 = ((long) blockIdx.x * blockDim.x + threadIdx.x) / 32 * 32;
  const int lane = threadIdx.x & 31;
  if (w0 >= n)
  {
    return;
  }
  double x = (double) w0;
  for (int f = 0; f < FRAMES; ++f)
  {
#pragma unroll
    for (int j = 0; j < 3; ++j)
    {
      const int e = 32 * j + lane;
      const int t = e / 3, c = e - 3 * t;
      double* p = out + (w0 + t) * recDoubles + 12 * f + 4 * c;
      store4(p, x, x + 1, x + 2, x + 3);
    }
    x += 1.0;
  }
}

/* "bulk": every thread stages its frame in its own 96-byte shared-memory slot and hands it to the
 * TMA engine as one asynchronous bulk store (cp.async.bulk.global.shared::cta): no exchange
 * between lanes, no warp barrier, the store does not occupy the thread */
template <int FRAMES>
__global__ void __launch_bounds__(128, 4) bulkKernel(double* out, long n, int recDoubles, int spin)
{
  __shared__ __align__(16) double slot[128 * 12];
  const long s = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
  {
    return;
  }
  double* mine = slot + 12 * threadIdx.x;
  const unsigned int src = (unsigned int) __cvta_generic_to_shared(mine);
  double* rec = out + s * recDoubles;
  double x = (double) s;
  for (int f = 0; f < FRAMES; ++f)
  {
    for (int k = 0; k < spin; ++k)
    {
      x = fma(x, 1.0000001, 1e-9);
    }
    /* the previous bulk store must have read the slot before it is overwritten */
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    double2* m2 = reinterpret_cast<double2*>(mine);
    m2[0] = make_double2(x, x + 1);
    m2[1] = make_double2(x + 2, x + 3);
    m2[2] = make_double2(x + 4, x + 5);
    m2[3] = make_double2(x + 6, x + 7);
    m2[4] = make_double2(x + 8, x + 9);
    m2[5] = make_double2(x + 10, x + 11);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 96;" ::"l"(rec + 12 * f), "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(128, 4) rowKernel(double* out, long n, int recDoubles)
{
  const long w0 = ((long) blockIdx.x * blockDim.x + threadIdx.x) / 32 * 32;
  const int lane = threadIdx.x & 31;
  if (w0 >= n)
  {
    return;
  }
  double* base = out + w0 * recDoubles;
  const long total = 32L * recDoubles;
  for (long e = lane; e < total; e += 32)
  {
    base[e] = (double) e;
  }
}

static float timeIt(void (*launch)(double*, long, int), double* out, long n, int recDoubles)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i)
  {
    launch(out, n, recDoubles);
  }
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i)
  {
    launch(out, n, recDoubles);
  }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

static int g_spin = 0;
static void launchFrame12(double* out, long n, int rd)
{
  frameKernel<12><<<(unsigned) ((n + 127) / 128), 128>>>(out, n, rd, g_spin);
}
static void launchTriple12(double* out, long n, int rd)
{
  tripleKernel<12><<<(unsigned) ((n + 127) / 128), 128>>>(out, n, rd);
}
static void launchBulk12(double* out, long n, int rd)
{
  bulkKernel<12><<<(unsigned) ((n + 127) / 128), 128>>>(out, n, rd, g_spin);
}
static void launchRow(double* out, long n, int rd)
{
  rowKernel<<<(unsigned) ((n + 127) / 128), 128>>>(out, n, rd);
}

int main()
{
  const long n = 1 << 20;
  const int recDoubles = 144;   /* 12 frames x 12 doubles = 1152 B, the A_BI record of configs[1] */
  double* out = nullptr;
  if (cudaMalloc(&out, (size_t) n * recDoubles * sizeof(double)) != cudaSuccess)
  {
    printf("{\"error\": \"cudaMalloc\"}\n");
    return 1;
  }
  const double bytes = (double) n * recDoubles * 8.0;
  for (int spin = 0; spin <= 64; spin += 32)
  {
    g_spin = spin;
    const float ms = timeIt(launchFrame12, out, n, recDoubles);
    printf("{\"pattern\": \"frame: 32-byte sector stores, one record per thread\", \"spin\": %d, \"ms\": %.4f, "
           "\"GBps\": %.1f}\n", spin, ms, bytes / (ms * 1e-3) / 1e9);
  }
  const float mst = timeIt(launchTriple12, out, n, recDoubles);
  printf("{\"pattern\": \"triple: the 3 sectors of a frame from 3 neighbouring lanes of one store\", \"ms\": %.4f, "
         "\"GBps\": %.1f}\n", mst, bytes / (mst * 1e-3) / 1e9);
  for (int spin = 0; spin <= 64; spin += 64)
  {
    g_spin = spin;
    const float msb = timeIt(launchBulk12, out, n, recDoubles);
    printf("{\"pattern\": \"bulk: one 96-byte TMA bulk store per thread and frame\", \"spin\": %d, \"ms\": %.4f, "
           "\"GBps\": %.1f}\n", spin, msb, bytes / (msb * 1e-3) / 1e9);
  }
  const float msr = timeIt(launchRow, out, n, recDoubles);
  printf("{\"pattern\": \"row: coalesced 8-byte stores, 32 records per warp\", \"ms\": %.4f, \"GBps\": %.1f}\n", msr,
         bytes / (msr * 1e-3) / 1e9);
  cudaFree(out);
  return 0;
}
