// Checks rcsb::sincosBatch (rcs_b200/csrc/rcsb_device.cuh) bit for bit against the CUDA library's
// sincos() on 2^26 arguments: uniform in [-8, 8], uniform in [-1e5, 1e5], log-uniform up to 2^31,
// and the special values the chain kernels can meet (+-0, denormals, multiples of pi/2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rcs_b200/csrc -I include -o build/sincos_check tools/sincos_check.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "rcsb_device.cuh"

__global__ void checkKernel(const double* x, long n, unsigned long long* bad, unsigned long long* outOfRange)
{
  const long i = ((long) blockIdx.x * blockDim.x + threadIdx.x) * 7;
  if (i + 7 > n)
  {
    return;
  }
  double a[7], s[7], c[7];
  for (int k = 0; k < 7; ++k)
  {
    a[k] = x[i + k];
  }
  const bool ok = rcsb::sincosBatch<7>(a, s, c);
  if (!ok)
  {
    atomicAdd(outOfRange, 1ull);
    return;
  }
  for (int k = 0; k < 7; ++k)
  {
    double rs, rc;
    sincos(a[k], &rs, &rc);
    if (__double_as_longlong(rs) != __double_as_longlong(s[k]) ||
        __double_as_longlong(rc) != __double_as_longlong(c[k]))
    {
      atomicAdd(bad, 1ull);
    }
  }
}

int main()
{
  const long n = 7l << 23;
  std::vector<double> h((size_t) n);
  unsigned long long st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    return (double) (st >> 11) / 9007199254740992.0;
  };
  for (long i = 0; i < n; ++i)
  {
    const int kind = (int) (i % 4);
    double v;
    if (kind == 0) v = 16.0 * rnd() - 8.0;
    else if (kind == 1) v = 2e5 * rnd() - 1e5;
    else if (kind == 2) v = (rnd() < 0.5 ? -1.0 : 1.0) * exp2(62.0 * rnd() - 31.0) * 0.999;
    else v = (double) ((long) (rnd() * 4000.0) - 2000) * 1.5707963267948966 * (1.0 + (rnd() - 0.5) * 1e-15);
    h[(size_t) i] = v;
  }
  const double special[] = {0.0, -0.0, 4.9e-324, -4.9e-324, 1e-310, 2147483647.9, -2147483647.9, 105615.0, 105616.0};
  memcpy(h.data(), special, sizeof(special));
  double* d;
  unsigned long long *dBad, hBad[2] = {0, 0};
  cudaMalloc(&d, n * sizeof(double));
  cudaMalloc(&dBad, 16);
  cudaMemcpy(d, h.data(), n * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemset(dBad, 0, 16);
  const long threads = n / 7;
  checkKernel<<<(unsigned int) ((threads + 127) / 128), 128>>>(d, n, dBad, dBad + 1);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(hBad, dBad, 16, cudaMemcpyDeviceToHost);
  printf("{\"args\": %ld, \"mismatches\": %llu, \"batches_out_of_range\": %llu, \"cuda\": \"%s\"}\n", n, hBad[0], hBad[1],
         cudaGetErrorString(e));
  return (hBad[0] == 0 && hBad[1] == 0 && e == cudaSuccess) ? 0 : 1;
}
