/*
 * Summary statistics of a per-state result on the device: the payload of the one collective
 * of the multi-GPU path (every rank all-gathers 5 doubles per step).
 */
#include "rcsb_priv.h"

namespace
{

__global__ void __launch_bounds__(1024)
statsKernel(const double* __restrict__ x, long n, long stride, double* __restrict__ out)
{
  __shared__ double sh[5][32];
  double sum = 0.0, sq = 0.0, mn = 1.0 / 0.0, mx = -1.0 / 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x)
  {
    const double v = x[i * stride];
    sum += v;
    sq += v * v;
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
  {
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
    sq += __shfl_xor_sync(0xffffffffu, sq, off);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if (lane == 0)
  {
    sh[0][warp] = sum;
    sh[1][warp] = sq;
    sh[2][warp] = mn;
    sh[3][warp] = mx;
  }
  __syncthreads();
  if (warp == 0)
  {
    const int nw = (blockDim.x + 31) >> 5;
    sum = lane < nw ? sh[0][lane] : 0.0;
    sq = lane < nw ? sh[1][lane] : 0.0;
    mn = lane < nw ? sh[2][lane] : 1.0 / 0.0;
    mx = lane < nw ? sh[3][lane] : -1.0 / 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
      sum += __shfl_xor_sync(0xffffffffu, sum, off);
      sq += __shfl_xor_sync(0xffffffffu, sq, off);
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0)
    {
      out[0] = (double) n;
      out[1] = sum;
      out[2] = sq;
      out[3] = mn;
      out[4] = mx;
    }
  }
}

}   // namespace

extern "C" int rcsb_summary_stats(const double* x, long n, long stride, double* out5, void* stream)
{
  if (n < 0 || stride < 1 || !out5 || (n > 0 && !x))
  {
    rcsb_set_error("rcsb_summary_stats: invalid arguments (n = %ld, stride = %ld)", n, stride);
    return RCSB_EINVAL;
  }
  statsKernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(x, n, stride, out5);
  RCSB_CUDA(cudaGetLastError());
  rcsbhost::g_launches.fetch_add(1);
  return RCSB_OK;
}
