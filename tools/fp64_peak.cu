// FP64 pipe microbenchmark for the roofline denominators of the FP64-bound kernels
// ("measure a DFMA-chain microbenchmark on the box").
//   build/fp64_peak            prints one JSON line: sustained DFMA TFLOP/s (2 flop per FMA)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) dfmaKernel(double* out, int iters, double a, double b)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
  {
    x[i] = threadIdx.x * 1e-3 + i;
  }
  for (int it = 0; it < iters; ++it)
  {
#pragma unroll
    for (int r = 0; r < 16; ++r)
    {
#pragma unroll
      for (int i = 0; i < ILP; ++i)
      {
        x[i] = fma(x[i], a, b);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i)
  {
    s += x[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
static double run(int blocksPerSm, int sms, double* out)
{
  const int iters = 2000;
  const int blocks = blocksPerSm * sms;
  dfmaKernel<ILP><<<blocks, 256>>>(out, 10, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  dfmaKernel<ILP><<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double) blocks * 256 * iters * 16.0 * ILP;
  return 2.0 * fma / (ms * 1e-3) / 1e12;
}

int main()
{
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess)
  {
    printf("{\"error\": \"no CUDA device\"}\n");
    return 1;
  }
  double* out;
  cudaMalloc(&out, sizeof(double) * 256 * 8 * p.multiProcessorCount);
  double best = 0.0;
  double t[6];
  t[0] = run<1>(1, p.multiProcessorCount, out);   // 8 warps/SM, dependent chain: latency
  t[1] = run<4>(1, p.multiProcessorCount, out);
  t[2] = run<8>(1, p.multiProcessorCount, out);
  t[3] = run<4>(4, p.multiProcessorCount, out);
  t[4] = run<8>(4, p.multiProcessorCount, out);
  t[5] = run<8>(8, p.multiProcessorCount, out);
  for (int i = 0; i < 6; ++i)
  {
    best = t[i] > best ? t[i] : best;
  }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_tflops\": %.3f, \"ilp1_8warps\": %.3f, \"ilp4_8warps\": %.3f, "
         "\"ilp8_8warps\": %.3f, \"ilp4_32warps\": %.3f, \"ilp8_32warps\": %.3f, \"ilp8_64warps\": %.3f, "
         "\"sm_clock_khz_nominal\": %d}\n",
         p.name, p.multiProcessorCount, best, t[0], t[1], t[2], t[3], t[4], t[5], p.clockRate);
  return 0;
}
