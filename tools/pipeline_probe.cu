// What can a host-buffer pipeline reach on this box? Models the end-to-end path of the IK workload
// (BASELINE.json configs[3]: 104 B in, 72 B out per solve, 4M solves, kernel 6.3 ms in all) with plain
// CUDA: `slots` streams, per stage two H2D copies, one kernel that spins for the stage's share of the
// kernel time, two D2H copies. Prints ms per pass for several stage sizes and stream counts, so that
// rcsb's runStaged (rcsb_api.cu) can be compared with the ceiling of the same schedule.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipeline_probe tools/pipeline_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void spin(long long cycles, double* sink)
{
  const long long t0 = clock64();
  while (clock64() - t0 < cycles)
  {
  }
  if (sink && threadIdx.x == 0 && blockIdx.x == 0 && cycles < 0)
  {
    *sink = 1.0;
  }
}

int main(int argc, char** argv)
{
  const long N = 4194304;
  const double kernelMsTotal = argc > 1 ? atof(argv[1]) : 6.3;
  int clockKHz = 0;
  cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, 0);
  double *hq, *hx, *hdet, *dq[4], *dx[4], *ddet[4];
  cudaHostAlloc(&hq, N * 7 * 8, cudaHostAllocDefault);
  cudaHostAlloc(&hx, N * 6 * 8, cudaHostAllocDefault);
  cudaHostAlloc(&hdet, N * 2 * 8, cudaHostAllocDefault);
  cudaStream_t st[4];
  for (int i = 0; i < 4; ++i)
  {
    cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    cudaMalloc(&dq[i], (1 << 20) * 7 * 8);
    cudaMalloc(&dx[i], (1 << 20) * 6 * 8);
    cudaMalloc(&ddet[i], (1 << 20) * 2 * 8);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const long chunks[] = {65536, 262144, 524288, 1048576};
  for (long chunk : chunks)
  {
    for (int slots = 1; slots <= 4; ++slots)
    {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep)
      {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, 0);
        cudaStreamSynchronize(0);
        int s = 0;
        for (long s0 = 0; s0 < N; s0 += chunk, s = (s + 1) % slots)
        {
          const long n = N - s0 < chunk ? N - s0 : chunk;
          cudaMemcpyAsync(dq[s], hq + s0 * 7, n * 7 * 8, cudaMemcpyHostToDevice, st[s]);
          cudaMemcpyAsync(dx[s], hx + s0 * 6, n * 6 * 8, cudaMemcpyHostToDevice, st[s]);
          const long long cyc = (long long) (kernelMsTotal * 1e-3 * (double) n / (double) N * clockKHz * 1e3);
          spin<<<296, 128, 0, st[s]>>>(cyc, nullptr);
          cudaMemcpyAsync(hq + s0 * 7, dq[s], n * 7 * 8, cudaMemcpyDeviceToHost, st[s]);
          cudaMemcpyAsync(hdet + s0 * 2, ddet[s], n * 2 * 8, cudaMemcpyDeviceToHost, st[s]);
        }
        for (int i = 0; i < slots; ++i)
        {
          cudaStreamSynchronize(st[i]);
        }
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      printf("{\"stage_states\": %ld, \"streams\": %d, \"ms_per_pass\": %.3f, \"solves_per_s\": %.4g, \"gbs_both_ways\": %.1f}\n",
             chunk, slots, best, N / (best * 1e-3), N * 176.0 / (best * 1e-3) * 1e-9);
    }
  }
  return 0;
}
