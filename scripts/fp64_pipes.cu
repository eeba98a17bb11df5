// Microbenchmark behind DESIGN.md "k_iekf reduction": how fast are the two ways of summing the
// per-point 6x6 normal-equation blocks on a B200 SM?
//   (1) vector DFMA  (per-thread accumulators)
//   (2) DMMA  mma.sync.aligned.m8n8k4.f64  (warp-level J^T W J over 4 points per instruction)
// Prints achieved TFLOP/s and cycles per warp-instruction per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_pipes scripts/fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters)
{
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; i++)
  {
    a0 = fma(a0, x, y);
    a1 = fma(a1, x, y);
    a2 = fma(a2, x, y);
    a3 = fma(a3, x, y);
    a4 = fma(a4, x, y);
    a5 = fma(a5, x, y);
    a6 = fma(a6, x, y);
    a7 = fma(a7, x, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void k_dmma(double* out, int iters)
{
  double c[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; i++)
  {
    dmma(c[0], c[1], a, b);
    dmma(c[2], c[3], a, b);
    dmma(c[4], c[5], a, b);
    dmma(c[6], c[7], a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7];
}

// dependent chain: one accumulator, the shape the reduction actually has (c += a_s b_s for 8 steps)
__global__ void k_dmma_chain(double* out, int iters)
{
  double c0 = 0, c1 = 0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; i++)
  {
    dmma(c0, c1, a, b);
    dmma(c0, c1, b, a);
    dmma(c0, c1, a, a);
    dmma(c0, c1, b, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1;
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : { 128, 256, 512, 1024 })
  {
    for (int which = 0; which < 3; which++)
    {
      const int blocks = sms * (1024 / threads);
      float best = 1e30f;
      for (int rep = 0; rep < 4; rep++)
      {
        cudaEventRecord(e0);
        if (which == 0) k_dfma<<<blocks, threads>>>(out, iters);
        if (which == 1) k_dmma<<<blocks, threads>>>(out, iters);
        if (which == 2) k_dmma_chain<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      const double warps = (double)blocks * threads / 32;
      const double winst = warps * iters * (which == 0 ? 8 : 4);
      const double flop = winst * (which == 0 ? 64.0 : 512.0);
      const double cyc_per_inst_smsp = (best * 1e-3) * (clk_khz * 1e3) / (winst / (sms * 4.0));
      printf("%-11s threads/block %4d (1024 thr/SM): %8.3f ms  %7.2f TFLOP/s  %6.2f cycles per warp-instr per SMSP (at %d MHz)\n",
             which == 0 ? "DFMA" : which == 1 ? "DMMA x4 ind" : "DMMA chain", threads, best, flop / best * 1e-9,
             cyc_per_inst_smsp, clk_khz / 1000);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
