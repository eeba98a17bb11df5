// Development tool: dependent-chain latency of DADD / DMUL / DFMA on sm_100a (unrolled chains, one warp).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false scripts/lat2_bench.cu -o scripts/lat2_bench
#include <cstdio>
__global__ void k(double* out, long long* cyc, const double* in)
{
  double a = in[0], b = in[1], c = in[2];
  long long t0, t1;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) a = __dadd_rn(a, b);
  t1 = clock64();
  cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) a = __dmul_rn(a, c);
  t1 = clock64();
  cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) a = __fma_rn(b, 1.0, a);
  t1 = clock64();
  cyc[2] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) a = __fma_rn(b, c, a);
  t1 = clock64();
  cyc[3] = t1 - t0;
  // two independent DADD chains interleaved
  double a2 = in[3];
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++)
  {
    a = __dadd_rn(a, b);
    a2 = __dadd_rn(a2, c);
  }
  t1 = clock64();
  cyc[4] = t1 - t0;
  // float add for reference
  float f = (float)in[0], g = (float)in[1];
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) f = __fadd_rn(f, g);
  t1 = clock64();
  cyc[5] = t1 - t0;
  out[threadIdx.x] = a + a2 + f;
}
int main()
{
  double h[4] = { 1.0000001, 1e-9, 1.0000000001, 0.5 };
  double *d, *o;
  long long* c;
  cudaMalloc(&d, 64); cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 64);
  cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
  for (int threads = 1; threads <= 32; threads *= 32)
    for (int rep = 0; rep < 2; rep++)
    {
      k<<<1, threads>>>(o, c, d);
      long long hc[8];
      cudaMemcpy(hc, c, 64, cudaMemcpyDeviceToHost);
      printf("threads %d: cycles per dependent op: dadd %.1f dmul %.1f dfma(b,1,a) %.1f dfma %.1f | 2 dadd chains: %.1f per pair | fadd %.1f\n", threads,
             hc[0] / 128.0, hc[1] / 128.0, hc[2] / 128.0, hc[3] / 128.0, hc[4] / 128.0, hc[5] / 128.0);
    }
  return 0;
}
