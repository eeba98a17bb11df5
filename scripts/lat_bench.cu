// Development tool: single-thread latencies (clock64) of the fp64 building blocks the map kernels chain together.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I include scripts/lat_bench.cu -o scripts/lat_bench
#include <cstdio>
#include "../vina_slam_b200/csrc/vn_math.cuh"

__global__ void k_lat(double* out, long long* cyc, const double* in)
{
  double a = in[0], b = in[1];
  long long t0, t1;
  // dependent DADD chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) a = da(a, b);
  t1 = clock64();
  cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) a = dm(a, b);
  t1 = clock64();
  cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) a = a / b;
  t1 = clock64();
  cyc[2] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) a = sqrt(a + 3.0);
  t1 = clock64();
  cyc[3] = t1 - t0;
  // eig3_sym on a planar covariance
  double L[6] = { in[2], in[3], in[4], in[5], in[6], in[7] }, ev[3], Q[9];
  t0 = clock64();
  eig3_sym(L, ev, Q);
  t1 = clock64();
  cyc[4] = t1 - t0;
  a += ev[0] + Q[3];
  // again (instruction cache warm)
  L[0] += 1e-3 * ev[1];
  t0 = clock64();
  eig3_sym(L, ev, Q);
  t1 = clock64();
  cyc[5] = t1 - t0;
  a += ev[0] + Q[3];
  // one cluster transform
  Cluster c, o;
  for (int k = 0; k < 6; k++) c.P[k] = in[2 + k];
  for (int k = 0; k < 3; k++) c.v[k] = in[8 + k];
  c.N = 17;
  double R[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, p[3] = { in[8], in[9], in[10] };
  R[1] = in[11];
  t0 = clock64();
  cluster_transform(o, c, R, p);
  t1 = clock64();
  cyc[6] = t1 - t0;
  a += o.P[3] + o.v[1];
  // global load latency (dependent pointer chase through `in`, cold)
  const double* q = in + 4096;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 8; i++) q = in + 4096 + ((long long)(*q)) ;
  t1 = clock64();
  cyc[7] = t1 - t0;
  out[0] = a + *q;
}

int main()
{
  const int N = 1 << 22;
  double* h = (double*)calloc(N, 8);
  h[0] = 1.0000001; h[1] = 0.9999999;
  h[2] = 2.0; h[3] = 0.3; h[4] = 0.1; h[5] = 1.5; h[6] = 0.2; h[7] = 0.01;
  h[8] = 3.0; h[9] = -2.0; h[10] = 1.0; h[11] = 0.01;
  for (int i = 0; i < 8; i++) h[4096 + i * 300000 % (N - 4096)] = 0;  // placeholder
  // pointer chase: element at offset k holds the next offset (spread over 32 MB)
  long long offs[9] = { 0, 700001, 1400003, 2100007, 2800009, 3500011, 350003, 1050005, 5 };
  for (int i = 0; i < 8; i++) h[4096 + offs[i]] = (double)offs[i + 1];
  double *d, *o;
  long long* c;
  cudaMalloc(&d, N * 8); cudaMalloc(&o, 64); cudaMalloc(&c, 64 * 8);
  cudaMemcpy(d, h, N * 8, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++)
  {
    // flush L2
    void* f; cudaMalloc(&f, 256 << 20); cudaMemset(f, rep, 256 << 20); cudaFree(f);
    k_lat<<<1, 1>>>(o, c, d);
    long long hc[8];
    cudaMemcpy(hc, c, 64, cudaMemcpyDeviceToHost);
    printf("rep %d: cycles per op: dadd %.1f dmul %.1f ddiv %.1f dsqrt %.1f | eig3_sym %lld (again %lld) | cluster_transform %lld | dependent global load (L2 flushed) %.0f\n",
           rep, hc[0] / 256.0, hc[1] / 256.0, hc[2] / 64.0, hc[3] / 64.0, hc[4], hc[5], hc[6], hc[7] / 8.0);
  }
  return 0;
}
