// Micro-benchmarks that size the sweep kernel's design: DFMA latency/throughput, PRMT+DFMA issue
// rate, SHFL latency, fp64 exp throughput on one SM and chip-wide.  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat_dfma(double* out, int iters, long long* cyc) {
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a = fma(a, b, c);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lat_shfl(double* out, int iters, long long* cyc) {
  double a = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a = __shfl_sync(0xffffffffu, a, (k + 1) & 31) + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: each thread 8 independent DFMA chains
template <int PRMT>
__global__ void k_tput(double* out, const unsigned* in, int iters, long long* cyc) {
  double acc[8];
  for (int k = 0; k < 8; ++k) acc[k] = out[k];
  unsigned w = in[threadIdx.x];
  double r = out[9];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double x;
      if (PRMT) x = __hiloint2double((int)__byte_perm(w + i, 0u, 0x4044 + 0x100 * (k & 3)), 0);
      else x = r;
      acc[k] = fma(x, r, acc[k]);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_exp(double* out, int iters, long long* cyc) {
  double a = out[threadIdx.x] * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = exp(a) * 1e-3;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; unsigned* in; long long* cyc;
  cudaMalloc(&out, 1 << 24); cudaMemset(out, 0, 1 << 24);
  cudaMalloc(&in, 1 << 16); cudaMemset(in, 1, 1 << 16);
  cudaMallocManaged(&cyc, 64);
  const int it = 2000;
  k_lat_dfma<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize();
  printf("DFMA dependent latency: %.1f cycles\n", (double)cyc[0] / (it * 16));
  k_lat_shfl<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize();
  printf("SHFL(64-bit)+DADD dependent latency: %.1f cycles\n", (double)cyc[0] / (it * 16));
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    k_tput<0><<<1, 32 * warps>>>(out, in, it, cyc); cudaDeviceSynchronize();
    double d0 = (double)cyc[0];
    k_tput<1><<<1, 32 * warps>>>(out, in, it, cyc); cudaDeviceSynchronize();
    double d1 = (double)cyc[0];
    printf("%2d warps/SM: DFMA %.2f lanes/clk/SM   PRMT+DFMA %.2f lanes/clk/SM\n", warps, warps * 32.0 * it * 8 / d0, warps * 32.0 * it * 8 / d1);
  }
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_tput<1><<<148, 512>>>(out, in, it * 10, cyc); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_tput<1><<<148, 512>>>(out, in, it * 10, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("chip PRMT+DFMA: %.2f T genotype-FMA/s (148 CTAs x 512 thr)\n", 148.0 * 512 * it * 10 * 8 / (ms * 1e-3) / 1e12);
  }
  for (int warps : {1, 8, 16}) {
    k_exp<<<1, 32 * warps>>>(out, 500, cyc); cudaDeviceSynchronize();
    printf("exp(double) %2d warps: %.1f cycles per dependent exp, %.3f exp/clk/SM\n", warps, (double)cyc[0] / 500, warps * 32.0 * 500 / cyc[0]);
  }
  return 0;
}
