// Micro-benchmark of the candidate-chain step: broadcast one lane's double to the warp and fma.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_shfl(double* out, int iters, long long* cyc) {
  double e = out[threadIdx.x], h = out[32 + threadIdx.x] * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 32; ++q) { const double d = __shfl_sync(0xffffffffu, e, q); e = fma(h, d, e); }
  }
  long long t1 = clock64();
  out[threadIdx.x] = e;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// same through shared memory: the lane whose turn it is stores, everybody loads
__global__ void k_smem(double* out, int iters, long long* cyc) {
  __shared__ double box[64];
  double e = out[threadIdx.x], h = out[32 + threadIdx.x] * 1e-9;
  const int lane = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      if (lane == q) box[q] = e;
      __syncwarp();
      const double d = box[q];
      e = fma(h, d, e);
    }
    __syncwarp();
  }
  long long t1 = clock64();
  out[threadIdx.x] = e;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// 32-bit shuffle pair issued explicitly
__global__ void k_shfl2(double* out, int iters, long long* cyc) {
  double e = out[threadIdx.x], h = out[32 + threadIdx.x] * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int lo = __shfl_sync(0xffffffffu, __double2loint(e), q), hi = __shfl_sync(0xffffffffu, __double2hiint(e), q);
      e = fma(h, __hiloint2double(hi, lo), e);
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = e;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// plain dependent DFMA for reference
__global__ void k_fma(double* out, int iters, long long* cyc) {
  double e = out[threadIdx.x], h = out[32 + threadIdx.x] * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 32; ++q) e = fma(h, e, e);
  }
  long long t1 = clock64();
  out[threadIdx.x] = e;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096);
  cudaMallocManaged(&cyc, 64);
  const int it = 1000;
  k_fma<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize();
  printf("dependent DFMA: %.1f cycles/step\n", (double)cyc[0] / (it * 32));
  k_shfl<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize();
  printf("shfl(double)+DFMA: %.1f cycles/step\n", (double)cyc[0] / (it * 32));
  k_shfl2<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize();
  printf("2x shfl(int)+DFMA: %.1f cycles/step\n", (double)cyc[0] / (it * 32));
  k_smem<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize();
  printf("smem box+DFMA: %.1f cycles/step\n", (double)cyc[0] / (it * 32));
  return 0;
}
