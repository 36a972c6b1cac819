// does a single working warp slow down when the other warps of its CTA wait at a barrier? (B200)
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ double dsm[];
__global__ void one_warp_works(double* out, long long* cyc, double a, double b, int mode) {
  double x = threadIdx.x;
  int j = threadIdx.x & 31;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (threadIdx.x < 32) {
    t0 = clock64();
    if (mode == 0) {
#pragma unroll 16
      for (int i = 0; i < 4096; i++) x = fma(x, a, b);
    } else if (mode == 1) {
#pragma unroll 16
      for (int i = 0; i < 4096; i++) j = __shfl_sync(0xffffffffu, j, (j + 1) & 31);
    } else {
#pragma unroll 16
      for (int i = 0; i < 4096; i++) { dsm[j] = x; x = dsm[(j + 1) & 31] + 1.0; }
    }
    t1 = clock64();
  }
  __syncthreads();
  out[threadIdx.x] = x + j; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  long long h;
  cudaFuncSetAttribute(one_warp_works, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[3] = {"DFMA chain", "SHFL chain", "STS+LDS chain"};
  for (int mode = 0; mode < 3; mode++)
    for (int thr : {32, 128, 256, 512, 1024})
      for (size_t sm : {(size_t)1024, (size_t)195 * 1024}) {
        one_warp_works<<<1, thr, sm>>>(out, cyc, 0.999, 1e-3, mode);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-14s threads %4d smem %3zu KB: %.2f clk per step\n", names[mode], thr, sm / 1024, h / 4096.0);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
