// nm_peak.cu -- FMA issue-rate microbenchmark: the roofline denominator of the force kernel.
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the LJ path is bound by the FP64
// (or FP32) FMA pipe, so its peak is measured here with 8 independent FMA chains per thread.
#include <cuda_runtime.h>
#include <stdint.h>

#include "nm_b200.h"

extern int nm_fail_msg(int code, const char* fmt, ...);

namespace {
template <typename R>
__global__ void __launch_bounds__(256) k_fma_peak(R* out, int iters, R a, R b) {
  R c0 = (R)threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      c0 = c0 * a + b; c1 = c1 * a + b; c2 = c2 * a + b; c3 = c3 * a + b;
      c4 = c4 * a + b; c5 = c5 * a + b; c6 = c6 * a + b; c7 = c7 * a + b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
}

template <typename R>
int run_peak(int device, double* flops, double* ms_out) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return nm_fail_msg(NM_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, device);
  const int blocks = pr.multiProcessorCount * 8, threads = 256, iters = sizeof(R) == 8 ? 4096 : 8192;
  R* out = nullptr;
  if (cudaMalloc(&out, sizeof(R) * blocks * threads) != cudaSuccess) return nm_fail_msg(NM_ENOMEM, "cudaMalloc failed");
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a);
    k_fma_peak<R><<<blocks, threads>>>(out, iters, (R)0.999999, (R)1e-6);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  e = cudaGetLastError();
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  if (e != cudaSuccess) return nm_fail_msg(NM_ECUDA, "fma peak kernel: %s", cudaGetErrorString(e));
  const double fl = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
  if (flops) *flops = fl / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return NM_OK;
}
}  // namespace

extern "C" int nm_measure_fma_peak(int device, int precision, double* flops_per_s, double* ms) {
  int n = nm_device_count();
  if (n < 0) return n;
  if (device < 0 || device >= n) return nm_fail_msg(NM_ENODEV, "nm_measure_fma_peak: device %d not in [0,%d)", device, n);
  if (precision == 32) return run_peak<float>(device, flops_per_s, ms);
  if (precision == 64) return run_peak<double>(device, flops_per_s, ms);
  return nm_fail_msg(NM_EINVAL, "nm_measure_fma_peak: precision must be 32 or 64");
}
