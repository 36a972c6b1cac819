// micro-benchmarks for the FP64 path on B200: dependent latency and throughput vs warps x ILP
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_dfma(double* out, long long* cyc, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 4096; i++) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_dadd(double* out, long long* cyc, double a) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 4096; i++) x = x + a;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_rcp(double* out, long long* cyc) {
  double x = 1.5 + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 4096; i++) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_lds(double* out, long long* cyc) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 4096; i++) j = idx[j];
  long long t1 = clock64();
  out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// rcp seed accuracy
__global__ void rcp_acc(double* maxerr) {
  double worst = 0;
  for (int i = 0; i < 100000; i++) {
    double x = 0.5 + (i + threadIdx.x * 100000.0) * (5.75 / (100000.0 * 32));
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fabs(x * y - 1.0);
    if (e > worst) worst = e;
  }
  maxerr[threadIdx.x] = worst;
}
template <int ILP>
__global__ void thr_dfma(double* out, int iters, double a, double b) {
  double c[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) c[k] = threadIdx.x + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int k = 0; k < ILP; k++) c[k] = fma(c[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: DP ops with integer ops interleaved (2 int per DP) to see co-issue
template <int ILP>
__global__ void thr_mixed(double* out, int iters, double a, double b) {
  double c[ILP]; int m[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) { c[k] = threadIdx.x + k; m[k] = threadIdx.x + k; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int k = 0; k < ILP; k++) { c[k] = fma(c[k], a, b); m[k] = (m[k] ^ 0x5bd1e995) + (m[k] >> 3); }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += c[k] + m[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 64);
  long long h;
  lat_dfma<<<1, 32>>>(out, cyc, 0.999, 1e-3); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency %.2f clk\n", h / 4096.0);
  lat_dadd<<<1, 32>>>(out, cyc, 1e-3); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DADD dependent latency %.2f clk\n", h / 4096.0);
  lat_rcp<<<1, 32>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("MUFU.RCP64H dependent latency %.2f clk\n", h / 4096.0);
  lat_lds<<<1, 32>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("LDS dependent latency %.2f clk\n", h / 4096.0);
  double* me; cudaMalloc(&me, 32 * 8); rcp_acc<<<1, 32>>>(me); double hm[32]; cudaMemcpy(hm, me, 256, cudaMemcpyDeviceToHost);
  double w = 0; for (int i = 0; i < 32; i++) if (hm[i] > w) w = hm[i]; printf("rcp.approx.ftz.f64 max rel err %.3e (2^%.1f)\n", w, log2(w));
  int dev_sms = 148;
  for (int warps : {1, 2, 4, 8, 16}) {
    int iters = 2048;
    auto run = [&](int ilp) {
      float ms;
      if (ilp == 1) ms = timeit([&] { thr_dfma<1><<<dev_sms, warps * 32>>>(out, iters, 0.999, 1e-3); });
      else if (ilp == 2) ms = timeit([&] { thr_dfma<2><<<dev_sms, warps * 32>>>(out, iters, 0.999, 1e-3); });
      else if (ilp == 4) ms = timeit([&] { thr_dfma<4><<<dev_sms, warps * 32>>>(out, iters, 0.999, 1e-3); });
      else ms = timeit([&] { thr_dfma<8><<<dev_sms, warps * 32>>>(out, iters, 0.999, 1e-3); });
      double fl = 2.0 * 8 * ilp * (double)iters * dev_sms * warps * 32;
      printf("  warps/SM %2d ILP %d: %.2f TFLOP/s\n", warps, ilp, fl / ms / 1e9);
    };
    for (int ilp : {1, 2, 4, 8}) run(ilp);
  }
  for (int warps : {4, 16}) {
    float ms = timeit([&] { thr_mixed<4><<<148, warps * 32>>>(out, 2048, 0.999, 1e-3); });
    double fl = 2.0 * 8 * 4 * 2048.0 * 148 * warps * 32;
    printf("  mixed (1 DFMA + 3 int) warps/SM %2d ILP 4: %.2f TFLOP/s DP\n", warps, fl / ms / 1e9);
  }
  return 0;
}
