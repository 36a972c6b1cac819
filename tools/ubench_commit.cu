// isolates the ordered-commit loop of iter_position_mc (one warp, serial chain) to find what an iteration costs
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int WS = 10;
extern __shared__ double dsm[];
template <int VARIANT>
__device__ __noinline__ void commit_window(const double* win, const double* corr, double* sp, int CS, int nwin, int k, double s, double rl, double rcg,
                                           double et, int* res, double* resd) {
  const int lane = threadIdx.x & 31;
  const bool mine = lane < nwin;
  double de = win[WS * (mine ? lane : 0) + 3];
  double umax = resd[0];
  int t = 0, need = 0, na = 0;
  double c_next = mine ? corr[lane] : 0.0;
  for (; t < nwin; t++) {
    const double* wt = win + WS * t;
    const int fl = reinterpret_cast<const int*>(wt + 8)[0];
    if (fl) { need = fl == 1; break; }
    const double un_t = wt[6];
    if (VARIANT != 1 && t > 0) {
      const bool ok = s * (rl - un_t - umax) >= rcg && s * (rl - 2.0 * umax) >= rcg;
      if (!ok) break;
    }
    const double de_t = __shfl_sync(0xffffffffu, de, t), thr_t = wt[4];
    bool acc;
    if (VARIANT == 2) acc = de_t < thr_t;
    else {
      const double gap = 1e-9 * (fabs(thr_t) + fabs(de_t)) + 1e-290;
      if (de_t > -600.0 * et && de_t < thr_t - gap) acc = true;
      else if (de_t > -600.0 * et && de_t > thr_t + gap) acc = false;
      else {
        const double m = exp(-(de_t / et)), uacc = wt[5];
        acc = !(isinf(m) || isnan(m)) && uacc <= (m < 1.0 ? m : 1.0);
      }
    }
    const double c_t = c_next;
    if (t + 1 < nwin) c_next = mine ? corr[(t + 1) * CS + lane] : 0.0;
    if (acc) {
      if (lane == t) { const int i = k + t; sp[3 * i] = wt[0]; sp[3 * i + 1] = wt[1]; sp[3 * i + 2] = wt[2]; }
      umax = fmax(umax, un_t);
      na++;
      if (lane > t) de += c_t;
    }
  }
  __syncwarp();
  if (lane == 0) { res[0] = k + t; res[1] = need; resd[0] = umax; resd[2] += (double)t; resd[3] += (double)na; }
}
template <int VARIANT>
__global__ void run(long long* cyc, double* out, double pacc) {
  double* win = dsm; double* corr = dsm + 512; double* sp = dsm + 2048; double* resd = dsm + 1800; int* res = (int*)(dsm + 1900);
  for (int i = threadIdx.x; i < 32; i += blockDim.x) {
    double* t = win + WS * i;
    t[0] = i; t[1] = i; t[2] = i; t[3] = (i * 37 % 11) * 0.1 - 0.5; t[4] = pacc; t[5] = 0.5; t[6] = 0.01 * (i % 5); t[7] = 0;
    ((int*)(t + 8))[0] = 0; ((int*)(t + 8))[1] = 3; ((int*)(t + 9))[0] = 0;
  }
  for (int i = threadIdx.x; i < 33 * 32; i += blockDim.x) corr[i] = 1e-3 * (i % 7);
  if (threadIdx.x < 8) resd[threadIdx.x] = 0;
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    t0 = clock64();
    for (int rep = 0; rep < 16; rep++) commit_window<VARIANT>(win, corr, sp, 33, 32, 0, 1.0, 2.8, 2.5, 1.0, res, resd);
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = resd[2]; out[1] = resd[3]; }
}
int main() {
  long long* cyc; double* out; cudaMalloc(&cyc, 64); cudaMalloc(&out, 64);
  long long h; double ho[2];
  cudaFuncSetAttribute(run<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(run<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(run<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (double pacc : {-10.0, 0.0, 10.0})
    for (int thr : {32, 1024}) {
      run<0><<<1, thr, 100 * 1024>>>(cyc, out, pacc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
      printf("full      thr %4d thr %.0f: %.1f clk/trial (trials %.0f acc %.0f)\n", thr, pacc, h / ho[0], ho[0], ho[1]);
      run<1><<<1, thr, 100 * 1024>>>(cyc, out, pacc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
      printf("no-reval  thr %4d thr %.0f: %.1f clk/trial (trials %.0f acc %.0f)\n", thr, pacc, h / ho[0], ho[0], ho[1]);
      run<2><<<1, thr, 100 * 1024>>>(cyc, out, pacc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
      printf("plain-cmp thr %4d thr %.0f: %.1f clk/trial (trials %.0f acc %.0f)\n", thr, pacc, h / ho[0], ho[0], ho[1]);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
