// nm_device.cuh -- device-side building blocks shared by the engine kernels (sm_100a).
//
// Layout in HBM (per configuration c, N atoms, Npad = roundup(N+1, 32)):
//   x, v, f, xs, vs, fs : double[3][Npad]   SoA (x-block, y-block, z-block); *s = saved copy for reverts
//   x0                  : double[3][Npad]   fractional coordinates at the last Verlet-list build
//   list                : ushort4[maxq][Npad]  full neighbour list, 4 neighbours per 8-byte word,
//                                             column i = atom i (coalesced across a warp)
//   nnb                 : uint16[Npad]      neighbour count of atom i
// On chip (per CTA = one configuration): positions double[3][Npad] in shared memory (the only
// randomly gathered array); velocities / forces stream through the owning thread.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nm {

// ---------------------------------------------------------------- RNG (same convention as oracle/nm_oracle.c)
enum { P_ROLL = 0, P_HMC_VEL = 1, P_HMC_ACC = 2, P_VMC_PROP = 3, P_VMC_ACC = 4,
       P_BULK_DISP = 5, P_BULK_ACC = 6, P_ITER_DISP = 7, P_ITER_ACC = 8, P_EXCH = 9 };
#define NM_EXCH_KEY 0xE8C4A93Bu

struct Rng { uint32_t k0, k1, m_lo, m_hi; };

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                              uint32_t c2, uint32_t c3, uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double u53_open(uint32_t hi, uint32_t lo) {
  return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ Rng rng_make(uint32_t seed_lo, uint32_t seed_hi, uint32_t slot, uint64_t m) {
  Rng r = { seed_lo, seed_hi ^ slot, (uint32_t)m, (uint32_t)(m >> 32) };
  return r;
}
__device__ __forceinline__ double rng_uniform(const Rng& r, uint32_t index, uint32_t purpose) {
  uint32_t w[4]; philox4x32_10(r.k0, r.k1, index, purpose, r.m_lo, r.m_hi, w);
  return u53(w[0], w[1]);
}
__device__ __forceinline__ void rng_uniform3(const Rng& r, uint32_t i, uint32_t purpose, double (&u)[3]) {
  uint32_t a[4], b[4];
  philox4x32_10(r.k0, r.k1, 2 * i, purpose, r.m_lo, r.m_hi, a);
  philox4x32_10(r.k0, r.k1, 2 * i + 1, purpose, r.m_lo, r.m_hi, b);
  u[0] = u53(a[0], a[1]); u[1] = u53(a[2], a[3]); u[2] = u53(b[0], b[1]);
}
__device__ __forceinline__ void rng_gauss3(const Rng& r, uint32_t i, uint32_t purpose, double (&g)[3]) {
  uint32_t a[4], b[4];
  philox4x32_10(r.k0, r.k1, 2 * i, purpose, r.m_lo, r.m_hi, a);
  philox4x32_10(r.k0, r.k1, 2 * i + 1, purpose, r.m_lo, r.m_hi, b);
  double r0 = sqrt(-2.0 * log(u53_open(a[0], a[1]))), t0 = 6.283185307179586477 * u53(a[2], a[3]);
  double r1 = sqrt(-2.0 * log(u53_open(b[0], b[1]))), t1 = 6.283185307179586477 * u53(b[2], b[3]);
  double s0, c0, s1, c1;
  sincos(t0, &s0, &c0); sincos(t1, &s1, &c1);
  g[0] = r0 * c0; g[1] = r0 * s0; g[2] = r1 * c1;
}

// ---------------------------------------------------------------- small math
// '%f' text round trip of the values the reference passes to LAMMPS as strings
// = strtod(sprintf("%f", x)): correctly rounded to 6 decimals, ties (on the EXACT binary value) to even.
// x*1e6 = p + err exactly (FMA residual); only when p sits exactly on a half does err decide.
__device__ __forceinline__ double round6(double x) {
  const double p = x * 1e6, err = fma(x, 1e6, -p);
  double k = rint(p);
  const double frac = p - k;
  if (frac == 0.5 && err > 0.0) k += 1.0;
  else if (frac == -0.5 && err < 0.0) k -= 1.0;
  return k / 1e6;
}

// LAMMPS remap into [0, L)
__device__ __forceinline__ double wrap1(double x, double L) {
  if (x < 0.0) x += L;
  if (x >= L) x -= L;
  if (x < 0.0) x = 0.0;
  return x;
}
// general re-wrap of a coordinate that may be any number of boxes away
__device__ __forceinline__ double wrapg(double x, double L) { return wrap1(x - floor(x / L) * L, L); }
// exact minimum image for wrapped coordinates (|d| < L)
__device__ __forceinline__ double mic_exact(double d, double L, double hL) {
  if (d > hL) d -= L; else if (d < -hL) d += L;
  return d;
}
// minimum image with the |d| > L/2 test done on the high 32 bits (integer pipe, not the FP64 pipe).
// Inexact only for |d| within 2^-20 of L/2, where both images lie beyond the cutoff (L >= 2 rc (1+1e-5)).
__device__ __forceinline__ double mic_fast(double d, int L_hi, int L_lo, int hL_hi) {
  int hi = __double2hiint(d);
  if ((hi & 0x7fffffff) > hL_hi) d -= __hiloint2double(L_hi | (hi & 0x80000000), L_lo);
  return d;
}
// FP64 reciprocal: MUFU.RCP64H seed + Newton steps on the FMA pipe (no slow-path branch).
// rsq is finite and positive here; 0 gives inf as IEEE division would.
__device__ __forceinline__ double rcp_nr(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// ---------------------------------------------------------------- block reductions (deterministic order)
// red: shared scratch of 32*K doubles that no thread is still reading (callers alternate between two scratch areas,
// so one barrier per reduction suffices). Warp butterflies, one partial per warp, then every warp reduces the
// partials with the same butterfly: all threads get bitwise identical totals, independent of scheduling.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) red[k * 32 + wid] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    double p = lane < nw ? red[k * 32 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    v[k] = p;
  }
}

}  // namespace nm
