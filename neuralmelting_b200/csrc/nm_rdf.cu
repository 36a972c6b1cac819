// nm_rdf.cu -- bit-exact radial-distribution histogram (calculate_rdf, lammps_distr.py:123-135).
//
// The reference forms, for each of the 27 image vectors s in {-1,0,1}^3 and every ordered pair
// (a, b), the float32 distance d = sqrt((dx^2 + dy^2) + dz^2) with d* = pos_a - (pos_b + box*s),
// every operation rounded separately (NumPy array passes), and np.histogram's it on float64
// edges. This kernel visits every ordered pair once, keeps per dimension only the shifts that can
// land inside the last edge (normally exactly one), forms those distances with the SAME float32
// operation sequence (no FMA contraction), and bins them with float32 thresholds that are exactly
// equivalent to the float64 edge comparisons. Thread-private histogram columns in shared memory
// (no atomics, no bank conflicts); integer / FP32 ALU bound, HBM traffic is 12 N bytes per sample.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "nm_b200.h"

namespace nmrdf {

constexpr int T = 256;

struct RdfParams {
  const float* pos; const float* box; uint32_t* counts;
  const float* thr;            // [2][nb]: float32 lower thresholds, (double)d >= edge[k]  <=>  d >= thr[k]; then the same
                               // thresholds on the SQUARED distance: fl(sqrt(s2)) >= thr[k]  <=>  s2 >= thr[nb + k] (exact: sqrt.rn is
                               // monotone, the host finds the smallest such float), thr[2 nb - 1] = +inf (last bin right-closed)
  int N, nb, a_chunk, private_hist;
  float t_top, cut, inv_dr;    // d <= t_top <=> (double)d <= edge[nb-1]; prune |delta| > cut
  float t2_top;                // fl(sqrt(s2)) <= t_top  <=>  s2 <= t2_top
};

__device__ __forceinline__ void count_one(float dx, float dy, float dz, const RdfParams& p, const float* sthr,
                                          uint32_t* hist) {
  const float s2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const float d = __fsqrt_rn(s2);
  const float t0 = sthr[0];
  if (d >= t0 && d <= p.t_top) {
    int k = (int)((d - t0) * p.inv_dr);
    k = max(0, min(k, p.nb - 1));
    while (k > 0 && d < sthr[k]) k--;
    while (k < p.nb - 1 && d >= sthr[k + 1]) k++;
    if (k == p.nb - 1) k = p.nb - 2;                       // last bin is right-closed
    if (p.private_hist) reinterpret_cast<unsigned short*>(hist)[k * T + threadIdx.x]++;
    else atomicAdd(&hist[k], 1u);
  }
}

__global__ void __launch_bounds__(T) k_rdf(RdfParams p) {
  extern __shared__ __align__(16) unsigned char sm[];
  float4* a4 = reinterpret_cast<float4*>(sm);
  float* sthr = reinterpret_cast<float*>(a4 + p.a_chunk);
  float* st2 = sthr + p.nb;
  uint32_t* hist = reinterpret_cast<uint32_t*>(st2 + p.nb);
  const int s = blockIdx.y, a0 = blockIdx.x * p.a_chunk, na = min(p.N, a0 + p.a_chunk) - a0;
  const float* ps = p.pos + (size_t)s * p.N * 3;
  for (int a = threadIdx.x; a < na; a += T) a4[a] = make_float4(ps[3 * (a0 + a)], ps[3 * (a0 + a) + 1], ps[3 * (a0 + a) + 2], 0.f);
  for (int k = threadIdx.x; k < 2 * p.nb; k += T) sthr[k] = p.thr[k];
  const int nh = p.private_hist ? p.nb * T / 2 : p.nb;     // private columns are 16-bit (two per word)
  for (int k = threadIdx.x; k < nh; k += T) hist[k] = 0u;
  __syncthreads();
  const float bx = p.box[s], nbx = __fmul_rn(bx, -1.0f), cut = p.cut;
  // FAST path (box >= 2 cut (1 + 1e-5), private columns): along an axis at most ONE of the three images can land inside the
  // last edge, the one nearest to the unshifted difference d0 = a - b (the others are at least box/2 - rounding > cut away),
  // so it is picked from d0 alone and only ITS difference is formed -- with the reference's own operations,
  // a - fl(b + fl(box * s)). Every pair then goes through the squared distance s2 (same three products and two sums as the
  // reference) and is binned ON s2 against thresholds that are exactly equivalent to the reference's tests on
  // fl(sqrt(s2)) (no sqrt.rn); the first guess of the bin comes from an approximate square root and is verified.
  if (p.private_hist && bx >= 2.0f * cut * (1.0f + 1e-5f)) {
    const float hb = 0.5f * bx, t0 = sthr[0], t2lo = st2[0], t2top = p.t2_top, inv_dr = p.inv_dr;
    const int kmax = p.nb - 2;
    unsigned short* col = reinterpret_cast<unsigned short*>(hist) + threadIdx.x;
    for (int b = threadIdx.x; b < p.N; b += T) {
      const float px = ps[3 * b], py = ps[3 * b + 1], pz = ps[3 * b + 2];
      const float xm = __fadd_rn(px, nbx), xp = __fadd_rn(px, bx);
      const float ym = __fadd_rn(py, nbx), yp = __fadd_rn(py, bx);
      const float zm = __fadd_rn(pz, nbx), zp = __fadd_rn(pz, bx);
#pragma unroll 2
      for (int a = 0; a < na; a++) {
        const float4 q = a4[a];
        const float dx0 = __fsub_rn(q.x, px), dy0 = __fsub_rn(q.y, py), dz0 = __fsub_rn(q.z, pz);
        const float xs = fabsf(dx0) > hb ? (dx0 > 0.f ? xp : xm) : px;
        const float ys = fabsf(dy0) > hb ? (dy0 > 0.f ? yp : ym) : py;
        const float zs = fabsf(dz0) > hb ? (dz0 > 0.f ? zp : zm) : pz;
        const float dx = __fsub_rn(q.x, xs), dy = __fsub_rn(q.y, ys), dz = __fsub_rn(q.z, zs);
        const float s2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (s2 >= t2lo && s2 <= t2top) {
          float dq; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(dq) : "f"(s2));
          int k = (int)((dq - t0) * inv_dr);
          k = max(0, min(k, kmax));
          if (s2 < st2[k] || s2 >= st2[k + 1]) {             // the guess is off (rounding at an edge, non-uniform edges)
            while (k > 0 && s2 < st2[k]) k--;
            while (k < kmax && s2 >= st2[k + 1]) k++;
          }
          col[k * T]++;
        }
      }
    }
  } else
  for (int b = threadIdx.x; b < p.N; b += T) {
    // the three images of atom b along each axis: pos_b + box*s, s = -1, 0, +1 (lammps_distr.py:130)
    const float px = ps[3 * b], py = ps[3 * b + 1], pz = ps[3 * b + 2];
    const float xm = __fadd_rn(px, nbx), xp = __fadd_rn(px, bx);
    const float ym = __fadd_rn(py, nbx), yp = __fadd_rn(py, bx);
    const float zm = __fadd_rn(pz, nbx), zp = __fadd_rn(pz, bx);
    for (int a = 0; a < na; a++) {
      const float qx = a4[a].x, qy = a4[a].y, qz = a4[a].z;
      const float dx0 = __fsub_rn(qx, px), dxm = __fsub_rn(qx, xm), dxp = __fsub_rn(qx, xp);
      const float dy0 = __fsub_rn(qy, py), dym = __fsub_rn(qy, ym), dyp = __fsub_rn(qy, yp);
      const float dz0 = __fsub_rn(qz, pz), dzm = __fsub_rn(qz, zm), dzp = __fsub_rn(qz, zp);
      const bool x0 = fabsf(dx0) <= cut, xmk = fabsf(dxm) <= cut, xpk = fabsf(dxp) <= cut;
      const bool y0 = fabsf(dy0) <= cut, ymk = fabsf(dym) <= cut, ypk = fabsf(dyp) <= cut;
      const bool z0 = fabsf(dz0) <= cut, zmk = fabsf(dzm) <= cut, zpk = fabsf(dzp) <= cut;
      const int cx = x0 + xmk + xpk, cy = y0 + ymk + ypk, cz = z0 + zmk + zpk;
      if (cx == 0 || cy == 0 || cz == 0) continue;
      if (cx == 1 && cy == 1 && cz == 1) {
        const float dx = x0 ? dx0 : (xmk ? dxm : dxp), dy = y0 ? dy0 : (ymk ? dym : dyp), dz = z0 ? dz0 : (zmk ? dzm : dzp);
        count_one(dx, dy, dz, p, sthr, hist);
      } else {                                              // rare: an axis with two admissible images
        const float vx[3] = { dxm, dx0, dxp }, vy[3] = { dym, dy0, dyp }, vz[3] = { dzm, dz0, dzp };
        const bool kx[3] = { xmk, x0, xpk }, ky[3] = { ymk, y0, ypk }, kz[3] = { zmk, z0, zpk };
        for (int i = 0; i < 3; i++) if (kx[i])
          for (int j = 0; j < 3; j++) if (ky[j])
            for (int k = 0; k < 3; k++) if (kz[k]) count_one(vx[i], vy[j], vz[k], p, sthr, hist);
      }
    }
  }
  __syncthreads();
  uint32_t* out = p.counts + (size_t)s * p.nb;
  if (p.private_hist) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = wid; k < p.nb - 1; k += T / 32) {
      uint32_t v = 0;
      for (int j = lane; j < T; j += 32) v += reinterpret_cast<unsigned short*>(hist)[k * T + j];
      v = __reduce_add_sync(0xffffffffu, v);
      if (lane == 0 && v) atomicAdd(&out[1 + k], v);
    }
  } else {
    for (int k = threadIdx.x; k < p.nb - 1; k += T) if (hist[k]) atomicAdd(&out[1 + k], hist[k]);
  }
}

}  // namespace nmrdf


extern "C" const char* nm_last_error(void);
// error text is routed through the engine's thread-local buffer
extern int nm_fail_msg(int code, const char* fmt, ...);

#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = nm_fail_msg(NM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); goto done; } } while (0)

// Scratch comes from the device's stream-ordered pool with the release threshold lifted: after the first call the
// allocations are served from the pool (no cudaMalloc / cudaFree, no device-wide synchronisation per call).
static cudaError_t keep_pool(int device) {
  static bool done[64] = { false };
  if (device < 0 || device >= 64 || done[device]) return cudaSuccess;
  cudaMemPool_t pool;
  cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
  if (e != cudaSuccess) return e;
  unsigned long long keep = ~0ull;
  e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  if (e == cudaSuccess) done[device] = true;
  return e;
}

extern "C" int nm_rdf_counts(int device, void* cuda_stream, int dev_ptrs, const float* pos, const float* box,
                             int32_t natoms, int64_t nsamples, const double* edges, int32_t nbins, uint32_t* counts) {
  using namespace nmrdf;
  if (!pos || !box || !edges || !counts) return nm_fail_msg(NM_EINVAL, "nm_rdf_counts: null argument");
  if (natoms < 1 || nsamples < 0 || nbins < 2 || nbins > 4096) return nm_fail_msg(NM_EINVAL, "nm_rdf_counts: bad sizes (natoms=%d nsamples=%lld nbins=%d)", natoms, (long long)nsamples, nbins);
  for (int k = 1; k < nbins; k++) if (!(edges[k] > edges[k - 1])) return nm_fail_msg(NM_EINVAL, "nm_rdf_counts: edges must increase strictly");
  if (nsamples == 0) return NM_OK;
  int ndev = nm_device_count();
  if (ndev < 0) return ndev;
  if (device < 0 || device >= ndev) return nm_fail_msg(NM_ENODEV, "nm_rdf_counts: device %d not in [0,%d)", device, ndev);
  int rc = NM_OK;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  float *d_pos = nullptr, *d_box = nullptr, *d_thr = nullptr; uint32_t* d_cnt = nullptr;
  std::vector<float> thr(2 * (size_t)nbins);
  // float32 thresholds exactly equivalent to the float64 edge tests of np.histogram
  for (int k = 0; k < nbins; k++) {
    float f = (float)edges[k];
    if ((double)f < edges[k]) f = nextafterf(f, INFINITY);
    thr[k] = f;
  }
  float t_top = (float)edges[nbins - 1];
  if ((double)t_top > edges[nbins - 1]) t_top = nextafterf(t_top, -INFINITY);
  // ... and on the squared distance: sqrtf is correctly rounded, hence monotone, so "fl(sqrt(s2)) >= t" holds from one float on
  // (smallest s2 found by stepping around t*t), and "fl(sqrt(s2)) <= t_top" up to one float
  auto t2_ge = [](float t) -> float {
    if (!(t > 0.0f)) return 0.0f;
    float s2 = t * t;
    while (s2 > 0.0f && sqrtf(nextafterf(s2, -INFINITY)) >= t) s2 = nextafterf(s2, -INFINITY);
    while (sqrtf(s2) < t) s2 = nextafterf(s2, INFINITY);
    return s2;
  };
  for (int k = 0; k < nbins - 1; k++) thr[nbins + k] = t2_ge(thr[k]);
  thr[2 * (size_t)nbins - 1] = INFINITY;
  float t2_top = t_top > 0.0f ? t_top * t_top : 0.0f;
  if (t_top > 0.0f) {
    while (sqrtf(nextafterf(t2_top, INFINITY)) <= t_top) t2_top = nextafterf(t2_top, INFINITY);
    while (t2_top > 0.0f && sqrtf(t2_top) > t_top) t2_top = nextafterf(t2_top, -INFINITY);
  }
  RdfParams p;
  p.N = natoms; p.nb = nbins; p.t_top = t_top; p.t2_top = t2_top;
  p.cut = t_top * (1.0f + 4e-6f) + 1e-30f;
  p.inv_dr = (float)((nbins - 1) / (edges[nbins - 1] - edges[0]));
  {
    RCK(cudaSetDevice(device));
    RCK(keep_pool(device));
    const size_t nposb = sizeof(float) * 3 * (size_t)natoms * nsamples, ncntb = sizeof(uint32_t) * (size_t)nbins * nsamples;
    RCK(cudaMallocAsync(&d_thr, sizeof(float) * thr.size(), st));
    RCK(cudaMemcpyAsync(d_thr, thr.data(), sizeof(float) * thr.size(), cudaMemcpyHostToDevice, st));
    if (dev_ptrs) { d_pos = const_cast<float*>(pos); d_box = const_cast<float*>(box); d_cnt = counts; }
    else {
      RCK(cudaMallocAsync(&d_pos, nposb, st)); RCK(cudaMallocAsync(&d_box, sizeof(float) * nsamples, st)); RCK(cudaMallocAsync(&d_cnt, ncntb, st));
      RCK(cudaMemcpyAsync(d_pos, pos, nposb, cudaMemcpyHostToDevice, st));
      RCK(cudaMemcpyAsync(d_box, box, sizeof(float) * nsamples, cudaMemcpyHostToDevice, st));
    }
    RCK(cudaMemsetAsync(d_cnt, 0, ncntb, st));
    // split the 'a' range of a sample over several CTAs when there are too few samples to fill the GPU
    int nsplit = 1;
    if (nsamples < 2 * 148) nsplit = (int)((2 * 148 + nsamples - 1) / nsamples);
    int a_chunk = (natoms + nsplit - 1) / nsplit;
    if (a_chunk < 64) a_chunk = natoms < 64 ? natoms : 64;
    if (a_chunk > 1024) a_chunk = 1024;      // 16 KB of positions + 32 KB of 16-bit private histogram columns: four CTAs per SM
    // a thread counts at most a_chunk * ceil(natoms / T) pairs: keep that inside 16 bits
    { const int per_thread = (natoms + T - 1) / T; if ((long long)a_chunk * per_thread > 65535) a_chunk = 65535 / per_thread; if (a_chunk < 1) a_chunk = 1; }
    nsplit = (natoms + a_chunk - 1) / a_chunk;
    p.a_chunk = a_chunk;
    const size_t base = sizeof(float) * (4 * (size_t)a_chunk + 2 * (size_t)nbins);
    p.private_hist = (base + sizeof(unsigned short) * (size_t)nbins * T) <= 200 * 1024 && (nbins * T) % 2 == 0;
    const size_t smem = base + (p.private_hist ? sizeof(unsigned short) * (size_t)nbins * T : sizeof(uint32_t) * (size_t)nbins);
    RCK(cudaFuncSetAttribute(k_rdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.pos = d_pos; p.box = d_box; p.counts = d_cnt; p.thr = d_thr;
    for (int64_t s0 = 0; s0 < nsamples; s0 += 65535) {       // gridDim.y limit
      const int ns = (int)((nsamples - s0) < 65535 ? (nsamples - s0) : 65535);
      RdfParams q = p;
      q.pos = d_pos + 3 * (size_t)natoms * s0; q.box = d_box + s0; q.counts = d_cnt + (size_t)nbins * s0;
      k_rdf<<<dim3(nsplit, ns), T, smem, st>>>(q);
      RCK(cudaGetLastError());
    }
    if (!dev_ptrs) {
      RCK(cudaMemcpyAsync(counts, d_cnt, ncntb, cudaMemcpyDeviceToHost, st));
      RCK(cudaStreamSynchronize(st));
    }
  }
done:
  if (rc != NM_OK) cudaStreamSynchronize(st);
  else if (!dev_ptrs) { /* already synchronised */ }
  if (d_thr) cudaFreeAsync(d_thr, st);              // stream-ordered: after the kernels that read it
  if (!dev_ptrs) { if (d_pos) cudaFreeAsync(d_pos, st); if (d_box) cudaFreeAsync(d_box, st); if (d_cnt) cudaFreeAsync(d_cnt, st); }

  return rc;
}

// =================================================================== N1: Cartesian pair-vector density
// calculate_cdf (lammps_distr.py:161-171): for each of the 27 image vectors, np.histogramdd of the float32 pair
// vectors dvm[b, a, :] = pos_a - (pos_b + box*s) on the float64 edges rv (three rows of nb+1 edges). Bit-exact:
// same float32 operation sequence as the RDF kernel, float32 thresholds exactly equivalent to the float64 edge
// tests of np.searchsorted(side='right') with the right-most edge closed, outliers dropped.
namespace nmcdf {

constexpr int T = 256;

struct CdfParams {
  const float* pos; const float* box; uint32_t* counts;
  const float* thr;            // [3][nb+1] lower thresholds, thr[d][k]: (double)x >= edge[d][k]  <=>  x >= thr[d][k]
  int N, nb, a_chunk;
  float t_top[3], inv_d[3];    // x <= t_top[d] <=> (double)x <= edge[d][nb]
};

// bin of x along one axis, or -1 when outside [edge[0], edge[nb]]
__device__ __forceinline__ int bin1(float x, const float* t, int nb, float t_top, float inv_d) {
  if (!(x >= t[0] && x <= t_top)) return -1;
  int k = (int)((x - t[0]) * inv_d);
  k = max(0, min(k, nb));
  while (k > 0 && x < t[k]) k--;
  while (k < nb && x >= t[k + 1]) k++;
  return k == nb ? nb - 1 : k;       // the right-most edge belongs to the last bin
}

__global__ void __launch_bounds__(T) k_cdf(CdfParams p) {
  extern __shared__ __align__(16) unsigned char sm[];
  float* ax = reinterpret_cast<float*>(sm);
  float* ay = ax + p.a_chunk; float* az = ay + p.a_chunk;
  float* sthr = az + p.a_chunk;                       // 3 * (nb + 1)
  uint32_t* hist = reinterpret_cast<uint32_t*>(sthr + 3 * (p.nb + 1));
  const int nb = p.nb, nb3 = nb * nb * nb;
  const int s = blockIdx.y, a0 = blockIdx.x * p.a_chunk, na = min(p.N, a0 + p.a_chunk) - a0;
  const float* ps = p.pos + (size_t)s * p.N * 3;
  for (int a = threadIdx.x; a < na; a += T) { ax[a] = ps[3 * (a0 + a)]; ay[a] = ps[3 * (a0 + a) + 1]; az[a] = ps[3 * (a0 + a) + 2]; }
  for (int k = threadIdx.x; k < 3 * (nb + 1); k += T) sthr[k] = p.thr[k];
  for (int k = threadIdx.x; k < nb3; k += T) hist[k] = 0u;
  __syncthreads();
  const float* tx = sthr; const float* ty = sthr + (nb + 1); const float* tz = sthr + 2 * (nb + 1);
  const float bx = p.box[s], nbx = __fmul_rn(bx, -1.0f);
  for (int b = threadIdx.x; b < p.N; b += T) {
    const float px = ps[3 * b], py = ps[3 * b + 1], pz = ps[3 * b + 2];
    const float im[3][3] = { { __fadd_rn(px, nbx), px, __fadd_rn(px, bx) }, { __fadd_rn(py, nbx), py, __fadd_rn(py, bx) },
                             { __fadd_rn(pz, nbx), pz, __fadd_rn(pz, bx) } };
    for (int a = 0; a < na; a++) {
      const float q[3] = { ax[a], ay[a], az[a] };
      int bins[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        bins[0][i] = bin1(__fsub_rn(q[0], im[0][i]), tx, nb, p.t_top[0], p.inv_d[0]);
        bins[1][i] = bin1(__fsub_rn(q[1], im[1][i]), ty, nb, p.t_top[1], p.inv_d[1]);
        bins[2][i] = bin1(__fsub_rn(q[2], im[2][i]), tz, nb, p.t_top[2], p.inv_d[2]);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) if (bins[0][i] >= 0)
#pragma unroll
        for (int j = 0; j < 3; j++) if (bins[1][j] >= 0)
#pragma unroll
          for (int k = 0; k < 3; k++) if (bins[2][k] >= 0)
            atomicAdd(&hist[(bins[0][i] * nb + bins[1][j]) * nb + bins[2][k]], 1u);
    }
  }
  __syncthreads();
  uint32_t* out = p.counts + (size_t)s * nb3;
  for (int k = threadIdx.x; k < nb3; k += T) if (hist[k]) atomicAdd(&out[k], hist[k]);
}

}  // namespace nmcdf

extern "C" int nm_cdf_counts(int device, void* cuda_stream, int dev_ptrs, const float* pos, const float* box,
                             int32_t natoms, int64_t nsamples, const double* edges, int32_t nb, uint32_t* counts) {
  using namespace nmcdf;
  if (!pos || !box || !edges || !counts) return nm_fail_msg(NM_EINVAL, "nm_cdf_counts: null argument");
  if (natoms < 1 || nsamples < 0 || nb < 1 || nb > 32) return nm_fail_msg(NM_EINVAL, "nm_cdf_counts: bad sizes (natoms=%d nsamples=%lld bins=%d; at most 32 bins per axis)", natoms, (long long)nsamples, nb);
  for (int d = 0; d < 3; d++) for (int k = 1; k <= nb; k++)
    if (!(edges[d * (nb + 1) + k] > edges[d * (nb + 1) + k - 1])) return nm_fail_msg(NM_EINVAL, "nm_cdf_counts: edges must increase strictly");
  if (nsamples == 0) return NM_OK;
  int ndev = nm_device_count();
  if (ndev < 0) return ndev;
  if (device < 0 || device >= ndev) return nm_fail_msg(NM_ENODEV, "nm_cdf_counts: device %d not in [0,%d)", device, ndev);
  int rc = NM_OK;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  float *d_pos = nullptr, *d_box = nullptr, *d_thr = nullptr; uint32_t* d_cnt = nullptr;
  std::vector<float> thr(3 * (nb + 1));
  CdfParams p;
  p.N = natoms; p.nb = nb;
  for (int d = 0; d < 3; d++) {
    const double* e = edges + d * (nb + 1);
    for (int k = 0; k <= nb; k++) {
      float f = (float)e[k];
      if ((double)f < e[k]) f = nextafterf(f, INFINITY);
      thr[d * (nb + 1) + k] = f;
    }
    float top = (float)e[nb];
    if ((double)top > e[nb]) top = nextafterf(top, -INFINITY);
    p.t_top[d] = top;
    p.inv_d[d] = (float)(nb / (e[nb] - e[0]));
  }
  {
    RCK(cudaSetDevice(device));
    RCK(keep_pool(device));
    const size_t nb3 = (size_t)nb * nb * nb;
    const size_t nposb = sizeof(float) * 3 * (size_t)natoms * nsamples, ncntb = sizeof(uint32_t) * nb3 * nsamples;
    RCK(cudaMallocAsync(&d_thr, sizeof(float) * thr.size(), st));
    RCK(cudaMemcpyAsync(d_thr, thr.data(), sizeof(float) * thr.size(), cudaMemcpyHostToDevice, st));
    if (dev_ptrs) { d_pos = const_cast<float*>(pos); d_box = const_cast<float*>(box); d_cnt = counts; }
    else {
      RCK(cudaMallocAsync(&d_pos, nposb, st)); RCK(cudaMallocAsync(&d_box, sizeof(float) * nsamples, st)); RCK(cudaMallocAsync(&d_cnt, ncntb, st));
      RCK(cudaMemcpyAsync(d_pos, pos, nposb, cudaMemcpyHostToDevice, st));
      RCK(cudaMemcpyAsync(d_box, box, sizeof(float) * nsamples, cudaMemcpyHostToDevice, st));
    }
    RCK(cudaMemsetAsync(d_cnt, 0, ncntb, st));
    int nsplit = 1;
    if (nsamples < 2 * 148) nsplit = (int)((2 * 148 + nsamples - 1) / nsamples);
    int a_chunk = (natoms + nsplit - 1) / nsplit;
    if (a_chunk < 64) a_chunk = natoms < 64 ? natoms : 64;
    if (a_chunk > 2048) a_chunk = 2048;
    nsplit = (natoms + a_chunk - 1) / a_chunk;
    p.a_chunk = a_chunk;
    const size_t smem = sizeof(float) * (3 * (size_t)a_chunk + 3 * (nb + 1)) + sizeof(uint32_t) * nb3;
    RCK(cudaFuncSetAttribute(k_cdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.thr = d_thr;
    for (int64_t s0 = 0; s0 < nsamples; s0 += 65535) {
      const int ns = (int)((nsamples - s0) < 65535 ? (nsamples - s0) : 65535);
      CdfParams q = p;
      q.pos = d_pos + 3 * (size_t)natoms * s0; q.box = d_box + s0; q.counts = d_cnt + nb3 * s0;
      k_cdf<<<dim3(nsplit, ns), T, smem, st>>>(q);
      RCK(cudaGetLastError());
    }
    if (!dev_ptrs) {
      RCK(cudaMemcpyAsync(counts, d_cnt, ncntb, cudaMemcpyDeviceToHost, st));
      RCK(cudaStreamSynchronize(st));
    }
  }
done:
  if (rc != NM_OK) cudaStreamSynchronize(st);
  if (d_thr) cudaFreeAsync(d_thr, st);              // stream-ordered: after the kernels that read it
  if (!dev_ptrs) { if (d_pos) cudaFreeAsync(d_pos, st); if (d_box) cudaFreeAsync(d_box, st); if (d_cnt) cudaFreeAsync(d_cnt, st); }
  return rc;
}
