// nm_engine.cu -- replica-exchange NPT Monte Carlo engine for sm_100a (B200), C-ABI in include/nm_b200.h.
//
// One CTA owns one replica configuration for a whole collection cycle (MOD moves): positions
// stay in shared memory, the LJ lj/cut evaluation runs off a Verlet list held in HBM/L2, all
// NSTPS velocity-Verlet steps of an HMC trajectory, the Metropolis tests and the counters
// happen on chip. Replaces the per-replica LAMMPS instance of lammps_remcmc.py:459-691.
// FP64 FMA pipe is the roofline (no tensor cores: not a dense contraction).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "nm_b200.h"
#include "nm_device.cuh"

// Helper-capable variant of the 1024-thread cycle kernel (see help_request). It is compiled as a kernel of its own
// (translation unit NM_TU=1025, symbol k_cycle_h) so that the plain kernels keep their register allocation: with the
// helper code inlined into the one k_cycle<1024>, ptxas' allocation of the ordered-commit loop of the iterative sweeps
// changed and C4 ran 19 % slower. The host launches k_cycle_h only when helpers are enabled (d.nhelp > 0).
#if defined(NM_TU) && NM_TU == 1025
#define NM_HELPERS 1
#define k_cycle k_cycle_h
#elif defined(NM_TU)
#define NM_HELPERS 0
#else
#define NM_HELPERS 1
#endif
#ifndef NM_BUILD_COST
#define NM_BUILD_COST 0.25      // SMALL-mode list build in units of one listed-pair evaluation, / N^2 (97 k clocks against 1.57 per pair at N = 500)
#endif
#ifndef NM_UNR
#define NM_UNR 1
#endif
#ifndef NM_CTAS_PER_SM
#define NM_CTAS_PER_SM(T) (1024 / (T))
#endif

namespace nm {

constexpr int LIST_SPARE_ROWS = 4;       // rows past the last quad of a list buffer that the force loop may prefetch
constexpr int NCMAX = 10;                // cell grid is at most 10^3 (N = 4000 must stay below the 196 KB shared-memory carve-out: L1 keeps 60 KB)
constexpr int RED_HALF = 32 * 9;           // one block_sum scratch area (K <= 9)
constexpr int RED_DOUBLES = 2 * RED_HALF;
constexpr int BC_DOUBLES = 32;
constexpr int SHT_DOUBLES = 84;          // 27 x 3 image shifts (+ padding)
constexpr int HELP_STRIDE = 16;          // ints per helper record: [0] helper attached, [1] last command issued (-1: the chain is finished),
                                         // [2] last command completed, [3] flags (1 energy / virial, 2 fused kick, 4 per-pair minimum image,
                                         // 8 inner list rows, 16 outer list rows), [4] list buffer, [5] claimed by a helper, [6..7] in-cutoff
                                         // ordered pairs of the helper's rows (force-only commands), [8] the owner's estimate of its remaining
                                         // clocks / 1024, [9..10] cell grid of an outer build (cells per axis, stencil half width)
constexpr int NSMALL = 768;              // largest N handled by the all-pairs hit-matrix build (SMALL mode: one atom per thread)
// SMALL mode resolves periodic images with GHOST atoms: the shared position array is extended by the shifted copies of
// the atoms within the list radius of a box face (up to 7 per atom), and the list stores the index of the copy to use.
// Capacity: expected (1 + 2 r/L)^3 - 1 <= 5.45 copies per atom while r/L < 0.43; beyond the capacity (or in boxes
// below 2 r_list) the build falls back to the per-pair minimum image.
__host__ __device__ inline int ghost_cap(int N) { return ((int)(5.6 * N) + 33) & ~1; }
constexpr int ST_BOX = 1, ST_NEIGH = 2;  // status bits

// ------------------------------------------------------------------ device-side engine description
struct Dev {
  int N, Npad, nrep, nrep_global, rep_offset, nt, maxq, maxqo, maxnbo;   // inner / outer list capacity in quads, outer scratch entries
  int row0, row_stride;                    // local pressure row lr is global row row0 + lr * row_stride (row u -> rank u mod G: row0 = rank, stride = G)
  int nstps, mod, bulk, text_rounding;
  int nsm, per_sm;                         // SM count and CTAs that fit per SM (cost-balanced placement)
  int f32;                                 // precision = 32: pair arithmetic in FP32 on the fractional float copies
  int small;                               // 1: N <= NSMALL: single-level list built from an all-pairs hit matrix in shared memory
  double ppos, pvol, lat, mass, rc, skin, oskin;
  uint32_t seed_lo, seed_hi;
  // per configuration
  double *x, *v, *f, *xs, *vs, *fs;        // [nrep][3][Npad]
  double* x0;                              // [nrep][2][3][Npad]  fractional coordinates at the build of list buffer 0 / 1
  ushort4* list;                           // [nrep][2][maxq + 1][Npad]  neighbour quads, grouped by periodic image; TWO buffers per
                                           // configuration: a build inside a move goes to the other buffer, so a rejected move
                                           // switches back to the list that was valid for the saved positions
  int* lcur;                               // [nrep] which buffer holds the current list
  uint32_t* hbT;                           // SMALL mode: [nrep][Npad/32][Npad] OUTER list = symmetric hit bit matrix (radius rc+skin+oskin),
                                           // word-major ("transposed": word w of row i at [w][i], coalesced over atoms)
  uint32_t* ginfo;                         // SMALL mode: [nrep][2][Npad] ghost table of each list buffer: base << 8 | lower-half bits << 3 | near-face bits
  uint32_t* ltmp;                          // [nrep][maxnbo][Npad] outer-build scratch: j | code << 16 in discovery order
  ushort4* olist;                          // [nrep][maxqo][Npad] OUTER list (radius rc+skin+oskin), same grouped format
  uint8_t* ocode;                          // [nrep][maxqo][Npad]
  uint16_t* onq;                           // [nrep][Npad]
  double* x0o;                             // [nrep][3][Npad] fractional coordinates at the last outer build
  double* L0o;                             // [nrep] box at the last outer build
  uint16_t* nnb;                           // [nrep][2][Npad]     number of quads of atom i (per list buffer)
  int* micmode;                            // [nrep] 1: box < 2(rc+skin) at build, images resolved per pair
  double *box, *pe, *w, *ke, *L0;          // [nrep]
  double *step;                            // [nrep][3]  dx dv dt
  double *cnt;                             // [nrep][6]  ntp nap ntv nav nth nah
  double *list_pairs;                      // [nrep] listed unordered pairs of the current list
  int *cfg_slot, *slot_cfg;                // local permutation
  unsigned long long* cta_clk;             // [nrep] SM clocks the configuration's CTA spent in the last cycle
  unsigned long long* cost;                // [nrep][2] last cycle: contention-independent work estimate (listed pairs evaluated + 0.3 N^2 per list
                                           // build, in units of one listed pair), force evaluations
  unsigned long long* rep_ct;              // [nrep][NM_COUNTER_WIDTH] the last cycle's counters of each configuration
  unsigned long long* mv_clk;              // [nrep][4] last cycle: SM clocks in PMC / VMC / HMC moves (solo-corrected), [3] = moves of each kind packed 3 x 16 bit
  int* order;                              // [nrep] ticket (within a segment) -> configuration, cheapest first
  int* sched;                              // work queue of the persistent cycle kernel (reset by k_schedule before every launch):
                                           // [0] next ticket, [1 + c] segments of configuration c that are complete,
                                           // [1 + nrep] CTAs arrived, [2 + nrep] placement invalid, [3 + nrep + smid] CTAs on SM smid,
                                           // [3 + nrep + SMID_MAX] chains finished, [4 + nrep + SMID_MAX] chains started (helpers)
  int nseg, seg_moves;                     // a cycle is cut into nseg segments of seg_moves moves (the unit of scheduling)
  int place;                               // 1: first ticket of every CTA from the SM-aware placement (placement_rank)
  double build_cost;                       // SMALL-mode list build in units of one listed-pair evaluation, / N^2 (cost ranks of the placement)
  double* skinc;                           // [nrep] list skin of each configuration (SMALL mode: tuned per configuration by k_adapt; d.skin otherwise)
  int adapt_skin;                          // 1: k_adapt moves skinc one step of 0.025 per cycle towards the cheaper side (see there)
  double skin_lo, skin_hi, skin_pow;       // its range and the exponent of the rebuild-count model
  double inner_cost;                       // LARGE mode: one inner list build in listed-pair evaluations, / N
  // force helpers (LARGE mode, fewer configurations than CTA slots): CTAs without a chain of their own evaluate the
  // upper half of the force rows of a running chain (see helper_serve)
  int nhelp;                               // CTAs launched beyond nrep (0: off)
  int help_quantum;                        // commands a helper serves before it looks for the chain that needs help most
  int* help;                               // [nrep][HELP_STRIDE] hand-shake record of configuration c (reset by k_schedule)
  double* helpd;                           // [nrep][4] command parameters: box, dtf
  double* hpart;                           // [nrep][4][nthr] the helper's per-thread partial sums (energy, virial, pairs, kinetic)
  int list_pf;                             // force loop: prefetch the list row this many rows past the one being loaded into L1 (-1: off)
  int *status;                             // [nrep]
  // per local slot
  double *label;                           // [nrep][4] et pf temp temp_vel
  double *thermo;                          // [nrep][18]
  unsigned long long* counters;            // [NM_COUNTER_WIDTH]
};

// global slot index k = i*NT + j of local slot k_local (the RNG streams and the exchange draws are keyed on it, so results
// do not depend on how the pressure rows are spread over GPUs)
__host__ __device__ inline int gslot(const Dev& d, int k) { return (d.row0 + (k / d.nt) * d.row_stride) * d.nt + k % d.nt; }

// per-CTA context (registers + shared-memory carve)
struct Ctx {
  int N, Npad, c;
  double L, L0, thr2;           // box, list build box, squared displacement budget (build-box units)
  double* sp;                   // shared positions, AoS: atom j at sp[3j..3j+2] (one address register per gather)
  float4* sf;                   // shared float32 fractional positions (list build prefilter only)
  uint32_t* hbT;                // global (SMALL mode): N x N hit bit matrix (outer list), word-major
  uint32_t* ginfo;              // shared (SMALL mode): ghost table of the current list, one word per atom
  uint32_t* ginfo_g;            // global copies of the ghost table, one per list buffer
  uint8_t* gtbl;                // shared (SMALL mode): rank of image subset g among the subsets of a near-face mask, [8][8]
  uint16_t* gidx;               // shared (SMALL mode): [Npad][8] index of the copy of atom j for image subset g (g = 0: j itself)
  uint4* hp;                    // shared (SMALL mode): [Npad/2] half-precision fractional coordinates of atoms 2k, 2k+1 packed (x2, y2, z2, -): tile tests
  int* iscan;                   // shared: block-scan scratch (34 ints)
  double *red, *bc;             // reduction scratch, broadcast scratch
  int *cell_cnt, *cell_start, *ibc;
  uint16_t* cell_atoms;
  uint16_t* gcur;               // shared (LARGE mode): [8][nthr] per-thread group counters / cursors of the outer build
  unsigned long long* s_pairs;  // shared: in-cutoff ordered pairs of force-only evaluations
  // global views of this configuration
  double *gx, *gv, *gf, *gxs, *gvs, *gfs, *gx0;
  ushort4* list; uint32_t* ltmp; uint16_t* nnb;
  ushort4* olist; uint8_t* ocode; uint16_t* onq; double* gx0o;
  double L0o, thro2;            // outer list: build box and squared displacement budget
  double* sht;                  // shared: 27 image shift vectors (k*L) for the current box
  int mic;                      // minimum image per pair (small boxes) instead of stored image codes / ghost copies
  int ghost;                    // SMALL mode, current list uses ghost indices: position updates maintain the copies
  unsigned long long* ct;       // shared: NM_COUNTER_WIDTH counters, touched by thread 0 only
  double list_pairs;
  // double-buffered lists: lbuf = buffer of the current list; inside a move (in_move) the first build switches to
  // the other buffer and leaves the list of the saved positions (sv_*) intact for a revert
  int lbuf, in_move, sv_lbuf, sv_mic, outer_in_move;
  double sv_L0, sv_list_pairs;
  ushort4* list_base; uint16_t* nnb_base; double* gx0_base;
  size_t list_stride;           // quads per list buffer
  int status;
  int redflip;                  // which half of `red` the next block_sum uses
  double skin;                  // list skin of this configuration
  int* help;                    // this configuration's helper record (nullptr: no helpers -- every kernel but k_cycle in LARGE mode)
  int help_seq;                 // commands issued to the helper so far in this cycle
  double *helpd, *hpart;
};

// shared-memory carve (ctx_init): doubles first (positions [+ ghost copies], reduction / broadcast scratch, image shifts,
// counters), then the float4 fractional copies (16-byte aligned: every count above is even), then the mode's integers
__host__ __device__ inline size_t smem_bytes(int Npad, int N, int small, int nthr) {
  size_t b = sizeof(double) * (3 * (size_t)Npad + RED_DOUBLES + BC_DOUBLES + SHT_DOUBLES);
  if (small) b += sizeof(double) * 3 * (size_t)ghost_cap(N);
  b += sizeof(unsigned long long) * (2 + NM_COUNTER_WIDTH);
  b += sizeof(float4) * (size_t)Npad;
  b += sizeof(int) * (8 + 2 + 34);                              // ibc, iscan (44 ints: keeps 16-byte alignment)
  if (small) b += sizeof(uint32_t) * (size_t)Npad + 64 + sizeof(uint16_t) * 8 * (size_t)Npad + sizeof(uint4) * (size_t)(Npad / 2);   // ghost table, subset-rank table, copy index table, packed half copies
  else {
    b += sizeof(int) * 2 * (NCMAX * NCMAX * NCMAX + 1);         // cell counts / starts
    b += sizeof(uint16_t) * (size_t)Npad;                       // cell members
    b += sizeof(uint16_t) * 8 * (size_t)nthr;                   // outer build: per-thread group counters / cursors
  }
  return b;
}

namespace {   // device functions: internal linkage (this file is compiled as several translation units, see NM_TU)
constexpr bool kHelpers = NM_HELPERS != 0;

// gpu-scope acquire load / release store (segment hand-over of the persistent kernel, force helpers). ptxas follows every
// acquire load with CCTL.IVALL: the SM's L1 is dropped, so plain loads issued after a barrier see the other SM's writes.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

// block_sum on alternating scratch halves (see nm_device.cuh): one barrier per reduction
template <int K>
__device__ __forceinline__ void bsum(double (&v)[K], Ctx& cx) {
  static_assert(32 * K <= RED_HALF, "block_sum scratch");
  cx.redflip ^= 1;
  block_sum<K>(v, cx.red + cx.redflip * RED_HALF);
}

// point the list views at buffer cx.lbuf
__device__ __forceinline__ void select_list(Ctx& cx) {
  cx.list = cx.list_base + (size_t)cx.lbuf * cx.list_stride;
  cx.nnb = cx.nnb_base + (size_t)cx.lbuf * cx.Npad;
  cx.gx0 = cx.gx0_base + (size_t)cx.lbuf * 3 * cx.Npad;
}

__device__ __forceinline__ void ctx_init(const Dev& d, Ctx& cx, int c, unsigned char* smem) {
  cx.redflip = 0;
  cx.N = d.N; cx.Npad = d.Npad; cx.c = c;
  double* p = reinterpret_cast<double*>(smem);
  cx.sp = p; p += 3 * d.Npad;
  if (d.small) p += 3 * ghost_cap(d.N);      // ghost copies: entries Npad .. Npad + ghost_cap - 1 of sp
  cx.red = p; p += RED_DOUBLES;
  cx.bc = p; p += BC_DOUBLES;
  cx.sht = p; p += SHT_DOUBLES;
  cx.s_pairs = reinterpret_cast<unsigned long long*>(p); p += 2;
  cx.ct = reinterpret_cast<unsigned long long*>(p); p += NM_COUNTER_WIDTH;
  cx.sf = reinterpret_cast<float4*>(p);
  int* q = reinterpret_cast<int*>(cx.sf + d.Npad);
  cx.ibc = q; q += 8 + 2;
  cx.iscan = q; q += 34;
  if (d.small) {
    cx.gidx = reinterpret_cast<uint16_t*>(q); q += 4 * d.Npad;        // first: 16-byte aligned rows
    cx.hp = reinterpret_cast<uint4*>(q); q += 2 * d.Npad;             // (Npad / 2) x 16 bytes
    cx.ginfo = reinterpret_cast<uint32_t*>(q); q += d.Npad;
    cx.gtbl = reinterpret_cast<uint8_t*>(q);
    cx.cell_cnt = cx.cell_start = nullptr; cx.cell_atoms = nullptr; cx.gcur = nullptr;
    // rank of the image subset g (bit a: shifted along axis a) among the non-empty subsets of the near-face mask nb
    if (threadIdx.x < 64) {
      const int nb = threadIdx.x >> 3, g = threadIdx.x & 7;
      int r = 0, sh = 0;
      for (int a = 0; a < 3; a++) if ((nb >> a) & 1) { r |= ((g >> a) & 1) << sh; sh++; }
      cx.gtbl[threadIdx.x] = (uint8_t)r;
    }
  } else {
    cx.cell_cnt = q; q += NCMAX * NCMAX * NCMAX + 1;
    cx.cell_start = q; q += NCMAX * NCMAX * NCMAX + 1;
    cx.cell_atoms = reinterpret_cast<uint16_t*>(q);
    cx.gcur = cx.cell_atoms + d.Npad;
    cx.ginfo = nullptr; cx.gtbl = nullptr; cx.gidx = nullptr; cx.hp = nullptr;
  }
  cx.hbT = d.small ? d.hbT + (size_t)c * (d.Npad / 32) * d.Npad : nullptr;
  cx.ginfo_g = d.small ? d.ginfo + (size_t)c * 2 * d.Npad : nullptr;
  const size_t off = (size_t)c * 3 * d.Npad;
  cx.gx = d.x + off; cx.gv = d.v + off; cx.gf = d.f + off;
  cx.gxs = d.xs + off; cx.gvs = d.vs + off; cx.gfs = d.fs + off;
  cx.list_stride = (size_t)(d.maxq + LIST_SPARE_ROWS) * d.Npad;
  cx.list_base = d.list + (size_t)c * 2 * cx.list_stride;
  cx.nnb_base = d.nnb + (size_t)c * 2 * d.Npad;
  cx.gx0_base = d.x0 + 2 * off;
  cx.lbuf = d.lcur[c]; cx.in_move = 0; cx.sv_lbuf = cx.lbuf; cx.outer_in_move = 0;
  select_list(cx);
  cx.ltmp = d.ltmp + (size_t)c * ((d.maxnbo + 3) & ~3) * d.Npad;     // [entry][atom]
  cx.olist = d.olist + (size_t)c * d.maxqo * d.Npad;
  cx.ocode = d.ocode + (size_t)c * d.maxqo * d.Npad;
  cx.onq = d.onq + (size_t)c * d.Npad;
  cx.gx0o = d.x0o + off;
  cx.L0o = d.L0o[c];
  cx.mic = d.micmode[c];
  cx.L = d.box[c]; cx.L0 = d.L0[c]; cx.list_pairs = d.list_pairs[c];
  cx.status = 0;
  cx.skin = d.adapt_skin ? d.skinc[c] : d.skin;
  cx.help = nullptr; cx.help_seq = 0; cx.helpd = nullptr; cx.hpart = nullptr;
  if (threadIdx.x == 0) { cx.s_pairs[0] = 0; cx.s_pairs[1] = 0; for (int k = 0; k < NM_COUNTER_WIDTH; k++) cx.ct[k] = 0; }
}

// displacement budgets for box L (s = L/L0): inner list complete while s*(rl - 2u) >= rc; the outer list can
// regenerate a complete inner list while s*(rlo - 2u) >= rl
__device__ __forceinline__ void update_thr(const Dev& d, Ctx& cx) {
  const double rl = d.rc + cx.skin, rlo = rl + d.oskin;
  if (cx.L0 <= 0.0) cx.thr2 = -1.0;
  else {
    const double s = cx.L / cx.L0, thr = 0.5 * (rl - d.rc / s) * (1.0 - 1e-9);
    cx.thr2 = thr > 0.0 ? thr * thr : -1.0;
  }
  if (cx.L0o <= 0.0) cx.thro2 = -1.0;
  else {
    const double s = cx.L / cx.L0o, thr = 0.5 * (rlo - rl * (1.0 + 1e-4) / s) * (1.0 - 1e-9);
    cx.thro2 = thr > 0.0 ? thr * thr : -1.0;
  }
}
// squared displacement of (x,y,z) from the inner / outer list reference of atom i, in build-box length units
__device__ __forceinline__ double disp2(const Ctx& cx, int i, double x, double y, double z, double invL) {
  double ux = x * invL - cx.gx0[i], uy = y * invL - cx.gx0[cx.Npad + i], uz = z * invL - cx.gx0[2 * cx.Npad + i];
  ux -= rint(ux); uy -= rint(uy); uz -= rint(uz);
  return (ux * ux + uy * uy + uz * uz) * cx.L0 * cx.L0;
}
__device__ __forceinline__ double disp2o(const Ctx& cx, int i, double x, double y, double z, double invL) {
  double ux = x * invL - cx.gx0o[i], uy = y * invL - cx.gx0o[cx.Npad + i], uz = z * invL - cx.gx0o[2 * cx.Npad + i];
  ux -= rint(ux); uy -= rint(uy); uz -= rint(uz);
  return (ux * ux + uy * uy + uz * uz) * cx.L0o * cx.L0o;
}

// positions global -> shared, plus the far-away dummy atom that pads the list. Positions are CONTINUOUS between
// list builds (an atom may sit slightly outside [0,L)): the list stores the periodic image of every pair, so
// atoms are re-wrapped only when the list is rebuilt.
// GHOST copies of atom i (SMALL mode, cx.ghost): for every non-empty subset g of the faces the atom is near (nb), the
// copy shifted by +L along the axes of g where the atom sits in the lower half of the box, -L in the upper half; the
// copies of one atom are consecutive, ordered by g (= by gtbl rank). Called by whoever rewrites sp[3i..3i+2].
__device__ __forceinline__ void write_ghosts(Ctx& cx, int i, double x, double y, double z) {
  const uint32_t gi = cx.ginfo[i];
  const unsigned nb = gi & 7u;
  if (nb == 0u) return;
  const double L = cx.L;
  const double xs = x + ((gi & 8u) ? L : -L), ys = y + ((gi & 16u) ? L : -L), zs = z + ((gi & 32u) ? L : -L);
  double* q = cx.sp + 3 * (size_t)(cx.Npad + (gi >> 8));
#pragma unroll
  for (unsigned g = 1; g < 8; g++)
    if ((g & ~nb) == 0u) { q[0] = (g & 1u) ? xs : x; q[1] = (g & 2u) ? ys : y; q[2] = (g & 4u) ? zs : z; q += 3; }
}
// the one way to move atom i
__device__ __forceinline__ void store_pos(Ctx& cx, int i, double x, double y, double z) {
  cx.sp[3 * i] = x; cx.sp[3 * i + 1] = y; cx.sp[3 * i + 2] = z;
  if (cx.ghost) write_ghosts(cx, i, x, y, z);
}
// ghost table of list buffer cx.lbuf -> shared (the owner's entry: read by the owner only until the next build)
__device__ __forceinline__ void load_ginfo(Ctx& cx) {
  for (int i = threadIdx.x; i < cx.N; i += blockDim.x) cx.ginfo[i] = cx.ginfo_g[(size_t)cx.lbuf * cx.Npad + i];
}

__device__ void load_positions(const Dev& d, Ctx& cx) {
  cx.ghost = d.small && !cx.mic && cx.L0 > 0.0;
  if (cx.ghost) load_ginfo(cx);
  for (int i = threadIdx.x; i < cx.N; i += blockDim.x) store_pos(cx, i, cx.gx[i], cx.gx[cx.Npad + i], cx.gx[2 * cx.Npad + i]);
  for (int i = cx.N + threadIdx.x; i < cx.Npad; i += blockDim.x) { cx.sp[3 * i] = 1e9; cx.sp[3 * i + 1] = 1e9; cx.sp[3 * i + 2] = 1e9; }
}
__device__ void store_positions(Ctx& cx) {
  for (int i = threadIdx.x; i < cx.N; i += blockDim.x) {
    cx.gx[i] = cx.sp[3 * (i)]; cx.gx[cx.Npad + i] = cx.sp[3 * (i) + 1]; cx.gx[2 * cx.Npad + i] = cx.sp[3 * (i) + 2];
  }
}

// force-loop quads carry their image code (0..26) in the 3 spare top bits of the first two indices (N <= 8191)
__device__ __forceinline__ ushort4 pack_code(ushort4 v, int code) {
  v.x = (unsigned short)(v.x | ((code & 7) << 13)); v.y = (unsigned short)(v.y | ((code >> 3) << 13));
  return v;
}

// ---- force helpers (LARGE mode). With fewer configurations than SMs (C3: 128 chains on 148 SMs) the spare CTAs of the
// grid, and every CTA whose own chain has finished, attach themselves to the running chain with the most work left (the
// owners publish an estimate after every move) and evaluate the force rows [2 * blockDim, N) of its evaluations while the
// owner does rows [0, 2 * blockDim): the owner publishes its positions (shared -> global x) and a command (release), the
// helper gathers them into its own shared memory, walks the same list rows with the same arithmetic, writes f (and the
// kicked v) of its atoms and its per-thread partial sums, and answers (release). After help_quantum commands the helper
// detaches (it clears the attached flag BEFORE its last answer, so the owner cannot address it again) and chooses anew:
// helped chains fall back in the ranking, and all chains finish at about the same time. The owner never waits for a
// helper that has not announced itself, so no CTA depends on another one being resident. Results do not depend on
// whether, when or by whom a chain is helped: per-atom forces are independent, and the energy / virial / kinetic sums
// are formed as (rows below the split) + (rows above it) per thread in either case. The same split serves the list
// builds of a helped chain: rows [2 * blockDim, N) of the outer search (the helper bins the cells itself: the cell order
// is sorted, hence identical) and of the inner regeneration are independent per atom.
__device__ __forceinline__ bool help_request(Ctx& cx, int flags, double dtf, int a = 0, int b = 0) {
  if (threadIdx.x == 0) cx.ibc[4] = ld_acquire_gpu(cx.help);
  __syncthreads();
  if (cx.ibc[4] != 1) return false;
  store_positions(cx);
  cx.help_seq++;
  __threadfence();                                       // (measured: free; the barrier + thread 0's release would do by cumulativity)
  __syncthreads();
  if (threadIdx.x == 0) {
    cx.help[3] = flags; cx.help[4] = cx.lbuf; cx.help[9] = a; cx.help[10] = b;
    cx.helpd[0] = cx.L; cx.helpd[1] = dtf;
    st_release_gpu(cx.help + 1, cx.help_seq);
  }
  return true;
}
// wait for the helper's answer; the barrier that follows the acquire (and its L1 invalidation) publishes it to the CTA
__device__ __forceinline__ void help_wait(Ctx& cx) {
  if (threadIdx.x == 0) while (ld_acquire_gpu(cx.help + 2) != cx.help_seq) __nanosleep(64);
  __syncthreads();
}

// ------------------------------------------------------------------ list construction helpers
// Per-atom list rows grouped by periodic image. A pair enters when its float32 distance is below the list radius
// times (1+margin); the margin covers the float32 rounding of the fractional coordinates (<= 2^-24 each, < 4e-6
// relative on r^2), so the row is a superset of the exact list. With box >= 2 r_list the image of a neighbour along
// one axis is either 0 or one fixed sign per atom (-1 for atoms in the lower half of the box, +1 in the upper
// half): at most 8 image groups, g = (kx != 0) | (ky != 0) << 1 | (kz != 0) << 2, counted in packed registers.
// Pass A streams the hits (discovery order) into the owner's scratch row with 16-byte stores; pass B sweeps that row
// once per non-empty group and emits whole quads (8 bytes).
// Candidates: all atoms (nc == 1), the 27 stencil cells (nc >= 3), or -- BITS -- the set bits of the atom's row in the
// shared-memory hit matrix. Output: OUTER rows (row-major per atom) or, BITS mode, the [quad][atom] force-loop layout.
__device__ __forceinline__ int cell_of(const float4 p, int nc) {
  const int a = min(nc - 1, (int)(p.x * nc)), b = min(nc - 1, (int)(p.y * nc)), e = min(nc - 1, (int)(p.z * nc));
  return (a * nc + b) * nc + e;
}
__device__ int outer_rows(const Dev& d, Ctx& cx, float rl2f, int nc, int sw, int i0, int i1) {
  const int N = cx.N, Npad = cx.Npad, tid = threadIdx.x, nthr = blockDim.x;
  const double invL = 1.0 / cx.L;
  const bool grouped = !cx.mic;
  const int capq = d.maxqo, maxnbo = d.maxnbo;
  const unsigned sf_s = (unsigned)__cvta_generic_to_shared(cx.sf);
  const unsigned cur_s = (unsigned)__cvta_generic_to_shared(cx.gcur + tid), gstep = 2u * (unsigned)nthr;
  uint16_t* ol16 = reinterpret_cast<uint16_t*>(cx.olist);
  int over = 0;
  for (int i = i0 + tid; i < i1; i += nthr) {
    const float4 pi = cx.sf[i];
    uint32_t* trow = cx.ltmp + i;                          // scratch entry t of atom i: trow[t * Npad] (coalesced over atoms)
#pragma unroll
    for (int g = 0; g < 8; g++) asm volatile("st.shared.u16 [%0], %1;" :: "r"(cur_s + g * gstep), "h"((unsigned short)0) : "memory");
    int cnt = 0;
    // pass A: candidates in a fixed order (stencil cell by stencil cell, ascending index inside a cell); hits go to the
    // scratch column with their image group, group sizes to the shared counters
    auto test = [&](int j) {
      float4 pj;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pj.x), "=f"(pj.y), "=f"(pj.z), "=f"(pj.w) : "r"(sf_s + 16u * (unsigned)j));
      const float ax = fabsf(pi.x - pj.x), ay = fabsf(pi.y - pj.y), az = fabsf(pi.z - pj.z);
      const float mx = fminf(ax, 1.f - ax), my = fminf(ay, 1.f - ay), mz = fminf(az, 1.f - az);
      if (fmaf(mz, mz, fmaf(my, my, mx * mx)) < rl2f && j != i) {
        if (cnt < maxnbo) {
          const unsigned g = grouped ? ((ax > 0.5f) | ((ay > 0.5f) << 1) | ((az > 0.5f) << 2)) : 0u;
          trow[(size_t)cnt * Npad] = (uint32_t)j | (g << 16);
          const unsigned ca = cur_s + g * gstep;
          unsigned short c16;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c16) : "r"(ca));
          asm volatile("st.shared.u16 [%0], %1;" :: "r"(ca), "h"((unsigned short)(c16 + 1)) : "memory");
        }
        cnt++;
      }
    };
    if (nc == 1) {
      for (int j = 0; j < N; j++) test(j);
    } else {
      const int ci = cell_of(pi, nc), a = ci / (nc * nc), b = (ci / nc) % nc, e = ci % nc;
      // the stencil cells along the fastest axis are contiguous in the cell order: one or (across the boundary) two
      // ranges per (da, db) instead of 2 sw + 1 cells
      int lo1 = e - sw, hi1 = e + sw, lo2 = 0, hi2 = -1;
      if (lo1 < 0) { lo2 = 0; hi2 = hi1; hi1 = nc - 1; lo1 += nc; }
      else if (hi1 >= nc) { lo2 = 0; hi2 = hi1 - nc; hi1 = nc - 1; }
      for (int da = -sw; da <= sw; da++) {
        const int am = (a + da + nc) % nc;
        for (int db = -sw; db <= sw; db++) {
          const int rowc = (am * nc + (b + db + nc) % nc) * nc;
          for (int p = cx.cell_start[rowc + lo1], en = cx.cell_start[rowc + hi1 + 1]; p < en; p++) test(cx.cell_atoms[p]);
          if (hi2 >= 0) for (int p = cx.cell_start[rowc + lo2], en = cx.cell_start[rowc + hi2 + 1]; p < en; p++) test(cx.cell_atoms[p]);
        }
      }
    }
    if (cnt > maxnbo) { over = 1; cnt = maxnbo; }
    // group sizes -> start cursors (quad-padded), image codes
    const int sx = pi.x < 0.5f ? -1 : 1, sy = pi.y < 0.5f ? -1 : 1, sz = pi.z < 0.5f ? -1 : 1;
    int ng[8], run = 0;
#pragma unroll
    for (int g = 0; g < 8; g++) {
      unsigned short c16;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c16) : "r"(cur_s + g * gstep));
      ng[g] = c16;
      asm volatile("st.shared.u16 [%0], %1;" :: "r"(cur_s + g * gstep), "h"((unsigned short)run) : "memory");
      run += (ng[g] + 3) & ~3;
    }
    const int nq = run >> 2;
    if (nq > capq) { over = 1; cx.onq[i] = 0; continue; }
    uint16_t* oi = ol16 + (size_t)i * 4;
    const unsigned qstride = (unsigned)Npad * 4u;
    // pass B: one sweep over the scratch column, every index straight to its slot of the [quad][atom] layout
    for (int t = 0; t < cnt; t++) {
      const uint32_t en = trow[(size_t)t * Npad];
      const unsigned ca = cur_s + (en >> 16) * gstep;
      unsigned short pos;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(pos) : "r"(ca));
      asm volatile("st.shared.u16 [%0], %1;" :: "r"(ca), "h"((unsigned short)(pos + 1)) : "memory");
      oi[(pos >> 2) * qstride + (pos & 3)] = (uint16_t)(en & 0xffffu);
    }
    int q0 = 0;
#pragma unroll
    for (int g = 0; g < 8; g++) {
      if (ng[g] == 0) continue;
      const int code = 13 + 9 * (g & 1) * sx + 3 * ((g >> 1) & 1) * sy + ((g >> 2) & 1) * sz;
      for (unsigned pos = (unsigned)(4 * q0 + ng[g]); pos & 3u; pos++) oi[(pos >> 2) * qstride + (pos & 3u)] = (uint16_t)N;   // dummy padding
      const int nqg = (ng[g] + 3) >> 2;
      for (int q = q0; q < q0 + nqg; q++) cx.ocode[(size_t)q * Npad + i] = (uint8_t)code;
      q0 += nqg;
    }
    cx.onq[i] = (uint16_t)nq;
    cx.gx0o[i] = cx.sp[3 * i] * invL; cx.gx0o[Npad + i] = cx.sp[3 * i + 1] * invL; cx.gx0o[2 * Npad + i] = cx.sp[3 * i + 2] * invL;
  }
  return over;
}

__device__ __forceinline__ uint32_t lds_u32(unsigned a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_u64(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void lds_f64x3(unsigned a, double& x, double& y, double& z) {
  asm volatile("ld.shared.f64 %0, [%3];\n\tld.shared.f64 %1, [%3+8];\n\tld.shared.f64 %2, [%3+16];" : "=d"(x), "=d"(y), "=d"(z) : "r"(a));
}
__device__ __forceinline__ void sts_u32(unsigned a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// re-wrap every atom into [0,L) and refresh the float32 fractional copies; entries N..Npad-1 are parked far away so
// that padded indices never test positive. The revert copy of a move in flight is NOT shifted: a build inside a move
// goes to the other list buffer, and a rejected move returns to the saved positions together with the list (and its
// periodic images) that was valid for them.
__device__ void wrap_and_refresh(Ctx& cx, bool wrap) {
  const int N = cx.N, Npad = cx.Npad;
  const double L = cx.L, invL = 1.0 / L;
  for (int i = threadIdx.x; i < Npad; i += blockDim.x) {
    if (i < N) {
      if (wrap) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const double x = cx.sp[3 * i + a], xw = wrap1(x - floor(x * invL) * L, L);
          if (xw != x) cx.sp[3 * i + a] = xw;
        }
      }
      cx.sf[i] = make_float4((float)(cx.sp[3 * i] * invL), (float)(cx.sp[3 * i + 1] * invL), (float)(cx.sp[3 * i + 2] * invL), 0.f);
    } else cx.sf[i] = make_float4(1e6f, 1e6f, 1e6f, 1e30f);     // .w: added to r^2 by the FP32-mode loop (the minimum image would fold the dummy back)
  }
}

// ------------------------------------------------------------------ SMALL mode (N <= NSMALL, one atom per thread)
// Single-level list (radius rl = rc + skin), rebuilt from scratch each time in two steps.
// (1) Symmetric N x N hit BIT MATRIX: every unordered pair is tested ONCE, 32 x 32 tile by tile, on the FP32 pipe
//     (0.7 instructions per pair test); the warp ballot of each column gives the transposed bits, so both rows of the
//     matrix are written without atomics (word-major in global memory: coalesced, 32 KB per configuration at N = 500).
// (2) Every thread walks the set bits of its row, four per iteration (four independent gathers in flight), and stores
//     for each the index of the COPY to use -- the atom itself, or its ghost shifted along the axes where the pair
//     wraps (|d| > 1/2 of the box) -- straight into the [quad][atom] layout. No image codes, no groups: the only
//     padding is the last quad.
// Atoms are re-wrapped here, and the ghost table is rebuilt: atom j gets a copy for every non-empty subset of the
// faces it is within rl (1 + 1e-3) of. Boxes below 2 rl (or more ghosts than the shared array holds, or FP32 mode)
// use plain indices and the per-pair minimum image (cx.mic).
template <bool GHOST>
__device__ __forceinline__ void walk_hit_row(const Dev& d, Ctx& cx, int i, const float4 pi, int& over, double& tot) {
  const int N = cx.N, Npad = cx.Npad, W = (N + 31) / 32;
  char* li = reinterpret_cast<char*>(cx.list + i);
  const unsigned rowbytes = (unsigned)Npad * 8u, cap = 4u * (unsigned)d.maxq;
  const unsigned sf_s = (unsigned)__cvta_generic_to_shared(cx.sf), gx_s = (unsigned)__cvta_generic_to_shared(cx.gidx);
  const uint32_t* hrow = cx.hbT + i;
  // Hits are pushed into a 128-bit shift register (stored XOR N, so that empty fields read as the parked dummy atom);
  // one 8-byte store per completed quad: a warp's rows advance at different rates, and a 2-byte store per hit costs a
  // sector write per lane -- the walk was bound by those.
  const unsigned long long dummy4 = 0x0001000100010001ull * (unsigned long long)N;
  unsigned pos = 0, fill = 0, oq = 0;
  unsigned long long acc = 0ull;
  int w = 0;
  uint32_t m = hrow[0], mnext = W > 1 ? hrow[Npad] : 0u;            // the next word is in flight while this one is walked
  for (;;) {
    while (m == 0u && ++w < W) { m = mnext; mnext = w + 1 < W ? hrow[(size_t)(w + 1) * Npad] : 0u; }
    if (w >= W) break;
    const unsigned nv = min(4u, (unsigned)__popc(m));
    unsigned jj[4], idx[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const unsigned bit = (unsigned)__ffs(m) - 1u;                  // m == 0: replaced by the parked dummy atom
      jj[t] = m ? 32u * (unsigned)w + bit : (unsigned)N;
      m &= m - 1u;
    }
    if (GHOST) {
      float4 pj[4];
#pragma unroll
      for (int t = 0; t < 4; t++)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pj[t].x), "=f"(pj[t].y), "=f"(pj[t].z), "=f"(pj[t].w) : "r"(sf_s + 16u * jj[t]));
#pragma unroll
      for (int t = 0; t < 4; t++) {
        // the pair wraps along an axis <=> the fractional separation exceeds 1/2 (listed pairs: < rl/L or > 1 - rl/L)
        const unsigned g = (fabsf(pi.x - pj[t].x) > 0.5f ? 2u : 0u) + (fabsf(pi.y - pj[t].y) > 0.5f ? 4u : 0u) + (fabsf(pi.z - pj[t].z) > 0.5f ? 8u : 0u);
        unsigned short r;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(gx_s + 16u * jj[t] + g));
        idx[t] = r;
      }
    } else {
#pragma unroll
      for (int t = 0; t < 4; t++) idx[t] = jj[t];
    }
    // invalid slots (t >= nv) hold the dummy N (plain) or its table entry (ghost): XOR N of slot t is forced to 0 there
    const unsigned n16 = (unsigned)N;
    const unsigned e0 = idx[0] ^ n16, e1 = nv > 1u ? idx[1] ^ n16 : 0u, e2 = nv > 2u ? idx[2] ^ n16 : 0u, e3 = nv > 3u ? idx[3] ^ n16 : 0u;
    const unsigned long long packed = (unsigned long long)(e0 | (e1 << 16)) | ((unsigned long long)(e2 | (e3 << 16)) << 32);
    const unsigned sh = 16u * fill;
    const unsigned long long lo = acc | (packed << sh);
    const unsigned long long hi = fill ? packed >> (64u - sh) : 0ull;
    fill += nv; pos += nv;
    if (fill >= 4u) {
      if (oq < (unsigned)d.maxq) *reinterpret_cast<unsigned long long*>(li + oq * rowbytes) = lo ^ dummy4;
      oq++; acc = hi; fill -= 4u;
    } else acc = lo;
  }
  tot = (double)pos;
  if (fill) { if (oq < (unsigned)d.maxq) *reinterpret_cast<unsigned long long*>(li + oq * rowbytes) = acc ^ dummy4; oq++; }
  if (pos > cap) over = 1;
  cx.nnb[i] = (uint16_t)min(oq, (unsigned)d.maxq);
  const double invL = 1.0 / cx.L;
  cx.gx0[i] = cx.sp[3 * i] * invL; cx.gx0[Npad + i] = cx.sp[3 * i + 1] * invL; cx.gx0[2 * Npad + i] = cx.sp[3 * i + 2] * invL;
}

__device__ void build_small(const Dev& d, Ctx& cx) {
  const int N = cx.N, Npad = cx.Npad, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const double L = cx.L, rl = d.rc + cx.skin, invL = 1.0 / L;
  const int W = (N + 31) / 32;
  __syncthreads();
  wrap_and_refresh(cx, true);
  __syncthreads();
  // ---- (1) tiles, two pair tests per instruction: the columns of a tile are taken two at a time from half-precision
  // copies of the (centred) fractional coordinates, packed per atom pair. The half arithmetic errs by < 1.2 % on the
  // squared separation of a pair near the list radius (inputs 2^-13 each, three roundings of 2^-12 per axis, the
  // products and sums 2^-11 relative), so the threshold is raised by that much: the matrix stays a superset of the
  // exact list (1.7 % more listed pairs than the float32 test), and it is symmetric bit for bit (|a - b| = |b - a|).
  const long long t_tiles0 = clock64();
  for (int k = tid; k < Npad / 2; k += blockDim.x) {
    const float4 a = cx.sf[2 * k], b = cx.sf[2 * k + 1];
    const bool da = 2 * k >= N, db = 2 * k + 1 >= N;                 // parked entries: far away, finite
    const __half2 hx = __floats2half2_rn(da ? 100.f : a.x - 0.5f, db ? 100.f : b.x - 0.5f);
    const __half2 hy = __floats2half2_rn(da ? 100.f : a.y - 0.5f, db ? 100.f : b.y - 0.5f);
    const __half2 hz = __floats2half2_rn(da ? 100.f : a.z - 0.5f, db ? 100.f : b.z - 0.5f);
    cx.hp[k] = make_uint4(*reinterpret_cast<const unsigned*>(&hx), *reinterpret_cast<const unsigned*>(&hy), *reinterpret_cast<const unsigned*>(&hz), 0u);
  }
  __syncthreads();
  const float rl2f = (float)(rl * rl * invL * invL * (1.0 + 2e-5));
  const int ntile = W * (W + 1) / 2;
  const uint32_t lastmask = (N & 31) ? (1u << (N & 31)) - 1u : 0xffffffffu;
  unsigned thr2;
  {
    __half t = __float2half_ru((float)(rl * rl * invL * invL * 1.012));
    const __half2 t2 = __halves2half2(t, t);
    thr2 = *reinterpret_cast<const unsigned*>(&t2);
  }
  const __half2 one2 = __floats2half2_rn(1.f, 1.f);
  uint32_t* colbuf = reinterpret_cast<uint32_t*>(cx.red) + 32 * wid;   // the warp's 32 column ballots of a tile (scratch is free here)
  for (int t = wid; t < ntile; t += nw) {
    int ti = 0, rem = t;                      // tile (ti, tj), ti <= tj, enumerated row by row
    while (rem >= W - ti) { rem -= W - ti; ti++; }
    const int tj = ti + rem;
    const int i = ti * 32 + lane;             // < Npad (Npad >= 32 W); rows >= N are never read
    __half2 xi, yi, zi;
    {
      const uint4 me = cx.hp[i >> 1];
      const __half2 mx = *reinterpret_cast<const __half2*>(&me.x), my = *reinterpret_cast<const __half2*>(&me.y), mz = *reinterpret_cast<const __half2*>(&me.z);
      xi = (i & 1) ? __high2half2(mx) : __low2half2(mx); yi = (i & 1) ? __high2half2(my) : __low2half2(my); zi = (i & 1) ? __high2half2(mz) : __low2half2(mz);
    }
    const unsigned pj_s = (unsigned)__cvta_generic_to_shared(cx.hp + tj * 16);
    const unsigned cb_s = (unsigned)__cvta_generic_to_shared(colbuf);
    uint32_t mask = 0;
    __syncwarp();                             // the previous tile's column reads are done
#pragma unroll
    for (int j2 = 0; j2 < 16; j2++) {
      unsigned px, py, pz, pw;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(px), "=r"(py), "=r"(pz), "=r"(pw) : "r"(pj_s + 16u * j2));
      // nearest-image separation along an axis = min(|d|, 1 - |d|) (coordinates in [-1/2, 1/2])
      __half2 ax = __habs2(__hsub2(xi, *reinterpret_cast<const __half2*>(&px)));
      __half2 ay = __habs2(__hsub2(yi, *reinterpret_cast<const __half2*>(&py)));
      __half2 az = __habs2(__hsub2(zi, *reinterpret_cast<const __half2*>(&pz)));
      ax = __hmin2(ax, __hsub2(one2, ax)); ay = __hmin2(ay, __hsub2(one2, ay)); az = __hmin2(az, __hsub2(one2, az));
      const __half2 r2 = __hfma2(az, az, __hfma2(ay, ay, __hmul2(ax, ax)));
      // columns 2 j2 and 2 j2 + 1 of the tile = rows (tj*32 + 2 j2 [+ 1]), word ti: two ballots, own bits into the row mask,
      // lane 0 parks the ballots for the transposed write
      asm volatile("{\n\t.reg .pred p, q, z;\n\t.reg .b32 c0, c1;\n\tsetp.lt.f16x2 p|q, %1, %2;\n\tsetp.ne.b32 z, %5, 0;\n\t"
                   "vote.sync.ballot.b32 c0, p, 0xffffffff;\n\tvote.sync.ballot.b32 c1, q, 0xffffffff;\n\t"
                   "@p or.b32 %0, %0, %3;\n\t@q or.b32 %0, %0, %4;\n\t"
                   "@z st.shared.v2.u32 [%6], {c0, c1};\n\t}"
                   : "+r"(mask) : "r"(*reinterpret_cast<const unsigned*>(&r2)), "r"(thr2), "r"(1u << (2 * j2)), "r"(2u << (2 * j2)),
                     "r"((int)(lane == 0)), "r"(cb_s + 8u * j2) : "memory");
    }
    __syncwarp();
    const uint32_t mycol = colbuf[lane];
    if (tj == W - 1) mask &= lastmask;        // padded columns
    if (ti == tj) mask &= ~(1u << lane);      // self
    cx.hbT[(size_t)tj * Npad + i] = mask;                                  // row i, word tj
    if (ti != tj) cx.hbT[(size_t)ti * Npad + tj * 32 + lane] = mycol;      // row tj*32+lane, word ti (rows >= N: never read)
  }
  (void)rl2f;
  // ---- ghost table (independent of the tiles)
  const bool own = tid < N;
  const float rg = (float)(rl * invL * (1.0 + 1e-3));
  float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned nb = 0, hb = 0;
  if (own) {
    pi = cx.sf[tid];
    hb = (pi.x < 0.5f ? 1u : 0u) | (pi.y < 0.5f ? 2u : 0u) | (pi.z < 0.5f ? 4u : 0u);
    nb = ((hb & 1u) ? pi.x < rg : pi.x > 1.f - rg) ? 1u : 0u;
    nb |= ((hb & 2u) ? pi.y < rg : pi.y > 1.f - rg) ? 2u : 0u;
    nb |= ((hb & 4u) ? pi.z < rg : pi.z > 1.f - rg) ? 4u : 0u;
  }
  const int cnt = own ? (1 << __popc(nb)) - 1 : 0;      // exclusive block scan of the ghost counts
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) cx.iscan[wid] = incl;
  __syncthreads();                                       // also: bit matrix complete
  if (threadIdx.x == 0) cx.ct[NM_CT_CLK_OUTER] += (unsigned long long)(clock64() - t_tiles0);
  if (wid == 0) {
    const int v = lane < nw ? cx.iscan[lane] : 0;
    int wsum = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wsum, o); if (lane >= o) wsum += t; }
    cx.iscan[lane] = wsum - v;                            // exclusive prefix of the warp totals
    if (lane == 31) cx.iscan[32] = wsum;                  // number of ghosts
  }
  __syncthreads();
  const int base = cx.iscan[wid] + incl - cnt, nghost = cx.iscan[32];
  cx.mic = d.f32 || L < 2.0 * rl * (1.0 + 1e-3) || nghost > ghost_cap(N);
  cx.ghost = !cx.mic;
  const bool ghost = cx.ghost;
  if (own) {
    const uint32_t gi = ghost ? ((uint32_t)base << 8) | (hb << 3) | nb : 0u;
    cx.ginfo[tid] = gi;
    cx.ginfo_g[(size_t)cx.lbuf * Npad + tid] = gi;
    if (ghost) {
      // index of the copy of this atom for every image subset g (0: the atom itself; subsets outside nb are never asked for)
      uint32_t e[4];
#pragma unroll
      for (unsigned g = 0; g < 8; g++) {
        const unsigned v = g == 0u ? (unsigned)tid : (unsigned)Npad + (unsigned)base + (unsigned)cx.gtbl[(nb << 3) + g] - 1u;
        if (g & 1u) e[g >> 1] |= v << 16; else e[g >> 1] = v & 0xffffu;
      }
      *reinterpret_cast<uint4*>(cx.gidx + 8 * tid) = make_uint4(e[0], e[1], e[2], e[3]);
      write_ghosts(cx, tid, cx.sp[3 * tid], cx.sp[3 * tid + 1], cx.sp[3 * tid + 2]);
    }
  }
  __syncthreads();
  // ---- (2) rows
#ifdef NM_DEBUG_CLOCKS
  const long long t_rows0 = clock64();
  if (threadIdx.x == 0) cx.ct[NM_CT_DBG_LOOPIT] += (unsigned long long)(t_rows0 - t_tiles0);      // tiles + ghost table
#endif
  int over = 0; double tot = 0.0;
  if (own) { if (ghost) walk_hit_row<true>(d, cx, tid, pi, over, tot); else walk_hit_row<false>(d, cx, tid, pi, over, tot); }
#ifdef NM_DEBUG_CLOCKS
  __syncthreads();
  if (threadIdx.x == 0) cx.ct[NM_CT_DBG_LOOPCLK] += (unsigned long long)(clock64() - t_rows0);    // row walk (slowest warp)
#endif
  double r[2] = { tot, (double)over };
  bsum<2>(r, cx);
  cx.list_pairs = 0.5 * r[0];
  if (r[1] > 0.0) cx.status |= ST_NEIGH;
  cx.L0 = L; cx.L0o = L;
  update_thr(d, cx);
}

// atoms binned into nc^3 cells (counting sort in shared memory; ascending ids inside a cell, so the order -- and with it
// the order of every list row -- does not depend on which thread got there first)
__device__ void bin_cells(Ctx& cx, int nc) {
  const int N = cx.N, tid = threadIdx.x, nthr = blockDim.x, ncell = nc * nc * nc;
  for (int c = tid; c < ncell; c += nthr) cx.cell_cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += nthr) {
    atomicAdd(&cx.cell_cnt[cell_of(cx.sf[i], nc)], 1);
  }
  __syncthreads();
  if (tid < 32) {
    const int per = (ncell + 31) / 32, base = tid * per;
    int s = 0;
    for (int q = 0; q < per; q++) if (base + q < ncell) s += cx.cell_cnt[base + q];
    int incl = s;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
    int run = incl - s;
    for (int q = 0; q < per; q++) if (base + q < ncell) { cx.cell_start[base + q] = run; run += cx.cell_cnt[base + q]; }
    if (tid == 31) cx.cell_start[ncell] = incl;
  }
  __syncthreads();
  for (int c = tid; c < ncell; c += nthr) cx.cell_cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += nthr) {
    const int c = cell_of(cx.sf[i], nc);
    int p = atomicAdd(&cx.cell_cnt[c], 1);
    cx.cell_atoms[cx.cell_start[c] + p] = (uint16_t)i;
  }
  __syncthreads();
  for (int c = tid; c < ncell; c += nthr) {          // ascending ids inside each cell -> deterministic list order
    const int s = cx.cell_start[c], e = cx.cell_start[c + 1];
    for (int p = s + 1; p < e; p++) {
      uint16_t key = cx.cell_atoms[p]; int q = p - 1;
      while (q >= s && cx.cell_atoms[q] > key) { cx.cell_atoms[q + 1] = cx.cell_atoms[q]; q--; }
      cx.cell_atoms[q + 1] = key;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ two-level Verlet lists (deterministic)
// OUTER list (radius rlo = rc + skin + oskin): cell-binned search on the FP32 pipe, rebuilt rarely. Per atom i it
// holds neighbour quads GROUPED BY PERIODIC IMAGE: every quad carries one image code (kx+1)*9+(ky+1)*3+(kz+1),
// k = rint((x_i - x_j)/L), groups padded to whole quads with the far-away dummy atom N.
// INNER list (radius rl = rc + skin): the subset of the outer list currently within rl, regenerated often by a
// single pass over the outer quads (it inherits the grouping). The force loop walks the inner list and subtracts
// the image shift once per quad -- no per-pair minimum-image arithmetic.
// Atoms are re-wrapped into [0,L) at outer builds only; the saved copy used for move reverts is shifted by the
// same lattice vector so that a revert stays consistent with the stored image codes.
__device__ void build_outer(const Dev& d, Ctx& cx) {
  const int N = cx.N, Npad = cx.Npad, tid = threadIdx.x, nthr = blockDim.x;
  const double L = cx.L, rlo = d.rc + cx.skin + d.oskin, invL = 1.0 / L;
  // cells of side >= rlo/2 searched with a 5^3 stencil when the box allows it (fewer candidates per atom than
  // cells of side >= rlo with a 3^3 stencil); all atoms when the box is below 3 rlo
  int sw = 2, nc = (int)floor(2.0 * L / (rlo * (1.0 + 1e-4)));
  if (nc > NCMAX) nc = NCMAX;
  if (nc < 5) { sw = 1; nc = (int)floor(L / (rlo * (1.0 + 1e-4))); if (nc < 3) nc = 1; }
  const int ncell = nc * nc * nc;
  cx.mic = L < 2.0 * rlo * (1.0 + 1e-3);    // small box: the nearest image of a listed pair may change between builds
  __syncthreads();
  wrap_and_refresh(cx, true);
  __syncthreads();
  const bool assisted = kHelpers && cx.help && help_request(cx, 16 | (cx.mic ? 4 : 0), 0.0, nc, sw);      // a helper searches the rows above 2 * blockDim
  if (nc > 1) bin_cells(cx, nc);
  const float rl2f = (float)(rlo * rlo * invL * invL * (1.0 + 2e-5));
  int over = outer_rows(d, cx, rl2f, nc, sw, 0, assisted ? 2 * nthr : N);
  if (assisted) { help_wait(cx); over |= cx.helpd[3] > 0.0; }
  if (__syncthreads_or(over)) cx.status |= ST_NEIGH;
  cx.L0o = L;
  update_thr(d, cx);
  if (tid == 0) cx.ct[NM_CT_OUTER_BUILDS]++;
}

// inner list = the outer entries currently within rl (float32 test with the stored image; minimum image in MIC mode).
// The owning thread streams its outer row (already ordered by image group, so the inner row inherits the grouping),
// packs the surviving indices four at a time into one 64-bit register (stored XOR the dummy index, so that untouched
// fields read as the dummy atom) and emits whole 8-byte quads into the [quad][atom] layout the force loop reads.
// One pass over the owner's outer quads ([quad][atom] rows walked with byte pointers). Hits are pushed from the top
// into a 128-bit shift register (predicated, no branch per candidate); once per outer quad the four oldest entries
// are emitted as one 8-byte quad. Entries are stored XOR N so that the zero bits of a partial quad read as the dummy.
template <bool MIC>
__device__ void inner_rows(const Dev& d, Ctx& cx, int i0, int i1, double& tot, int& over) {
  const int N = cx.N, Npad = cx.Npad, tid = threadIdx.x, nthr = blockDim.x;
  const double L = cx.L, rl = d.rc + cx.skin, invL = 1.0 / L;
  const float rl2f = (float)(rl * rl * invL * invL * (1.0 + 2e-5));
  const unsigned long long dummy4 = 0x0001000100010001ull * (unsigned long long)N;
  const unsigned sf_s = (unsigned)__cvta_generic_to_shared(cx.sf);
  const unsigned rowbytes = (unsigned)Npad * 8u;
  for (int i = i0 + tid; i < i1; i += nthr) {
    const float4 pi = cx.sf[i];
    const int nqo = cx.onq[i];
    const char* op = reinterpret_cast<const char*>(cx.olist + i);
    const uint8_t* cp = cx.ocode + i;
    char* lq = reinterpret_cast<char*>(cx.list + i);
    unsigned long long hi = 0ull, lo = 0ull;
    int oq = 0, fill = 0, cnt = 0;
    unsigned curcode = 13u;
    auto emit = [&](unsigned long long quad) {
      if (oq < d.maxq) {
        const unsigned long long codebits = ((unsigned long long)(curcode & 7u) << 13) | ((unsigned long long)(curcode >> 3) << 29);
        *reinterpret_cast<unsigned long long*>(lq) = (quad ^ dummy4) | codebits;
        lq += rowbytes;
      } else over = 1;
      oq++;
    };
    uint2 e = *reinterpret_cast<const uint2*>(op);
    unsigned code = *cp;
    for (int q = 0; q < nqo; q++) {
      op += rowbytes; cp += Npad;
      const uint2 en = *reinterpret_cast<const uint2*>(op);      // rows past nqo are allocated and never used
      const unsigned coden = *cp;
      if (code != curcode) {                                     // new image group: close the partial quad
        if (fill) { emit(hi >> (64 - 16 * fill)); fill = 0; }
        curcode = code;
      }
      float px = pi.x, py = pi.y, pz = pi.z;
      if (!MIC) { px -= (float)((int)(code / 9u) - 1); py -= (float)((int)((code / 3u) % 3u) - 1); pz -= (float)((int)(code % 3u) - 1); }
      const unsigned jj[4] = { e.x & 0xffffu, e.x >> 16, e.y & 0xffffu, e.y >> 16 };
#pragma unroll
      for (int t = 0; t < 4; t++) {
        float4 pj;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pj.x), "=f"(pj.y), "=f"(pj.z), "=f"(pj.w) : "r"(sf_s + 16u * jj[t]));
        float dx = px - pj.x, dy = py - pj.y, dz = pz - pj.z;
        if (MIC) { dx = fabsf(dx); dy = fabsf(dy); dz = fabsf(dz); dx = fminf(dx, 1.f - dx); dy = fminf(dy, 1.f - dy); dz = fminf(dz, 1.f - dz); }
        if (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < rl2f) {        // predicated push from the top
          lo = (lo >> 16) | (hi << 48);
          hi = (hi >> 16) | ((unsigned long long)(jj[t] ^ (unsigned)N) << 48);
          fill++; cnt++;
        }
      }
      if (fill >= 4) {                                           // the four oldest entries start 16*fill bits from the top
        const int sh = 16 * fill - 64;                           // 0, 16, 32 or 48
        emit(sh ? (hi << sh) | (lo >> (64 - sh)) : hi);
        fill -= 4;
      }
      e = en; code = coden;
    }
    if (fill) emit(hi >> (64 - 16 * fill));
    cx.nnb[i] = (uint16_t)min(oq, d.maxq);
    tot += cnt;
    cx.gx0[i] = cx.sp[3 * i] * invL; cx.gx0[Npad + i] = cx.sp[3 * i + 1] * invL; cx.gx0[2 * Npad + i] = cx.sp[3 * i + 2] * invL;
  }
}
template <bool MIC>
__device__ void build_inner_t(const Dev& d, Ctx& cx) {
  const bool assisted = kHelpers && cx.help && help_request(cx, 8 | (MIC ? 4 : 0), 0.0);      // a helper regenerates the rows above 2 * blockDim
  double tot = 0.0; int over = 0;
  inner_rows<MIC>(d, cx, 0, assisted ? 2 * (int)blockDim.x : cx.N, tot, over);
  double r[2] = { tot, (double)over };
  bsum<2>(r, cx);
  if (assisted) { help_wait(cx); r[0] += cx.helpd[2]; r[1] += cx.helpd[3]; }      // (pair counts: exact in either order)
  cx.list_pairs = 0.5 * r[0];
  if (r[1] > 0.0) cx.status |= ST_NEIGH;
  cx.L0 = cx.L;
  update_thr(d, cx);
}
__device__ void build_inner(const Dev& d, Ctx& cx) {
  __syncthreads();
  wrap_and_refresh(cx, false);
  __syncthreads();
  if (cx.mic) build_inner_t<true>(d, cx); else build_inner_t<false>(d, cx);
}

// (re)build: make sure the outer list can still supply every pair within rl, then regenerate the inner list
__device__ void build_list(const Dev& d, Ctx& cx) {
  const long long t_build0 = clock64();
  if (cx.in_move && cx.lbuf == cx.sv_lbuf) { cx.lbuf ^= 1; select_list(cx); }   // keep the list of the saved positions
  if (d.small) {
    build_small(d, cx);
    if (threadIdx.x == 0) { cx.ct[NM_CT_LIST_BUILDS]++; const unsigned long long dt = (unsigned long long)(clock64() - t_build0); cx.ct[NM_CT_CLK_BUILD] += dt; cx.ct[NM_CT_CLK_INNER] += dt; }
    return;
  }
  int flag = cx.thro2 < 0.0;
  if (!flag) {
    const double invL = 1.0 / cx.L;
    for (int i = threadIdx.x; i < cx.N; i += blockDim.x)
      flag |= disp2o(cx, i, cx.sp[3 * i], cx.sp[3 * i + 1], cx.sp[3 * i + 2], invL) > cx.thro2;
  }
  if (__syncthreads_or(flag)) { const long long t0 = clock64(); cx.outer_in_move = cx.in_move; build_outer(d, cx); if (threadIdx.x == 0) cx.ct[NM_CT_CLK_OUTER] += (unsigned long long)(clock64() - t0); }
  { const long long t0 = clock64(); build_inner(d, cx); if (threadIdx.x == 0) cx.ct[NM_CT_CLK_INNER] += (unsigned long long)(clock64() - t0); }
  if (threadIdx.x == 0) { cx.ct[NM_CT_LIST_BUILDS]++; cx.ct[NM_CT_CLK_BUILD] += (unsigned long long)(clock64() - t_build0); }
}

// barrier after a position update; rebuilds the list if any thread saw its budget exceeded
__device__ __forceinline__ void sync_and_maybe_build(const Dev& d, Ctx& cx, int flag) {
  if (!d.small && threadIdx.x < 27) {      // image shift vectors k*L of the current box (readers are past a barrier)
    const int c = threadIdx.x;
    cx.sht[3 * c] = (c / 9 - 1) * cx.L; cx.sht[3 * c + 1] = ((c / 3) % 3 - 1) * cx.L; cx.sht[3 * c + 2] = (c % 3 - 1) * cx.L;
  }
  if (__syncthreads_or(flag)) build_list(d, cx);
  if (d.f32) { wrap_and_refresh(cx, false); __syncthreads(); }   // FP32 mode: the force loop reads the float copies
}
// generic pass: is every atom still inside the displacement budget?
__device__ void check_list(const Dev& d, Ctx& cx) {
  int flag = cx.thr2 < 0.0;
  if (!flag) {
    const double invL = 1.0 / cx.L;
    for (int i = threadIdx.x; i < cx.N; i += blockDim.x)
      flag |= disp2(cx, i, cx.sp[3 * (i)], cx.sp[3 * (i) + 1], cx.sp[3 * (i) + 2], invL) > cx.thr2;
  }
  sync_and_maybe_build(d, cx, flag);
}

// ------------------------------------------------------------------ a-1: LJ lj/cut evaluation off the list
// one listed pair, image already resolved (xs = x_i - image shift of the quad): rsq, reciprocal from the
// MUFU.RCP64H seed + one cubic FP64 step, LJ force. The cutoff test is a 64-bit INTEGER compare (positive doubles
// order like integers) and the masking a single select on the high word, so only arithmetic reaches the FP64 pipe:
// 17 FP64-pipe instructions per pair for forces, +4 for energy and virial.
// MIC: small boxes -- the minimum image is taken per pair (high-word test + FP64 subtract) instead.
template <bool EW, bool MIC, bool S32>
__device__ __forceinline__ void lj_pair(double xj, double yj, double zj, double xs, double ys, double zs,
                                        int L_hi, int L_lo, int hL_hi, long long rc2_bits,
                                        double& fx, double& fy, double& fz, int& np, double& e, double& vir) {
  double dx = xs - xj, dy = ys - yj, dz = zs - zj;
  if (MIC) { dx = mic_fast(dx, L_hi, L_lo, hL_hi); dy = mic_fast(dy, L_hi, L_lo, hL_hi); dz = mic_fast(dz, L_hi, L_lo, hL_hi); }
  const double rsq = fma(dz, dz, fma(dy, dy, dx * dx));
  // reciprocal, two variants chosen per kernel instantiation (S32 = the 1024-thread kernels):
  //  * MUFU.RCP64H seed (2^-19.9, measured) + one cubic step (3 DFMA), relative error ~ 2^-59: best when two CTAs
  //    share an SM and are usually in different phases (the conversions of the other variant cost two issue slots each);
  //  * FP32 MUFU.RCP seed + one Newton step (2 DFMA + 2 conversions), 6e-14: best when one 1024-thread CTA keeps all
  //    32 warps in the loop at once and the FP64 pipe itself is the contended unit (N = 4000: 131 vs 136 ms per step).
  //  -DNM_RCP_EXACT adds a quadratic step to the first variant (< 1 ulp).
  double r2inv;
  if (S32) {
    float yf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)rsq));
    const double y = (double)yf, t = fma(-rsq, y, 1.0);
    r2inv = fma(y, t, y);
  } else {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(rsq));
    double t = fma(-rsq, y, 1.0);
    t = fma(t, t, t);
#if defined(NM_RCP_EXACT)
    y = fma(y, t, y);
    t = fma(-rsq, y, 1.0);
#endif
    r2inv = fma(y, t, y);
  }
  const double r6inv = r2inv * r2inv * r2inv;
  double fpair = r6inv * fma(48.0, r6inv, -24.0) * r2inv;
  // outside the cutoff the high word is zeroed: the operand becomes a denormal (< 1e-308) whose products vanish.
  // One predicate drives the select(s) and the pair count.
  int fh = __double2hiint(fpair);
  if (EW) {
    const double ep = r6inv * fma(4.0, r6inv, -4.0);
    int eh = __double2hiint(ep);
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s64 p, %3, %4;\n\tselp.b32 %0, %0, 0, p;\n\tselp.b32 %1, %1, 0, p;\n\t@p add.s32 %2, %2, 1;\n\t}"
        : "+r"(fh), "+r"(eh), "+r"(np) : "l"(__double_as_longlong(rsq)), "l"(rc2_bits));
    fpair = __hiloint2double(fh, __double2loint(fpair));
    e += __hiloint2double(eh, __double2loint(ep));
    vir = fma(rsq, fpair, vir);
  } else {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s64 p, %2, %3;\n\tselp.b32 %0, %0, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
        : "+r"(fh), "+r"(np) : "l"(__double_as_longlong(rsq)), "l"(rc2_bits));
    fpair = __hiloint2double(fh, __double2loint(fpair));
  }
  fx = fma(dx, fpair, fx); fy = fma(dy, fpair, fy); fz = fma(dz, fpair, fz);
}

// the force rows [i0, i1) of one evaluation (thread t owns atoms i0 + t, i0 + t + blockDim, ...): forces (and the kicked
// velocities) go to global memory, the energy / virial / kinetic / pair sums of the rows to the caller's accumulators
template <bool EW, bool KICK, int IMG, bool S32>
__device__ __forceinline__ void force_rows(const Dev& d, Ctx& cx, double dtf, int i0, int i1, double& e, double& vir, double& ke, int& np) {
  constexpr bool MIC = IMG == 1;
  const int Npad = cx.Npad;
  const long long rc2_bits = __double_as_longlong(d.rc * d.rc);
  const int L_hi = __double2hiint(cx.L), L_lo = __double2loint(cx.L), hL_hi = __double2hiint(0.5 * cx.L);
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
#ifdef NM_DEBUG_LOOPCLOCKS
    const long long t_atom0 = clock64();
#endif
    const double xi = cx.sp[3 * i], yi = cx.sp[3 * i + 1], zi = cx.sp[3 * i + 2];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    const int nq = cx.nnb[i];
    // the list is walked with one byte pointer (row stride Npad quads); quads are loaded unconditionally:
    // the allocation carries two spare rows, rows past nq are never used
    const char* lp = reinterpret_cast<const char*>(cx.list + i);
    const unsigned stride = (unsigned)Npad * 8u, sp_s = (unsigned)__cvta_generic_to_shared(cx.sp);
    auto do_quad = [&](const uint2 cur) {
      double xs = xi, ys = yi, zs = zi;
      if (IMG == 0) {
        const unsigned code = ((cur.x >> 13) & 7u) | ((cur.x >> 26) & 0x38u);
        const double* sh = cx.sht + 3 * code;
        xs -= sh[0]; ys -= sh[1]; zs -= sh[2];
      }
      // explicit 32-bit shared addresses (one IMAD per neighbour). ptxas places each gather next to its use
      // whatever the source order (volatile loads up front were tried: same schedule)
      const unsigned a0 = sp_s + 24u * (IMG == 0 ? cur.x & 0x1fffu : cur.x & 0xffffu), a1 = sp_s + 24u * (IMG == 0 ? (cur.x >> 16) & 0x1fffu : cur.x >> 16),
                     a2 = sp_s + 24u * (cur.y & 0xffffu), a3 = sp_s + 24u * (cur.y >> 16);
      double p[12];
#ifdef NM_FAKE_GATHER   // timing experiment only (wrong physics): conflict-free gathers, same instruction stream
      { const unsigned fk = 24u * ((threadIdx.x + (cur.x & 63u)) & 511u);
        lds_f64x3(sp_s + fk + (a0 & 0u), p[0], p[1], p[2]); lds_f64x3(sp_s + fk + 24u + (a1 & 0u), p[3], p[4], p[5]);
        lds_f64x3(sp_s + fk + 48u + (a2 & 0u), p[6], p[7], p[8]); lds_f64x3(sp_s + fk + 72u + (a3 & 0u), p[9], p[10], p[11]); }
#elif defined(NM_PLAIN_GATHER)
      { const double *q0 = cx.sp + (a0 - sp_s) / 8u, *q1 = cx.sp + (a1 - sp_s) / 8u, *q2 = cx.sp + (a2 - sp_s) / 8u, *q3 = cx.sp + (a3 - sp_s) / 8u;
        p[0] = q0[0]; p[1] = q0[1]; p[2] = q0[2]; p[3] = q1[0]; p[4] = q1[1]; p[5] = q1[2];
        p[6] = q2[0]; p[7] = q2[1]; p[8] = q2[2]; p[9] = q3[0]; p[10] = q3[1]; p[11] = q3[2]; }
#else
      lds_f64x3(a0, p[0], p[1], p[2]); lds_f64x3(a1, p[3], p[4], p[5]); lds_f64x3(a2, p[6], p[7], p[8]); lds_f64x3(a3, p[9], p[10], p[11]);
#endif
      lj_pair<EW, MIC, S32>(p[0], p[1], p[2], xs, ys, zs, L_hi, L_lo, hL_hi, rc2_bits, fx, fy, fz, np, e, vir);
      lj_pair<EW, MIC, S32>(p[3], p[4], p[5], xs, ys, zs, L_hi, L_lo, hL_hi, rc2_bits, fx, fy, fz, np, e, vir);
      lj_pair<EW, MIC, S32>(p[6], p[7], p[8], xs, ys, zs, L_hi, L_lo, hL_hi, rc2_bits, fx, fy, fz, np, e, vir);
      lj_pair<EW, MIC, S32>(p[9], p[10], p[11], xs, ys, zs, L_hi, L_lo, hL_hi, rc2_bits, fx, fy, fz, np, e, vir);
    };
#if NM_UNR == 2
    // two quads per iteration: the same operations in the same order (the force sums stay sequential), twice the
    // independent work in flight per warp -- for CTAs that have their SM to themselves
    const uint2 dq = make_uint2((unsigned)N * 0x10001u, (unsigned)N * 0x10001u);
    uint2 c0 = *reinterpret_cast<const uint2*>(lp), c1 = *reinterpret_cast<const uint2*>(lp + stride);
    lp += 2 * stride;
    for (int q = 0; q < nq; q += 2) {
      const uint2 n0 = *reinterpret_cast<const uint2*>(lp), n1 = *reinterpret_cast<const uint2*>(lp + stride);
      lp += 2 * stride;
      if (q + 1 >= nq) c1 = dq;
      do_quad(c0); do_quad(c1);
      c0 = n0; c1 = n1;
    }
#else
    uint2 cur = *reinterpret_cast<const uint2*>(lp);
    lp += stride;
    // ptxas sinks the load of the next quad to the end of the iteration (its registers serve as temporaries in between), so
    // the load itself hides nothing: the row list_pf rows further down is pulled into L1 by a prefetch (no destination
    // register, nothing to sink), and the late load hits L1 instead of waiting for L2 / HBM
    if (d.list_pf >= 0) {
      const size_t pf_off = (size_t)d.list_pf * stride;
      for (int q = 0; q < nq; q++) {
        asm volatile("prefetch.global.L1 [%0];" :: "l"(lp + pf_off));
        const uint2 nxt = *reinterpret_cast<const uint2*>(lp);
        lp += stride;
        do_quad(cur);
        cur = nxt;
      }
    } else {
      for (int q = 0; q < nq; q++) {
        const uint2 nxt = *reinterpret_cast<const uint2*>(lp);   // for the next iteration
        lp += stride;
        do_quad(cur);
        cur = nxt;
      }
    }
#endif
#ifdef NM_DEBUG_LOOPCLOCKS   // per-atom loop clocks of thread 0 (tools/probe.py); compiled out of the product build
    if (threadIdx.x == 0) { cx.ct[NM_CT_DBG_LOOPCLK] += (unsigned long long)(clock64() - t_atom0); cx.ct[NM_CT_DBG_LOOPIT] += (unsigned long long)nq; }
#endif
    cx.gf[i] = fx; cx.gf[Npad + i] = fy; cx.gf[2 * Npad + i] = fz;
    if (KICK) {
      const double vx = fma(dtf, fx, cx.gv[i]), vy = fma(dtf, fy, cx.gv[Npad + i]), vz = fma(dtf, fz, cx.gv[2 * Npad + i]);
      cx.gv[i] = vx; cx.gv[Npad + i] = vy; cx.gv[2 * Npad + i] = vz;
      if (EW) ke += vx * vx + vy * vy + vz * vz;
    }
  }
}

// EW: also energy / virial / pair count (block-reduced into out[0..2]); KICK: fused second velocity-Verlet
// half kick v += dtf*f of the owning thread, KE returned in out[3] when EW.
// Ends with a barrier: shared positions may be rewritten afterwards.
// IMG: how a list entry names the periodic image of its atom. 0: one image code per quad (LARGE mode), 1: per-pair
// minimum image (small boxes), 2: the entry is the index of the copy to use (SMALL mode ghost atoms: nothing to do).
template <bool EW, bool KICK, int IMG, bool S32>
__device__ void eval_forces_t(const Dev& d, Ctx& cx, double dtf, double (&out)[4]) {
  const int N = cx.N;
  const long long t_eval0 = clock64();
  double e = 0.0, vir = 0.0, ke = 0.0, e1 = 0.0, vir1 = 0.0, ke1 = 0.0, np1 = 0.0; int np = 0;
  bool assisted = false;
  if (S32 && IMG != 2) {
    // rows above the split are summed separately (see the helper note above); the split depends on N alone
    const int Ns = (!d.small && N > 2 * (int)blockDim.x) ? 2 * (int)blockDim.x : N;
    if (kHelpers && Ns < N && cx.help) assisted = help_request(cx, (EW ? 1 : 0) | (KICK ? 2 : 0) | (IMG == 1 ? 4 : 0), dtf);
    const int nparts = (Ns < N && !assisted) ? 2 : 1;
#pragma unroll 1
    for (int part = 0; part < nparts; part++) {
      double pe = 0.0, pv = 0.0, pk = 0.0;
      force_rows<EW, KICK, IMG, S32>(d, cx, dtf, part ? Ns : 0, part ? N : Ns, pe, pv, pk, np);
      if (part == 0) { e = pe; vir = pv; ke = pk; } else { e1 = pe; vir1 = pv; ke1 = pk; }
    }
    if (kHelpers && assisted) {
      help_wait(cx);
#if !defined(NM_DEBUG_CLOCKS) && !defined(NM_DEBUG_SMID)     // (the debug builds keep other figures in this column)
      if (threadIdx.x == 0) cx.ct[NM_CT_HELPED_EVALS]++;
#endif
      if (EW) {
        const int nt = (int)blockDim.x, t = (int)threadIdx.x;
        e1 = cx.hpart[t]; vir1 = cx.hpart[nt + t]; np1 = cx.hpart[2 * nt + t]; ke1 = cx.hpart[3 * nt + t];
      } else if (threadIdx.x == 0) atomicAdd(cx.s_pairs, *reinterpret_cast<const unsigned long long*>(cx.help + 6));
    }
  } else {
    force_rows<EW, KICK, IMG, S32>(d, cx, dtf, 0, N, e, vir, ke, np);
  }
  if (EW) {
    double r[4] = { 0.5 * (e + e1), 0.5 * (vir + vir1), (double)np + np1, ke + ke1 };
    bsum<4>(r, cx);
    out[0] = r[0]; out[1] = r[1]; out[2] = 0.5 * r[2]; out[3] = 0.5 * d.mass * r[3];
    if (threadIdx.x == 0) {
      cx.ct[NM_CT_FORCE_EVALS]++; cx.ct[NM_CT_PAIRS_FULL] += (unsigned long long)out[2];
      cx.ct[NM_CT_LIST_PAIRS] += (unsigned long long)cx.list_pairs;
      cx.ct[NM_CT_CLK_EVAL] += (unsigned long long)(clock64() - t_eval0);
    }
  } else {
    np = __reduce_add_sync(0xffffffffu, np);
    if ((threadIdx.x & 31) == 0) atomicAdd(cx.s_pairs, (unsigned long long)np);
    __syncthreads();
    if (threadIdx.x == 0) { cx.ct[NM_CT_FORCE_EVALS]++; cx.ct[NM_CT_LIST_PAIRS] += (unsigned long long)cx.list_pairs;
                            cx.ct[NM_CT_CLK_EVAL] += (unsigned long long)(clock64() - t_eval0); }
  }
}
// FP32 mode (nm_config.precision = 32; tolerance 1e-5 relative): the pair arithmetic runs on the FP32 pipe from the
// float32 FRACTIONAL copies of the positions (one 16-byte gather per pair), image shifts are -1/0/+1 in box units,
// forces are accumulated per atom in float and scaled by L at the end; energy / virial partial sums are per-thread
// floats reduced in double. State, integration and Metropolis arithmetic stay FP64.
template <bool EW, bool KICK, bool MIC>
__device__ void eval_forces_f32(const Dev& d, Ctx& cx, double dtf, double (&out)[4]) {
  const int N = cx.N, Npad = cx.Npad;
  const float Lf = (float)cx.L, L2f = Lf * Lf, rc2f = (float)(d.rc * d.rc), magic = 12582912.f;
  float e = 0.f, vir = 0.f; double ke = 0.0; int np = 0;
  const long long t_eval0 = clock64();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float4 pi = cx.sf[i];
    float fx = 0.f, fy = 0.f, fz = 0.f;
    const int nq = cx.nnb[i];
    const char* lp = reinterpret_cast<const char*>(cx.list + i);
    const unsigned stride = (unsigned)Npad * 8u;
    uint2 cur = *reinterpret_cast<const uint2*>(lp);
    lp += stride;
    for (int q = 0; q < nq; q++) {
      const uint2 nxt = *reinterpret_cast<const uint2*>(lp);
      lp += stride;
      float xs = pi.x, ys = pi.y, zs = pi.z;
      if (!MIC) {
        const int code = (int)(((cur.x >> 13) & 7u) | ((cur.x >> 26) & 0x38u));
        xs -= (float)(code / 9 - 1); ys -= (float)((code / 3) % 3 - 1); zs -= (float)(code % 3 - 1);
      }
      const unsigned jj[4] = { cur.x & 0x1fffu, (cur.x >> 16) & 0x1fffu, cur.y & 0xffffu, cur.y >> 16 };
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const float4 pj = cx.sf[jj[t]];
        float dx = xs - pj.x, dy = ys - pj.y, dz = zs - pj.z;
        if (MIC) { dx -= __fadd_rn(__fadd_rn(dx, magic), -magic); dy -= __fadd_rn(__fadd_rn(dy, magic), -magic); dz -= __fadd_rn(__fadd_rn(dz, magic), -magic); }
        const float rsq = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, pj.w))) * L2f;
        const bool in = rsq < rc2f;
        const float r2inv = __frcp_rn(rsq);
        const float r6inv = r2inv * r2inv * r2inv;
        const float fpair = in ? r6inv * fmaf(48.f, r6inv, -24.f) * r2inv : 0.f;
        fx = fmaf(dx, fpair, fx); fy = fmaf(dy, fpair, fy); fz = fmaf(dz, fpair, fz);
        np += in;
        if (EW) { e += in ? r6inv * fmaf(4.f, r6inv, -4.f) : 0.f; vir = fmaf(rsq, fpair, vir); }
      }
      cur = nxt;
    }
    const double Fx = (double)(fx * Lf), Fy = (double)(fy * Lf), Fz = (double)(fz * Lf);
    cx.gf[i] = Fx; cx.gf[Npad + i] = Fy; cx.gf[2 * Npad + i] = Fz;
    if (KICK) {
      const double vx = fma(dtf, Fx, cx.gv[i]), vy = fma(dtf, Fy, cx.gv[Npad + i]), vz = fma(dtf, Fz, cx.gv[2 * Npad + i]);
      cx.gv[i] = vx; cx.gv[Npad + i] = vy; cx.gv[2 * Npad + i] = vz;
      if (EW) ke += vx * vx + vy * vy + vz * vz;
    }
  }
  if (EW) {
    double r[4] = { 0.5 * (double)e, 0.5 * (double)vir, (double)np, ke };
    bsum<4>(r, cx);
    out[0] = r[0]; out[1] = r[1]; out[2] = 0.5 * r[2]; out[3] = 0.5 * d.mass * r[3];
    if (threadIdx.x == 0) {
      cx.ct[NM_CT_FORCE_EVALS]++; cx.ct[NM_CT_PAIRS_FULL] += (unsigned long long)out[2];
      cx.ct[NM_CT_LIST_PAIRS] += (unsigned long long)cx.list_pairs;
      cx.ct[NM_CT_CLK_EVAL] += (unsigned long long)(clock64() - t_eval0);
    }
  } else {
    np = __reduce_add_sync(0xffffffffu, np);
    if ((threadIdx.x & 31) == 0) atomicAdd(cx.s_pairs, (unsigned long long)np);
    __syncthreads();
    if (threadIdx.x == 0) { cx.ct[NM_CT_FORCE_EVALS]++; cx.ct[NM_CT_LIST_PAIRS] += (unsigned long long)cx.list_pairs;
                            cx.ct[NM_CT_CLK_EVAL] += (unsigned long long)(clock64() - t_eval0); }
  }
}

// S32 marks the 1024-thread kernels: the only ones that can meet a LARGE system (N > NSMALL)
template <bool EW, bool KICK, bool S32>
__device__ __forceinline__ void eval_forces(const Dev& d, Ctx& cx, double dtf, double (&out)[4]) {
  if (d.f32) {
    if (cx.mic) eval_forces_f32<EW, KICK, true>(d, cx, dtf, out);           // SMALL mode: always (no float ghost copies)
    else eval_forces_f32<EW, KICK, false>(d, cx, dtf, out);
  } else if (cx.mic) eval_forces_t<EW, KICK, 1, S32>(d, cx, dtf, out);
  else if (!S32 || d.small) eval_forces_t<EW, KICK, 2, S32>(d, cx, dtf, out);
  else eval_forces_t<EW, KICK, 0, S32>(d, cx, dtf, out);
}

// the acceptance rule shared by all moves (lammps_remcmc.py:487-500, 532-547, 578-593, 623-638)
__device__ __forceinline__ bool metropolis(double de, const Rng& r, uint32_t index, uint32_t purpose) {
  const double m = exp(-de);
  if (isinf(m) || isnan(m)) return false;
  const double u = rng_uniform(r, index, purpose);
  return u <= (m < 1.0 ? m : 1.0);
}
// thread 0 decides, everybody learns
__device__ __forceinline__ bool broadcast_flag(Ctx& cx, bool v) {
  __syncthreads();
  if (threadIdx.x == 0) cx.ibc[0] = v;
  __syncthreads();
  return cx.ibc[0] != 0;
}

struct Energy { double pe, w; };

__device__ void save_xf(Ctx& cx, bool with_v) {
  cx.in_move = 1; cx.sv_lbuf = cx.lbuf; cx.sv_L0 = cx.L0; cx.sv_mic = cx.mic; cx.sv_list_pairs = cx.list_pairs; cx.outer_in_move = 0;
  for (int i = threadIdx.x; i < cx.N; i += blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int o = a * cx.Npad + i;
      cx.gxs[o] = cx.sp[3 * i + a];
      cx.gfs[o] = cx.gf[o];
      if (with_v) cx.gvs[o] = cx.gv[o];
    }
  }
}
// rejected move: back to the saved positions (cx.L already restored by the caller) and to the list that was current when
// they were saved; an outer list rebuilt inside the move refers to re-wrapped positions and is dropped
__device__ void restore_xf(const Dev& d, Ctx& cx, bool with_v) {
  if (cx.lbuf != cx.sv_lbuf) {
    cx.lbuf = cx.sv_lbuf; select_list(cx);
    cx.L0 = cx.sv_L0; cx.mic = cx.sv_mic; cx.list_pairs = cx.sv_list_pairs;
    if (cx.outer_in_move) cx.L0o = -1.0;
    cx.ghost = d.small && !cx.mic && cx.L0 > 0.0;
    if (cx.ghost) load_ginfo(cx);                 // thread i reloads entry i and is its only reader until the next build
  }
  cx.in_move = 0;
  update_thr(d, cx);
  for (int i = threadIdx.x; i < cx.N; i += blockDim.x) {
    store_pos(cx, i, cx.gxs[i], cx.gxs[cx.Npad + i], cx.gxs[2 * cx.Npad + i]);
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int o = a * cx.Npad + i;
      cx.gf[o] = cx.gfs[o];
      if (with_v) cx.gv[o] = cx.gvs[o];
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ a-7 bulk_position_mc (lammps_remcmc.py:477-502)
template <bool S32>
__device__ void bulk_position_mc(const Dev& d, Ctx& cx, const Rng& r, double et, double dxs, Energy& en, double* cnt) {
  save_xf(cx, false);
  const double dmax = d.text_rounding ? round6(dxs * d.lat) : dxs * d.lat;     // 'displace_atoms all random %f'
  const double invL = 1.0 / cx.L;
  int flag = cx.thr2 < 0.0;
  for (int i = threadIdx.x; i < cx.N; i += blockDim.x) {
    double u[3]; rng_uniform3(r, (uint32_t)i, P_BULK_DISP, u);
    const double x = cx.sp[3 * i] + dmax * 2.0 * (u[0] - 0.5), y = cx.sp[3 * i + 1] + dmax * 2.0 * (u[1] - 0.5),
                 z = cx.sp[3 * i + 2] + dmax * 2.0 * (u[2] - 0.5);
    store_pos(cx, i, x, y, z);
    if (!flag) flag = disp2(cx, i, x, y, z, invL) > cx.thr2;
  }
  sync_and_maybe_build(d, cx, flag);
  double o[4]; eval_forces<true, false, S32>(d, cx, 0.0, o);
  bool acc = false;
  if (threadIdx.x == 0) {
    const double de = o[0] / et - en.pe / et;
    acc = metropolis(de, r, 0, P_BULK_ACC);
    cnt[0] += 1.0; if (acc) cnt[1] += 1.0;
    cx.ct[NM_CT_PMC_MOVES]++; cx.ct[NM_CT_PMC_TRIALS]++;
  }
  acc = broadcast_flag(cx, acc);
  if (acc) { en.pe = o[0]; en.w = o[1]; cx.in_move = 0; } else restore_xf(d, cx, false);
}

// ------------------------------------------------------------------ a-6 volume_mc (lammps_remcmc.py:552-595)
template <bool S32>
__device__ void volume_mc(const Dev& d, Ctx& cx, const Rng& r, double et, double pf, double dvs, Energy& en, double* cnt) {
  save_xf(cx, false);
  const double box = cx.L;
  if (threadIdx.x == 0) {
    const double vol = pow(box, 3.0);
    const double volnew = exp(log(vol) + 2 * (rng_uniform(r, 0, P_VMC_PROP) - 0.5) * dvs);
    const double boxnew = cbrt(volnew);
    cx.bc[0] = vol; cx.bc[1] = volnew; cx.bc[2] = boxnew / box;
    cx.bc[3] = d.text_rounding ? round6(boxnew) : boxnew;                      // 'change_box ... %f'
  }
  __syncthreads();
  const double vol = cx.bc[0], volnew = cx.bc[1], scale = cx.bc[2], Lnew = cx.bc[3];
  const bool box_ok = Lnew >= 2.0 * d.rc * (1.0 + 1e-5);
  bool acc = false;
  double o[4] = { 0, 0, 0, 0 };
  if (box_ok) {
    cx.L = Lnew; update_thr(d, cx);
    const double invL = 1.0 / Lnew, invLold = 1.0 / box, dround = scale * box - Lnew;
    int flag = cx.thr2 < 0.0;
    for (int i = threadIdx.x; i < cx.N; i += blockDim.x) {
      // the reference scales WRAPPED coordinates by boxnew/box and then periodises with the '%f'-rounded box:
      // an atom represented k boxes away from [0,L) must land on the same lattice image, hence the k*(boxnew-Lnew)
      const double kx = floor(cx.sp[3 * i] * invLold), ky = floor(cx.sp[3 * i + 1] * invLold), kz = floor(cx.sp[3 * i + 2] * invLold);
      const double x = scale * cx.sp[3 * i] - kx * dround, y = scale * cx.sp[3 * i + 1] - ky * dround, z = scale * cx.sp[3 * i + 2] - kz * dround;
      store_pos(cx, i, x, y, z);
      if (!flag) flag = disp2(cx, i, x, y, z, invL) > cx.thr2;
    }
    sync_and_maybe_build(d, cx, flag);
    eval_forces<true, false, S32>(d, cx, 0.0, o);
  } else {
    cx.status |= ST_BOX;
  }
  if (threadIdx.x == 0) {
    if (box_ok) {
      const double dh = (o[0] / et - en.pe / et) + pf * (volnew - vol) - (cx.N + 1) * log(volnew / vol);
      acc = metropolis(dh, r, 0, P_VMC_ACC);
    }
    cnt[2] += 1.0; if (acc) cnt[3] += 1.0;
    cx.ct[NM_CT_VMC_MOVES]++;
  }
  acc = broadcast_flag(cx, acc);
  if (acc) { en.pe = o[0]; en.w = o[1]; cx.in_move = 0; }
  else { cx.L = box; if (box_ok) restore_xf(d, cx, false); else { cx.in_move = 0; update_thr(d, cx); } }
}

// ------------------------------------------------------------------ a-4 velocity create + zero linear + zero angular
// LAMMPS 'velocity all create T seed dist gaussian' (loop all, mom yes, rot no), then 'zero linear',
// 'zero angular' (lammps_remcmc.py:604-606). Returns KE = 0.5 m sum v^2.
// one atom per thread (N <= blockDim.x): the same arithmetic with the atom's velocity and wrapped position held in
// registers across the six reductions (one global store at the end instead of five read-modify-write passes)
__device__ double velocity_create_1(const Dev& d, Ctx& cx, const Rng& r, double t_target) {
  const int N = cx.N, Npad = cx.Npad, i = threadIdx.x;
  const bool own = i < N;
  const double m = d.mass, inv = 1.0 / sqrt(m), inv_mN = 1.0 / (m * cx.N);
  double vx = 0.0, vy = 0.0, vz = 0.0, px = 0.0, py = 0.0, pz = 0.0;
  double s[4] = { 0, 0, 0, 0 };
  if (own) {
    double g[3]; rng_gauss3(r, (uint32_t)i, P_HMC_VEL, g);
    vx = g[0] * inv; vy = g[1] * inv; vz = g[2] * inv;
    s[0] += m * vx; s[1] += m * vy; s[2] += m * vz;
    px = wrapg(cx.sp[3 * i], cx.L); py = wrapg(cx.sp[3 * i + 1], cx.L); pz = wrapg(cx.sp[3 * i + 2], cx.L);
  }
  bsum<4>(s, cx);
  double vcm[3] = { s[0] * inv_mN, s[1] * inv_mN, s[2] * inv_mN };
  double t[1] = { 0 };
  if (own) { vx -= vcm[0]; vy -= vcm[1]; vz -= vcm[2]; t[0] += vx * vx + vy * vy + vz * vz; }
  bsum<1>(t, cx);
  const double tinst = m * t[0] / (3.0 * N - 3.0), fac = sqrt(t_target / tinst);
  s[0] = s[1] = s[2] = s[3] = 0;
  if (own) { vx *= fac; vy *= fac; vz *= fac; s[0] += m * vx; s[1] += m * vy; s[2] += m * vz; }
  bsum<4>(s, cx);                                                   // 'velocity all zero linear'
  vcm[0] = s[0] * inv_mN; vcm[1] = s[1] * inv_mN; vcm[2] = s[2] * inv_mN;
  double xc[3] = { 0, 0, 0 };
  if (own) { vx -= vcm[0]; vy -= vcm[1]; vz -= vcm[2]; xc[0] += m * px; xc[1] += m * py; xc[2] += m * pz; }
  bsum<3>(xc, cx);                                                  // 'velocity all zero angular'
  xc[0] *= inv_mN; xc[1] *= inv_mN; xc[2] *= inv_mN;
  double a[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };   // L(3), Ixx Iyy Izz Ixy Iyz Ixz
  const double dx = px - xc[0], dy = py - xc[1], dz = pz - xc[2];
  if (own) {
    a[0] += m * (dy * vz - dz * vy); a[1] += m * (dz * vx - dx * vz); a[2] += m * (dx * vy - dy * vx);
    a[3] += m * (dy * dy + dz * dz); a[4] += m * (dx * dx + dz * dz); a[5] += m * (dx * dx + dy * dy);
    a[6] -= m * dx * dy; a[7] -= m * dy * dz; a[8] -= m * dx * dz;
  }
  bsum<9>(a, cx);
  const double I00 = a[3], I11 = a[4], I22 = a[5], I01 = a[6], I12 = a[7], I02 = a[8];
  const double det = I00 * (I11 * I22 - I12 * I12) - I01 * (I01 * I22 - I12 * I02) + I02 * (I01 * I12 - I11 * I02);
  double w0 = 0, w1 = 0, w2 = 0;
  if (det > 0.0) {
    const double id = 1.0 / det;
    const double i00 = (I11 * I22 - I12 * I12) * id, i01 = -(I01 * I22 - I02 * I12) * id, i02 = (I01 * I12 - I02 * I11) * id;
    const double i11 = (I00 * I22 - I02 * I02) * id, i12 = -(I00 * I12 - I02 * I01) * id, i22 = (I00 * I11 - I01 * I01) * id;
    w0 = i00 * a[0] + i01 * a[1] + i02 * a[2];
    w1 = i01 * a[0] + i11 * a[1] + i12 * a[2];
    w2 = i02 * a[0] + i12 * a[1] + i22 * a[2];
  }
  t[0] = 0;
  if (own) {
    vx -= (w1 * dz - w2 * dy); vy -= (w2 * dx - w0 * dz); vz -= (w0 * dy - w1 * dx);
    cx.gv[i] = vx; cx.gv[Npad + i] = vy; cx.gv[2 * Npad + i] = vz;
    t[0] += vx * vx + vy * vy + vz * vz;
  }
  bsum<1>(t, cx);
  return 0.5 * m * t[0];
}

__device__ double velocity_create(const Dev& d, Ctx& cx, const Rng& r, double t_target) {
  if (cx.N <= (int)blockDim.x) return velocity_create_1(d, cx, r, t_target);
  const int N = cx.N, Npad = cx.Npad;
  const double m = d.mass, inv = 1.0 / sqrt(m), inv_mN = 1.0 / (m * cx.N);
  double s[4] = { 0, 0, 0, 0 };
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double g[3]; rng_gauss3(r, (uint32_t)i, P_HMC_VEL, g);
    const double vx = g[0] * inv, vy = g[1] * inv, vz = g[2] * inv;
    cx.gv[i] = vx; cx.gv[Npad + i] = vy; cx.gv[2 * Npad + i] = vz;
    s[0] += m * vx; s[1] += m * vy; s[2] += m * vz;
  }
  bsum<4>(s, cx);
  double vcm[3] = { s[0] * inv_mN, s[1] * inv_mN, s[2] * inv_mN };
  double t[1] = { 0 };
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double vx = cx.gv[i] - vcm[0], vy = cx.gv[Npad + i] - vcm[1], vz = cx.gv[2 * Npad + i] - vcm[2];
    cx.gv[i] = vx; cx.gv[Npad + i] = vy; cx.gv[2 * Npad + i] = vz;
    t[0] += vx * vx + vy * vy + vz * vz;
  }
  bsum<1>(t, cx);
  const double tinst = m * t[0] / (3.0 * N - 3.0), fac = sqrt(t_target / tinst);
  s[0] = s[1] = s[2] = s[3] = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double vx = cx.gv[i] * fac, vy = cx.gv[Npad + i] * fac, vz = cx.gv[2 * Npad + i] * fac;
    cx.gv[i] = vx; cx.gv[Npad + i] = vy; cx.gv[2 * Npad + i] = vz;
    s[0] += m * vx; s[1] += m * vy; s[2] += m * vz;
  }
  bsum<4>(s, cx);                                                   // 'velocity all zero linear'
  vcm[0] = s[0] * inv_mN; vcm[1] = s[1] * inv_mN; vcm[2] = s[2] * inv_mN;
  double xc[3] = { 0, 0, 0 };
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    cx.gv[i] -= vcm[0]; cx.gv[Npad + i] -= vcm[1]; cx.gv[2 * Npad + i] -= vcm[2];
    xc[0] += m * wrapg(cx.sp[3 * i], cx.L); xc[1] += m * wrapg(cx.sp[3 * i + 1], cx.L); xc[2] += m * wrapg(cx.sp[3 * i + 2], cx.L);
  }
  bsum<3>(xc, cx);                                                  // 'velocity all zero angular'
  xc[0] *= inv_mN; xc[1] *= inv_mN; xc[2] *= inv_mN;
  double a[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };   // L(3), Ixx Iyy Izz Ixy Iyz Ixz
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double dx = wrapg(cx.sp[3 * i], cx.L) - xc[0], dy = wrapg(cx.sp[3 * i + 1], cx.L) - xc[1], dz = wrapg(cx.sp[3 * i + 2], cx.L) - xc[2];
    const double vx = cx.gv[i], vy = cx.gv[Npad + i], vz = cx.gv[2 * Npad + i];
    a[0] += m * (dy * vz - dz * vy); a[1] += m * (dz * vx - dx * vz); a[2] += m * (dx * vy - dy * vx);
    a[3] += m * (dy * dy + dz * dz); a[4] += m * (dx * dx + dz * dz); a[5] += m * (dx * dx + dy * dy);
    a[6] -= m * dx * dy; a[7] -= m * dy * dz; a[8] -= m * dx * dz;
  }
  bsum<9>(a, cx);
  const double I00 = a[3], I11 = a[4], I22 = a[5], I01 = a[6], I12 = a[7], I02 = a[8];
  const double det = I00 * (I11 * I22 - I12 * I12) - I01 * (I01 * I22 - I12 * I02) + I02 * (I01 * I12 - I11 * I02);
  double w0 = 0, w1 = 0, w2 = 0;
  if (det > 0.0) {
    const double id = 1.0 / det;
    const double i00 = (I11 * I22 - I12 * I12) * id, i01 = -(I01 * I22 - I02 * I12) * id, i02 = (I01 * I12 - I02 * I11) * id;
    const double i11 = (I00 * I22 - I02 * I02) * id, i12 = -(I00 * I12 - I02 * I01) * id, i22 = (I00 * I11 - I01 * I01) * id;
    w0 = i00 * a[0] + i01 * a[1] + i02 * a[2];
    w1 = i01 * a[0] + i11 * a[1] + i12 * a[2];
    w2 = i02 * a[0] + i12 * a[1] + i22 * a[2];
  }
  t[0] = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double dx = wrapg(cx.sp[3 * i], cx.L) - xc[0], dy = wrapg(cx.sp[3 * i + 1], cx.L) - xc[1], dz = wrapg(cx.sp[3 * i + 2], cx.L) - xc[2];
    const double vx = cx.gv[i] - (w1 * dz - w2 * dy), vy = cx.gv[Npad + i] - (w2 * dx - w0 * dz), vz = cx.gv[2 * Npad + i] - (w0 * dy - w1 * dx);
    cx.gv[i] = vx; cx.gv[Npad + i] = vy; cx.gv[2 * Npad + i] = vz;
    t[0] += vx * vx + vy * vy + vz * vz;
  }
  bsum<1>(t, cx);
  return 0.5 * m * t[0];
}

// ------------------------------------------------------------------ a-3 / a-5 hamiltonian_mc (lammps_remcmc.py:598-640)
template <bool S32>
__device__ void hamiltonian_mc(const Dev& d, Ctx& cx, const Rng& r, double et, double t_vel, double dts, Energy& en, double* cnt) {
  const int N = cx.N, Npad = cx.Npad;
  const long long t_vel0 = clock64();
  const double ke0 = velocity_create(d, cx, r, t_vel);
  if (threadIdx.x == 0) cx.ct[NM_CT_CLK_VEL] += (unsigned long long)(clock64() - t_vel0);
  save_xf(cx, true);
  const double dt = d.text_rounding ? round6(dts) : dts;                        // 'timestep %f'
  const double dtf = 0.5 * dt / d.mass;
  const double etot = en.pe / et + ke0 / et;
  double o[4] = { 0, 0, 0, 0 };
  for (int st = 0; st < d.nstps; st++) {
    const double invL = 1.0 / cx.L;
    int flag = cx.thr2 < 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double vx = fma(dtf, cx.gf[i], cx.gv[i]), vy = fma(dtf, cx.gf[Npad + i], cx.gv[Npad + i]),
                   vz = fma(dtf, cx.gf[2 * Npad + i], cx.gv[2 * Npad + i]);
      cx.gv[i] = vx; cx.gv[Npad + i] = vy; cx.gv[2 * Npad + i] = vz;
      const double x = fma(dt, vx, cx.sp[3 * i]), y = fma(dt, vy, cx.sp[3 * i + 1]), z = fma(dt, vz, cx.sp[3 * i + 2]);
      store_pos(cx, i, x, y, z);
      if (!flag) flag = disp2(cx, i, x, y, z, invL) > cx.thr2;
    }
    sync_and_maybe_build(d, cx, flag);
    if (st == d.nstps - 1) eval_forces<true, true, S32>(d, cx, dtf, o);
    else eval_forces<false, true, S32>(d, cx, dtf, o);
  }
  bool acc = false;
  if (threadIdx.x == 0) {
    const double etotnew = o[0] / et + o[3] / et;
    acc = metropolis(etotnew - etot, r, 0, P_HMC_ACC);
    cnt[4] += 1.0; if (acc) cnt[5] += 1.0;
    cx.ct[NM_CT_HMC_MOVES]++; cx.ct[NM_CT_HMC_ATOM_STEPS] += (unsigned long long)N * d.nstps;
  }
  acc = broadcast_flag(cx, acc);
  if (acc) { en.pe = o[0]; en.w = o[1]; cx.in_move = 0; } else restore_xf(d, cx, true);
}

// clock read that the compiler cannot move across barriers or memory operations (debug timing)
__device__ __forceinline__ long long clk_fenced() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }

// Step (C) of iter_position_mc (see there): ordered commit of a window of speculative single-atom trials by ONE warp,
// lane l holding trial k + l. Everything that does not depend on the acceptances is done for all lanes at once, before
// and after the serial loop, which is left with: broadcast the running dE of trial t, compare it with the trial's
// threshold band, and on acceptance add c(t, l) to the later lanes.
//  before: the window is cut at the first trial that was not evaluated or whose list column is not guaranteed complete
//          if EVERY earlier trial of the window is accepted (prefix maximum of the displacements: a bound of the
//          running maximum at its turn, so the test is conservative; trial 0 was tested against the exact value);
//  after : the accepted lanes move their atoms (all at once), the displacement maxima take the accepted trials in.
// win: window records (WS doubles per trial, see WS_*); corr: c(a, b) at [a * CS + b]; res: out {next k, need rebuild};
// resd: in/out {umax, umaxo}, accumulators {trials, accepted, visited}.
constexpr int WS = 12;
enum { WS_XN = 0, WS_YN, WS_ZN, WS_DE, WS_LO, WS_HI, WS_UACC, WS_UN, WS_UNO, WS_FLAG_VIS, WS_SRC };
// the reference's rule (lammps_remcmc.py:532-547) for a dE inside the threshold band (or not finite, or so negative
// that exp(-dE/et) could overflow: inf => reject without drawing)
__device__ __noinline__ bool decide_exact(double de, double et, double uacc) {
  const double m = exp(-(de / et));
  return !(isinf(m) || isnan(m)) && uacc <= (m < 1.0 ? m : 1.0);
}
__device__ __noinline__ void commit_window(const double* win, const double* corr, double* sp, const uint32_t* ginfo, int ghost, int Npad,
                                           int CS, int nwin, int k, double L, double s, double so, double rl, double rlo, double rcg,
                                           double et, int* res, double* resd) {
  const int lane = threadIdx.x & 31;
  const bool mine = lane < nwin;
  const double* wl = win + WS * (mine ? lane : 0);
  double de = wl[WS_DE];                                 // running dE of trial k + lane
  const double un = mine ? wl[WS_UN] : 0.0, uno = mine ? wl[WS_UNO] : 0.0;
  const int flag = mine ? reinterpret_cast<const int*>(wl + WS_FLAG_VIS)[0] : 3, src = reinterpret_cast<const int*>(wl + WS_SRC)[0];
  const double um = resd[0], umo = resd[1];
  // ---- before: how far the window can be committed
  double pm = un, pmo = uno;                             // exclusive prefix maxima of the displacements (with the maxima so far)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double a = __shfl_up_sync(0xffffffffu, pm, o), b = __shfl_up_sync(0xffffffffu, pmo, o);
    if (lane >= o) { pm = fmax(pm, a); pmo = fmax(pmo, b); }
  }
  pm = __shfl_up_sync(0xffffffffu, pm, 1); pmo = __shfl_up_sync(0xffffffffu, pmo, 1);
  pm = lane ? fmax(pm, um) : um; pmo = lane ? fmax(pmo, umo) : umo;
  const bool ok = lane == 0 || (src == 0 ? (s * (rl - un - pm) >= rcg && s * (rl - 2.0 * pm) >= rcg)
                                         : (so * (rlo - uno - pmo) >= rcg && so * (rlo - 2.0 * pmo) >= rcg));
  const unsigned stop = __ballot_sync(0xffffffffu, flag != 0 || !ok);
  const int tmax = stop ? __ffs(stop) - 1 : 32;          // <= nwin: lanes past the window carry flag 3
  const int need = tmax == 0 && __shfl_sync(0xffffffffu, flag, 0) == 1;
  // ---- the serial chain. Every lane tests ITS OWN trial against its own thresholds at every turn (only lane t's answer
  // counts at turn t), so what travels between the lanes is one ballot instead of a broadcast double, and the correction
  // row of the next turn is loaded while this one is decided: the chain is compare -> vote -> add.
  const double floor_ = -600.0 * et;
  const double lo = mine ? wl[WS_LO] : 0.0, hi = mine ? wl[WS_HI] : 0.0;
  unsigned accmask = 0u;
  double c_t = corr[lane];                               // c(0, lane)
  for (int t = 0; t < tmax; t++) {
    const double c_next = corr[(t + 1 < tmax ? t + 1 : t) * CS + lane];
    const unsigned b_acc = __ballot_sync(0xffffffffu, de > floor_ && de < lo), b_rej = __ballot_sync(0xffffffffu, de > hi);
    bool acc;
    if ((b_acc >> t) & 1u) acc = true;
    else if ((b_rej >> t) & 1u) acc = false;
    else acc = decide_exact(__shfl_sync(0xffffffffu, de, t), et, win[WS * t + WS_UACC]);      // inside the band: the exact rule
    if (acc) { accmask |= 1u << t; if (lane > t) de += c_t; }
    c_t = c_next;
  }
  // ---- after: move the accepted atoms (store_pos on the values passed in), update the maxima and the counters
  const bool moved = (accmask >> lane) & 1u;
  if (moved) {
    const int i = k + lane;
    const double x = wl[WS_XN], y = wl[WS_YN], z = wl[WS_ZN];
    sp[3 * i] = x; sp[3 * i + 1] = y; sp[3 * i + 2] = z;
    if (ghost) {
      const uint32_t gi = ginfo[i];
      const unsigned nb = gi & 7u;
      if (nb) {
        const double xs = x + ((gi & 8u) ? L : -L), ys = y + ((gi & 16u) ? L : -L), zs = z + ((gi & 32u) ? L : -L);
        double* q = sp + 3 * (size_t)(Npad + (gi >> 8));
        for (unsigned g = 1; g < 8; g++)
          if ((g & ~nb) == 0u) { q[0] = (g & 1u) ? xs : x; q[1] = (g & 2u) ? ys : y; q[2] = (g & 4u) ? zs : z; q += 3; }
      }
    }
  }
  double mx = moved ? un : 0.0, mxo = moved ? uno : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); mxo = fmax(mxo, __shfl_xor_sync(0xffffffffu, mxo, o)); }
  int v = lane < tmax ? reinterpret_cast<const int*>(wl + WS_FLAG_VIS)[1] : 0;
  v = __reduce_add_sync(0xffffffffu, v);
  if (lane == 0) {
    res[0] = k + tmax; res[1] = need;
    resd[0] = fmax(um, mx); resd[1] = fmax(umo, mxo); resd[2] += (double)tmax; resd[3] += (double)__popc(accmask); resd[4] += (double)v;
  }
}

// ------------------------------------------------------------------ a-8 iter_position_mc (lammps_remcmc.py:505-549)
// The reference re-evaluates the whole system for each of the N sequential single-atom trials; here the same sequential
// chain is walked with the single-atom energy change summed over the atom's list column, a WINDOW of W trials (one per
// warp) at a time, in three steps:
//  (A) every warp w draws the proposal of trial k + w (position, acceptance uniform, displacement from the list
//      references) and publishes it;
//  (B) every warp evaluates its trial SPECULATIVELY against the configuration at the start of the round (dE over the
//      list column: the lanes split the quads), and lane a < w evaluates the CORRECTION trial w would need if the
//      earlier trial a of the window were accepted -- the four pair terms that involve atom a,
//          c(a, w) = [u(w_new, a_new) - u(w_old, a_new)] - [u(w_new, a_old) - u(w_old, a_old)]
//      (a's proposal is known from step A whether or not it will be accepted);
//  (C) warp 0 commits the window IN ORDER, lane l holding trial k + l: at turn t the running dE of trial t (its
//      speculative dE plus the corrections of the accepted trials before it, added in window order) is broadcast and
//      decided; if accepted, atom k + t moves and every later lane adds its c(t, l). A rejected trial changes nothing.
//      This is the reference's chain term by term (each trial sees exactly the positions its predecessors left); only
//      the order of the floating-point additions differs from a fresh sum. The serial part is ~100 clocks per trial.
// The Metropolis test U <= min(1, exp(-dE/et)) is decided by comparing dE with the threshold -et ln U prepared in
// step A; the exponential itself is evaluated only when dE lies within 1e-9 of the threshold (or could overflow), so
// the decisions are those of the exact rule.
// Which column is complete for a trial: the inner list (radius rc + skin) while the trial's and the largest
// displacement since its build fit the skin; otherwise (LARGE mode) the OUTER list column (radius rc + skin + outer
// skin, three times as many entries) -- a single-atom sweep at the adapted step size would otherwise rebuild the
// inner list dozens of times. A window ends early at a trial whose column is not guaranteed complete for the
// displacement budget at ITS turn (re-checked at commit time with the running maxima); that trial opens the next
// round, where the first trial may also take the slow paths (list rebuild, or an all-atom sum when the step exceeds
// the skin). Deterministic: the schedule depends on the chain state only.
// Scratch: the window records live in the reduction scratch, the W x W correction table in the float32 fractional
// copies (cx.sf), which only list builds and the FP32 force loop read -- both refresh them first.
template <bool S32>
__device__ void iter_position_mc(const Dev& d, Ctx& cx, const Rng& r, double et, double dxs, Energy& en, double* cnt) {
  const int N = cx.N, Npad = cx.Npad, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int nw = blockDim.x >> 5;              // window: one trial per warp, as long as the correction table fits its scratch
  while (nw > 1 && nw * (nw + 1) > 2 * Npad) nw--;
  const int CS = nw + 1;                 // row stride of the correction table (odd: conflict-free): c(a, b) at corr[a * CS + b]
  const double rc = d.rc, rc2 = rc * rc, rl = rc + cx.skin, rlo = rl + d.oskin, L = cx.L, hL = 0.5 * L, invL = 1.0 / L;
  const double rcg = rc * (1 + 1e-9);
  const int L_hi = __double2hiint(L), L_lo = __double2loint(L), hL_hi = __double2hiint(hL);
  const bool two_level = !d.small;
  check_list(d, cx);
  // um / umo: largest displacement (build units) of any atom from its inner / outer list reference
  double um = 0.0, umo = 0.0;
  auto max_displacements = [&]() {
    double mx = 0.0, mxo = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double x = cx.sp[3 * i], y = cx.sp[3 * i + 1], z = cx.sp[3 * i + 2];
      mx = fmax(mx, disp2(cx, i, x, y, z, invL));
      if (two_level) mxo = fmax(mxo, disp2o(cx, i, x, y, z, invL));
    }
    for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); mxo = fmax(mxo, __shfl_xor_sync(0xffffffffu, mxo, o)); }
    __syncthreads();
    if (lane == 0) { cx.red[wid] = mx; cx.red[32 + wid] = mxo; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double q = 0.0, qo = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) { q = fmax(q, cx.red[w]); qo = fmax(qo, cx.red[32 + w]); }
      cx.bc[8] = sqrt(q); cx.bc[9] = sqrt(qo);        // bc[8..12]: umax, umaxo, and the sweep's trials / accepted / visited (commit_window)
    }
    __syncthreads();
    um = cx.bc[8]; umo = cx.bc[9];
    __syncthreads();
  };
  max_displacements();
  // LARGE mode: start the sweep from a fresh inner list unless the current one is all but fresh. A trial may use its inner
  // column while un + um <= skin; with the budget already half spent by earlier moves most trials of the sweep would fall
  // back to the outer column (three passes of the warp instead of one): a rebuild costs ~20 rounds of a 125-round sweep.
  if (two_level && um > 0.05 * cx.skin) { build_list(d, cx); max_displacements(); }
  if (threadIdx.x == 0) { cx.bc[10] = 0.0; cx.bc[11] = 0.0; cx.bc[12] = 0.0; }
  double* win = cx.red;
  double* corr = reinterpret_cast<double*>(cx.sf);
  static_assert(32 * WS <= RED_DOUBLES, "window scratch");
  // pair energy with the cutoff from a separation (minimum image on the integer pipe)
  auto u_lj = [&](double ax, double ay, double az) -> double {
    ax = mic_fast(ax, L_hi, L_lo, hL_hi); ay = mic_fast(ay, L_hi, L_lo, hL_hi); az = mic_fast(az, L_hi, L_lo, hL_hi);
    const double rr = ax * ax + ay * ay + az * az, i2 = rcp_nr(rr), s6 = i2 * i2 * i2;
    return rr < rc2 ? s6 * (4.0 * s6 - 4.0) : 0.0;
  };
  int k = 0;
  while (k < N) {
    const double s = L / cx.L0, so = two_level ? L / cx.L0o : 1.0;
    const int kk = k + wid, nwin = min(nw, N - k);
#ifdef NM_DEBUG_CLOCKS
    const long long t_round0 = clk_fenced();
#endif
    // ---- (A) proposal of trial kk
    int nq = 0, flag = 3, src = 0;
    ushort4 e4c = make_ushort4(0, 0, 0, 0);
    double xo = 0, yo = 0, zo = 0, xn = 0, yn = 0, zn = 0;
    if (wid < nwin) {
      nq = cx.nnb[kk];
      e4c = lane < nq ? cx.list[(size_t)lane * Npad + kk] : make_ushort4(0, 0, 0, 0);
      double x0c = lane < 3 ? cx.gx0[lane * Npad + kk] : 0.0;
      if (two_level && lane >= 3 && lane < 6) x0c = cx.gx0o[(lane - 3) * Npad + kk];
      // the trial's three Philox blocks (two for the displacement, one for the acceptance uniform) in ONE pass: lanes 0, 1, 2
      // take one counter each, the words are broadcast (same streams as rng_uniform3 / rng_uniform)
      double u[3], uacc;
      {
        const uint32_t c0 = lane == 2 ? (uint32_t)kk : 2u * (uint32_t)kk + (lane == 1 ? 1u : 0u), c1 = lane == 2 ? (uint32_t)P_ITER_ACC : (uint32_t)P_ITER_DISP;
        uint32_t w[4]; philox4x32_10(r.k0, r.k1, c0, c1, r.m_lo, r.m_hi, w);
        const double ua = u53(w[0], w[1]), ub = u53(w[2], w[3]);
        u[0] = __shfl_sync(0xffffffffu, ua, 0); u[1] = __shfl_sync(0xffffffffu, ub, 0); u[2] = __shfl_sync(0xffffffffu, ua, 1);
        uacc = __shfl_sync(0xffffffffu, ua, 2);
      }
      // accept <=> dE <= thr = -et ln U up to rounding: decided by comparison outside the band thr -+ 2e-9 |thr|, by the exact rule inside
      const double thr = -log(uacc) * et, band = 2e-9 * fabs(thr) + 1e-290;
      xo = cx.sp[3 * kk]; yo = cx.sp[3 * kk + 1]; zo = cx.sp[3 * kk + 2];
      xn = xo + 2 * (u[0] - 0.5) * dxs * d.lat; yn = yo + 2 * (u[1] - 0.5) * dxs * d.lat; zn = zo + 2 * (u[2] - 0.5) * dxs * d.lat;
      double un, uno = 0.0;
      {
        const double rx = __shfl_sync(0xffffffffu, x0c, 0), ry = __shfl_sync(0xffffffffu, x0c, 1), rz = __shfl_sync(0xffffffffu, x0c, 2);
        double ux = xn * invL - rx, uy = yn * invL - ry, uz = zn * invL - rz;
        ux -= rint(ux); uy -= rint(uy); uz -= rint(uz);
        un = sqrt(ux * ux + uy * uy + uz * uz) * cx.L0;
      }
      if (two_level) {
        const double rx = __shfl_sync(0xffffffffu, x0c, 3), ry = __shfl_sync(0xffffffffu, x0c, 4), rz = __shfl_sync(0xffffffffu, x0c, 5);
        double ux = xn * invL - rx, uy = yn * invL - ry, uz = zn * invL - rz;
        ux -= rint(ux); uy -= rint(uy); uz -= rint(uz);
        uno = sqrt(ux * ux + uy * uy + uz * uz) * cx.L0o;
      }
      // every atom within rc of the old or the new position must be in the column that is summed
      const bool inner_ok = s * (rl - un - um) >= rcg && s * (rl - 2.0 * um) >= rcg;
      const bool outer_ok = two_level && cx.thro2 >= 0.0 && so * (rlo - uno - umo) >= rcg && so * (rlo - 2.0 * umo) >= rcg;
      flag = 0;                          // 0: evaluated, 1: a fresh list would do (rebuild, then retry), 3: not evaluated (opens a later round)
      src = inner_ok ? 0 : (outer_ok ? 1 : 2);          // column summed: inner list, outer list, all atoms
      if (src == 2) {
        if (wid == 0) {
          const double ddx = mic_exact(xn - xo, L, hL), ddy = mic_exact(yn - yo, L, hL), ddz = mic_exact(zn - zo, L, hL);
          const double step = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
          if (rl - step >= rcg) flag = 1;               // else: step larger than the skin, all-atom sum
        } else flag = 3;
      }
      if (lane == 0) {
        double* t = win + WS * wid;
        t[WS_XN] = xn; t[WS_YN] = yn; t[WS_ZN] = zn; t[WS_LO] = thr - band; t[WS_HI] = thr + band; t[WS_UACC] = uacc; t[WS_UN] = un; t[WS_UNO] = uno;
        reinterpret_cast<int*>(t + WS_FLAG_VIS)[0] = flag; reinterpret_cast<int*>(t + WS_SRC)[0] = src;
      }
    }
    __syncthreads();
#ifdef NM_DEBUG_CLOCKS
    if (threadIdx.x == 0) cx.ct[NM_CT_CLK_VEL] += (unsigned long long)(clk_fenced() - t_round0);
#endif
    // ---- (B) speculative dE of trial kk, and its corrections for the earlier trials of the window
    if (wid < nwin && flag == 0) {
      double de = 0.0; int vis = 0;
      auto pair = [&](int j) {
        const double ax = mic_exact(xn - cx.sp[3 * j], L, hL), ay = mic_exact(yn - cx.sp[3 * j + 1], L, hL), az = mic_exact(zn - cx.sp[3 * j + 2], L, hL);
        const double bx = mic_exact(xo - cx.sp[3 * j], L, hL), by = mic_exact(yo - cx.sp[3 * j + 1], L, hL), bz = mic_exact(zo - cx.sp[3 * j + 2], L, hL);
        const double rn = ax * ax + ay * ay + az * az, ro = bx * bx + by * by + bz * bz;
        const double r2n = rcp_nr(rn), r2o = rcp_nr(ro);
        const double r6n = r2n * r2n * r2n, r6o = r2o * r2o * r2o;
        if (rn < rc2) { de += r6n * (4.0 * r6n - 4.0); vis++; }
        if (ro < rc2) { de -= r6o * (4.0 * r6o - 4.0); vis++; }
      };
      if (src == 2) {
        for (int j = lane; j < N; j += 32) if (j != kk) pair(j);
      } else if (src == 1) {
        const int nqo = cx.onq[kk];
        for (int q = lane; q < nqo; q += 32) {
          const ushort4 e4 = cx.olist[(size_t)q * Npad + kk];
          pair(e4.x); pair(e4.y); pair(e4.z); pair(e4.w);
        }
      } else {
        if (lane < nq) { pair(e4c.x & 0x1fff); pair(e4c.y & 0x1fff); pair(e4c.z); pair(e4c.w); }
        for (int q = lane + 32; q < nq; q += 32) {
          const ushort4 e4 = cx.list[(size_t)q * Npad + kk];
          pair(e4.x & 0x1fff); pair(e4.y & 0x1fff); pair(e4.z); pair(e4.w);
        }
      }
      if (lane < wid) {                  // c(a = lane, b = wid)
        const double* ta = win + WS * lane;
        const double anx = ta[WS_XN], any_ = ta[WS_YN], anz = ta[WS_ZN];
        const double aox = cx.sp[3 * (k + lane)], aoy = cx.sp[3 * (k + lane) + 1], aoz = cx.sp[3 * (k + lane) + 2];
        corr[lane * CS + wid] = (u_lj(xn - anx, yn - any_, zn - anz) - u_lj(xo - anx, yo - any_, zo - anz))
                              - (u_lj(xn - aox, yn - aoy, zn - aoz) - u_lj(xo - aox, yo - aoy, zo - aoz));
      }
      for (int o = 16; o > 0; o >>= 1) { de += __shfl_xor_sync(0xffffffffu, de, o); vis += __shfl_xor_sync(0xffffffffu, vis, o); }
      if (lane == 0) { win[WS * wid + WS_DE] = de; reinterpret_cast<int*>(win + WS * wid + WS_FLAG_VIS)[1] = vis; }
    }
    __syncthreads();
#ifdef NM_DEBUG_CLOCKS
    const long long t_round1 = clk_fenced();
    if (threadIdx.x == 0) { cx.ct[NM_CT_DBG_LOOPCLK] += (unsigned long long)(t_round1 - t_round0); cx.ct[NM_CT_HELPED_EVALS]++; }
#endif
    // ---- (C) ordered commit: lane l holds the running dE of trial k + l
    if (wid == 0) {
      __syncwarp();                      // the warp enters the serial chain converged (a diverged warp pays a re-synchronisation per shuffle)
      commit_window(win, corr, cx.sp, cx.ginfo, cx.ghost, Npad, CS, nwin, k, L, s, so, rl, rlo, rcg, et, cx.ibc + 1, cx.bc + 8);
#ifdef NM_DEBUG_CLOCKS
      if (lane == 0) cx.ct[NM_CT_DBG_LOOPIT] += (unsigned long long)(clk_fenced() - t_round1);
#endif
    }
    __syncthreads();
    k = cx.ibc[1];
    const int need = cx.ibc[2];
    um = cx.bc[8]; umo = cx.bc[9];
    if (need) { __syncthreads(); build_list(d, cx); max_displacements(); }   // (the build's reductions use the window scratch)
  }
  if (threadIdx.x == 0) {
    cnt[0] += cx.bc[10]; cnt[1] += cx.bc[11];
    cx.ct[NM_CT_PMC_MOVES]++; cx.ct[NM_CT_PMC_TRIALS] += (unsigned long long)cx.bc[10]; cx.ct[NM_CT_PAIRS_DELTA] += (unsigned long long)cx.bc[12];
  }
  // the state the last 'run 0' of the sweep leaves: a fresh full evaluation
  check_list(d, cx);
  double o[4]; eval_forces<true, false, S32>(d, cx, 0.0, o);
  en.pe = o[0]; en.w = o[1];
}

}  // anonymous namespace


// ------------------------------------------------------------------ kernels
// 'run 0' on the resident configurations: wrap, (re)build list, evaluate; optionally export.
template <int NTHR>
__global__ void __launch_bounds__(NTHR, NM_CTAS_PER_SM(NTHR))
k_eval(Dev d, double* pe_out, double* w_out, double* f_out_aos, long long* npairs_out) {
  constexpr bool S32 = NTHR == 1024;
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx cx; ctx_init(d, cx, blockIdx.x, smem);
  if (cx.L < 2.0 * d.rc * (1.0 + 1e-5)) { if (threadIdx.x == 0) d.status[cx.c] |= ST_BOX; return; }
  load_positions(d, cx);
  update_thr(d, cx);
  __syncthreads();
  check_list(d, cx);
  double o[4]; eval_forces<true, false, S32>(d, cx, 0.0, o);
  store_positions(cx);
  double t[1] = { 0 };
  for (int i = threadIdx.x; i < cx.N; i += blockDim.x) {
    const double vx = cx.gv[i], vy = cx.gv[cx.Npad + i], vz = cx.gv[2 * cx.Npad + i];
    t[0] += vx * vx + vy * vy + vz * vz;
  }
  bsum<1>(t, cx);
  const int slot = d.cfg_slot[cx.c];
  if (f_out_aos) {
    double* fo = f_out_aos + (size_t)slot * 3 * cx.N;
    for (int i = threadIdx.x; i < cx.N; i += blockDim.x) {
      fo[3 * i] = cx.gf[i]; fo[3 * i + 1] = cx.gf[cx.Npad + i]; fo[3 * i + 2] = cx.gf[2 * cx.Npad + i];
    }
  }
  if (threadIdx.x == 0) {
    d.pe[cx.c] = o[0]; d.w[cx.c] = o[1]; d.ke[cx.c] = 0.5 * d.mass * t[0]; d.L0[cx.c] = cx.L0; d.L0o[cx.c] = cx.L0o; d.micmode[cx.c] = cx.mic; d.list_pairs[cx.c] = cx.list_pairs; d.lcur[cx.c] = cx.lbuf;
    if (pe_out) pe_out[slot] = o[0];
    if (w_out) w_out[slot] = o[1];
    if (npairs_out) npairs_out[slot] = (long long)o[2];
    if (cx.status) d.status[cx.c] |= cx.status;
    for (int k = 0; k < NM_COUNTER_WIDTH; k++) if (cx.ct[k]) atomicAdd(&d.counters[k], cx.ct[k]);
  }
}

// gen_sample (lammps_remcmc.py:665-691): MOD x move_mc (:643-658), then lammps_extract (:377-391).
//
// PERSISTENT kernel with a move-granular work queue. The MOD moves of a configuration form a sequential chain, but the
// chains differ in cost (a melt rebuilds its list twice per trajectory, a cold solid never; 45 .. 93 M clocks at C2) and
// there are fewer CTA slots than would balance them: with one CTA per configuration the slowest chain sets the kernel time
// while SMs that hold a single CTA (which runs 1.6x faster alone) finish early and idle. Here a cycle is cut into nseg
// SEGMENTS of seg_moves moves. Tickets t = segment * nrep + rank are handed out in order (atomic counter); the CTA that
// draws (segment s, configuration c) waits until segment s - 1 of c is published (acquire on sched[1 + c]; its ticket is
// older, hence held by a running CTA: no deadlock as long as every CTA of the grid is resident, which the launch bounds to
// the occupancy), runs the moves out of shared memory exactly as before, writes the state back and publishes s + 1
// (release). A configuration thus migrates between CTAs / SMs from segment to segment: whoever is free continues the
// cheapest unfinished chain, so every SM stays busy to the end of the cycle. The state that crosses a segment boundary is
// exactly the state that crosses a cycle boundary (positions, box, energies, step counters, list bookkeeping, all in
// global memory), so results are bit-identical to the unsegmented cycle and independent of the schedule.

// A CTA without a chain of its own serves configuration c (which it has claimed) for up to help_quantum commands or
// until that chain is finished (see the note at help_request).
template <int NTHR>
__device__ __noinline__ void helper_serve(const Dev& d, unsigned char* smem, int c) {      // (out of line: the owner's code keeps its registers)
  Ctx cx; ctx_init(d, cx, c, smem);
  int* hs = d.help + (size_t)HELP_STRIDE * c;
  const double* hd = d.helpd + 4 * (size_t)c;
  double* hp = d.hpart + (size_t)c * 4 * NTHR;
  const int N = cx.N, Npad = cx.Npad, tid = threadIdx.x, Ns = 2 * NTHR;
  for (int i = N + tid; i < Npad; i += NTHR) { cx.sp[3 * i] = 1e9; cx.sp[3 * i + 1] = 1e9; cx.sp[3 * i + 2] = 1e9; }
  if (tid == 0) cx.ibc[5] = ld_acquire_gpu(hs + 1);      // every command up to this one is complete (the claim is exclusive)
  __syncthreads();
  int seq = cx.ibc[5];
  if (seq == -1) return;                                 // finished in the meantime (the claim stays: nobody needs it)
  __syncthreads();
  if (tid == 0) st_release_gpu(hs, 1);                   // attached: the owner may send commands from now on
  for (int served = 1;; served++) {
    seq++;
    if (tid == 0) {
      int v;
      while ((v = ld_acquire_gpu(hs + 1)) != seq && v != -1) __nanosleep(128);
      cx.ibc[5] = v;
    }
    __syncthreads();
    if (cx.ibc[5] == -1) return;
    const int flags = hs[3];
    cx.lbuf = hs[4]; select_list(cx);
    cx.L = hd[0]; cx.mic = (flags >> 2) & 1;
    const double dtf = hd[1];
    if (tid < 27) { cx.sht[3 * tid] = (tid / 9 - 1) * cx.L; cx.sht[3 * tid + 1] = ((tid / 3) % 3 - 1) * cx.L; cx.sht[3 * tid + 2] = (tid % 3 - 1) * cx.L; }
    for (int i0 = 0; i0 < N; i0 += 4 * NTHR) {           // positions global (L2) -> shared: twelve loads in flight per thread
      double px[4], py[4], pz[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * NTHR + tid;
        if (i < N) { px[u] = __ldcg(cx.gx + i); py[u] = __ldcg(cx.gx + Npad + i); pz[u] = __ldcg(cx.gx + 2 * Npad + i); }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * NTHR + tid;
        if (i < N) { cx.sp[3 * i] = px[u]; cx.sp[3 * i + 1] = py[u]; cx.sp[3 * i + 2] = pz[u]; }
      }
    }
    if (tid == 0) cx.s_pairs[0] = 0ull;
    __syncthreads();
    double e = 0.0, vir = 0.0, ke = 0.0; int np = 0;
    if (flags & 24) {                                    // list build: the rows above the split of the outer search / the inner regeneration
      wrap_and_refresh(cx, false);                       // float32 fractional copies (the owner has wrapped the atoms already)
      __syncthreads();
      double r[2] = { 0.0, 0.0 };
      if (flags & 16) {
        const int nc = hs[9], sw = hs[10];
        const double rlo = d.rc + cx.skin + d.oskin, invL = 1.0 / cx.L;
        if (nc > 1) bin_cells(cx, nc);
        r[1] = (double)outer_rows(d, cx, (float)(rlo * rlo * invL * invL * (1.0 + 2e-5)), nc, sw, Ns, N);
      } else {
        int over = 0;
        if (cx.mic) inner_rows<true>(d, cx, Ns, N, r[0], over); else inner_rows<false>(d, cx, Ns, N, r[0], over);
        r[1] = (double)over;
      }
      bsum<2>(r, cx);
      if (tid == 0) { d.helpd[4 * (size_t)c + 2] = r[0]; d.helpd[4 * (size_t)c + 3] = r[1]; }
    } else
    switch (flags & 7) {
      case 1: force_rows<true, false, 0, true>(d, cx, dtf, Ns, N, e, vir, ke, np); break;
      case 2: force_rows<false, true, 0, true>(d, cx, dtf, Ns, N, e, vir, ke, np); break;
      case 3: force_rows<true, true, 0, true>(d, cx, dtf, Ns, N, e, vir, ke, np); break;
      case 5: force_rows<true, false, 1, true>(d, cx, dtf, Ns, N, e, vir, ke, np); break;
      case 6: force_rows<false, true, 1, true>(d, cx, dtf, Ns, N, e, vir, ke, np); break;
      case 7: force_rows<true, true, 1, true>(d, cx, dtf, Ns, N, e, vir, ke, np); break;
      default: break;
    }
    if (flags & 24) { }
    else if (flags & 1) { hp[tid] = e; hp[NTHR + tid] = vir; hp[2 * NTHR + tid] = (double)np; hp[3 * NTHR + tid] = ke; }
    else {
      np = __reduce_add_sync(0xffffffffu, np);
      if ((tid & 31) == 0) atomicAdd(cx.s_pairs, (unsigned long long)np);
    }
    const bool leave = served >= d.help_quantum;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      if (!(flags & 25)) *reinterpret_cast<unsigned long long*>(hs + 6) = cx.s_pairs[0];
      if (leave) st_release_gpu(hs, 0);                  // detached before the answer: the owner will not address this CTA again
      st_release_gpu(hs + 2, seq);
      if (leave) st_release_gpu(hs + 5, 0);              // the chain may be claimed again
    }
    if (leave) return;
  }
}

// SM-aware placement (thread 0 of every CTA, once per launch; nsm < nrep <= 2 nsm, two CTAs fit per SM, whole grid resident).
// The block scheduler deals the CTAs breadth-first: 2 nsm - nrep SMs end up with ONE CTA, which then runs ~1.4x faster than a
// CTA that shares its SM, and two expensive chains that happen to share an SM set the kernel time. Every CTA registers on
// its SM (%smid), waits until the whole grid has registered, and derives its rank in the cost order (order[] is sorted most
// expensive first) from the final occupancy map, the same for everybody: the k-th single-CTA SM takes rank k (the most
// expensive chains run alone), the k-th shared SM pairs rank nsolo + k with rank nrep - 1 - k (expensive with cheap: the
// cheap chain ends early and leaves the SM to its partner). Any other occupancy pattern: arrival order.
constexpr int SMID_MAX = 256;
static __device__ int placement_rank(const Dev& d) {
  int* arrived = d.sched + 1 + d.nrep; int* invalid = arrived + 1; int* smcount = arrived + 2;
  unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  int j = 0;
  if (smid < (unsigned)SMID_MAX) j = atomicAdd(&smcount[smid], 1); else atomicExch(invalid, 1);
  __threadfence();
  const int a = atomicAdd(arrived, 1);
  while (ld_acquire_gpu(arrived) < (int)gridDim.x) __nanosleep(64);
  int nsolo = 0, npair = 0, ksolo = 0, kpair = 0, bad = ld_acquire_gpu(invalid);
  for (int i = 0; i < SMID_MAX; i++) {
    const int n = ld_acquire_gpu(&smcount[i]);
    if (n == 1) { nsolo++; if (i < (int)smid) ksolo++; }
    else if (n == 2) { npair++; if (i < (int)smid) kpair++; }
    else if (n != 0) bad = 1;
  }
  if (bad || nsolo + 2 * npair != d.nrep) return a;
  const int n_me = ld_acquire_gpu(&smcount[smid]);
  return n_me == 1 ? ksolo : (j == 0 ? nsolo + kpair : d.nrep - 1 - kpair);
}

template <int NTHR>
__global__ void __launch_bounds__(NTHR, NM_CTAS_PER_SM(NTHR))
k_cycle(Dev d, long long cycle) {
  constexpr bool S32 = NTHR == 1024;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_ticket;
  const int ntickets = d.nseg * d.nrep;
  bool first = true;
  for (;;) {
    __syncthreads();                                  // everybody is done with the previous segment (and its s_ticket)
    if (threadIdx.x == 0) s_ticket = (first && d.place) ? placement_rank(d) : atomicAdd(&d.sched[0], 1);
    first = false;
    __syncthreads();
    const int ticket = s_ticket;
    if (ticket >= ntickets) {
      // no chain left: help the running chain with the most work left (claims are exclusive), until every chain is finished
      if constexpr (kHelpers && NTHR == 1024) for (; d.nhelp > 0;) {
        __syncthreads();
        if (threadIdx.x == 0) {
          int c = -1;
          const int* finished = d.sched + 3 + d.nrep + SMID_MAX;
          // only a chain that is RUNNING may be waited for: if some chains have not started within ~100 us, the grid is not
          // fully resident (the device is shared) and this CTA makes room instead of holding an SM
          const int* started = finished + 1;
          for (int tries = 0; ld_acquire_gpu(started) < d.nrep && tries < 400; tries++) __nanosleep(250);
          if (ld_acquire_gpu(started) >= d.nrep)
          while (ld_acquire_gpu(finished) < d.nrep) {
            int best = -1, bestrem = 0;
            for (int o = 0; o < d.nrep; o++) {
              const volatile int* r = d.help + (size_t)HELP_STRIDE * o;
              if (r[5] == 0 && r[1] != -1) { const int rem = r[8]; if (rem > bestrem) { bestrem = rem; best = o; } }
            }
            if (best < 0) { __nanosleep(2000); continue; }
            if (atomicCAS(d.help + (size_t)HELP_STRIDE * best + 5, 0, 1) == 0) { c = best; break; }
          }
          s_ticket = c;
        }
        __syncthreads();
        const int c = s_ticket;
        if (c < 0) break;
        helper_serve<NTHR>(d, smem, c);
      }
      break;
    }
    const int seg = ticket / d.nrep, c = d.order[ticket - seg * d.nrep];
    if (seg > 0) {
      if (threadIdx.x == 0) while (ld_acquire_gpu(&d.sched[1 + c]) < seg) __nanosleep(256);
      __syncthreads();
    }
    Ctx cx; ctx_init(d, cx, c, smem);
    if (kHelpers && NTHR == 1024 && d.nhelp > 0 && threadIdx.x == 0) atomicAdd(d.sched + 4 + d.nrep + SMID_MAX, 1);   // this chain is running
    if (kHelpers && NTHR == 1024 && d.nhelp > 0) { cx.help = d.help + (size_t)HELP_STRIDE * c; cx.helpd = d.helpd + 4 * (size_t)c; cx.hpart = d.hpart + (size_t)c * 4 * NTHR; }
    const int slot = d.cfg_slot[c], N = cx.N, Npad = cx.Npad;
    const long long t_seg0 = clock64();
    const double et = d.label[4 * slot], pf = d.label[4 * slot + 1], t_vel = d.label[4 * slot + 3];
    const double dxs = d.step[3 * c], dvs = d.step[3 * c + 1], dts = d.step[3 * c + 2];
    load_positions(d, cx);
    update_thr(d, cx);
    Energy en = { d.pe[c], d.w[c] };
    double cnt[6];
    for (int k = 0; k < 6; k++) cnt[k] = d.cnt[6 * c + k];
    __syncthreads();
    unsigned long long kclk[3] = { 0ull, 0ull, 0ull }; unsigned kcnt[3] = { 0u, 0u, 0u };   // per move kind (thread 0)
    const int mv0 = seg * d.seg_moves, mv1 = min(d.mod, mv0 + d.seg_moves);
    for (int mv = mv0; mv < mv1; mv++) {
      const Rng r = rng_make(d.seed_lo, d.seed_hi, (uint32_t)gslot(d, slot), (uint64_t)cycle * (uint64_t)d.mod + (uint64_t)mv);
      const double roll = rng_uniform(r, 0, P_ROLL);
      const long long t_mv0 = clock64();
      const int kind = roll <= d.ppos ? 0 : (roll <= (d.ppos + d.pvol) ? 1 : 2);
      if (kind == 0) {
        if (d.bulk) bulk_position_mc<S32>(d, cx, r, et, dxs, en, cnt);
        else iter_position_mc<S32>(d, cx, r, et, dxs, en, cnt);
      } else if (kind == 1) volume_mc<S32>(d, cx, r, et, pf, dvs, en, cnt);
      else hamiltonian_mc<S32>(d, cx, r, et, t_vel, dts, en, cnt);
      if (threadIdx.x == 0) {
        const long long t_now = clock64();
        cx.ct[NM_CT_SWEEPS]++; kclk[kind] += (unsigned long long)(t_now - t_mv0); kcnt[kind]++;
        if (kHelpers && cx.help) {                       // remaining clocks at the pace so far (helpers rank the chains by it)
          const long long rem = (t_now - t_seg0) / (mv - mv0 + 1) * (mv1 - mv - 1);
          *reinterpret_cast<volatile int*>(cx.help + 8) = (int)min(rem >> 10, 0x7fffffffll);
        }
      }
    }
    const bool last = seg == d.nseg - 1;
    double t[1] = { 0 };
    if (last) {                                         // lammps_extract
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double vx = cx.gv[i], vy = cx.gv[Npad + i], vz = cx.gv[2 * Npad + i];
        t[0] += vx * vx + vy * vy + vz * vz;
      }
      bsum<1>(t, cx);
    }
    store_positions(cx);
    if (threadIdx.x == 0) {
      for (int k = 0; k < 6; k++) d.cnt[6 * c + k] = cnt[k];
      if (last) {
        const double ke = 0.5 * d.mass * t[0], dof = 3.0 * N - 3.0, temp = 2.0 * ke / dof, vol = pow(cx.L, 3.0);
        double* th = d.thermo + (size_t)slot * NM_THERMO_WIDTH;
        th[NM_TH_TEMP] = temp; th[NM_TH_PE] = en.pe; th[NM_TH_KE] = ke;
        th[NM_TH_VIRIAL] = (dof * temp + en.w) / 3.0 * (1.0 / vol);
        th[NM_TH_BOX] = cx.L; th[NM_TH_VOL] = vol; th[NM_TH_DX] = dxs; th[NM_TH_DV] = dvs; th[NM_TH_DT] = dts;
        for (int k = 0; k < 6; k++) th[NM_TH_NTP + k] = cnt[k];
        for (int k = 0; k < 3; k++) {
          const float a = (float)cnt[2 * k + 1] / (float)cnt[2 * k];        // float32 ratio, nan_to_num (0/0 -> 0)
          th[NM_TH_AP + k] = isnan(a) ? 0.0 : (double)a;
        }
        d.ke[c] = ke;
      }
      d.box[c] = cx.L; d.pe[c] = en.pe; d.w[c] = en.w; d.L0[c] = cx.L0; d.L0o[c] = cx.L0o; d.micmode[c] = cx.mic; d.list_pairs[c] = cx.list_pairs; d.lcur[c] = cx.lbuf;
      if (cx.status) d.status[c] |= cx.status;
      cx.ct[NM_CT_PAIRS_FORCE] += cx.s_pairs[0] / 2;
#ifdef NM_DEBUG_SMID
      { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); cx.ct[NM_CT_HELPED_EVALS] = smid; }
#endif
      const unsigned long long dt_seg = (unsigned long long)(clock64() - t_seg0);
      cx.ct[NM_CT_CLK_TOTAL] += dt_seg;
      // per-configuration bookkeeping of the cycle (cost ranks for the next schedule, diagnostics): reset by segment 0
      const unsigned long long old_cnt = seg ? d.mv_clk[4 * c + 3] : 0ull;
      d.cta_clk[c] = (seg ? d.cta_clk[c] : 0ull) + dt_seg;
      // work estimate in units of one listed pair: evaluations + list builds (SMALL: all-pairs tiles + row walk; LARGE, measured
      // at N = 4000: an inner build ~ 75 N, an outer build ~ 650 N pair evaluations)
      d.cost[2 * c] = (seg ? d.cost[2 * c] : 0ull) + cx.ct[NM_CT_LIST_PAIRS] +
                      (d.small ? cx.ct[NM_CT_LIST_BUILDS] * (unsigned long long)(d.build_cost * N * N)
                               : (cx.ct[NM_CT_LIST_BUILDS] * 75ull + cx.ct[NM_CT_OUTER_BUILDS] * 650ull) * (unsigned long long)N);
      d.cost[2 * c + 1] = (seg ? d.cost[2 * c + 1] : 0ull) + cx.ct[NM_CT_FORCE_EVALS];
      for (int k = 0; k < 3; k++) d.mv_clk[4 * c + k] = (seg ? d.mv_clk[4 * c + k] : 0ull) + kclk[k];
      unsigned long long packed = 0ull;
      for (int k = 0; k < 3; k++) packed |= (unsigned long long)min((unsigned)((old_cnt >> (16 * k)) & 0xffffull) + kcnt[k], 65535u) << (16 * k);
      d.mv_clk[4 * c + 3] = packed;
      for (int k = 0; k < NM_COUNTER_WIDTH; k++) {
        d.rep_ct[(size_t)c * NM_COUNTER_WIDTH + k] = (seg ? d.rep_ct[(size_t)c * NM_COUNTER_WIDTH + k] : 0ull) + cx.ct[k];
        if (cx.ct[k]) atomicAdd(&d.counters[k], cx.ct[k]);
      }
    }
    __threadfence();                                    // this thread's state writes are visible device-wide ...
    __syncthreads();                                    // ... for every thread of the CTA, before thread 0 publishes the segment
    if (threadIdx.x == 0) {
      st_release_gpu(&d.sched[1 + c], seg + 1);
      if (kHelpers && cx.help) { st_release_gpu(cx.help + 1, -1); atomicAdd(d.sched + 3 + d.nrep + SMID_MAX, 1); }   // dismiss the helper; one chain fewer to help
    }
  }
}

// 'velocity all create T seed dist gaussian' + zero linear + zero angular on the resident configurations, outside a move:
// the draw the reference's init_sample keeps in STATE with -is (lammps_remcmc.py:420-425; the 'run 1024' that follows has no
// integrator defined and changes nothing). Stream: (seed, global slot, move counter 2^64 - 1 - tag).
template <int NTHR>
__global__ void __launch_bounds__(NTHR, NM_CTAS_PER_SM(NTHR))
k_velinit(Dev d, long long tag) {
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx cx; ctx_init(d, cx, blockIdx.x, smem);
  const int slot = d.cfg_slot[cx.c];
  load_positions(d, cx);
  __syncthreads();
  const Rng r = rng_make(d.seed_lo, d.seed_hi, (uint32_t)gslot(d, slot), ~0ull - (uint64_t)tag);
  const double ke = velocity_create(d, cx, r, d.label[4 * slot + 3]);
  if (threadIdx.x == 0) d.ke[cx.c] = ke;
}

// ---- launchers, one set per thread count. build.py compiles this file four times (-DNM_TU=256 / 512 / 1024: the
// kernels of that thread count only; -DNM_TU=0: the small kernels and the host C-ABI) so that the three heavy
// instantiations build in parallel; without NM_TU everything lands in one translation unit.
#define NM_LAUNCHERS(T)                                                                                              \
  cudaError_t set_smem_##T(size_t sm) {                                                                              \
    cudaError_t e = cudaFuncSetAttribute(k_cycle<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);          \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eval<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_velinit<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    return e;                                                                                                        \
  }                                                                                                                  \
  void launch_cycle_##T(const Dev& d, long long cycle, int grid, size_t sm, cudaStream_t st) { k_cycle<T><<<grid, T, sm, st>>>(d, cycle); } \
  int occupancy_##T(size_t sm) { int n = 0; return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_cycle<T>, T, sm) == cudaSuccess ? n : 1; } \
  void launch_velinit_##T(const Dev& d, long long tag, size_t sm, cudaStream_t st) { k_velinit<T><<<d.nrep, T, sm, st>>>(d, tag); } \
  void launch_eval_##T(const Dev& d, double* pe, double* w, double* f, long long* np_, size_t sm, cudaStream_t st) { \
    k_eval<T><<<d.nrep, T, sm, st>>>(d, pe, w, f, np_);                                                              \
  }
#define NM_LAUNCHER_DECLS(T)                                                                                         \
  cudaError_t set_smem_##T(size_t sm);                                                                               \
  void launch_cycle_##T(const Dev& d, long long cycle, int grid, size_t sm, cudaStream_t st);                        \
  int occupancy_##T(size_t sm);                                                                                      \
  void launch_velinit_##T(const Dev& d, long long tag, size_t sm, cudaStream_t st);                                  \
  void launch_eval_##T(const Dev& d, double* pe, double* w, double* f, long long* np_, size_t sm, cudaStream_t st);
NM_LAUNCHER_DECLS(256) NM_LAUNCHER_DECLS(512) NM_LAUNCHER_DECLS(1024)
cudaError_t set_smem_1024h(size_t sm);
void launch_cycle_1024h(const Dev& d, long long cycle, int grid, size_t sm, cudaStream_t st);
#if !defined(NM_TU) || NM_TU == 1025      // the helper-capable cycle kernel (k_cycle_h; the one k_cycle<1024> in a single-unit build)
cudaError_t set_smem_1024h(size_t sm) { return cudaFuncSetAttribute(k_cycle<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); }
void launch_cycle_1024h(const Dev& d, long long cycle, int grid, size_t sm, cudaStream_t st) { k_cycle<1024><<<grid, 1024, sm, st>>>(d, cycle); }
#endif
#if !defined(NM_TU) || NM_TU == 256
NM_LAUNCHERS(256)
#endif
#if !defined(NM_TU) || NM_TU == 512
NM_LAUNCHERS(512)
#endif
#if !defined(NM_TU) || NM_TU == 1024
NM_LAUNCHERS(1024)
#endif

#if !defined(NM_TU) || NM_TU == 0
// Queue set-up for the next cycle: reset the ticket counter, the per-configuration progress and the SM occupancy map, and rank
// the configurations by predicted cost. One segment per cycle (default): most expensive first (SM-aware placement, or
// longest-first tickets when there are more configurations than CTA slots). Several segments (NM_SEG_MOVES): cheapest
// first, so that the k-th CTA to become free continues the k-th cheapest chain, whose previous segment is the k-th to
// have finished (short waits).
__global__ void k_schedule(Dev d, long long cycle, int do_sort) {
  extern __shared__ unsigned long long sclk[];
  int* sorted = reinterpret_cast<int*>(sclk + d.nrep);
  for (int c = threadIdx.x; c < 5 + d.nrep + SMID_MAX; c += blockDim.x) d.sched[c] = (c == 0 && d.place) ? d.nrep : 0;   // placement hands out the first nrep tickets
  if (d.nhelp > 0) for (int c = threadIdx.x; c < HELP_STRIDE * d.nrep; c += blockDim.x)      // helper records; first ranking: last cycle's clocks
    d.help[c] = (c % HELP_STRIDE) == 8 ? (int)min((d.cta_clk[c / HELP_STRIDE] >> 10) + 1ull, 0x7fffffffull) : 0;
  if (!do_sort) { for (int c = threadIdx.x; c < d.nrep; c += blockDim.x) d.order[c] = c; return; }
  // predicted cost of the coming cycle: last cycle's work estimate of the configuration (listed pairs evaluated + list builds:
  // independent of which SM it ran on and with whom), scaled by the number of force evaluations the coming cycle will make
  // -- the move kinds are known in advance (counter-based RNG: the same rolls k_cycle will draw)
  for (int c = threadIdx.x; c < d.nrep; c += blockDim.x) {
    unsigned long long cost = d.cost[2 * c];
    const unsigned long long evals_last = d.cost[2 * c + 1];
    if (evals_last) {
      unsigned long long evals = 0;
      const int slot = d.cfg_slot[c];
      for (int mv = 0; mv < d.mod; mv++) {
        const Rng r = rng_make(d.seed_lo, d.seed_hi, (uint32_t)gslot(d, slot), (uint64_t)cycle * (uint64_t)d.mod + (uint64_t)mv);
        const double roll = rng_uniform(r, 0, P_ROLL);
        evals += roll <= d.ppos + d.pvol ? 1 : d.nstps;
      }
      cost = (unsigned long long)((double)cost * (double)evals / (double)evals_last);
    }
    sclk[c] = cost;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d.nrep; c += blockDim.x) {
    const unsigned long long v = sclk[c];
    int rank = 0;
    if (d.nseg > 1) { for (int o = 0; o < d.nrep; o++) rank += (sclk[o] < v) || (sclk[o] == v && o < c); }      // cheapest first
    else { for (int o = 0; o < d.nrep; o++) rank += (sclk[o] > v) || (sclk[o] == v && o < c); }                   // one segment: longest first
    d.order[rank] = c;
  }
}

// gen_mc_param (lammps_remcmc.py:726-745): ratios are the float32 values stored in the thermo record
__global__ void k_adapt(Dev d) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d.nrep) return;
  const int c = d.slot_cfg[k];
  const double* th = d.thermo + (size_t)k * NM_THERMO_WIDTH;
  for (int a = 0; a < 3; a++) {
    const double ratio = th[NM_TH_AP + a];
    double s = d.step[3 * c + a];
    if (ratio < 0.5) s = 0.9375 * s;
    if (ratio > 0.5) s = 1.0625 * s;
    d.step[3 * c + a] = s;
  }
  for (int a = 0; a < 6; a++) d.cnt[6 * c + a] = 0.0;
  // List skin of the configuration (not part of the reference: listed pairs outside the cutoff contribute
  // exact zeros and the rows keep their ascending order, so the skin reaches the results only through the moment at which
  // a rebuild re-wraps an atom that has left the box -- last bits). A cold solid hardly ever rebuilds
  // and pays for every listed pair, a fluid rebuilds once per move: from the last cycle's exact counters, the cost per
  // evaluation J(s) = listed pairs x ((rc + s) / (rc + s_now))^3 + build cost x builds x (s_now / s) (rebuild counts were
  // measured to fall as s^-1.1) is compared one step of 0.025 up and down, and the skin moves there if that saves more
  // than 1 %. The lists of a configuration whose skin changed are dropped (L0 = -1: rebuilt at its next position check).
  if (d.adapt_skin) {
    const unsigned long long* ct = d.rep_ct + (size_t)c * NM_COUNTER_WIDTH;
    const double pairs = (double)ct[NM_CT_LIST_PAIRS], builds = (double)ct[NM_CT_LIST_BUILDS];
    if (ct[NM_CT_FORCE_EVALS] > 0 && pairs > 0.0) {
      // cost of one (inner) list build in listed-pair evaluations: SMALL N^2 tiles + row walk; LARGE ~ 75 N (measured at N = 4000;
      // the outer searches depend on the outer skin and on diffusion, hardly on this one)
      const double s0 = d.skinc[c], bc = d.small ? d.build_cost * (double)d.N * (double)d.N : d.inner_cost * (double)d.N, step = 0.025;
      auto J = [&](double sn) { const double q = (d.rc + sn) / (d.rc + s0); return pairs * q * q * q + bc * builds * pow(s0 / sn, d.skin_pow); };
      const double j0 = J(s0);
      double sn = s0;
      if (s0 + step <= d.skin_hi + 1e-9 && J(s0 + step) < 0.99 * j0) sn = s0 + step;
      else if (s0 - step >= d.skin_lo - 1e-9 && J(s0 - step) < 0.99 * j0) sn = s0 - step;
      if (sn != s0) { d.skinc[c] = sn; d.L0[c] = -1.0; if (!d.small) d.L0o[c] = -1.0; }     // (the outer radius is rc + skin + outer skin)
    }
  }
}

// exchange payload: (pe + ke, vol) per local slot, as lammps_remcmc.py:791 reads them from STATE
__global__ void k_exchange_pack(Dev d, double* dst) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d.nrep) return;
  const int c = d.slot_cfg[k];
  dst[2 * k] = d.pe[c] + d.ke[c];
  dst[2 * k + 1] = pow(d.box[c], 3.0);
}

// replica_exchange (lammps_remcmc.py:776-803): one thread replays the sequential sweep of one LOCAL pressure row (exchanges never
// cross rows, :782-789, so a rank needs nothing from the other ranks to decide its swaps). table: (pe + ke, vol) per slot,
// either the local pack (global_table = 0, local slot order) or the all-gathered job-wide table (global slot order); et / pf
// come from the labels uploaded by nm_set_labels. The uniform of a pair is addressed by its GLOBAL draw index (row u, pair
// n of the row: u * NT (NT-1) / 2 + n), so the decisions do not depend on the row -> rank map. perm[k] = local source slot
// of local slot k.
__global__ void k_exchange_sweep(Dev d, const double* table, int global_table, const double* uniforms, long long cycle,
                                 double* scratch, int* perm, unsigned long long* swaps) {
  const int lr = blockIdx.x * blockDim.x + threadIdx.x, nt = d.nt;
  if (lr >= d.nrep / nt) return;
  const int u = d.row0 + lr * d.row_stride;
  double* e = scratch; double* vv = scratch + d.nrep;
  for (int t = 0; t < nt; t++) {
    const int k = lr * nt + t, ks = global_table ? u * nt + t : k;
    e[k] = table[2 * ks]; vv[k] = table[2 * ks + 1]; perm[k] = k;
  }
  unsigned long long draw = (unsigned long long)u * (unsigned long long)(nt * (nt - 1) / 2), sw = 0;
  for (int v = nt - 1; v >= 0; v--)
    for (int w = 0; w < v; w++) {
      const int i = lr * nt + v, j = lr * nt + w;
      const double de = e[i] - e[j], dvol = vv[i] - vv[j];
      const double dh = de * (1. / d.label[4 * i] - 1. / d.label[4 * j]) + (d.label[4 * i + 1] - d.label[4 * j + 1]) * dvol;
      const double m = exp(dh), crit = isnan(m) ? m : (m < 1.0 ? m : 1.0);
      double un;
      if (uniforms) un = uniforms[draw];
      else {
        uint32_t wd[4];
        philox4x32_10(d.seed_lo, d.seed_hi ^ NM_EXCH_KEY, (uint32_t)draw, P_EXCH, (uint32_t)cycle, (uint32_t)((unsigned long long)cycle >> 32), wd);
        un = u53(wd[0], wd[1]);
      }
      draw++;
      if (un <= crit) {
        sw++;
        double t = e[i]; e[i] = e[j]; e[j] = t; t = vv[i]; vv[i] = vv[j]; vv[j] = t;
        int q = perm[i]; perm[i] = perm[j]; perm[j] = q;
      }
    }
  atomicAdd(swaps, sw);
}
// apply the (row-local) permutation to the local slot -> configuration labels
__global__ void k_exchange_permute(Dev d, const int* perm, int* tmp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < d.nrep) tmp[k] = d.slot_cfg[perm[k]];
}
__global__ void k_exchange_commit(Dev d, const int* tmp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < d.nrep) { d.slot_cfg[k] = tmp[k]; d.cfg_slot[tmp[k]] = k; }
}

// host AoS (slot order) <-> device SoA (configuration order)
__global__ void k_scatter_state(Dev d, const double* x_aos, const double* v_aos, const double* box,
                                const double* dx, const double* dv, const double* dt) {
  const int k = blockIdx.x, c = d.slot_cfg[k], N = d.N, Npad = d.Npad;
  const size_t off = (size_t)c * 3 * Npad;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    for (int a = 0; a < 3; a++) {
      if (x_aos) d.x[off + a * Npad + i] = x_aos[((size_t)k * N + i) * 3 + a];   // re-wrapped by the list build that follows
      if (v_aos) d.v[off + a * Npad + i] = v_aos[((size_t)k * N + i) * 3 + a];
    }
  if (threadIdx.x == 0) {
    if (box) d.box[c] = box[k];
    if (dx) d.step[3 * c] = dx[k];
    if (dv) d.step[3 * c + 1] = dv[k];
    if (dt) d.step[3 * c + 2] = dt[k];
    if (x_aos || box) { d.L0[c] = -1.0; d.L0o[c] = -1.0; }   // new configuration: the old lists are meaningless
  }
}
__global__ void k_gather_state(Dev d, double* x_aos, double* v_aos, double* box, double* dx, double* dv, double* dt) {
  const int k = blockIdx.x, c = d.slot_cfg[k], N = d.N, Npad = d.Npad;
  const size_t off = (size_t)c * 3 * Npad;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    for (int a = 0; a < 3; a++) {
      if (x_aos) x_aos[((size_t)k * N + i) * 3 + a] = wrapg(d.x[off + a * Npad + i], d.box[c]);
      if (v_aos) v_aos[((size_t)k * N + i) * 3 + a] = d.v[off + a * Npad + i];
    }
  if (threadIdx.x == 0) {
    if (box) box[k] = d.box[c];
    if (dx) dx[k] = d.step[3 * c];
    if (dv) dv[k] = d.step[3 * c + 1];
    if (dt) dt[k] = d.step[3 * c + 2];
  }
}

#endif  // small kernels
}  // namespace nm

#if !defined(NM_TU) || NM_TU == 0

// =================================================================== host side / C-ABI
using namespace nm;

static thread_local char g_err[512] = "";
// shared by every translation unit of the library (not part of the C-ABI)
int nm_fail_msg(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return code;
}
#define fail nm_fail_msg
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return fail(NM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct nm_engine {
  nm_config cfg;
  Dev d;
  cudaStream_t stream; bool own_stream;
  int threads; size_t smem; int nsm; int grid;   // grid: CTAs of the persistent cycle kernel (all resident)
  std::vector<void*> allocs;
  double *stage_a, *stage_b, *stage_s;     // device staging: x/v AoS [nrep][3N], scalars [nrep][8]
  double *ex_table, *ex_uni, *ex_scratch; int *ex_perm, *ex_tmp; unsigned long long* ex_swaps;
  double* h_thermo; int* h_status;        // pinned host staging of nm_get_thermo (one synchronisation per cycle)
  long long* np_out;
  bool have_state, have_labels, have_thermo;
  int64_t launches;
};

template <typename T>
static int dev_alloc(nm_engine* h, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) return fail(NM_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
  e = cudaMemset(q, 0, n * sizeof(T));
  if (e != cudaSuccess) return fail(NM_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
  h->allocs.push_back(q); *p = static_cast<T*>(q);
  return NM_OK;
}
#define DA(ptr, n) do { int r_ = dev_alloc(h, &(ptr), (n)); if (r_) { nm_destroy(h); return r_; } } while (0)

extern "C" {

const char* nm_last_error(void) { return g_err; }
int nm_abi_version(void) { return NM_ABI_VERSION; }
int nm_device_count(void) {
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(NM_ENODEV, "no CUDA device: %s", cudaGetErrorString(e)); }
  return n;
}

int nm_create(const nm_config* cfg, nm_engine** out) {
  if (!cfg || !out) return fail(NM_EINVAL, "nm_create: null argument");
  if (cfg->struct_size != (int32_t)sizeof(nm_config)) return fail(NM_EINVAL, "nm_create: nm_config size mismatch (%d vs %zu)", cfg->struct_size, sizeof(nm_config));
  if (cfg->natoms < 2 || cfg->natoms > 8000) return fail(NM_EINVAL, "nm_create: natoms %d out of range [2, 8000] (13-bit neighbour indices; shared memory holds ~5000 atoms)", cfg->natoms);
  const int row_stride = cfg->row_stride > 0 ? cfg->row_stride : 1;
  if (cfg->n_rep < 1 || cfg->nt < 1 || cfg->n_rep % cfg->nt || cfg->rep_offset < 0 || cfg->rep_offset % cfg->nt || cfg->n_rep_global % cfg->nt ||
      (cfg->rep_offset / cfg->nt + (cfg->n_rep / cfg->nt - 1) * row_stride + 1) * cfg->nt > cfg->n_rep_global)
    return fail(NM_EINVAL, "nm_create: local slots must be whole pressure rows of the global grid (n_rep=%d rep_offset=%d row_stride=%d nt=%d global=%d)", cfg->n_rep, cfg->rep_offset, row_stride, cfg->nt, cfg->n_rep_global);
  if (cfg->precision != 64 && cfg->precision != 32 && cfg->precision != 0) return fail(NM_EINVAL, "nm_create: precision must be 64 or 32 (got %d)", cfg->precision);
  if (cfg->nstps < 1 || cfg->mod < 0 || cfg->ppos < 0 || cfg->pvol < 0 || cfg->ppos + cfg->pvol > 1.0 + 1e-12) return fail(NM_EINVAL, "nm_create: bad move parameters");
  if (!(cfg->rc > 0) || !(cfg->mass > 0)) return fail(NM_EINVAL, "nm_create: rc and mass must be positive");
  int ndev = nm_device_count();
  if (ndev < 0) return ndev;
  if (cfg->device < 0 || cfg->device >= ndev) return fail(NM_ENODEV, "nm_create: device %d not in [0,%d)", cfg->device, ndev);
  CK(cudaSetDevice(cfg->device));
  nm_engine* h = new (std::nothrow) nm_engine();
  if (!h) return fail(NM_ENOMEM, "nm_create: host allocation failed");
  h->cfg = *cfg;
  h->own_stream = cfg->stream == nullptr;
  if (h->own_stream) { cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking); if (e != cudaSuccess) { delete h; return fail(NM_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); } }
  else h->stream = (cudaStream_t)cfg->stream;
  Dev& d = h->d;
  memset(&d, 0, sizeof d);
  const int N = cfg->natoms, nrep = cfg->n_rep;
  d.N = N; d.Npad = ((N + 1) + 31) & ~31; d.nrep = nrep; d.nrep_global = cfg->n_rep_global; d.rep_offset = cfg->rep_offset; d.nt = cfg->nt; d.row0 = cfg->rep_offset / cfg->nt; d.row_stride = row_stride;
  d.nstps = cfg->nstps; d.mod = cfg->mod; d.bulk = cfg->bulk_move; d.text_rounding = cfg->text_rounding;
  d.ppos = cfg->ppos; d.pvol = cfg->pvol; d.lat = cfg->lat_scale; d.mass = cfg->mass; d.rc = cfg->rc;
  // default skin: tuned at the stationary state of the default workload (step sizes adapted to 50 % acceptance, 1.6
  // rebuilds per move): 0.4 where a rebuild costs 3 evaluations (hit-matrix builds), 0.3 with the cheaper two-level lists
  // Iterative single-atom sweeps at N > 768 (bulk_move = 0): 0.5, so that a trial's displacement plus the largest one of the
  // sweep so far (both up to sqrt(3) dx lat ~ 0.23 at the adapted step size) stay inside the skin of a list built at the
  // start of the sweep, and the inner column (31 quads: one pass of a warp) serves every trial.
  d.skin = cfg->skin > 0 ? cfg->skin : (N <= NSMALL ? 0.4 : (cfg->bulk_move ? 0.3 : 0.5));
  if (const char* ev = getenv("NM_SKIN")) { const double v = atof(ev); if (cfg->skin <= 0 && v > 0.05 && v < 1.0) d.skin = v; }   // experiments: another default
  // SMALL mode, default skin: tuned per configuration between 0.2 and 0.5 (k_adapt; NM_SKIN_RANGE=lo,hi[,p] for experiments); NM_FIXED_SKIN=1 keeps it fixed
  // (LARGE mode: measured on the C3 shard with 0.2 .. 0.4 around the default 0.3 and inner-build costs of 75 .. 600 N pair
  // evaluations: 121.5 .. 125.2 ms against 119.3 ms with the fixed skin -- a skin change also drops the outer list, and the
  // hot chains that set the kernel time are the ones that keep changing; NM_ADAPT_SKIN_LARGE=1 turns it on for experiments)
  d.adapt_skin = cfg->skin <= 0 && (N <= NSMALL || (cfg->bulk_move && getenv("NM_ADAPT_SKIN_LARGE"))) && cfg->precision != 32 && !getenv("NM_FIXED_SKIN");
  d.skin_lo = 0.2; d.skin_hi = N <= NSMALL ? 0.5 : 0.4; d.skin_pow = 1.0;
  d.inner_cost = 75.0;
  if (const char* ev = getenv("NM_INNER_COST")) { const double v = atof(ev); if (v > 0) d.inner_cost = v; }
  if (const char* ev = getenv("NM_SKIN_RANGE")) { double a = 0, b = 0, c = 0; const int n = sscanf(ev, "%lf,%lf,%lf", &a, &b, &c); if (n >= 2 && a > 0.05 && b >= a && b <= 0.6) { d.skin_lo = a; d.skin_hi = b; } if (n >= 3 && c > 0) d.skin_pow = c; }
  // outer skin (LARGE mode only): stationary N = 4000 grid: 245 ms per cycle at 1.0, 215 at 1.3, 221 at 1.6
  d.oskin = cfg->skin_outer > 0 ? cfg->skin_outer : 1.3;
  d.seed_lo = (uint32_t)cfg->seed; d.seed_hi = (uint32_t)(cfg->seed >> 32);
  {
    // list capacities: neighbours inside the list radius at the densest state we expect (rho* 1.6) plus slack,
    // plus padding of the image groups (at most 8 per atom when the box is >= 2 rlo) to whole quads
    const double rl = d.rc + (d.adapt_skin ? d.skin_hi : d.skin), rlo = rl + d.oskin;      // capacity for the largest skin in use
    int maxnb = (int)(4.18879 * rl * rl * rl * 1.6) + 16, maxnbo = (int)(4.18879 * rlo * rlo * rlo * 1.6) + 16;
    if (maxnb > N - 1) maxnb = N - 1;
    if (maxnbo > N - 1) maxnbo = N - 1;
    if (maxnb > maxnbo) maxnb = maxnbo;
    if (maxnb < 1) maxnb = 1;
    if (maxnbo < 1) maxnbo = 1;
    d.maxnbo = maxnbo;
    d.maxq = (maxnb + 3) / 4 + 8;
    d.maxqo = ((maxnbo + 3) / 4 + 8 + 1) & ~1;
  }
  h->threads = N <= 256 ? 256 : (N <= 512 ? 512 : 1024);   // 64 registers/thread: 32 warps per SM hide the FP64 latency
  d.f32 = cfg->precision == 32;
  d.small = N <= NSMALL;
  h->smem = smem_bytes(d.Npad, N, d.small, h->threads);
  { cudaDeviceProp pr; if (cudaGetDeviceProperties(&pr, cfg->device) == cudaSuccess) h->nsm = pr.multiProcessorCount; else h->nsm = 148; }
  d.nsm = h->nsm; d.per_sm = 1;
  if (h->smem > 227 * 1024) { nm_destroy(h); return fail(NM_EINVAL, "nm_create: natoms %d needs %zu B of shared memory per CTA (> 227 KB)", N, h->smem); }
  const size_t per = (size_t)nrep * 3 * d.Npad;
  DA(d.x, per); DA(d.v, per); DA(d.f, per); DA(d.xs, per); DA(d.vs, per); DA(d.fs, per); DA(d.x0, 2 * per);
  DA(d.list, (size_t)nrep * 2 * (d.maxq + LIST_SPARE_ROWS) * d.Npad); DA(d.lcur, nrep);
  if (d.small) { DA(d.hbT, (size_t)nrep * (d.Npad / 32) * d.Npad); DA(d.ginfo, (size_t)nrep * 2 * d.Npad); }       // one spare row per buffer: the loop prefetches one quad ahead
  DA(d.ltmp, (size_t)nrep * ((d.maxnbo + 3) & ~3) * d.Npad); DA(d.nnb, (size_t)nrep * 2 * d.Npad); DA(d.micmode, nrep);
  DA(d.olist, ((size_t)nrep * d.maxqo + 2) * d.Npad); DA(d.ocode, ((size_t)nrep * d.maxqo + 2) * d.Npad); DA(d.onq, (size_t)nrep * d.Npad);
  DA(d.x0o, per); DA(d.L0o, nrep);
  DA(d.box, nrep); DA(d.pe, nrep); DA(d.w, nrep); DA(d.ke, nrep); DA(d.L0, nrep); DA(d.list_pairs, nrep);
  DA(d.step, 3 * (size_t)nrep); DA(d.cnt, 6 * (size_t)nrep);
  DA(d.cfg_slot, nrep); DA(d.slot_cfg, nrep); DA(d.status, nrep); DA(d.cta_clk, nrep); DA(d.cost, 2 * (size_t)nrep); DA(d.rep_ct, (size_t)nrep * NM_COUNTER_WIDTH); DA(d.mv_clk, (size_t)nrep * 4); DA(d.order, nrep); DA(d.sched, (size_t)nrep + 5 + SMID_MAX);
  DA(d.skinc, nrep);
  DA(d.label, 4 * (size_t)nrep); DA(d.thermo, (size_t)nrep * NM_THERMO_WIDTH); DA(d.counters, NM_COUNTER_WIDTH);
  DA(h->stage_a, (size_t)nrep * 3 * N); DA(h->stage_b, (size_t)nrep * 3 * N); DA(h->stage_s, (size_t)nrep * 8);
  const int nsg = cfg->n_rep_global;
  DA(h->ex_table, 2 * (size_t)nrep); DA(h->ex_uni, (size_t)nsg * cfg->nt); DA(h->ex_scratch, 2 * (size_t)nrep);
  DA(h->ex_perm, nrep); DA(h->ex_tmp, nrep); DA(h->ex_swaps, 1); DA(h->np_out, nrep);
  if (cudaHostAlloc((void**)&h->h_thermo, sizeof(double) * (size_t)nrep * NM_THERMO_WIDTH, cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc((void**)&h->h_status, sizeof(int) * (size_t)nrep, cudaHostAllocDefault) != cudaSuccess) { nm_destroy(h); return fail(NM_ENOMEM, "nm_create: pinned host staging allocation failed"); }
  {
    std::vector<int> id(nrep); for (int k = 0; k < nrep; k++) id[k] = k;
    std::vector<double> neg(nrep, -1.0);
    { std::vector<double> sk(nrep, d.skin); if (cudaMemcpy(d.skinc, sk.data(), sizeof(double) * nrep, cudaMemcpyHostToDevice) != cudaSuccess) { nm_destroy(h); return fail(NM_ECUDA, "nm_create: init copy failed"); } }
    cudaError_t e1 = cudaMemcpy(d.cfg_slot, id.data(), sizeof(int) * nrep, cudaMemcpyHostToDevice);
    cudaError_t e2 = cudaMemcpy(d.slot_cfg, id.data(), sizeof(int) * nrep, cudaMemcpyHostToDevice);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(d.order, id.data(), sizeof(int) * nrep, cudaMemcpyHostToDevice);
    cudaError_t e3 = cudaMemcpy(d.L0, neg.data(), sizeof(double) * nrep, cudaMemcpyHostToDevice);
    if (e3 == cudaSuccess) e3 = cudaMemcpy(d.L0o, neg.data(), sizeof(double) * nrep, cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { nm_destroy(h); return fail(NM_ECUDA, "nm_create: init copy failed"); }
  }
  cudaError_t e = cudaSuccess;
  const int sm = (int)h->smem;
  e = h->threads == 256 ? set_smem_256(h->smem) : (h->threads == 512 ? set_smem_512(h->smem) : set_smem_1024(h->smem));
  if (e == cudaSuccess && h->threads == 1024) e = set_smem_1024h(h->smem);
  if (e != cudaSuccess) { nm_destroy(h); return fail(NM_ECUDA, "cudaFuncSetAttribute(smem=%zu): %s", h->smem, cudaGetErrorString(e)); }
  {
    // persistent cycle kernel: every CTA of the grid must be resident (a CTA may wait for a segment held by another one)
    int occ = h->threads == 256 ? occupancy_256(h->smem) : (h->threads == 512 ? occupancy_512(h->smem) : occupancy_1024(h->smem));
    if (occ < 1) occ = 1;
    d.per_sm = occ;
    const long long slots = (long long)occ * h->nsm;
    h->grid = (int)(nrep < slots ? nrep : slots);
    // segment length: NM_SEG_MOVES moves (default: the whole cycle in one segment; measured at C2: 4-move segments
    // without placement 46.2 ms, one segment 49 ms, SM-aware placement of whole cycles: see DESIGN 4.1)
    int seg_moves = 0;
    if (const char* ev = getenv("NM_SEG_MOVES")) seg_moves = atoi(ev);
    if (seg_moves <= 0 || seg_moves > d.mod) seg_moves = d.mod > 0 ? d.mod : 1;
    d.seg_moves = seg_moves;
    d.nseg = d.mod > 0 ? (d.mod + seg_moves - 1) / seg_moves : 1;
    // measured (B200): N = 4000 (lists streamed from HBM by every evaluation) 140.7 -> 135.5 ms per C3 step with the row being
    // loaded prefetched; N = 500 (lists mostly L2-resident) 43.4 -> 43.7 ms: off there
    d.list_pf = d.small ? -1 : 0;
    if (const char* ev = getenv("NM_LIST_PF")) d.list_pf = atoi(ev);
    if (d.list_pf > LIST_SPARE_ROWS - 2) d.list_pf = LIST_SPARE_ROWS - 2;      // stays inside the spare rows of the list buffer
    d.build_cost = NM_BUILD_COST;
    if (const char* ev = getenv("NM_BUILD_COST")) { const double v = atof(ev); if (v > 0) d.build_cost = v; }
    d.place = occ == 2 && nrep > h->nsm && nrep <= slots && h->nsm <= SMID_MAX && !getenv("NM_NO_PLACEMENT");
    // force helpers: LARGE mode with rows above 2 * blockDim, one segment per cycle, spare CTA slots (NM_NO_HELPERS: off;
    // NM_HELPERS=n: at most n). The grid grows by the helpers; all of it fits the device at once.
    d.nhelp = 0;
    // (not for iterative single-atom sweeps or very short cycles: few full evaluations per launch, the hand-shakes would cost more than they save)
    if (!d.small && !d.f32 && h->threads == 1024 && N > 2 * h->threads && d.nseg == 1 && nrep < slots && d.bulk && d.mod >= 4 && !getenv("NM_NO_HELPERS")) {
      long long nh = slots - nrep;
      if (nh > nrep) nh = nrep;
      if (const char* ev = getenv("NM_HELPERS")) { const long long lim = atoll(ev); if (lim >= 0 && lim < nh) nh = lim; }
      d.nhelp = (int)nh;
    }
    d.help_quantum = 32;
    if (const char* ev = getenv("NM_HELP_QUANTUM")) { const int q = atoi(ev); if (q > 0) d.help_quantum = q; }
    if (d.nhelp > 0) {
      DA(d.help, (size_t)HELP_STRIDE * nrep); DA(d.helpd, 4 * (size_t)nrep); DA(d.hpart, (size_t)nrep * 4 * h->threads);
      h->grid += d.nhelp;
    }
  }
  *out = h;
  return NM_OK;
}

int nm_destroy(nm_engine* h) {
  if (!h) return NM_OK;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  if (h->h_thermo) cudaFreeHost(h->h_thermo);
  if (h->h_status) cudaFreeHost(h->h_status);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return NM_OK;
}

int nm_set_stream(nm_engine* h, void* s) {
  if (!h) return fail(NM_EINVAL, "null engine");
  CK(cudaStreamSynchronize(h->stream));
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  h->stream = (cudaStream_t)s;
  return NM_OK;
}
int nm_synchronize(nm_engine* h) {
  if (!h) return fail(NM_EINVAL, "null engine");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  return NM_OK;
}

static int check_status(nm_engine* h) {
  int* st = h->h_status;
  CK(cudaMemcpyAsync(st, h->d.status, sizeof(int) * h->d.nrep, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < h->d.nrep; c++) {
    if (st[c] & ST_BOX) return fail(NM_EBOX, "configuration %d: box side below 2*rc (minimum image invalid)", c);
    if (st[c] & ST_NEIGH) return fail(NM_ENEIGH, "configuration %d: neighbour list capacity (%d inner quads / %d outer entries) exceeded", c, h->d.maxq, h->d.maxnbo);
  }
  return NM_OK;
}

static int launch_eval(nm_engine* h, double* pe, double* w, double* f_aos, long long* np_) {
  if (h->threads == 256) launch_eval_256(h->d, pe, w, f_aos, np_, h->smem, h->stream);
  else if (h->threads == 512) launch_eval_512(h->d, pe, w, f_aos, np_, h->smem, h->stream);
  else launch_eval_1024(h->d, pe, w, f_aos, np_, h->smem, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return NM_OK;
}

int nm_set_state(nm_engine* h, const double* x, const double* v, const double* box,
                 const double* dx, const double* dv, const double* dt) {
  if (!h) return fail(NM_EINVAL, "null engine");
  CK(cudaSetDevice(h->cfg.device));
  const int nrep = h->d.nrep; const size_t n3 = (size_t)nrep * 3 * h->d.N;
  if (!h->have_state && (!x || !box)) return fail(NM_ESTATE, "nm_set_state: the first upload needs positions and box");
  if (x) CK(cudaMemcpyAsync(h->stage_a, x, sizeof(double) * n3, cudaMemcpyHostToDevice, h->stream));
  if (v) CK(cudaMemcpyAsync(h->stage_b, v, sizeof(double) * n3, cudaMemcpyHostToDevice, h->stream));
  const double* src[4] = { box, dx, dv, dt };
  for (int q = 0; q < 4; q++) if (src[q]) CK(cudaMemcpyAsync(h->stage_s + (size_t)q * nrep, src[q], sizeof(double) * nrep, cudaMemcpyHostToDevice, h->stream));
  k_scatter_state<<<nrep, 256, 0, h->stream>>>(h->d, x ? h->stage_a : nullptr, v ? h->stage_b : nullptr,
      box ? h->stage_s : nullptr, dx ? h->stage_s + nrep : nullptr, dv ? h->stage_s + 2 * nrep : nullptr, dt ? h->stage_s + 3 * nrep : nullptr);
  h->launches++;
  CK(cudaGetLastError());
  h->have_state = true;
  if (x || box) {                       // the 'run 0' of init_lammps
    int r = launch_eval(h, nullptr, nullptr, nullptr, nullptr); if (r) return r;
    return check_status(h);
  }
  CK(cudaStreamSynchronize(h->stream));
  return NM_OK;
}

int nm_get_state(nm_engine* h, double* x, double* v, double* box, double* dx, double* dv, double* dt) {
  if (!h) return fail(NM_EINVAL, "null engine");
  if (!h->have_state) return fail(NM_ESTATE, "nm_get_state: no state uploaded");
  CK(cudaSetDevice(h->cfg.device));
  const int nrep = h->d.nrep; const size_t n3 = (size_t)nrep * 3 * h->d.N;
  k_gather_state<<<nrep, 256, 0, h->stream>>>(h->d, x ? h->stage_a : nullptr, v ? h->stage_b : nullptr,
      h->stage_s, h->stage_s + nrep, h->stage_s + 2 * nrep, h->stage_s + 3 * nrep);
  h->launches++;
  CK(cudaGetLastError());
  if (x) CK(cudaMemcpyAsync(x, h->stage_a, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
  if (v) CK(cudaMemcpyAsync(v, h->stage_b, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
  double* dst[4] = { box, dx, dv, dt };
  for (int q = 0; q < 4; q++) if (dst[q]) CK(cudaMemcpyAsync(dst[q], h->stage_s + (size_t)q * nrep, sizeof(double) * nrep, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NM_OK;
}

int nm_set_labels(nm_engine* h, const double* et, const double* pf, const double* temp, const double* temp_vel) {
  if (!h || !et || !pf || !temp || !temp_vel) return fail(NM_EINVAL, "nm_set_labels: null argument");
  CK(cudaSetDevice(h->cfg.device));
  std::vector<double> lab(4 * (size_t)h->d.nrep);
  for (int k = 0; k < h->d.nrep; k++) {
    if (!(et[k] > 0)) return fail(NM_EINVAL, "nm_set_labels: et[%d] must be positive", k);
    lab[4 * k] = et[k]; lab[4 * k + 1] = pf[k]; lab[4 * k + 2] = temp[k]; lab[4 * k + 3] = temp_vel[k];
  }
  CK(cudaMemcpyAsync(h->d.label, lab.data(), sizeof(double) * lab.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_labels = true;
  return NM_OK;
}

int nm_eval(nm_engine* h, double* pe, double* w, double* f, int64_t* npairs) {
  if (!h) return fail(NM_EINVAL, "null engine");
  if (!h->have_state) return fail(NM_ESTATE, "nm_eval: no state uploaded");
  CK(cudaSetDevice(h->cfg.device));
  const int nrep = h->d.nrep; const size_t n3 = (size_t)nrep * 3 * h->d.N;
  int r = launch_eval(h, h->stage_s, h->stage_s + nrep, f ? h->stage_a : nullptr, h->np_out); if (r) return r;
  if (pe) CK(cudaMemcpyAsync(pe, h->stage_s, sizeof(double) * nrep, cudaMemcpyDeviceToHost, h->stream));
  if (w) CK(cudaMemcpyAsync(w, h->stage_s + nrep, sizeof(double) * nrep, cudaMemcpyDeviceToHost, h->stream));
  if (f) CK(cudaMemcpyAsync(f, h->stage_a, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
  if (npairs) CK(cudaMemcpyAsync(npairs, h->np_out, sizeof(long long) * nrep, cudaMemcpyDeviceToHost, h->stream));
  return check_status(h);
}

int nm_run_cycle(nm_engine* h, int64_t cycle) {
  if (!h) return fail(NM_EINVAL, "null engine");
  if (!h->have_state || !h->have_labels) return fail(NM_ESTATE, "nm_run_cycle: state and labels must be uploaded first");
  CK(cudaSetDevice(h->cfg.device));
  {                                                       // queue reset + cost ranks from the last cycle's clocks and this cycle's move kinds
    const int do_sort = h->d.nrep > 1 && h->d.nrep <= 4096 && (h->d.nseg > 1 || h->d.nrep > h->nsm);
    k_schedule<<<1, 1024, do_sort ? h->d.nrep * (sizeof(unsigned long long) + sizeof(int)) : 0, h->stream>>>(h->d, (long long)cycle, do_sort);
    h->launches++;
    CK(cudaGetLastError());
  }
  if (h->threads == 256) launch_cycle_256(h->d, (long long)cycle, h->grid, h->smem, h->stream);
  else if (h->threads == 512) launch_cycle_512(h->d, (long long)cycle, h->grid, h->smem, h->stream);
  else if (h->d.nhelp > 0 || getenv("NM_FORCE_HELPER_KERNEL"))      // (the variable: experiments with the helper-capable kernel alone)
    launch_cycle_1024h(h->d, (long long)cycle, h->grid, h->smem, h->stream);
  else launch_cycle_1024(h->d, (long long)cycle, h->grid, h->smem, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  h->have_thermo = true;
  return NM_OK;
}

int nm_velocity_create(nm_engine* h, int64_t tag) {
  if (!h) return fail(NM_EINVAL, "null engine");
  if (!h->have_state || !h->have_labels) return fail(NM_ESTATE, "nm_velocity_create: state and labels must be uploaded first");
  CK(cudaSetDevice(h->cfg.device));
  if (h->threads == 256) launch_velinit_256(h->d, (long long)tag, h->smem, h->stream);
  else if (h->threads == 512) launch_velinit_512(h->d, (long long)tag, h->smem, h->stream);
  else launch_velinit_1024(h->d, (long long)tag, h->smem, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return NM_OK;
}

int nm_get_thermo(nm_engine* h, double* out) {
  if (!h || !out) return fail(NM_EINVAL, "nm_get_thermo: null argument");
  if (!h->have_thermo) return fail(NM_ESTATE, "nm_get_thermo: no cycle has run");
  CK(cudaSetDevice(h->cfg.device));
  // thermo + status into pinned staging, one synchronisation, then a host copy into the caller's (pageable) array
  const size_t nb = sizeof(double) * (size_t)h->d.nrep * NM_THERMO_WIDTH;
  CK(cudaMemcpyAsync(h->h_thermo, h->d.thermo, nb, cudaMemcpyDeviceToHost, h->stream));
  int r = check_status(h); if (r) return r;
  memcpy(out, h->h_thermo, nb);
  return NM_OK;
}

int nm_adapt(nm_engine* h) {
  if (!h) return fail(NM_EINVAL, "null engine");
  if (!h->have_thermo) return fail(NM_ESTATE, "nm_adapt: no cycle has run");
  CK(cudaSetDevice(h->cfg.device));
  k_adapt<<<(h->d.nrep + 127) / 128, 128, 0, h->stream>>>(h->d);
  h->launches++;
  CK(cudaGetLastError());
  return NM_OK;
}

int nm_exchange_pack(nm_engine* h, void* dev_dst) {
  if (!h || !dev_dst) return fail(NM_EINVAL, "nm_exchange_pack: null argument");
  if (!h->have_state) return fail(NM_ESTATE, "nm_exchange_pack: no state uploaded");
  CK(cudaSetDevice(h->cfg.device));
  k_exchange_pack<<<(h->d.nrep + 127) / 128, 128, 0, h->stream>>>(h->d, (double*)dev_dst);
  h->launches++;
  CK(cudaGetLastError());
  return NM_OK;
}

// sweep of the local rows from `table` (local pack or job-wide table) + permutation of the local labels
static int exchange_sweep(nm_engine* h, const double* dev_table, int global_table, const double* uniforms, int64_t cycle,
                          int32_t* perm_out, int64_t* swaps_out) {
  const int nt = h->d.nt, nrows = h->d.nrep / nt, npg = h->d.nrep_global / nt;
  const size_t nu = (size_t)npg * nt * (nt - 1) / 2;
  if (uniforms && nu) CK(cudaMemcpyAsync(h->ex_uni, uniforms, sizeof(double) * nu, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->ex_swaps, 0, sizeof(unsigned long long), h->stream));
  k_exchange_sweep<<<(nrows + 31) / 32, 32, 0, h->stream>>>(h->d, dev_table, global_table, uniforms ? h->ex_uni : nullptr, (long long)cycle,
                                                            h->ex_scratch, h->ex_perm, h->ex_swaps);
  k_exchange_permute<<<(h->d.nrep + 127) / 128, 128, 0, h->stream>>>(h->d, h->ex_perm, h->ex_tmp);
  k_exchange_commit<<<(h->d.nrep + 127) / 128, 128, 0, h->stream>>>(h->d, h->ex_tmp);
  h->launches += 3;
  CK(cudaGetLastError());
  if (perm_out || swaps_out) {
    unsigned long long sw = 0;
    if (perm_out) CK(cudaMemcpyAsync(perm_out, h->ex_perm, sizeof(int) * h->d.nrep, cudaMemcpyDeviceToHost, h->stream));
    if (swaps_out) CK(cudaMemcpyAsync(&sw, h->ex_swaps, sizeof sw, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (swaps_out) *swaps_out = (int64_t)sw;
  }
  return NM_OK;
}

int nm_exchange_apply(nm_engine* h, const void* dev_table_global, const double* uniforms, int64_t cycle,
                      int32_t* perm_out, int64_t* swaps_out) {
  if (!h || !dev_table_global) return fail(NM_EINVAL, "nm_exchange_apply: null argument");
  if (!h->have_labels) return fail(NM_ESTATE, "nm_exchange_apply: labels not set");
  CK(cudaSetDevice(h->cfg.device));
  return exchange_sweep(h, (const double*)dev_table_global, 1, uniforms, cycle, perm_out, swaps_out);
}

int nm_exchange(nm_engine* h, const double* uniforms, int64_t cycle, int32_t* perm_out, int64_t* swaps_out) {
  if (!h) return fail(NM_EINVAL, "null engine");
  if (!h->have_labels) return fail(NM_ESTATE, "nm_exchange: labels not set");
  int r = nm_exchange_pack(h, h->ex_table); if (r) return r;
  return exchange_sweep(h, h->ex_table, 0, uniforms, cycle, perm_out, swaps_out);
}

int nm_get_counters(nm_engine* h, uint64_t* out) {
  if (!h || !out) return fail(NM_EINVAL, "nm_get_counters: null argument");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(out, h->d.counters, sizeof(uint64_t) * NM_COUNTER_WIDTH, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return NM_OK;
}
int nm_reset_counters(nm_engine* h) {
  if (!h) return fail(NM_EINVAL, "null engine");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemsetAsync(h->d.counters, 0, sizeof(uint64_t) * NM_COUNTER_WIDTH, h->stream));
  return NM_OK;
}
int64_t nm_launch_count(nm_engine* h) { return h ? h->launches : 0; }

int nm_get_cta_clocks(nm_engine* h, uint64_t* out) {
  if (!h || !out) return fail(NM_EINVAL, "nm_get_cta_clocks: null argument");
  CK(cudaSetDevice(h->cfg.device));
  std::vector<unsigned long long> clk(h->d.nrep); std::vector<int> cs(h->d.nrep);
  CK(cudaMemcpyAsync(clk.data(), h->d.cta_clk, sizeof(unsigned long long) * h->d.nrep, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(cs.data(), h->d.cfg_slot, sizeof(int) * h->d.nrep, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < h->d.nrep; c++) out[cs[c]] = clk[c];
  return NM_OK;
}

int nm_get_replica_counters(nm_engine* h, uint64_t* out) {
  if (!h || !out) return fail(NM_EINVAL, "nm_get_replica_counters: null argument");
  CK(cudaSetDevice(h->cfg.device));
  const size_t W = NM_COUNTER_WIDTH;
  std::vector<unsigned long long> ct(h->d.nrep * W); std::vector<int> cs(h->d.nrep);
  CK(cudaMemcpyAsync(ct.data(), h->d.rep_ct, sizeof(unsigned long long) * ct.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(cs.data(), h->d.cfg_slot, sizeof(int) * h->d.nrep, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < h->d.nrep; c++) for (size_t k = 0; k < W; k++) out[cs[c] * W + k] = ct[c * W + k];
  return NM_OK;
}

}  // extern "C"
#endif  // host C-ABI
