/*
 * nm_oracle.c -- CPU restatement of the neuralMelting hot path. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (neuralmelting_b200/) never does.
 *
 * What is restated (reference = /root/reference/scripts/, file:line cited per function):
 *   - lj/cut energy / force / virial as the LAMMPS deck of lammps_remcmc.py:364-372 defines it
 *   - the Monte Carlo moves of lammps_remcmc.py:477-658, the cycle :665-691,
 *     the adaptation :726-745, the replica exchange :776-803
 *   - the RDF of lammps_distr.py:123-135
 *
 * PARITY UNPINNED for the LJ/MD physics: the arithmetic of that part lives in LAMMPS
 * (third-party, un-vendored, version not pinned by the reference, not installed here), so
 * it cannot be run to generate golden vectors. The restatement follows the published
 * pair_lj_cut / fix_nve / velocity / displace_atoms / change_box semantics and is pinned
 * by self-derived anchors (analytic fcc lattice sums, finite-difference forces, two
 * independent implementations below agreeing to 1e-13) -- see tests/test_oracle_lj.py.
 * The exchange, adaptation, text formats and the RDF ARE pinned against the reference's
 * own functions run in the build container (oracle/gen_golden.py -> tests/golden/).
 *
 * RNG: the reference draws from NumPy's global Mersenne twister and LAMMPS' Park-Miller
 * generator seeded by randint (not reproducible across workers). This restatement and the
 * CUDA engine share ONE counter-based convention (Philox4x32-10, documented below) so that
 * the two can be compared move by move.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ RNG */
/* Philox4x32-10 (Salmon et al. 2011). key = {seed_lo, seed_hi ^ slot_global};
 * counter = {index, purpose, m_lo, m_hi}, m = cycle*MOD + move_in_cycle.
 * For the exchange sweep: key = {seed_lo, seed_hi ^ 0xE8C4A93B}, m = cycle. */
enum { P_ROLL = 0, P_HMC_VEL = 1, P_HMC_ACC = 2, P_VMC_PROP = 3, P_VMC_ACC = 4,
       P_BULK_DISP = 5, P_BULK_ACC = 6, P_ITER_DISP = 7, P_ITER_ACC = 8, P_EXCH = 9 };
#define EXCH_KEY 0xE8C4A93Bu

static void philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                   uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u53(uint32_t hi, uint32_t lo) {          /* [0,1) with 53 bits, like np.random.rand */
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
static double u53_open(uint32_t hi, uint32_t lo) {     /* (0,1] for the log of Box-Muller */
  return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1) * (1.0 / 9007199254740992.0);
}
typedef struct { uint32_t k0, k1, m_lo, m_hi; } rng_t;
static rng_t rng_make(uint64_t seed, uint32_t slot, uint64_t m) {
  rng_t r = { (uint32_t)seed, (uint32_t)(seed >> 32) ^ slot, (uint32_t)m, (uint32_t)(m >> 32) };
  return r;
}
static void rng_words(const rng_t* r, uint32_t index, uint32_t purpose, uint32_t w[4]) {
  philox(r->k0, r->k1, index, purpose, r->m_lo, r->m_hi, w);
}
static double rng_uniform(const rng_t* r, uint32_t index, uint32_t purpose) {
  uint32_t w[4]; rng_words(r, index, purpose, w); return u53(w[0], w[1]);
}
/* three uniforms for atom i: words of index 2i give (u0,u1), index 2i+1 gives u2 */
static void rng_uniform3(const rng_t* r, uint32_t i, uint32_t purpose, double u[3]) {
  uint32_t a[4], b[4];
  rng_words(r, 2 * i, purpose, a); rng_words(r, 2 * i + 1, purpose, b);
  u[0] = u53(a[0], a[1]); u[1] = u53(a[2], a[3]); u[2] = u53(b[0], b[1]);
}
/* three standard normals for atom i (Box-Muller) */
static void rng_gauss3(const rng_t* r, uint32_t i, uint32_t purpose, double g[3]) {
  uint32_t a[4], b[4];
  rng_words(r, 2 * i, purpose, a); rng_words(r, 2 * i + 1, purpose, b);
  double r0 = sqrt(-2.0 * log(u53_open(a[0], a[1]))), t0 = 6.283185307179586477 * u53(a[2], a[3]);
  double r1 = sqrt(-2.0 * log(u53_open(b[0], b[1]))), t1 = 6.283185307179586477 * u53(b[2], b[3]);
  g[0] = r0 * cos(t0); g[1] = r0 * sin(t0); g[2] = r1 * cos(t1);
}
ORC_API void orc_philox(uint32_t k0, uint32_t k1, const uint32_t c[4], uint32_t out[4]) {
  philox(k0, k1, c[0], c[1], c[2], c[3], out);
}

/* '%f' text round trip: the reference hands values to LAMMPS as '%f' strings
 * (lammps_remcmc.py:403,407,416,466,483,570,604,607). */
ORC_API double orc_round6(double x) {
  char buf[64]; snprintf(buf, sizeof buf, "%f", x); return strtod(buf, NULL);
}

/* ------------------------------------------------------------------ a-1: lj/cut */
/* LAMMPS remap into [0,L): domain.cpp::remap semantics for an orthogonal periodic box */
static inline double wrap1(double x, double L) {
  if (x < 0.0) x += L;
  if (x >= L) x -= L;
  if (x < 0.0) x = 0.0;
  return x;
}
ORC_API void orc_wrap(int n, double* x, double L) {
  for (int i = 0; i < 3 * n; i++) x[i] = wrap1(x[i], L);
}

/* pair_style lj/cut rc; pair_coeff 1 1 1.0 1.0 rc (lammps_remcmc.py:365-367):
 * strict rsq < rc^2, no shift, no tail. Brute force, minimum image, any x. */
ORC_API void orc_lj_eval_n2(int n, const double* x, double L, double rc,
                            double* pe, double* w, double* f, int64_t* npairs) {
  double e = 0.0, vir = 0.0, rc2 = rc * rc; int64_t np_ = 0;
  if (f) memset(f, 0, sizeof(double) * 3 * n);
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      double d[3], rsq = 0.0;
      for (int c = 0; c < 3; c++) {
        d[c] = x[3 * i + c] - x[3 * j + c];
        d[c] -= L * nearbyint(d[c] / L);
        rsq += d[c] * d[c];
      }
      if (rsq < rc2) {
        double r2inv = 1.0 / rsq, r6inv = r2inv * r2inv * r2inv;
        double forcelj = r6inv * (48.0 * r6inv - 24.0), fpair = forcelj * r2inv;
        if (f) for (int c = 0; c < 3; c++) { f[3 * i + c] += d[c] * fpair; f[3 * j + c] -= d[c] * fpair; }
        e += r6inv * (4.0 * r6inv - 4.0);
        vir += rsq * fpair;
        np_++;
      }
    }
  if (pe) *pe = e; if (w) *w = vir; if (npairs) *npairs = np_;
}

/* Second, independent implementation: Verlet list from a cell grid. Used by the MC engine
 * (and so by the timed CPU baseline). Positions must be wrapped into [0,L). */
typedef struct {
  int n, cap;            /* atoms, list capacity per atom */
  int *nnb, *nb;         /* half list: j > i stored once, nb[i*cap + k] */
  double *x0;            /* positions at build, in units of the build box */
  double L0, rl, skin;   /* build box, list radius */
  int64_t builds;
} vlist_t;

static void vlist_init(vlist_t* vl, int n, double skin) {
  memset(vl, 0, sizeof *vl);
  vl->n = n; vl->cap = 0; vl->skin = skin;
  vl->nnb = (int*)calloc(n, sizeof(int));
  vl->x0 = (double*)calloc(3 * (size_t)n, sizeof(double));
  vl->L0 = -1.0;
}
static void vlist_free(vlist_t* vl) { free(vl->nnb); free(vl->nb); free(vl->x0); }

static inline double mic(double d, double L, double hL) {
  if (d > hL) d -= L; else if (d < -hL) d += L;
  return d;
}

static void vlist_build(vlist_t* vl, const double* x, double L, double rc) {
  int n = vl->n; double rl = rc + vl->skin, rl2 = rl * rl, hL = 0.5 * L;
  int nc = (int)floor(L / rl); if (nc < 3) nc = 1;
  int ncell = nc * nc * nc;
  int* head = (int*)malloc(sizeof(int) * ncell); int* next = (int*)malloc(sizeof(int) * n);
  for (int c = 0; c < ncell; c++) head[c] = -1;
  int* cidx = (int*)malloc(sizeof(int) * n);
  for (int i = n - 1; i >= 0; i--) {
    int cx = (int)(x[3 * i] / L * nc), cy = (int)(x[3 * i + 1] / L * nc), cz = (int)(x[3 * i + 2] / L * nc);
    if (cx >= nc) cx = nc - 1; if (cy >= nc) cy = nc - 1; if (cz >= nc) cz = nc - 1;
    int c = (cx * nc + cy) * nc + cz; cidx[i] = c; next[i] = head[c]; head[c] = i;
  }
  for (int pass = 0; pass < 2; pass++) {       /* pass 0 counts, pass 1 fills */
    int maxn = 0;
    for (int i = 0; i < n; i++) {
      int cnt = 0;
      int cx = cidx[i] / (nc * nc), cy = (cidx[i] / nc) % nc, cz = cidx[i] % nc;
      int span = (nc == 1) ? 0 : 1;
      for (int ax = -span; ax <= span; ax++) for (int ay = -span; ay <= span; ay++) for (int az = -span; az <= span; az++) {
        int c = (((cx + ax + nc) % nc) * nc + (cy + ay + nc) % nc) * nc + (cz + az + nc) % nc;
        for (int j = head[c]; j >= 0; j = next[j]) {
          if (j <= i) continue;
          double dx = mic(x[3 * i] - x[3 * j], L, hL), dy = mic(x[3 * i + 1] - x[3 * j + 1], L, hL),
                 dz = mic(x[3 * i + 2] - x[3 * j + 2], L, hL);
          if (dx * dx + dy * dy + dz * dz < rl2) { if (pass) vl->nb[(size_t)i * vl->cap + cnt] = j; cnt++; }
        }
      }
      vl->nnb[i] = cnt; if (cnt > maxn) maxn = cnt;
    }
    if (!pass) {
      if (maxn > vl->cap) { vl->cap = maxn + 8; free(vl->nb); vl->nb = (int*)malloc(sizeof(int) * (size_t)n * vl->cap); }
    }
  }
  for (int i = 0; i < 3 * n; i++) vl->x0[i] = x[i] / L;   /* fractional, so box rescales are tracked */
  vl->L0 = L; vl->rl = rl; vl->builds++;
  free(head); free(next); free(cidx);
}

/* list is valid while s*(rl - 2*umax) >= rc with s = L/L0 and umax measured in build-box units */
static int vlist_valid(const vlist_t* vl, const double* x, double L, double rc) {
  if (vl->L0 < 0) return 0;
  double s = L / vl->L0, umax2 = 0.0;
  for (int i = 0; i < vl->n; i++) {
    double u2 = 0.0;
    for (int c = 0; c < 3; c++) {
      double d = x[3 * i + c] / L - vl->x0[3 * i + c];
      d -= nearbyint(d); d *= vl->L0; u2 += d * d;
    }
    if (u2 > umax2) umax2 = u2;
  }
  return s * (vl->rl - 2.0 * sqrt(umax2)) >= rc * (1.0 + 1e-12);
}

static void lj_eval_list(vlist_t* vl, const double* x, double L, double rc,
                         double* pe, double* w, double* f, int64_t* npairs) {
  int n = vl->n; double rc2 = rc * rc, hL = 0.5 * L, e = 0.0, vir = 0.0; int64_t np_ = 0;
  if (!vlist_valid(vl, x, L, rc)) vlist_build(vl, x, L, rc);
  memset(f, 0, sizeof(double) * 3 * n);
  for (int i = 0; i < n; i++) {
    const int* nb = vl->nb + (size_t)i * vl->cap;
    double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2], fx = 0, fy = 0, fz = 0;
    for (int k = 0; k < vl->nnb[i]; k++) {
      int j = nb[k];
      double dx = mic(xi - x[3 * j], L, hL), dy = mic(yi - x[3 * j + 1], L, hL), dz = mic(zi - x[3 * j + 2], L, hL);
      double rsq = dx * dx + dy * dy + dz * dz;
      if (rsq < rc2) {
        double r2inv = 1.0 / rsq, r6inv = r2inv * r2inv * r2inv;
        double fpair = r6inv * (48.0 * r6inv - 24.0) * r2inv;
        fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
        f[3 * j] -= dx * fpair; f[3 * j + 1] -= dy * fpair; f[3 * j + 2] -= dz * fpair;
        e += r6inv * (4.0 * r6inv - 4.0); vir += rsq * fpair; np_++;
      }
    }
    f[3 * i] += fx; f[3 * i + 1] += fy; f[3 * i + 2] += fz;
  }
  *pe = e; *w = vir; if (npairs) *npairs = np_;
}

ORC_API void orc_lj_eval_list(int n, const double* x, double L, double rc, double skin,
                              double* pe, double* w, double* f, int64_t* npairs) {
  vlist_t vl; vlist_init(&vl, n, skin);
  double* xx = (double*)malloc(sizeof(double) * 3 * n); memcpy(xx, x, sizeof(double) * 3 * n);
  orc_wrap(n, xx, L);
  double* ff = f ? f : (double*)malloc(sizeof(double) * 3 * n);
  double e, vir; lj_eval_list(&vl, xx, L, rc, &e, &vir, ff, npairs);
  if (pe) *pe = e; if (w) *w = vir;
  if (!f) free(ff); free(xx); vlist_free(&vl);
}

/* energy change of moving atom k from its current position to xn (O(N), minimum image) */
static double lj_delta_atom(int n, const double* x, int k, const double xn[3], double L, double rc,
                            int64_t* nvis) {
  double rc2 = rc * rc, hL = 0.5 * L, de = 0.0;
  for (int j = 0; j < n; j++) {
    if (j == k) continue;
    double ro = 0, rn = 0;
    for (int c = 0; c < 3; c++) {
      double d0 = mic(x[3 * k + c] - x[3 * j + c], L, hL), d1 = mic(xn[c] - x[3 * j + c], L, hL);
      ro += d0 * d0; rn += d1 * d1;
    }
    if (rn < rc2) { double r2 = 1.0 / rn, r6 = r2 * r2 * r2; de += r6 * (4.0 * r6 - 4.0); if (nvis) (*nvis)++; }
    if (ro < rc2) { double r2 = 1.0 / ro, r6 = r2 * r2 * r2; de -= r6 * (4.0 * r6 - 4.0); if (nvis) (*nvis)++; }
  }
  return de;
}
ORC_API double orc_lj_delta_atom(int n, const double* x, int k, const double* xn, double L, double rc) {
  return lj_delta_atom(n, x, k, xn, L, rc, NULL);
}

/* ------------------------------------------------------------------ MC engine */
typedef struct {
  int32_t nstps, mod, bulk_move, text_rounding;
  double ppos, pvol, lat_scale, mass, rc, skin;
  uint64_t seed;
} orc_params;

/* per-replica statistics, same meaning as the engine's counters (include/nm_b200.h) */
enum { CT_SWEEPS = 0, CT_HMC_MOVES, CT_HMC_ATOM_STEPS, CT_VMC_MOVES, CT_PMC_MOVES, CT_PMC_TRIALS,
       CT_FORCE_EVALS, CT_PAIRS_FORCE, CT_PAIRS_FULL, CT_PAIRS_DELTA, CT_LIST_BUILDS, CT_LIST_PAIRS, CT_N };

typedef struct {
  const orc_params* p;
  int n; double *x, *v, *f, *xs, *vs, *fs;   /* state + saved copies */
  double box, pe, w;
  vlist_t vl;
  uint64_t ct[CT_N];
  int err;
} sim_t;

static double r6(const sim_t* s, double x) { return s->p->text_rounding ? orc_round6(x) : x; }

/* 'run 0' (LAMMPS init+setup: remap atoms, neighbour build, full evaluation) */
static void run0(sim_t* s) {
  orc_wrap(s->n, s->x, s->box);
  int64_t np_; uint64_t b0 = s->vl.builds;
  lj_eval_list(&s->vl, s->x, s->box, s->p->rc, &s->pe, &s->w, s->f, &np_);
  s->ct[CT_FORCE_EVALS]++; s->ct[CT_PAIRS_FULL] += (uint64_t)np_; s->ct[CT_LIST_BUILDS] += s->vl.builds - b0;
  if (s->box < 2.0 * s->p->rc) s->err = -5;
}
static double kinetic(const sim_t* s) {       /* compute ke: 0.5 * sum m v^2 */
  double a = 0.0; for (int i = 0; i < 3 * s->n; i++) a += s->v[i] * s->v[i];
  return 0.5 * s->p->mass * a;
}
/* the acceptance rule shared by all moves (lammps_remcmc.py:487-500 etc.):
 * metcrit = exp(-de); isinf -> reject WITHOUT drawing; else accept iff U <= min(1, metcrit);
 * NaN compares false -> reject. */
static int metropolis(double de, const rng_t* r, uint32_t index, uint32_t purpose) {
  double m = exp(-de);
  if (isinf(m) || isnan(m)) return 0;
  double u = rng_uniform(r, index, purpose);
  return u <= (m < 1.0 ? m : 1.0);
}

/* a-7 bulk_position_mc, lammps_remcmc.py:477-502 */
static void bulk_position_mc(sim_t* s, double et, double* ntp, double* nap, double dx, const rng_t* r) {
  int n = s->n;
  *ntp += 1;
  memcpy(s->xs, s->x, sizeof(double) * 3 * n); memcpy(s->fs, s->f, sizeof(double) * 3 * n);
  double pe_old = s->pe, w_old = s->w, pe = s->pe / et;
  double d = r6(s, dx * s->p->lat_scale);                       /* displace_atoms all random %f */
  for (int i = 0; i < n; i++) {
    double u[3]; rng_uniform3(r, (uint32_t)i, P_BULK_DISP, u);
    for (int c = 0; c < 3; c++) s->x[3 * i + c] += d * 2.0 * (u[c] - 0.5);
  }
  run0(s);
  double de = s->pe / et - pe;
  s->ct[CT_PMC_MOVES]++; s->ct[CT_PMC_TRIALS]++;
  if (metropolis(de, r, 0, P_BULK_ACC)) { *nap += 1; }
  else {                                                        /* scatter old x; run 0 */
    memcpy(s->x, s->xs, sizeof(double) * 3 * n); memcpy(s->f, s->fs, sizeof(double) * 3 * n);
    s->pe = pe_old; s->w = w_old;
  }
}

/* Per-sweep candidate lists for the single-atom energy change (so that the timed CPU baseline of the iterative sweep is
 * not an O(N^2) strawman): every atom moves at most once per sweep, by at most reach = sqrt(3) * dmax, so every atom
 * within rc of the old or the new position of atom k at k's turn lies within R = rc + 2 reach of k at the START of the
 * sweep. Rows hold those atoms in ASCENDING index order: the sum visits the same non-zero terms in the same order as
 * the all-atom loop of lj_delta_atom, i.e. the result is bitwise identical (tests/test_oracle_lj.py). */
static int g_delta_lists = 1;
ORC_API void orc_set_delta_lists(int on) { g_delta_lists = on; }
static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }
typedef struct { int* start; int* idx; } dlist_t;
static int dlist_build(dlist_t* dl, int n, const double* x, double L, double R) {
  int nc = (int)floor(L / R);
  if (nc < 3) return 0;
  double hL = 0.5 * L, R2 = R * R;
  int ncell = nc * nc * nc;
  int* head = (int*)malloc(sizeof(int) * ncell); int* next = (int*)malloc(sizeof(int) * n); int* cidx = (int*)malloc(sizeof(int) * n);
  for (int c = 0; c < ncell; c++) head[c] = -1;
  for (int i = n - 1; i >= 0; i--) {
    int cx = (int)(x[3 * i] / L * nc), cy = (int)(x[3 * i + 1] / L * nc), cz = (int)(x[3 * i + 2] / L * nc);
    if (cx >= nc) cx = nc - 1; if (cy >= nc) cy = nc - 1; if (cz >= nc) cz = nc - 1;
    if (cx < 0) cx = 0; if (cy < 0) cy = 0; if (cz < 0) cz = 0;
    int c = (cx * nc + cy) * nc + cz; cidx[i] = c; next[i] = head[c]; head[c] = i;
  }
  dl->start = (int*)malloc(sizeof(int) * (n + 1));
  dl->idx = NULL;
  for (int pass = 0; pass < 2; pass++) {
    int tot = 0;
    for (int i = 0; i < n; i++) {
      int cnt = 0, cx = cidx[i] / (nc * nc), cy = (cidx[i] / nc) % nc, cz = cidx[i] % nc;
      if (pass) tot = dl->start[i];
      for (int ax = -1; ax <= 1; ax++) for (int ay = -1; ay <= 1; ay++) for (int az = -1; az <= 1; az++) {
        int c = (((cx + ax + nc) % nc) * nc + (cy + ay + nc) % nc) * nc + (cz + az + nc) % nc;
        for (int j = head[c]; j >= 0; j = next[j]) {
          if (j == i) continue;
          double dx = mic(x[3 * i] - x[3 * j], L, hL), dy = mic(x[3 * i + 1] - x[3 * j + 1], L, hL), dz = mic(x[3 * i + 2] - x[3 * j + 2], L, hL);
          if (dx * dx + dy * dy + dz * dz < R2) { if (pass) dl->idx[tot + cnt] = j; cnt++; }
        }
      }
      if (!pass) { dl->start[i] = tot; tot += cnt; if (i == n - 1) dl->start[n] = tot; }
      else qsort(dl->idx + dl->start[i], cnt, sizeof(int), cmp_int);
    }
    if (!pass) dl->idx = (int*)malloc(sizeof(int) * (dl->start[n] > 0 ? dl->start[n] : 1));
  }
  free(head); free(next); free(cidx);
  return 1;
}
static double lj_delta_atom_rows(const dlist_t* dl, const double* x, int k, const double xn[3], double L, double rc, int64_t* nvis) {
  double rc2 = rc * rc, hL = 0.5 * L, de = 0.0;
  for (int q = dl->start[k]; q < dl->start[k + 1]; q++) {
    int j = dl->idx[q];
    double ro = 0, rn = 0;
    for (int c = 0; c < 3; c++) {
      double d0 = mic(x[3 * k + c] - x[3 * j + c], L, hL), d1 = mic(xn[c] - x[3 * j + c], L, hL);
      ro += d0 * d0; rn += d1 * d1;
    }
    if (rn < rc2) { double r2 = 1.0 / rn, r6 = r2 * r2 * r2; de += r6 * (4.0 * r6 - 4.0); if (nvis) (*nvis)++; }
    if (ro < rc2) { double r2 = 1.0 / ro, r6 = r2 * r2 * r2; de -= r6 * (4.0 * r6 - 4.0); if (nvis) (*nvis)++; }
  }
  return de;
}

/* a-8 iter_position_mc, lammps_remcmc.py:505-549. The reference re-evaluates the whole system
 * per trial; E_tot' - E_tot is the single-atom energy change, computed directly here. */
static void iter_position_mc(sim_t* s, double et, double* ntp, double* nap, double dx, const rng_t* r) {
  int n = s->n; double box = s->box;
  dlist_t dl; int have = 0;
  if (g_delta_lists && n >= 256) {
    double dmax = fabs(dx * s->p->lat_scale), R = s->p->rc + 2.0 * sqrt(3.0) * dmax * (1.0 + 1e-9) + 1e-9;
    have = dlist_build(&dl, n, s->x, box, R);
  }
  for (int k = 0; k < n; k++) {
    *ntp += 1;
    double u[3], nd[3]; rng_uniform3(r, (uint32_t)k, P_ITER_DISP, u);
    for (int c = 0; c < 3; c++) {
      nd[c] = s->x[3 * k + c] + 2 * (u[c] - 0.5) * dx * s->p->lat_scale;
      nd[c] -= floor(nd[c] / box) * box;
      nd[c] = wrap1(nd[c], box);                                /* the following 'run 0' remap */
    }
    int64_t nvis = 0;
    double de = (have ? lj_delta_atom_rows(&dl, s->x, k, nd, box, s->p->rc, &nvis) : lj_delta_atom(n, s->x, k, nd, box, s->p->rc, &nvis)) / et;
    s->ct[CT_PAIRS_DELTA] += (uint64_t)nvis; s->ct[CT_PMC_TRIALS]++;
    if (metropolis(de, r, (uint32_t)k, P_ITER_ACC)) {
      *nap += 1; for (int c = 0; c < 3; c++) s->x[3 * k + c] = nd[c];
    }
  }
  if (have) { free(dl.start); free(dl.idx); }
  s->ct[CT_PMC_MOVES]++;
  run0(s);                                                      /* state the last 'run 0' of the sweep leaves */
}

/* a-6 volume_mc, lammps_remcmc.py:552-595 */
static void volume_mc(sim_t* s, double et, double pf, double* ntv, double* nav, double dv, const rng_t* r) {
  int n = s->n;
  *ntv += 1;
  double box = s->box, vol = pow(box, 3.0);
  memcpy(s->xs, s->x, sizeof(double) * 3 * n); memcpy(s->fs, s->f, sizeof(double) * 3 * n);
  double pe_old = s->pe, w_old = s->w, pe = s->pe / et;
  double volnew = exp(log(vol) + 2 * (rng_uniform(r, 0, P_VMC_PROP) - 0.5) * dv);
  double boxnew = cbrt(volnew), scale = boxnew / box;
  for (int i = 0; i < 3 * n; i++) s->x[i] = scale * s->x[i];
  s->box = r6(s, boxnew);                                       /* change_box ... %f */
  run0(s);
  double penew = s->pe / et;
  double dh = (penew - pe) + pf * (volnew - vol) - (n + 1) * log(volnew / vol);
  s->ct[CT_VMC_MOVES]++;
  if (metropolis(dh, r, 0, P_VMC_ACC)) { *nav += 1; }
  else {
    s->box = r6(s, box);
    memcpy(s->x, s->xs, sizeof(double) * 3 * n); memcpy(s->f, s->fs, sizeof(double) * 3 * n);
    s->pe = pe_old; s->w = w_old;
  }
}

/* a-4: 'velocity all create T seed dist gaussian' (loop all, mom yes, rot no) followed by
 * 'velocity all zero linear' and 'velocity all zero angular' (lammps_remcmc.py:604-606).
 * LAMMPS velocity.cpp::create + group.cpp (vcm / xcm / angmom / inertia / omega) semantics;
 * wrapped coordinates are used for the angular part. */
static void velocity_create(sim_t* s, double t_target, const rng_t* r) {
  int n = s->n; double m = s->p->mass, *v = s->v, *x = s->x;
  double inv = 1.0 / sqrt(m);
  for (int i = 0; i < n; i++) { double g[3]; rng_gauss3(r, (uint32_t)i, P_HMC_VEL, g);
    for (int c = 0; c < 3; c++) v[3 * i + c] = g[c] * inv; }
  for (int rep = 0; rep < 2; rep++) {            /* mom yes ... scale ... then 'zero linear' */
    double vcm[3] = {0, 0, 0};
    for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) vcm[c] += m * v[3 * i + c];
    for (int c = 0; c < 3; c++) vcm[c] /= (m * n);
    for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) v[3 * i + c] -= vcm[c];
    if (rep == 0) {                              /* scale to exactly T with dof = 3N-3 */
      double t = 0.0; for (int i = 0; i < 3 * n; i++) t += v[i] * v[i];
      t = m * t / (3.0 * n - 3.0);
      double fac = sqrt(t_target / t);
      for (int i = 0; i < 3 * n; i++) v[i] *= fac;
    }
  }
  /* zero angular: omega = I^-1 L about the centre of mass; v -= omega x (x - xcm) */
  double xcm[3] = {0, 0, 0}, L[3] = {0, 0, 0}, I[3][3] = {{0}};
  for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) xcm[c] += m * x[3 * i + c];
  for (int c = 0; c < 3; c++) xcm[c] /= (m * n);
  for (int i = 0; i < n; i++) {
    double dx = x[3 * i] - xcm[0], dy = x[3 * i + 1] - xcm[1], dz = x[3 * i + 2] - xcm[2];
    double vx = v[3 * i], vy = v[3 * i + 1], vz = v[3 * i + 2];
    L[0] += m * (dy * vz - dz * vy); L[1] += m * (dz * vx - dx * vz); L[2] += m * (dx * vy - dy * vx);
    I[0][0] += m * (dy * dy + dz * dz); I[1][1] += m * (dx * dx + dz * dz); I[2][2] += m * (dx * dx + dy * dy);
    I[0][1] -= m * dx * dy; I[1][2] -= m * dy * dz; I[0][2] -= m * dx * dz;
  }
  I[1][0] = I[0][1]; I[2][1] = I[1][2]; I[2][0] = I[0][2];
  double det = I[0][0] * (I[1][1] * I[2][2] - I[1][2] * I[2][1]) - I[0][1] * (I[1][0] * I[2][2] - I[1][2] * I[2][0])
             + I[0][2] * (I[1][0] * I[2][1] - I[1][1] * I[2][0]);
  double w[3] = {0, 0, 0};
  if (det > 0.0) {
    double inv_[3][3];
    inv_[0][0] =  (I[1][1] * I[2][2] - I[1][2] * I[2][1]) / det; inv_[0][1] = -(I[0][1] * I[2][2] - I[0][2] * I[2][1]) / det;
    inv_[0][2] =  (I[0][1] * I[1][2] - I[0][2] * I[1][1]) / det; inv_[1][0] = -(I[1][0] * I[2][2] - I[1][2] * I[2][0]) / det;
    inv_[1][1] =  (I[0][0] * I[2][2] - I[0][2] * I[2][0]) / det; inv_[1][2] = -(I[0][0] * I[1][2] - I[0][2] * I[1][0]) / det;
    inv_[2][0] =  (I[1][0] * I[2][1] - I[1][1] * I[2][0]) / det; inv_[2][1] = -(I[0][0] * I[2][1] - I[0][1] * I[2][0]) / det;
    inv_[2][2] =  (I[0][0] * I[1][1] - I[0][1] * I[1][0]) / det;
    for (int a = 0; a < 3; a++) w[a] = inv_[a][0] * L[0] + inv_[a][1] * L[1] + inv_[a][2] * L[2];
  }
  for (int i = 0; i < n; i++) {
    double dx = x[3 * i] - xcm[0], dy = x[3 * i + 1] - xcm[1], dz = x[3 * i + 2] - xcm[2];
    v[3 * i] -= w[1] * dz - w[2] * dy; v[3 * i + 1] -= w[2] * dx - w[0] * dz; v[3 * i + 2] -= w[0] * dy - w[1] * dx;
  }
}

/* a-5: 'run NSTPS' under fix nve: velocity Verlet (fix_nve.cpp initial/final_integrate) */
static void run_nve(sim_t* s, int nsteps, double dt) {
  int n = s->n; double dtf = 0.5 * dt / s->p->mass;
  for (int st = 0; st < nsteps; st++) {
    for (int i = 0; i < 3 * n; i++) { s->v[i] += dtf * s->f[i]; s->x[i] += dt * s->v[i]; }
    orc_wrap(n, s->x, s->box);                  /* image relabelling only (see header) */
    int64_t np_; uint64_t b0 = s->vl.builds;
    lj_eval_list(&s->vl, s->x, s->box, s->p->rc, &s->pe, &s->w, s->f, &np_);
    s->ct[CT_FORCE_EVALS]++; s->ct[CT_LIST_BUILDS] += s->vl.builds - b0;
    if (st == nsteps - 1) s->ct[CT_PAIRS_FULL] += (uint64_t)np_; else s->ct[CT_PAIRS_FORCE] += (uint64_t)np_;
    for (int i = 0; i < 3 * n; i++) s->v[i] += dtf * s->f[i];
  }
}

/* a-3 hamiltonian_mc, lammps_remcmc.py:598-640 */
static void hamiltonian_mc(sim_t* s, double et, double t_vel, double* nth, double* nah, double dt, const rng_t* r) {
  int n = s->n;
  *nth += 1;
  velocity_create(s, t_vel, r);
  double dtu = r6(s, dt);                                       /* timestep %f */
  memcpy(s->xs, s->x, sizeof(double) * 3 * n); memcpy(s->vs, s->v, sizeof(double) * 3 * n);
  memcpy(s->fs, s->f, sizeof(double) * 3 * n);
  double pe_old = s->pe, w_old = s->w;
  double etot = s->pe / et + kinetic(s) / et;
  run_nve(s, s->p->nstps, dtu);
  double etotnew = s->pe / et + kinetic(s) / et;
  double de = etotnew - etot;
  s->ct[CT_HMC_MOVES]++; s->ct[CT_HMC_ATOM_STEPS] += (uint64_t)n * (uint64_t)s->p->nstps;
  if (metropolis(de, r, 0, P_HMC_ACC)) { *nah += 1; }
  else {
    memcpy(s->x, s->xs, sizeof(double) * 3 * n); memcpy(s->v, s->vs, sizeof(double) * 3 * n);
    memcpy(s->f, s->fs, sizeof(double) * 3 * n); s->pe = pe_old; s->w = w_old;
  }
}

/* gen_sample, lammps_remcmc.py:665-691: MOD x move_mc (:643-658), then lammps_extract (:377-391)
 * x, v: [3n] in/out; scal = {box, dx, dv, dt} (box in/out); counts = {ntp,nap,ntv,nav,nth,nah} in/out;
 * label = {et, pf, temp, temp_vel}; thermo[18] out (include/nm_b200.h order); ct[CT_N] accumulates.
 * energies_trace (optional, [mod][3]) records (pe, ke, box) after every move. */
ORC_API int orc_cycle(const orc_params* p, const double* label, int32_t slot_global, int64_t cycle,
                      int32_t n, double* x, double* v, double* scal, double* counts,
                      double* thermo, uint64_t* ct, double* energies_trace) {
  sim_t s; memset(&s, 0, sizeof s);
  s.p = p; s.n = n; s.x = x; s.v = v; s.box = r6(&s, scal[0]);   /* init_lammps: change_box %f */
  size_t sz = sizeof(double) * 3 * (size_t)n;
  s.f = (double*)malloc(sz); s.xs = (double*)malloc(sz); s.vs = (double*)malloc(sz); s.fs = (double*)malloc(sz);
  vlist_init(&s.vl, n, p->skin > 0 ? p->skin : 0.3);
  double et = label[0], pf = label[1], t_vel = label[3];
  double dx = scal[1], dv = scal[2], dt = scal[3];
  run0(&s);                                                     /* init_lammps: 'run 0' */
  for (int mv = 0; mv < p->mod; mv++) {
    rng_t r = rng_make(p->seed, (uint32_t)slot_global, (uint64_t)cycle * (uint64_t)p->mod + (uint64_t)mv);
    double roll = rng_uniform(&r, 0, P_ROLL);
    if (roll <= p->ppos) {
      if (p->bulk_move) bulk_position_mc(&s, et, &counts[0], &counts[1], dx, &r);
      else iter_position_mc(&s, et, &counts[0], &counts[1], dx, &r);
    } else if (roll <= (p->ppos + p->pvol)) volume_mc(&s, et, pf, &counts[2], &counts[3], dv, &r);
    else hamiltonian_mc(&s, et, t_vel, &counts[4], &counts[5], dt, &r);
    s.ct[CT_SWEEPS]++;
    if (energies_trace) { energies_trace[3 * mv] = s.pe; energies_trace[3 * mv + 1] = kinetic(&s); energies_trace[3 * mv + 2] = s.box; }
  }
  /* lammps_extract */
  double ke = kinetic(&s), dof = 3.0 * n - 3.0, temp = 2.0 * ke / dof, vol = pow(s.box, 3.0);
  scal[0] = s.box;
  if (thermo) {
    thermo[0] = temp; thermo[1] = s.pe; thermo[2] = ke;
    thermo[3] = (dof * temp + s.w) / 3.0 * (1.0 / vol);         /* compute pressure (lj: nktv2p = 1) */
    thermo[4] = s.box; thermo[5] = vol; thermo[6] = dx; thermo[7] = dv; thermo[8] = dt;
    for (int c = 0; c < 6; c++) thermo[9 + c] = counts[c];
    for (int c = 0; c < 3; c++) {                               /* float32 ratios, nan_to_num (0/0 -> 0) */
      float a = (float)counts[2 * c + 1] / (float)counts[2 * c];
      thermo[15 + c] = isnan(a) ? 0.0 : (double)a;
    }
  }
  if (ct) for (int c = 0; c < CT_N; c++) ct[c] += s.ct[c];
  int err = s.err;
  free(s.f); free(s.xs); free(s.vs); free(s.fs); vlist_free(&s.vl);
  return err;
}

/* a-10 gen_mc_param, lammps_remcmc.py:726-745. step = {dx,dv,dt} in/out, ratio = {ap,av,ah} (float32 values) */
ORC_API void orc_adapt(double* step, const double* ratio) {
  for (int c = 0; c < 3; c++) {
    if (ratio[c] < 0.5) step[c] = 0.9375 * step[c];
    if (ratio[c] > 0.5) step[c] = 1.0625 * step[c];
  }
}

/* a-11 replica_exchange, lammps_remcmc.py:776-803. etot = pe+ke, vol, et, pf: [np*nt] by slot.
 * uniforms: np*nt*(nt-1)/2 draws in loop order. perm[k] = original slot of the configuration
 * that ends in slot k. Returns the number of swaps. */
ORC_API int64_t orc_exchange(int32_t np_, int32_t nt, const double* etot_in, const double* vol_in,
                             const double* et, const double* pf, const double* uniforms, int32_t* perm) {
  int ns = np_ * nt; int64_t swaps = 0, draw = 0;
  double* e = (double*)malloc(sizeof(double) * ns); double* vv = (double*)malloc(sizeof(double) * ns);
  memcpy(e, etot_in, sizeof(double) * ns); memcpy(vv, vol_in, sizeof(double) * ns);
  for (int k = 0; k < ns; k++) perm[k] = k;
  for (int u = 0; u < np_; u++)
    for (int v = nt - 1; v >= 0; v--)
      for (int w = 0; w < v; w++) {
        int i = u * nt + v, j = u * nt + w;
        double de = e[i] - e[j], dvol = vv[i] - vv[j];
        double dh = de * (1. / et[i] - 1. / et[j]) + (pf[i] - pf[j]) * dvol;
        double m = exp(dh), crit = (isnan(m) ? m : (m < 1.0 ? m : 1.0));   /* np.min([1, m]) propagates NaN */
        if (uniforms[draw++] <= crit) {
          swaps++;
          double t = e[i]; e[i] = e[j]; e[j] = t; t = vv[i]; vv[i] = vv[j]; vv[j] = t;
          int q = perm[i]; perm[i] = perm[j]; perm[j] = q;
        }
      }
  free(e); free(vv);
  return swaps;
}
/* the engine's own uniform stream for the exchange of a given cycle */
ORC_API void orc_exchange_uniforms(uint64_t seed, int64_t cycle, int64_t n, double* out) {
  for (int64_t i = 0; i < n; i++) {
    uint32_t w[4];
    philox((uint32_t)seed, (uint32_t)(seed >> 32) ^ EXCH_KEY, (uint32_t)i, P_EXCH, (uint32_t)cycle, (uint32_t)((uint64_t)cycle >> 32), w);
    out[i] = u53(w[0], w[1]);
  }
}

/* ------------------------------------------------------------------ a-14: RDF */
/* calculate_rdf, lammps_distr.py:123-135, before the final '/natoms': for each of the 27
 * image vectors (br order of :100-103) every ordered pair's float32 distance, np.histogram'ed
 * on the float64 edges r. counts[0] = 0, counts[1+b] = bin b. Compile with -ffp-contract=off. */
ORC_API void orc_rdf_counts(int32_t n, const float* pos, float box, const double* r, int32_t nbins,
                            uint32_t* counts) {
  memset(counts, 0, sizeof(uint32_t) * nbins);
  static const int b[3] = {-1, 0, 1};
  for (int i0 = 0; i0 < 3; i0++) for (int i1 = 0; i1 < 3; i1++) for (int i2 = 0; i2 < 3; i2++) {
    float s[3] = { box * (float)b[i0], box * (float)b[i1], box * (float)b[i2] };
    for (int q = 0; q < n; q++) {                /* shifted atom (first axis of dvm) */
      float img[3] = { pos[3 * q] + s[0], pos[3 * q + 1] + s[1], pos[3 * q + 2] + s[2] };
      for (int a = 0; a < n; a++) {
        float dx = pos[3 * a] - img[0], dy = pos[3 * a + 1] - img[1], dz = pos[3 * a + 2] - img[2];
        float sx = dx * dx, sy = dy * dy, sz = dz * dz;
        float d = sqrtf((sx + sy) + sz);
        double dd = (double)d;
        if (!(dd >= r[0] && dd <= r[nbins - 1])) continue;
        int lo = 0, hi = nbins - 1;                 /* largest k with r[k] <= d */
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (dd < r[mid]) hi = mid - 1; else lo = mid; }
        if (lo == nbins - 1) lo = nbins - 2;        /* last bin right-closed */
        counts[1 + lo]++;
      }
    }
  }
}

/* N1: calculate_cdf, lammps_distr.py:161-171, before the final '/natoms': for each of the 27 image vectors
 * np.histogramdd of the float32 pair vectors on the float64 edges rv[3][nb+1] (searchsorted side='right', the
 * right-most edge closed, outliers dropped). counts[nb][nb][nb]. */
static int cdf_bin(float x, const double* e, int nb) {
  const double v = (double)x;
  if (!(v >= e[0] && v <= e[nb])) return -1;
  int lo = 0, hi = nb;                        /* largest k with e[k] <= v */
  while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (v < e[mid]) hi = mid - 1; else lo = mid; }
  return lo == nb ? nb - 1 : lo;
}
ORC_API void orc_cdf_counts(int32_t n, const float* pos, float box, const double* rv, int32_t nb, uint32_t* counts) {
  memset(counts, 0, sizeof(uint32_t) * (size_t)nb * nb * nb);
  static const int b[3] = {-1, 0, 1};
  for (int i0 = 0; i0 < 3; i0++) for (int i1 = 0; i1 < 3; i1++) for (int i2 = 0; i2 < 3; i2++) {
    float s[3] = { box * (float)b[i0], box * (float)b[i1], box * (float)b[i2] };
    for (int q = 0; q < n; q++) {
      float img[3] = { pos[3 * q] + s[0], pos[3 * q + 1] + s[1], pos[3 * q + 2] + s[2] };
      for (int a = 0; a < n; a++) {
        int bx = cdf_bin(pos[3 * a] - img[0], rv, nb); if (bx < 0) continue;
        int by = cdf_bin(pos[3 * a + 1] - img[1], rv + (nb + 1), nb); if (by < 0) continue;
        int bz = cdf_bin(pos[3 * a + 2] - img[2], rv + 2 * (nb + 1), nb); if (bz < 0) continue;
        counts[((size_t)bx * nb + by) * nb + bz]++;
      }
    }
  }
}

/* ------------------------------------------------------------------ replica farm (CPU baseline) */
typedef struct {
  const orc_params* p; const double* labels; int32_t slot0, nrep, n; int64_t cycle0, ncycles;
  double *x, *v, *scal, *counts, *thermo; uint64_t* ct; int next; pthread_mutex_t mu; int err;
} farm_t;
static void* farm_worker(void* arg) {
  farm_t* fm = (farm_t*)arg;
  uint64_t ct[CT_N]; memset(ct, 0, sizeof ct);
  for (;;) {
    pthread_mutex_lock(&fm->mu); int k = fm->next++; pthread_mutex_unlock(&fm->mu);
    if (k >= fm->nrep) break;
    for (int64_t c = 0; c < fm->ncycles; c++) {
      int e = orc_cycle(fm->p, fm->labels + 4 * k, fm->slot0 + k, fm->cycle0 + c, fm->n,
                        fm->x + 3 * (size_t)fm->n * k, fm->v + 3 * (size_t)fm->n * k, fm->scal + 4 * k,
                        fm->counts + 6 * k, fm->thermo ? fm->thermo + 18 * k : NULL, ct, NULL);
      if (e) fm->err = e;
      /* between cycles: gen_mc_param */
      if (c + 1 < fm->ncycles && fm->thermo) {
        orc_adapt(fm->scal + 4 * k + 1, fm->thermo + 18 * k + 15);
        for (int q = 0; q < 6; q++) fm->counts[6 * k + q] = 0.0;
      }
    }
  }
  pthread_mutex_lock(&fm->mu); for (int c = 0; c < CT_N; c++) fm->ct[c] += ct[c]; pthread_mutex_unlock(&fm->mu);
  return NULL;
}
/* one task per replica over nthreads workers -- the decomposition Dask uses (lammps_remcmc.py:698-700) */
ORC_API int orc_farm(const orc_params* p, const double* labels, int32_t slot0, int32_t nrep, int32_t n,
                     int64_t cycle0, int64_t ncycles, double* x, double* v, double* scal, double* counts,
                     double* thermo, uint64_t* ct, int32_t nthreads) {
  farm_t fm; memset(&fm, 0, sizeof fm);
  fm.p = p; fm.labels = labels; fm.slot0 = slot0; fm.nrep = nrep; fm.n = n; fm.cycle0 = cycle0; fm.ncycles = ncycles;
  fm.x = x; fm.v = v; fm.scal = scal; fm.counts = counts; fm.thermo = thermo; fm.ct = ct;
  pthread_mutex_init(&fm.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, farm_worker, &fm);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); pthread_mutex_destroy(&fm.mu);
  return fm.err;
}

typedef struct { int32_t n, nbins; const float* pos; const float* box; const double* r; uint32_t* counts;
                 int64_t ns; int64_t next; pthread_mutex_t mu; } rdf_farm_t;
static void* rdf_worker(void* arg) {
  rdf_farm_t* fm = (rdf_farm_t*)arg;
  for (;;) {
    pthread_mutex_lock(&fm->mu); int64_t s = fm->next++; pthread_mutex_unlock(&fm->mu);
    if (s >= fm->ns) break;
    orc_rdf_counts(fm->n, fm->pos + 3 * (size_t)fm->n * s, fm->box[s], fm->r, fm->nbins, fm->counts + (size_t)fm->nbins * s);
  }
  return NULL;
}
ORC_API void orc_rdf_farm(int32_t n, int64_t ns, const float* pos, const float* box, const double* r,
                          int32_t nbins, uint32_t* counts, int32_t nthreads) {
  rdf_farm_t fm = { n, nbins, pos, box, r, counts, ns, 0 };
  pthread_mutex_init(&fm.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, rdf_worker, &fm);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); pthread_mutex_destroy(&fm.mu);
}
