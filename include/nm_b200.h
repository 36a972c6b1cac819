/*
 * nm_b200.h -- C-ABI of the B200-native replica-exchange NPT Monte Carlo engine.
 *
 * This is the drop-in boundary for the hot path of walkernr/neuralMelting:
 * everything the reference's Python driver asks of its per-replica LAMMPS
 * instance (scripts/lammps_remcmc.py) and of its numba RDF routine
 * (scripts/lammps_distr.py) is reached through the entry points below.
 * Plain pointers and sizes only; no torch / numpy / C++ types cross the boundary.
 *
 * Conventions
 *   - every function returns 0 on success, a negative NM_E* code on failure;
 *     nm_last_error() returns a thread-local message for the last failure.
 *   - "slot" k = i*NT + j is the reference's replica index (pressure i,
 *     temperature j, C order; lammps_remcmc.py:117,300). Host-visible arrays
 *     are always in LOCAL SLOT order: local pressure row lr holds the NT slots of global
 *     row rep_offset/NT + lr*row_stride (row_stride 1: a contiguous block of rows,
 *     slot_local = k - rep_offset; row_stride G with rep_offset = rank*NT: row u -> rank u mod G).
 *   - positions/velocities: double[n_rep][3*natoms], atom-id major (x0 y0 z0 x1 ..),
 *     the layout of lammps.gather_atoms('x',1,3) (lammps_remcmc.py:381-382).
 *   - host pointers are caller-owned (NumPy arrays); device memory is owned by
 *     the engine and released by nm_destroy. One engine drives one GPU; the
 *     multi-GPU job is one process (rank) per GPU, each with its own engine
 *     holding whole pressure rows.
 *   - the library never falls back to the CPU: without a usable CUDA device
 *     nm_create / nm_rdf_counts fail with NM_ENODEV.
 */
#ifndef NM_B200_H
#define NM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NM_ABI_VERSION 2

/* error codes */
#define NM_OK         0
#define NM_EINVAL    -1   /* bad argument / configuration */
#define NM_ENODEV    -2   /* no CUDA device / driver */
#define NM_ECUDA     -3   /* CUDA runtime error (see nm_last_error) */
#define NM_ENOMEM    -4
#define NM_EBOX      -5   /* a box side fell below 2*rc (minimum image invalid) */
#define NM_ENEIGH    -6   /* neighbour-list capacity exceeded */
#define NM_ESTATE    -7   /* call sequence error (e.g. no state uploaded) */

/* per-slot thermo record written by nm_get_thermo: state slots [3..20] of the
 * reference's 21-slot list (lammps_remcmc.py:432-433,690-691) */
#define NM_THERMO_WIDTH 18
enum nm_thermo_col {
  NM_TH_TEMP = 0, NM_TH_PE, NM_TH_KE, NM_TH_VIRIAL, NM_TH_BOX, NM_TH_VOL,
  NM_TH_DX, NM_TH_DV, NM_TH_DT,
  NM_TH_NTP, NM_TH_NAP, NM_TH_NTV, NM_TH_NAV, NM_TH_NTH, NM_TH_NAH,
  NM_TH_AP, NM_TH_AV, NM_TH_AH
};

/* counters returned by nm_get_counters (uint64 each, summed over local replicas
 * since nm_create or the last nm_reset_counters) */
#define NM_COUNTER_WIDTH 22
enum nm_counter_col {
  NM_CT_SWEEPS = 0,        /* move_mc calls (lammps_remcmc.py:677-679)            */
  NM_CT_HMC_MOVES,         /* hamiltonian_mc calls                                */
  NM_CT_HMC_ATOM_STEPS,    /* natoms * NSTPS per HMC move  (the headline metric)  */
  NM_CT_VMC_MOVES,
  NM_CT_PMC_MOVES,         /* bulk moves or iterative sweeps                      */
  NM_CT_PMC_TRIALS,        /* single-atom trials (iterative) or bulk trials       */
  NM_CT_FORCE_EVALS,       /* full-system evaluations                             */
  NM_CT_PAIRS_FORCE,       /* in-cutoff unordered pairs, force-only evaluations   */
  NM_CT_PAIRS_FULL,        /* in-cutoff unordered pairs, force+energy+virial      */
  NM_CT_PAIRS_DELTA,       /* in-cutoff neighbours visited by single-atom dE      */
  NM_CT_LIST_BUILDS,       /* Verlet-list rebuilds                                */
  NM_CT_LIST_PAIRS,        /* listed (candidate) unordered pairs touched          */
  NM_CT_CLK_EVAL,          /* SM clocks spent in force evaluations (summed over CTAs) */
  NM_CT_CLK_BUILD,         /* SM clocks spent in list builds                      */
  NM_CT_CLK_TOTAL,         /* SM clocks of the cycle kernels                      */
  NM_CT_OUTER_BUILDS,      /* outer (cell-search) list builds                     */
  NM_CT_CLK_OUTER,         /* SM clocks in outer builds                           */
  NM_CT_CLK_INNER,         /* SM clocks in inner builds                           */
  NM_CT_CLK_VEL,           /* SM clocks in the HMC velocity draw                  */
  NM_CT_HELPED_EVALS,     /* force evaluations whose upper rows a helper CTA computed (LARGE mode)     */
  NM_CT_DBG_LOOPCLK,       /* diagnostics: clocks thread 0 spent in its own pair loop (first atom)  */
  NM_CT_DBG_LOOPIT         /* diagnostics: quad iterations of that loop                             */
};

typedef struct nm_engine nm_engine;   /* opaque */

/* Engine configuration. Mirrors the globals lammps_remcmc.py derives from its
 * CLI (lammps_remcmc.py:836-899). Zero-initialise, set struct_size, fill. */
typedef struct nm_config {
  int32_t  struct_size;     /* sizeof(nm_config) -- ABI guard                              */
  int32_t  device;          /* CUDA device ordinal                                         */
  int32_t  natoms;          /* atoms per replica (4*SZ^3 for fcc)                          */
  int32_t  n_rep;           /* replicas resident on this engine (local slots)              */
  int32_t  n_rep_global;    /* NS = NP*NT                                                  */
  int32_t  rep_offset;      /* global slot index of local slot 0 (multiple of nt)          */
  int32_t  nt;              /* NT: temperatures per pressure row                           */
  int32_t  precision;       /* 64 (default) or 32: arithmetic of the LJ/MD kernels         */
  int32_t  nstps;           /* NSTPS, HMC velocity-Verlet steps (-ts)                      */
  int32_t  mod;             /* MOD, moves per collection cycle (-sm)                       */
  int32_t  bulk_move;       /* BM (-bm): 1 bulk PMC, 0 iterative single-atom PMC           */
  int32_t  text_rounding;   /* 1: reproduce the '%f' (6-decimal) rounding the reference
                               applies to every value it passes to LAMMPS as text
                               (timestep, box side, bulk displacement, velocity T)     */
  int32_t  row_stride;      /* global pressure rows between consecutive local rows; <= 1 = contiguous block.
                               G ranks, cyclic: rep_offset = rank*nt, row_stride = G        */
  int32_t  reserved0;       /* set to 0                                                    */
  double   ppos, pvol;      /* PPOS, PVOL move probabilities (-pm, -vm)                    */
  double   lat_scale;       /* LAT[EL][1] (1.122): displacement scale factor               */
  double   mass;            /* MASS[EL]                                                    */
  double   rc;              /* lj/cut cutoff (2.5); epsilon = sigma = 1                    */
  double   skin;            /* inner Verlet-list skin; <= 0 selects the default (0.4 for natoms <= 768, else 0.3) */
  double   skin_outer;      /* extra radius of the outer list; <= 0 selects the default (1.3) */
  uint64_t seed;            /* counter-based RNG seed (reference: SEED = 256)              */
  void*    stream;          /* cudaStream_t to launch on; NULL = engine-owned stream       */
} nm_config;

const char* nm_last_error(void);
int  nm_abi_version(void);
/* number of visible CUDA devices, or NM_ENODEV */
int  nm_device_count(void);

/* ---- engine lifetime: replaces lammps(cmdargs=...) / lmps.file / lmps.close
 *      (lammps_remcmc.py:399-400,462-463,428,683): one engine holds ALL local
 *      replicas for the whole run instead of one LAMMPS instance per replica
 *      per cycle. */
int  nm_create(const nm_config* cfg, nm_engine** out);
int  nm_destroy(nm_engine* h);
int  nm_set_stream(nm_engine* h, void* cuda_stream);
int  nm_synchronize(nm_engine* h);

/* ---- state in/out: replaces change_box + scatter_atoms('x'/'v') + 'run 0'
 *      (init_lammps, lammps_remcmc.py:459-470) and gather_atoms / extract_global
 *      (lammps_extract, :377-391). Any pointer may be NULL (= leave / skip).
 *      x, v: [n_rep][3*natoms]; box, dx, dv, dt: [n_rep]. Uploading a state also
 *      evaluates pe / virial / forces for it (the 'run 0'). */
int  nm_set_state(nm_engine* h, const double* x, const double* v, const double* box,
                  const double* dx, const double* dv, const double* dt);
int  nm_get_state(nm_engine* h, double* x, double* v, double* box,
                  double* dx, double* dv, double* dt);

/* ---- 'velocity all create T[j] seed dist gaussian' + 'zero linear' + 'zero angular' for every local slot outside a move:
 *      the draw init_sample leaves in STATE with -is (lammps_remcmc.py:420-425). tag selects the RNG stream. */
int  nm_velocity_create(nm_engine* h, int64_t tag);

/* ---- thermodynamic labels of the local slots: (et, pf) of init_constant
 *      (lammps_remcmc.py:128-131), T[j] and the '%f'-rounded T LAMMPS receives
 *      in 'velocity all create %f' (:604). Each [n_rep]. */
int  nm_set_labels(nm_engine* h, const double* et, const double* pf,
                   const double* temp, const double* temp_vel);

/* ---- a-1: LJ lj/cut evaluation of the resident configurations: what
 *      'run 0' + extract_compute('thermo_pe'/'thermo_press') deliver
 *      (lammps_remcmc.py:384,386,405). pe, w (= sum r.f), npairs: [n_rep];
 *      f: [n_rep][3*natoms]. Any output may be NULL. */
int  nm_eval(nm_engine* h, double* pe, double* w, double* f, int64_t* npairs);

/* ---- a-2..a-9: one collection cycle = MOD moves for every local replica
 *      (gen_sample, lammps_remcmc.py:665-691), asynchronous on the stream.
 *      cycle = STEP index (selects the RNG streams). */
int  nm_run_cycle(nm_engine* h, int64_t cycle);
/* thermo record of every local slot after the last cycle: [n_rep][NM_THERMO_WIDTH] */
int  nm_get_thermo(nm_engine* h, double* out);
/* ---- a-10: gen_mc_param (lammps_remcmc.py:726-745): adapt dx, dv, dt; zero counters */
int  nm_adapt(nm_engine* h);

/* ---- a-11: replica_exchange (lammps_remcmc.py:776-803). Exchanges never cross pressure rows
 *      (:782-789) and an engine holds whole rows, so every engine decides the swaps of ITS rows
 *      from its own (pe+ke, vol) values: no collective sits on the critical path. et / pf are the
 *      labels of nm_set_labels; the uniform of row u's n-th pair is draw u*NT*(NT-1)/2 + n of the
 *      job-wide stream whatever the row -> rank map.
 *   nm_exchange      : pack + sweep of the local rows + permutation of the local
 *                      slot -> configuration labels (configurations never move).
 *                      uniforms: optional HOST array of NP*NT*(NT-1)/2 doubles in GLOBAL draw
 *                      order (injects the reference's np.random stream); NULL = the engine's
 *                      counter-based stream for this cycle.
 *                      perm_out: optional HOST int32[n_rep]; perm_out[k] = local slot whose
 *                      pre-exchange configuration now sits in local slot k.
 *                      swaps_out: optional HOST int64, accepted swaps of the local rows.
 *                      With both NULL the call is asynchronous on the stream.
 *   nm_exchange_pack : writes (pe+ke, vol) of every local slot to a DEVICE buffer
 *                      double[n_rep][2]: the payload of the NCCL all-gather that gives every
 *                      rank (and the log) the job-wide table, as the reference's client holds it.
 *   nm_exchange_apply: the same sweep of the local rows, reading a job-wide DEVICE table
 *                      double[n_rep_global][2] in GLOBAL slot order (an all-gathered table, or
 *                      one injected by a test) instead of the local pack. */
int  nm_exchange(nm_engine* h, const double* uniforms, int64_t cycle,
                 int32_t* perm_out, int64_t* swaps_out);
int  nm_exchange_pack(nm_engine* h, void* dev_dst);
int  nm_exchange_apply(nm_engine* h, const void* dev_table_global,
                       const double* uniforms, int64_t cycle,
                       int32_t* perm_out, int64_t* swaps_out);

/* counters (for the roofline figures): out[NM_COUNTER_WIDTH] */
int  nm_get_counters(nm_engine* h, uint64_t* out);
int  nm_reset_counters(nm_engine* h);
/* number of kernels this engine has launched since nm_create (bench.py: gpu_launches) */
int64_t nm_launch_count(nm_engine* h);
/* SM clocks each local slot's CTA spent in the last cycle kernel (load-balance diagnostics): out[n_rep] */
int  nm_get_cta_clocks(nm_engine* h, uint64_t* out);
/* the last cycle's counters of each local slot (diagnostics): out[n_rep][NM_COUNTER_WIDTH] */
int  nm_get_replica_counters(nm_engine* h, uint64_t* out);

/* ---- a-14: calculate_rdf (lammps_distr.py:123-135) over a batch of samples.
 *   pos   : HOST or DEVICE float32 [nsamples][natoms][3] (dev_ptrs selects which)
 *   box   : float32 [nsamples]
 *   edges : HOST float64 [nbins] (R = linspace(1e-16, .5, SBINS)*l, lammps_distr.py:86,94)
 *   counts: uint32 [nsamples][nbins]; counts[s][0] = 0, counts[s][1+b] = histogram bin b
 *           (the reference's rd before '/natoms', :134). HOST or DEVICE like pos.
 *   Bit-exact: every distance is formed in un-contracted float32 exactly as numba
 *   forms it, and binned by np.histogram's rules against the float64 edges. */
int  nm_rdf_counts(int device, void* cuda_stream, int dev_ptrs,
                   const float* pos, const float* box, int32_t natoms, int64_t nsamples,
                   const double* edges, int32_t nbins, uint32_t* counts);

/* ---- N1 (next row): calculate_cdf (lammps_distr.py:161-171) over a batch of samples.
 *   edges : HOST float64 [3][nb+1] (RV = linspace(0, l, CBINS+1) - l/2 per axis, lammps_distr.py:109-111)
 *   counts: uint32 [nsamples][nb][nb][nb], the reference's cd before '/natoms' (sum over the 27 images of
 *           np.histogramdd of the float32 pair vectors). nb <= 32. Bit-exact like nm_rdf_counts. */
int  nm_cdf_counts(int device, void* cuda_stream, int dev_ptrs,
                   const float* pos, const float* box, int32_t natoms, int64_t nsamples,
                   const double* edges, int32_t nb, uint32_t* counts);

/* ---- a-13: native '%.4E' text records (write_thrm / write_traj,
 *      lammps_remcmc.py:235-256). Pure host code; returns bytes written or <0.
 *      buf may be NULL to query the size. */
int64_t nm_format_thrm(const double* vals17, char* buf, int64_t cap);
int64_t nm_format_traj(int32_t natoms, double box, const double* x, char* buf, int64_t cap);
/* nrep trajectory records formatted by nthreads host threads and packed back to back in replica
 * order; out_off[nrep+1] receives the byte offsets. box: [nrep], x: [nrep][3*natoms]. */
int64_t nm_format_traj_batch(int32_t nrep, int32_t natoms, const double* box, const double* x,
                             char* buf, int64_t cap, int64_t* out_off, int32_t nthreads);

/* ---- N2: streaming writers. One call per recorded cycle appends every local replica's record to its own file
 *      (open(.., 'a') of write_thrm / write_traj), formatted once; NULL path = format only. pos_out / box_out /
 *      parsed_out (optional) receive what lammps_parse.py:45,88-93 reads back from that text (decimal -> double ->
 *      float32), i.e. the contents of the parser's .pos/.box/.<thermo>.npy files. Return bytes written or <0.
 *      vals: [nrep][17] (write_thrm column order); box: [nrep]; x: [nrep][3*natoms]. */
int64_t nm_append_traj_batch(int32_t nrep, int32_t natoms, const double* box, const double* x,
                             const char* const* paths, int32_t nthreads, float* pos_out, float* box_out);
int64_t nm_append_thrm_batch(int32_t nrep, const double* vals, const char* const* paths, float* parsed_out);

/* ---- roofline denominators: sustained FMA issue rate of this device, measured
 *      with a dependent-chain-free FMA kernel. Returns FLOP/s (2 per FMA). */
int  nm_measure_fma_peak(int device, int precision, double* flops_per_s, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* NM_B200_H */
