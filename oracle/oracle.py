"""CPU oracle for the neuralMelting hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module. The product package
(``neuralmelting_b200``) never does and fails loudly without its CUDA library.

Two layers:
  * ctypes bindings to ``oracle/libnm_oracle.so`` (``nm_oracle.c``: C restatement of the
    LJ ``lj/cut`` evaluation, the Monte Carlo moves / cycle / adaptation / exchange of
    ``/root/reference/scripts/lammps_remcmc.py`` and the RDF of ``lammps_distr.py``);
  * independent NumPy restatements (vectorised O(N^2)) used to cross-check the C code.

PARITY UNPINNED for the LJ/MD physics (LAMMPS is an un-vendored, unpinned third-party
dependency of the reference and is not installable here); pinned for exchange / adaptation /
text formats / RDF through tests/golden (made by oracle/gen_golden.py from the reference's
own functions).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnm_oracle.so")

THERMO_WIDTH = 18
CT_N = 12
(CT_SWEEPS, CT_HMC_MOVES, CT_HMC_ATOM_STEPS, CT_VMC_MOVES, CT_PMC_MOVES, CT_PMC_TRIALS,
 CT_FORCE_EVALS, CT_PAIRS_FORCE, CT_PAIRS_FULL, CT_PAIRS_DELTA, CT_LIST_BUILDS, CT_LIST_PAIRS) = range(12)


def build(force=False):
    """compile nm_oracle.c with gcc (oracle/Makefile)"""
    src = os.path.join(_HERE, "nm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libnm_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Params(C.Structure):
    _fields_ = [("nstps", C.c_int32), ("mod", C.c_int32), ("bulk_move", C.c_int32),
                ("text_rounding", C.c_int32),
                ("ppos", C.c_double), ("pvol", C.c_double), ("lat_scale", C.c_double),
                ("mass", C.c_double), ("rc", C.c_double), ("skin", C.c_double),
                ("seed", C.c_uint64)]


def make_params(nstps=8, mod=128, bulk_move=0, text_rounding=1, ppos=0.125, pvol=0.125,
                lat_scale=1.122, mass=1.0, rc=2.5, skin=0.3, seed=256):
    return Params(nstps, mod, bulk_move, text_rounding, ppos, pvol, lat_scale, mass, rc, skin, seed)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, u64p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        L.orc_round6.restype = C.c_double
        L.orc_round6.argtypes = [C.c_double]
        L.orc_lj_delta_atom.restype = C.c_double
        L.orc_exchange.restype = C.c_int64
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


# ----------------------------------------------------------------------------- LJ
def lj_eval_n2(x, box, rc=2.5):
    """brute-force lj/cut (C). x: (N,3) float64. returns pe, w, f(N,3), npairs"""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
    n = x.shape[0]
    pe, w, npairs = C.c_double(), C.c_double(), C.c_int64()
    f = np.zeros((n, 3))
    lib().orc_lj_eval_n2(C.c_int(n), _dp(x), C.c_double(box), C.c_double(rc),
                         C.byref(pe), C.byref(w), _dp(f), C.byref(npairs))
    return pe.value, w.value, f, npairs.value


def lj_eval_list(x, box, rc=2.5, skin=0.3):
    """Verlet-list lj/cut (C, second implementation)"""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
    n = x.shape[0]
    pe, w, npairs = C.c_double(), C.c_double(), C.c_int64()
    f = np.zeros((n, 3))
    lib().orc_lj_eval_list(C.c_int(n), _dp(x), C.c_double(box), C.c_double(rc), C.c_double(skin),
                           C.byref(pe), C.byref(w), _dp(f), C.byref(npairs))
    return pe.value, w.value, f, npairs.value


def lj_eval_numpy(x, box, rc=2.5):
    """independent NumPy restatement of pair_style lj/cut (deck: lammps_remcmc.py:364-367):
    strict rsq < rc^2, epsilon = sigma = 1, unshifted, no tail correction."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 3)
    d = x[:, None, :] - x[None, :, :]
    d -= box * np.rint(d / box)
    rsq = np.einsum("ijk,ijk->ij", d, d)
    iu = np.triu_indices(x.shape[0], 1)
    mask = np.zeros_like(rsq, dtype=bool)
    mask[iu] = rsq[iu] < rc * rc
    r2inv = np.zeros_like(rsq)
    r2inv[mask] = 1.0 / rsq[mask]
    r6inv = r2inv ** 3
    fpair = r6inv * (48.0 * r6inv - 24.0) * r2inv
    fij = d * fpair[:, :, None]
    f = fij.sum(1) - fij.sum(0)
    pe = float(np.sum(r6inv * (4.0 * r6inv - 4.0)))
    w = float(np.sum(rsq * fpair))
    return pe, w, f, int(mask.sum())


def set_delta_lists(on):
    """iterative PMC: per-sweep candidate lists (default, bitwise identical results) or the plain all-atom sum"""
    lib().orc_set_delta_lists(C.c_int(int(bool(on))))


def lj_delta_atom(x, k, xn, box, rc=2.5):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
    xn = np.ascontiguousarray(xn, dtype=np.float64)
    return lib().orc_lj_delta_atom(C.c_int(x.shape[0]), _dp(x), C.c_int(k), _dp(xn), C.c_double(box), C.c_double(rc))


def wrap(x, box):
    x = np.array(x, dtype=np.float64, order="C")
    lib().orc_wrap(C.c_int(x.size // 3), _dp(x), C.c_double(box))
    return x


def round6(v):
    """'%f' text round trip (what LAMMPS receives from the reference's command strings)"""
    return float("%f" % v)


# ----------------------------------------------------------------------------- lattice helpers
def fcc_positions(sz, box):
    """fcc lattice of sz^3 conventional cells in a cubic box of side `box`, LAMMPS create_atoms order
    (cells k-j-i outermost to innermost x fastest? -- order is irrelevant to every test that uses it)"""
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    cells = np.array([[i, j, k] for k in range(sz) for j in range(sz) for i in range(sz)], dtype=np.float64)
    frac = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) / sz
    return frac * box


def fcc_shell_sum(a, rc=2.5):
    """analytic per-atom lattice sums of the perfect fcc crystal with cubic constant a under lj/cut rc:
    returns (E/N, W/N) from explicit neighbour shells -- the independent anchor for the oracle."""
    m = int(np.ceil(rc / a * 2)) + 2
    e = w = 0.0
    nn = 0
    for i in range(-m, m + 1):
        for j in range(-m, m + 1):
            for k in range(-m, m + 1):
                if (i + j + k) % 2 or (i == 0 and j == 0 and k == 0):
                    continue
                rsq = (i * i + j * j + k * k) * (a / 2) ** 2
                if rsq < rc * rc:
                    r2 = 1.0 / rsq
                    r6 = r2 ** 3
                    e += 0.5 * r6 * (4 * r6 - 4)
                    w += 0.5 * rsq * (r6 * (48 * r6 - 24) * r2)
                    nn += 1
    return e, w, nn


# ----------------------------------------------------------------------------- MC engine
def cycle(params, label, slot_global, cycle_idx, x, v, scal, counts, trace=False):
    """gen_sample (lammps_remcmc.py:665-691). label=(et,pf,temp,temp_vel); scal=[box,dx,dv,dt];
    counts=[ntp,nap,ntv,nav,nth,nah]. Arrays are modified in place. returns thermo(18), ct(12)[, trace]"""
    n = x.size // 3
    label = np.ascontiguousarray(label, dtype=np.float64)
    thermo = np.zeros(THERMO_WIDTH)
    ct = np.zeros(CT_N, dtype=np.uint64)
    tr = np.zeros((params.mod, 3)) if trace else None
    err = lib().orc_cycle(C.byref(params), _dp(label), C.c_int32(slot_global), C.c_int64(cycle_idx), C.c_int32(n),
                          _dp(x), _dp(v), _dp(scal), _dp(counts), _dp(thermo),
                          ct.ctypes.data_as(C.POINTER(C.c_uint64)), _dp(tr))
    if err:
        raise RuntimeError("oracle cycle error %d" % err)
    return (thermo, ct, tr) if trace else (thermo, ct)


def farm(params, labels, slot0, x, v, scal, counts, cycle0=0, ncycles=1, nthreads=1):
    """one task per replica over a thread pool (the Dask decomposition, lammps_remcmc.py:698-700).
    x,v: (nrep,3N); scal: (nrep,4); counts: (nrep,6); labels: (nrep,4). In place. returns thermo, ct"""
    nrep = x.shape[0]
    n = x.shape[1] // 3
    thermo = np.zeros((nrep, THERMO_WIDTH))
    ct = np.zeros(CT_N, dtype=np.uint64)
    labels = np.ascontiguousarray(labels, dtype=np.float64)
    err = lib().orc_farm(C.byref(params), _dp(labels), C.c_int32(slot0), C.c_int32(nrep), C.c_int32(n),
                         C.c_int64(cycle0), C.c_int64(ncycles), _dp(x), _dp(v), _dp(scal), _dp(counts),
                         _dp(thermo), ct.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int32(nthreads))
    if err:
        raise RuntimeError("oracle farm error %d" % err)
    return thermo, ct


def adapt(step, ratio):
    """gen_mc_param (lammps_remcmc.py:726-745) on (dx,dv,dt) given float32 ratios (ap,av,ah)"""
    step = np.array(step, dtype=np.float64)
    ratio = np.ascontiguousarray(ratio, dtype=np.float64)
    lib().orc_adapt(_dp(step), _dp(ratio))
    return step


def exchange(np_, nt, etot, vol, et, pf, uniforms):
    """replica_exchange (lammps_remcmc.py:776-803). returns perm (perm[k] = source slot), swaps"""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (etot, vol, et, pf, uniforms)]
    perm = np.zeros(np_ * nt, dtype=np.int32)
    swaps = lib().orc_exchange(C.c_int32(np_), C.c_int32(nt), *[_dp(a) for a in arrs],
                               perm.ctypes.data_as(C.POINTER(C.c_int32)))
    return perm, int(swaps)


def exchange_uniforms(seed, cycle_idx, n):
    out = np.zeros(n)
    lib().orc_exchange_uniforms(C.c_uint64(seed), C.c_int64(cycle_idx), C.c_int64(n), _dp(out))
    return out


def philox(k0, k1, ctr):
    c = (C.c_uint32 * 4)(*ctr)
    o = (C.c_uint32 * 4)()
    lib().orc_philox(C.c_uint32(k0), C.c_uint32(k1), c, o)
    return list(o)


# ----------------------------------------------------------------------------- RDF
def rdf_edges(box_all, sbins):
    """R of calculate_spatial (lammps_distr.py:82-94): float64 edges scaled by the minimum box"""
    l = np.min(box_all)
    r = np.linspace(1e-16, 1 / 2, sbins)
    return r * l


def rdf_counts(pos, box, r):
    """C restatement of calculate_rdf before '/natoms' (lammps_distr.py:123-134)"""
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    r = np.ascontiguousarray(r, dtype=np.float64)
    counts = np.zeros(r.size, dtype=np.uint32)
    lib().orc_rdf_counts(C.c_int32(pos.shape[0]), pos.ctypes.data_as(C.POINTER(C.c_float)),
                         C.c_float(np.float32(box)), _dp(r), C.c_int32(r.size),
                         counts.ctypes.data_as(C.POINTER(C.c_uint32)))
    return counts


def rdf_counts_farm(pos, box, r, nthreads=1):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    ns, n = pos.shape[0], pos.shape[1]
    box = np.ascontiguousarray(box, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.float64)
    counts = np.zeros((ns, r.size), dtype=np.uint32)
    lib().orc_rdf_farm(C.c_int32(n), C.c_int64(ns), pos.ctypes.data_as(C.POINTER(C.c_float)),
                       box.ctypes.data_as(C.POINTER(C.c_float)), _dp(r), C.c_int32(r.size),
                       counts.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int32(nthreads))
    return counts


def rdf_counts_numpy(pos, box, r):
    """NumPy float32 restatement of calculate_rdf (one image at a time, exactly the reference's ops)"""
    pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    box = np.float32(box)
    b = [-1, 0, 1]
    br = np.array([[b[i], b[j], b[k]] for i in range(3) for j in range(3) for k in range(3)], dtype=np.int8)
    rd = np.zeros(r.size, dtype=np.float32)
    for j in range(br.shape[0]):
        shift = (box * br[j].astype(np.float32)).astype(np.float32)
        dvm = pos - (pos + shift.reshape(1, -1)).reshape(-1, 1, 3)
        sq = np.square(dvm)
        d = np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])
        rd[1:] += np.histogram(d, r)[0]
    return rd.astype(np.uint32)


# ----------------------------------------------------------------------------- CDF (next row N1)
def cdf_edges(box_all, cbins):
    """RV of calculate_spatial (lammps_distr.py:109-111): three rows of CBINS+1 float64 edges centred on 0
    (float64 as under the numpy-1 rules the script was written for)"""
    l = float(np.min(box_all))
    rv = np.array([np.linspace(0, l, cbins + 1) for _ in range(3)], dtype=np.float64)
    rv -= l / 2
    return rv


def cdf_counts(pos, box, rv):
    """C restatement of calculate_cdf before '/natoms' (lammps_distr.py:161-170)"""
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    rv = np.ascontiguousarray(rv, dtype=np.float64)
    nb = rv.shape[1] - 1
    counts = np.zeros((nb, nb, nb), dtype=np.uint32)
    lib().orc_cdf_counts(C.c_int32(pos.shape[0]), pos.ctypes.data_as(C.POINTER(C.c_float)), C.c_float(np.float32(box)),
                         _dp(rv), C.c_int32(nb), counts.ctypes.data_as(C.POINTER(C.c_uint32)))
    return counts


def cdf_counts_numpy(pos, box, rv):
    """NumPy restatement: np.histogramdd per image on the float32 pair vectors"""
    pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    box = np.float32(box)
    b = [-1, 0, 1]
    br = np.array([[b[i], b[j], b[k]] for i in range(3) for j in range(3) for k in range(3)], dtype=np.int8)
    nb = rv.shape[1] - 1
    cd = np.zeros((nb, nb, nb), dtype=np.float32)
    for j in range(br.shape[0]):
        shift = (box * br[j].astype(np.float32)).astype(np.float32)
        dvm = pos - (pos + shift.reshape(1, -1)).reshape(-1, 1, 3)
        cd += np.histogramdd(dvm.reshape(-1, 3), list(rv))[0]
    return cd.astype(np.uint32)
