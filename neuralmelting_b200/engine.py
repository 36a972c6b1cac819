"""ctypes binding of libnm_b200.so (C-ABI: include/nm_b200.h).

The host side mirrors what /root/reference/scripts/lammps_remcmc.py asks of its per-replica LAMMPS
object (lammps_remcmc.py:377-391, 459-470 and every move at :477-640), but for ALL local replicas
at once and with the state resident on the GPU. There is no CPU fallback: a missing library or a
missing CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NM_B200_LIB") or os.path.join(_HERE, "libnm_b200.so")      # NM_B200_LIB: development builds (tools/)

THERMO_WIDTH = 18
THERMO_COLS = ("temp", "pe", "ke", "virial", "box", "vol", "dx", "dv", "dt",
               "ntp", "nap", "ntv", "nav", "nth", "nah", "ap", "av", "ah")
COUNTER_WIDTH = 22
COUNTER_COLS = ("sweeps", "hmc_moves", "hmc_atom_steps", "vmc_moves", "pmc_moves", "pmc_trials",
                "force_evals", "pairs_force", "pairs_full", "pairs_delta", "list_builds", "list_pairs",
                "clk_eval", "clk_build", "clk_total", "outer_builds", "clk_outer", "clk_inner", "clk_vel", "helped_evals", "dbg_loopclk", "dbg_loopit")

NM_OK, NM_EINVAL, NM_ENODEV, NM_ECUDA, NM_ENOMEM, NM_EBOX, NM_ENEIGH, NM_ESTATE = 0, -1, -2, -3, -4, -5, -6, -7


class NmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nm_b200 error %d: %s" % (code, msg))
        self.code = code


class NmConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("natoms", C.c_int32),
                ("n_rep", C.c_int32), ("n_rep_global", C.c_int32), ("rep_offset", C.c_int32),
                ("nt", C.c_int32), ("precision", C.c_int32), ("nstps", C.c_int32), ("mod", C.c_int32),
                ("bulk_move", C.c_int32), ("text_rounding", C.c_int32), ("row_stride", C.c_int32), ("reserved0", C.c_int32),
                ("ppos", C.c_double), ("pvol", C.c_double), ("lat_scale", C.c_double),
                ("mass", C.c_double), ("rc", C.c_double), ("skin", C.c_double), ("skin_outer", C.c_double),
                ("seed", C.c_uint64), ("stream", C.c_void_p)]


_lib = None
_DP = C.POINTER(C.c_double)


def load_library():
    """dlopen the CUDA library; fails loudly when it has not been built (no fallback)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m neuralmelting_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.nm_last_error.restype = C.c_char_p
    L.nm_launch_count.restype = C.c_int64
    L.nm_format_thrm.restype = C.c_int64
    L.nm_format_traj.restype = C.c_int64
    L.nm_create.argtypes = [C.POINTER(NmConfig), C.POINTER(C.c_void_p)]
    for name in ("nm_destroy", "nm_synchronize", "nm_adapt", "nm_reset_counters"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.nm_launch_count.argtypes = [C.c_void_p]
    L.nm_get_cta_clocks.argtypes = [C.c_void_p, C.c_void_p]
    L.nm_get_replica_counters.argtypes = [C.c_void_p, C.c_void_p]
    L.nm_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.nm_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    L.nm_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    L.nm_set_labels.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.nm_eval.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.nm_run_cycle.argtypes = [C.c_void_p, C.c_int64]
    L.nm_get_thermo.argtypes = [C.c_void_p, C.c_void_p]
    L.nm_exchange_pack.argtypes = [C.c_void_p, C.c_void_p]
    L.nm_exchange_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.nm_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.nm_get_counters.argtypes = [C.c_void_p, C.c_void_p]
    L.nm_rdf_counts.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                C.c_void_p, C.c_int32, C.c_void_p]
    L.nm_cdf_counts.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                C.c_void_p, C.c_int32, C.c_void_p]
    L.nm_velocity_create.argtypes = [C.c_void_p, C.c_int64]
    L.nm_format_traj_batch.restype = C.c_int64
    L.nm_format_traj_batch.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]
    L.nm_append_traj_batch.restype = C.c_int64
    L.nm_append_traj_batch.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.nm_append_thrm_batch.restype = C.c_int64
    L.nm_append_thrm_batch.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nm_format_thrm.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.nm_format_traj.argtypes = [C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int64]
    L.nm_measure_fma_peak.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise NmError(rc, load_library().nm_last_error().decode(errors="replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.size != int(np.prod(shape)):
        raise ValueError("array of size %d where %s expected" % (a.size, (shape,)))
    return a


def device_count():
    n = load_library().nm_device_count()
    if n < 0:
        raise NmError(n, load_library().nm_last_error().decode(errors="replace"))
    return n


class Engine:
    """All local replicas of the (P, T) grid on one GPU.

    Host arrays are in local slot order (lammps_remcmc.py:117: k = i*NT + j): local pressure row lr is global row
    rep_offset/NT + lr*row_stride (row_stride 1: contiguous block; G ranks cyclic: rep_offset = rank*NT, row_stride = G).
    """

    def __init__(self, natoms, n_rep, nt, n_rep_global=None, rep_offset=0, row_stride=1, device=0, nstps=8, mod=128,
                 bulk_move=False, ppos=0.125, pvol=0.125, lat_scale=1.122, mass=1.0, rc=2.5, skin=0.0, skin_outer=0.0,
                 seed=256, text_rounding=True, precision=64, stream=None):
        L = load_library()
        self.natoms, self.n_rep, self.nt = int(natoms), int(n_rep), int(nt)
        self.n_rep_global = int(n_rep_global if n_rep_global is not None else n_rep)
        self.rep_offset, self.row_stride = int(rep_offset), max(1, int(row_stride))
        self.mod, self.nstps = int(mod), int(nstps)
        cfg = NmConfig(C.sizeof(NmConfig), device, natoms, n_rep, self.n_rep_global, rep_offset, nt, precision,
                       nstps, mod, int(bool(bulk_move)), int(bool(text_rounding)), self.row_stride, 0, ppos, pvol, lat_scale, mass, rc,
                       skin, skin_outer, seed, None)
        self._h = C.c_void_p()
        _check(L.nm_create(C.byref(cfg), C.byref(self._h)))
        self._L = L
        if stream is not None:      # 0 = the legacy default stream (what torch.cuda.current_stream() is by default)
            _check(L.nm_set_stream(self._h, C.c_void_p(stream)))

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.nm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream):
        _check(self._L.nm_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        _check(self._L.nm_synchronize(self._h))

    # -- state (init_lammps / lammps_extract)
    def set_state(self, x=None, v=None, box=None, dx=None, dv=None, dt=None):
        n3 = (self.n_rep, 3 * self.natoms)
        x, v = _f64(x, n3), _f64(v, n3)
        box, dx, dv, dt = (_f64(a, (self.n_rep,)) for a in (box, dx, dv, dt))
        _check(self._L.nm_set_state(self._h, _ptr(x), _ptr(v), _ptr(box), _ptr(dx), _ptr(dv), _ptr(dt)))

    def get_state(self, want_x=True, want_v=True, x_out=None, v_out=None):
        """state of every local slot; x_out / v_out: caller-owned C-contiguous float64 (n_rep, 3 natoms) arrays to fill
        in place (pinned memory makes the device-to-host copy run at PCIe speed and saves a host copy)"""
        n3 = (self.n_rep, 3 * self.natoms)
        for a in (x_out, v_out):
            if a is not None and not (a.dtype == np.float64 and a.shape == n3 and a.flags["C_CONTIGUOUS"]):
                raise ValueError("get_state: output arrays must be C-contiguous float64 of shape %s" % (n3,))
        x = x_out if x_out is not None else (np.empty(n3) if want_x else None)
        v = v_out if v_out is not None else (np.empty(n3) if want_v else None)
        box, dx, dv, dt = (np.empty(self.n_rep) for _ in range(4))
        _check(self._L.nm_get_state(self._h, _ptr(x), _ptr(v), _ptr(box), _ptr(dx), _ptr(dv), _ptr(dt)))
        return dict(x=x, v=v, box=box, dx=dx, dv=dv, dt=dt)

    def set_labels(self, et, pf, temp, temp_vel=None):
        sh = (self.n_rep,)
        temp = _f64(temp, sh)
        if temp_vel is None:      # '%f' text round trip of lammps_remcmc.py:604
            temp_vel = np.array([float("%f" % t) for t in temp])
        et, pf, temp_vel = _f64(et, sh), _f64(pf, sh), _f64(temp_vel, sh)
        self._labels = (et.copy(), pf.copy())
        _check(self._L.nm_set_labels(self._h, _ptr(et), _ptr(pf), _ptr(temp), _ptr(temp_vel)))

    def velocity_create(self, tag=0):
        """draw velocities at every slot's temperature (LAMMPS 'velocity create' + zero linear + zero angular)"""
        _check(self._L.nm_velocity_create(self._h, int(tag)))

    # -- a-1
    def eval(self, want_forces=True):
        pe, w = np.empty(self.n_rep), np.empty(self.n_rep)
        f = np.empty((self.n_rep, self.natoms, 3)) if want_forces else None
        npairs = np.empty(self.n_rep, dtype=np.int64)
        _check(self._L.nm_eval(self._h, _ptr(pe), _ptr(w), _ptr(f), _ptr(npairs)))
        return pe, w, f, npairs

    # -- a-2..a-10
    def run_cycle(self, cycle):
        _check(self._L.nm_run_cycle(self._h, int(cycle)))

    def get_thermo(self):
        out = np.empty((self.n_rep, THERMO_WIDTH))
        _check(self._L.nm_get_thermo(self._h, _ptr(out)))
        return out

    def adapt(self):
        _check(self._L.nm_adapt(self._h))

    # -- a-11
    def exchange_pack(self, dev_ptr):
        _check(self._L.nm_exchange_pack(self._h, C.c_void_p(dev_ptr)))

    def global_slots(self):
        """global slot index k = i*NT + j of every local slot"""
        lr, j = np.divmod(np.arange(self.n_rep), self.nt)
        return (self.rep_offset // self.nt + lr * self.row_stride) * self.nt + j

    def exchange(self, cycle, uniforms=None, want_perm=True):
        """replica exchange of the local pressure rows (rows never exchange with each other, lammps_remcmc.py:782-789).
        uniforms: optional job-wide array in global draw order. want_perm=False: asynchronous, returns (None, None)."""
        uniforms = _f64(uniforms)
        perm = np.empty(self.n_rep, dtype=np.int32) if want_perm else None
        swaps = C.c_int64(0)
        _check(self._L.nm_exchange(self._h, _ptr(uniforms), int(cycle), _ptr(perm), C.addressof(swaps) if want_perm else None))
        return (perm, swaps.value) if want_perm else (None, None)

    def exchange_apply(self, dev_table_ptr, cycle, uniforms=None, want_perm=True):
        """the same sweep reading a job-wide device table double[n_rep_global][2] (global slot order)"""
        uniforms = _f64(uniforms)
        perm = np.empty(self.n_rep, dtype=np.int32) if want_perm else None
        swaps = C.c_int64(0)
        _check(self._L.nm_exchange_apply(self._h, C.c_void_p(dev_table_ptr), _ptr(uniforms), int(cycle), _ptr(perm),
                                         C.addressof(swaps) if want_perm else None))
        return (perm, swaps.value) if want_perm else (None, None)

    # -- bookkeeping
    def counters(self):
        out = np.zeros(COUNTER_WIDTH, dtype=np.uint64)
        _check(self._L.nm_get_counters(self._h, _ptr(out)))
        return dict(zip(COUNTER_COLS, (int(v) for v in out)))

    def reset_counters(self):
        _check(self._L.nm_reset_counters(self._h))

    def launch_count(self):
        return int(self._L.nm_launch_count(self._h))

    def cta_clocks(self):
        out = np.zeros(self.n_rep, dtype=np.uint64)
        _check(self._L.nm_get_cta_clocks(self._h, _ptr(out)))
        return out

    def replica_counters(self):
        """last cycle's counters per local slot: (n_rep, NM_COUNTER_WIDTH) uint64, columns as COUNTER_COLS"""
        out = np.zeros((self.n_rep, COUNTER_WIDTH), dtype=np.uint64)
        _check(self._L.nm_get_replica_counters(self._h, _ptr(out)))
        return out


# ----------------------------------------------------------------------------- a-14 RDF
def rdf_counts(pos, box, edges, device=0, stream=None):
    """calculate_rdf (lammps_distr.py:123-135) before '/natoms' for a batch of samples on the GPU.
    pos (S,N,3) float32, box (S,) float32, edges (SBINS,) float64 -> counts (S,SBINS) uint32"""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    if pos.ndim == 2:
        pos = pos[None]
    ns, n = pos.shape[0], pos.shape[1]
    box = np.ascontiguousarray(box, dtype=np.float32).reshape(-1)
    if box.size != ns:
        raise ValueError("box must have one entry per sample")
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    counts = np.zeros((ns, edges.size), dtype=np.uint32)
    if ns == 0:
        return counts
    _check(load_library().nm_rdf_counts(device, C.c_void_p(stream or 0), 0, _ptr(pos), _ptr(box), n, ns,
                                        _ptr(edges), edges.size, _ptr(counts)))
    return counts


def rdf_counts_device(pos_ptr, box_ptr, natoms, nsamples, edges, counts_ptr, device=0, stream=None):
    """same, on device-resident buffers (pointers as ints, e.g. torch .data_ptr())"""
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    _check(load_library().nm_rdf_counts(device, C.c_void_p(stream or 0), 1, C.c_void_p(pos_ptr), C.c_void_p(box_ptr),
                                        natoms, nsamples, _ptr(edges), edges.size, C.c_void_p(counts_ptr)))


def cdf_counts(pos, box, edges, device=0, stream=None):
    """calculate_cdf (lammps_distr.py:161-171) before '/natoms' for a batch of samples on the GPU.
    pos (S,N,3) float32, box (S,) float32, edges (3, CBINS+1) float64 -> counts (S,CBINS,CBINS,CBINS) uint32"""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    if pos.ndim == 2:
        pos = pos[None]
    ns, n = pos.shape[0], pos.shape[1]
    box = np.ascontiguousarray(box, dtype=np.float32).reshape(-1)
    if box.size != ns:
        raise ValueError("box must have one entry per sample")
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    if edges.ndim != 2 or edges.shape[0] != 3:
        raise ValueError("edges must have shape (3, bins + 1)")
    nb = edges.shape[1] - 1
    counts = np.zeros((ns, nb, nb, nb), dtype=np.uint32)
    if ns == 0:
        return counts
    _check(load_library().nm_cdf_counts(device, C.c_void_p(stream or 0), 0, _ptr(pos), _ptr(box), n, ns,
                                        _ptr(edges), nb, _ptr(counts)))
    return counts


# ----------------------------------------------------------------------------- a-13 native text
def format_thrm(vals17):
    vals17 = _f64(vals17, (17,))
    L = load_library()
    n = L.nm_format_thrm(_ptr(vals17), None, 0)
    buf = C.create_string_buffer(n + 1)
    L.nm_format_thrm(_ptr(vals17), buf, n + 1)
    return buf.raw[:n]


def format_traj(natoms, box, x):
    x = _f64(x, (3 * natoms,))
    L = load_library()
    n = L.nm_format_traj(natoms, float(box), _ptr(x), None, 0)
    buf = C.create_string_buffer(n + 1)
    L.nm_format_traj(natoms, float(box), _ptr(x), buf, n + 1)
    return buf.raw[:n]


def _path_array(paths):
    arr = (C.c_char_p * len(paths))()
    for k, p in enumerate(paths):
        arr[k] = None if p is None else os.fsencode(p)
    return arr


def append_traj_batch(natoms, box, x, paths, nthreads=1, parse_back=False):
    """append one write_traj record per replica to paths[k] (None: format only); with parse_back also returns the float32
    (pos, box) lammps_parse.py would read from that text"""
    box = _f64(box)
    nrep = box.size
    x = _f64(x, (nrep, 3 * natoms))
    pos = np.empty((nrep, natoms, 3), dtype=np.float32) if parse_back else None
    bx = np.empty(nrep, dtype=np.float32) if parse_back else None
    n = load_library().nm_append_traj_batch(nrep, natoms, _ptr(box), _ptr(x), _path_array(paths), int(nthreads), _ptr(pos), _ptr(bx))
    if n < 0:
        raise NmError(int(n), load_library().nm_last_error().decode(errors="replace"))
    return pos, bx


def append_thrm_batch(vals17, paths, parse_back=False):
    vals17 = _f64(vals17)
    nrep = vals17.shape[0]
    if vals17.shape != (nrep, 17):
        raise ValueError("vals17 must be (nrep, 17)")
    out = np.empty((nrep, 17), dtype=np.float32) if parse_back else None
    n = load_library().nm_append_thrm_batch(nrep, _ptr(vals17), _path_array(paths), _ptr(out))
    if n < 0:
        raise NmError(int(n), load_library().nm_last_error().decode(errors="replace"))
    return out


def measure_fma_peak(device=0, precision=64):
    """sustained FMA rate of the device (FLOP/s, 2 per FMA) -- the roofline denominator of the force kernel"""
    fl, ms = C.c_double(), C.c_double()
    _check(load_library().nm_measure_fma_peak(device, precision, C.addressof(fl), C.addressof(ms)))
    return fl.value, ms.value
