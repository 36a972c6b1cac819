"""Host driver of the replica-exchange NPT Monte Carlo run: the B200 drop-in for
/root/reference/scripts/lammps_remcmc.py.

Same command line (every flag of lammps_remcmc.py:25-85 is accepted, the Dask/joblib/PBS ones are
kept for compatibility and ignored: the replica grid runs on the GPUs), same element dictionaries
(:873-893), same (P, T) grids (:895-897), same output files: <prefix>.virial.trgt.npy,
<prefix>.temp.trgt.npy, <prefix>.thrm, <prefix>.traj (text formats of :176-256, consumed unchanged
by lammps_parse.py) and <prefix>.rstrt.%04d.npy restart dumps (:821-828). The per-replica LAMMPS
work (:394-691) happens inside the CUDA engine (neuralmelting_b200.engine); no LAMMPS, Dask, numba
or CPU fallback is involved.

Multi-GPU: one process per GPU (torchrun); pressure row u lives on rank u mod G (all temperatures of a row together:
exchanges never cross rows, so every rank decides its own swaps). (pe + ke, vol) of every replica is all-gathered over
NCCL asynchronously, off the critical path, for the job-wide log. Output is streamed: every recorded cycle is appended
to per-replica files by a writer thread (bounded memory) and consolidated by concatenation at the end.
"""
import argparse
import os
import queue
import shutil
import sys
import threading

import numpy as np

from . import engine as nm

# material tables (lammps_remcmc.py:873-893)
UNITS = {"Ti": "metal", "Al": "metal", "Ni": "metal", "Cu": "metal", "LJ": "lj"}
LAT = {"Ti": ("bcc", 2.951), "Al": ("fcc", 4.046), "Ni": ("fcc", 3.524), "Cu": ("fcc", 3.615), "LJ": ("fcc", 1.122)}
MASS = {"Ti": 47.867, "Al": 29.982, "Ni": 58.693, "Cu": 63.546, "LJ": 1.0}
TIMESTEP = {"real": 4.0, "metal": 0.00390625, "lj": 0.00390625}
SEED = 256          # lammps_remcmc.py:851
RC = 2.5            # pair_style lj/cut 2.5 (lammps_remcmc.py:365)

# (short, long, kwargs) -- the flag surface of lammps_remcmc.py:25-85
_FLAGS = [
    ("-v", "--verbose", dict(action="store_true", help="verbose output mode")),
    ("-r", "--restart", dict(action="store_true", help="restart run mode")),
    ("-p", "--parallel", dict(action="store_true", help="parallel run mode (accepted, ignored: GPU engine)")),
    ("-c", "--client", dict(action="store_true", help="dask client run mode (accepted, ignored)")),
    ("-d", "--distributed", dict(action="store_true", help="distributed run mode (accepted, ignored)")),
    ("-is", "--interpolate_states", dict(action="store_true", help="interpolate initial states")),
    ("-bm", "--bulk_move", dict(action="store_true", help="bulk position monte carlo moves")),
    ("-rd", "--restart_dump", dict(type=int, default=128, help="restart dump frequency")),
    ("-rn", "--restart_name", dict(type=str, default="remcmc_init", help="restart dump simulation name")),
    ("-rs", "--restart_step", dict(type=int, default=1024, help="restart dump start step")),
    ("-q", "--queue", dict(type=str, default="jobqueue", help="job submission queue (ignored)")),
    ("-a", "--allocation", dict(type=str, default="startup", help="job submission allocation (ignored)")),
    ("-nn", "--nodes", dict(type=int, default=1, help="job node count (ignored)")),
    ("-np", "--procs_per_node", dict(type=int, default=20, help="number of processors per node (ignored)")),
    ("-w", "--walltime", dict(type=int, default=72, help="job walltime (ignored)")),
    ("-m", "--memory", dict(type=int, default=32, help="job memory (ignored)")),
    ("-nw", "--workers", dict(type=int, default=20, help="job worker count (ignored)")),
    ("-nt", "--threads", dict(type=int, default=1, help="threads per worker (host formatter threads)")),
    ("-mt", "--method", dict(type=str, default="fork", help="parallelization method (ignored)")),
    ("-n", "--name", dict(type=str, default="remcmc_init", help="simulation name")),
    ("-e", "--element", dict(type=str, default="LJ", help="simulation element")),
    ("-ss", "--supercell_size", dict(type=int, default=5, help="simulation supercell size")),
    ("-pn", "--pressure_number", dict(type=int, default=16, help="number of pressures")),
    ("-pr", "--pressure_range", dict(type=float, nargs=2, default=[1, 8], help="pressure range (low and high)")),
    ("-tn", "--temperature_number", dict(type=int, default=16, help="number of temperatures")),
    ("-tr", "--temperature_range", dict(type=float, nargs=2, default=[0.25, 2.5], help="temperature range (low and high)")),
    ("-sc", "--sample_cutoff", dict(type=int, default=0, help="sample recording cutoff")),
    ("-sn", "--sample_number", dict(type=int, default=1024, help="number of samples to generate")),
    ("-sm", "--sample_mod", dict(type=int, default=128, help="sample collection frequency")),
    ("-pm", "--position_move", dict(type=float, default=0.125, help="position monte carlo move probability")),
    ("-vm", "--volume_move", dict(type=float, default=0.125, help="volume monte carlo move probability")),
    ("-ts", "--timesteps", dict(type=int, default=8, help="hamiltonian monte carlo timesteps")),
    ("-dx", "--pos_displace", dict(type=float, default=0.03125, help="position displacement (lattice proportion)")),
    ("-dv", "--vol_displace", dict(type=float, default=0.03125, help="logarithmic volume displacement")),
]
# not in the reference: also write what lammps_parse.py would produce (.pos/.box/.natoms/.<thermo>.npy) during the run
_EXTRA_FLAGS = [
    ("-dn", "--direct_npy", dict(action="store_true", help="also emit the parser's .npy files directly (same 5-digit rounding)")),
]


def build_parser():
    parser = argparse.ArgumentParser(description="replica-exchange NPT Monte Carlo on B200 (lammps_remcmc.py drop-in)")
    for short, long_, kw in _FLAGS + _EXTRA_FLAGS:
        parser.add_argument(short, long_, **kw)
    return parser


def parse_args(argv=None):
    return build_parser().parse_args(argv)


def grids(lp, hp, npn, lt, ht, ntn):
    """P and T exactly as lammps_remcmc.py:895-897 builds them (float32 linspace)"""
    return np.linspace(lp, hp, npn, dtype=np.float32), np.linspace(lt, ht, ntn, dtype=np.float32)


def init_constants(P, T, units="lj"):
    """(et, pf) of init_constant for every slot k = i*NT + j (lammps_remcmc.py:114-141); float32 grid values
    promoted to float64 (the numpy-1 scalar rules the script was written for)"""
    if units != "lj":
        raise NotImplementedError("only the 'lj' unit branch of init_constant is on the hot path (MEAM potentials are not shipped)")
    Pd, Td = P.astype(np.float64), T.astype(np.float64)
    et = np.tile(1.0 * Td, P.size)
    pf = np.repeat(Pd, T.size) / et
    return et, pf


def text6(v):
    """the '%f' round trip every number takes on its way into a LAMMPS command string"""
    return float("%f" % v)


def fcc_fractional(sz):
    """fcc sites of an sz^3 supercell in box fractions, in LAMMPS create_atoms order (unit cells x-fastest
    inside z-slowest loops, 4 basis atoms per cell; lattice/create_atoms of the deck at lammps_remcmc.py:340-343)"""
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    cells = np.array([[i, j, k] for k in range(sz) for j in range(sz) for i in range(sz)], dtype=np.float64)
    return (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) / sz


def relaxed_boxes(pressures, sz, device=0):
    """zero-temperature pressure-relaxed fcc box side for each pressure: what 'fix box/relax iso P' + 'minimize'
    (lammps_remcmc.py:402-405) converge to for the perfect lattice, i.e. the root of W(L)/(3 L^3) = P, found by
    bracketing + bisection with the GPU evaluation (the lattice stays on its sites by symmetry)."""
    frac = fcc_fractional(sz)
    n = frac.shape[0]
    pressures = np.asarray(pressures, dtype=np.float64)
    a0 = (4.0 / LAT["LJ"][1]) ** (1.0 / 3.0)
    nscan = 96
    # box side brackets: density from ~0.8 to ~1.45
    lo, hi = np.full(pressures.size, sz * a0 * 0.91), np.full(pressures.size, sz * a0 * 1.12)
    with nm.Engine(natoms=n, n_rep=nscan, nt=nscan, device=device, mod=0) as eng:
        def virial_pressure(boxes):
            out = np.empty(boxes.size)
            for s0 in range(0, boxes.size, nscan):
                chunk = boxes[s0:s0 + nscan]
                pad = np.concatenate([chunk, np.full(nscan - chunk.size, chunk[-1])])
                x = (frac[None, :, :] * pad[:, None, None]).reshape(nscan, -1)
                eng.set_state(x=x, box=pad)
                _, w, _, _ = eng.eval(want_forces=False)
                out[s0:s0 + chunk.size] = (w / (3.0 * pad ** 3))[:chunk.size]
            return out
        # coarse scan to bracket the first crossing from the dense side, then bisection
        scan = np.linspace(lo[0], hi[0], nscan)
        pscan = virial_pressure(scan)
        for k, p in enumerate(pressures):
            idx = np.where((pscan[:-1] >= p) & (pscan[1:] < p))[0]
            if idx.size == 0:
                raise RuntimeError("no zero-temperature fcc state at pressure %g inside the scanned densities" % p)
            lo[k], hi[k] = scan[idx[0]], scan[idx[0] + 1]
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            pm = virial_pressure(mid)
            dense = pm >= pressures
            lo = np.where(dense, mid, lo)
            hi = np.where(dense, hi, mid)
    return 0.5 * (lo + hi)


def init_samples(P, T, sz, dx, seed, slots=None, interpolate=False, device=0):
    """init_sample for the given global slots k = i*NT + j (default: all; lammps_remcmc.py:394-456): relaxed fcc at P[i], then
    'displace_atoms all random' by +-DX*LAT (text-rounded) per axis; velocities zero (with -is the caller draws them on
    the GPU: nm_velocity_create). With -is the log-volume is raised by 0.75 (j+1)/NT (:412); the reference's subsequent
    'run 1024' has no integrator defined and leaves positions unchanged. The displacement stream of a slot is keyed on
    (seed, k): the start configuration does not depend on how the grid is spread over GPUs."""
    frac = fcc_fractional(sz)
    n = frac.shape[0]
    nt = T.size
    slots = np.arange(P.size * nt) if slots is None else np.asarray(slots)
    rows = np.unique(slots // nt)
    boxes_p = dict(zip(rows.tolist(), relaxed_boxes(P.astype(np.float64)[rows], sz, device=device)))
    x = np.empty((slots.size, 3 * n))
    box = np.empty(slots.size)
    d = text6(dx * LAT["LJ"][1])
    for q, k in enumerate(slots.tolist()):
        i, j = divmod(k, nt)
        L = boxes_p[i]
        rng = np.random.default_rng([int(seed), 1, k])
        pos = frac * L + d * 2.0 * (rng.random((n, 3)) - 0.5)
        pos -= np.floor(pos / L) * L
        if interpolate:
            Lnew = np.cbrt(np.exp(np.log(L ** 3) + 0.75 * (j + 1) / nt))
            pos *= Lnew / L
            L = text6(Lnew)
        x[q] = pos.reshape(-1)
        box[q] = L
    return x, np.zeros_like(x), box


# ----------------------------------------------------------------------------- output files (a-13)
def file_prefix(name, el):
    return os.path.join(os.getcwd(), "%s.%s.%s.lammps" % (name, el.lower(), LAT[el][0]))


def header_text(args, P_i, T_j, nsmpl, cutoff, mod, dt):
    """init_header (lammps_remcmc.py:176-209)"""
    rows = [("nsmpl", "%d" % nsmpl), ("cutoff", "%d" % cutoff), ("mod", "%d" % mod), ("nswps", "%d" % (nsmpl * mod)),
            ("ppos", "%f" % args.position_move), ("pvol", "%f" % args.volume_move),
            ("phmc", "%f" % (1 - args.position_move - args.volume_move)), ("nstps", "%d" % args.timesteps),
            ("seed", "%d" % SEED)]
    mat = [("element", "%s" % args.element), ("units", "%s" % UNITS[args.element]), ("lattice", "%s" % LAT[args.element][0]),
           ("latpar", "%f" % LAT[args.element][1]), ("size", "%d" % args.supercell_size), ("mass", "%f" % MASS[args.element]),
           ("press", "%f" % P_i), ("temp", "%f" % T_j), ("dx", "%f" % args.pos_displace), ("dv", "%f" % args.vol_displace),
           ("dt", "%f" % dt)]
    bar = "# ---------------------\n"
    out = [bar, "# simulation parameters\n", bar]
    out += ["# %-10s%s\n" % (k + ":", v) for k, v in rows]
    out += [bar, "# material properties\n", bar]
    out += ["# %-10s%s\n" % (k + ":", v) for k, v in mat]
    wide = "# " + "-" * 95 + "\n"
    out += [wide, "# | tmp | pe | ke | vir | vol | dx | dv | dt | ntp | nap | ntv | nav | nth | nah | ap | av | ah |\n", wide]
    return "".join(out)


_THRM_COLS = [0, 1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17]   # thermo record minus 'box'


def thrm_line(th_row):
    """write_thrm (lammps_remcmc.py:235-245): temp pe ke virial vol dx dv dt ntp nap ntv nav nth nah ap av ah"""
    return nm.format_thrm(np.ascontiguousarray(th_row[_THRM_COLS]))


def traj_records(natoms, box, x, nthreads=1):
    """write_traj (lammps_remcmc.py:248-256) for a batch of replicas: list of bytes, one record per replica (formatted once
    into a buffer sized by the per-record upper bound)"""
    import ctypes as C
    L = nm.load_library()
    nrep = box.size
    x = np.ascontiguousarray(x, dtype=np.float64)
    box = np.ascontiguousarray(box, dtype=np.float64)
    off = np.zeros(nrep + 1, dtype=np.int64)
    cap = nrep * (96 + natoms * (3 * 13 + 1)) + 1
    buf = C.create_string_buffer(cap)
    total = L.nm_format_traj_batch(nrep, natoms, box.ctypes.data, x.ctypes.data, buf, cap, off.ctypes.data, nthreads)
    if total < 0:
        raise nm.NmError(int(total), L.nm_last_error().decode())
    raw = buf.raw
    return [raw[off[k]:off[k + 1]] for k in range(nrep)]


THERMO_NAMES = ("temp", "pe", "ke", "virial", "vol", "dx", "dv", "dt", "ntp", "nap", "ntv", "nav", "nth", "nah", "ap", "av", "ah")


def replica_prefix(name, el, i, j):
    """file_prefix(i, j) of the reference (lammps_remcmc.py:148-151): the per-replica files written during the run"""
    return os.path.join(os.getcwd(), "%s.%s.%s.%02d.%02d.lammps" % (name, el.lower(), LAT[el][0], i, j))


class StreamWriter:
    """write_outputs (lammps_remcmc.py:259-286) as a pipeline stage: the main loop hands over one recorded cycle (thermo rows,
    boxes, positions of the local slots) and goes on with the next cycle; a writer thread formats it once (native, threaded)
    and appends it to the per-replica .thrm / .traj files, as the reference does, so a crash loses nothing that was recorded
    and host memory stays flat (at most `depth` cycles in flight). With direct_npy the same pass fills memory-mapped
    .pos/.box/.natoms/.<thermo>.npy files with the values lammps_parse.py:37-103 would read back from the text."""

    def __init__(self, args, pref, slots, npn, ntn, natoms, nrec, headers, nthreads=1, direct_npy=False, create=True, depth=2):
        self.slots = np.asarray(slots)
        self.natoms, self.nthreads, self.nrec = natoms, max(1, nthreads), nrec
        ij = [divmod(int(k), ntn) for k in self.slots]
        self.thrm_paths = [replica_prefix(args.name, args.element, i, j) + ".thrm" for i, j in ij]
        self.traj_paths = [replica_prefix(args.name, args.element, i, j) + ".traj" for i, j in ij]
        for p_thrm, p_traj, head in zip(self.thrm_paths, self.traj_paths, headers):
            with open(p_thrm, "wb") as fh:          # init_output + init_header: start clean (the stale-.traj bug is dropped)
                fh.write(head)
            open(p_traj, "wb").close()
        self.npy = None
        if direct_npy and nrec > 0:
            from numpy.lib.format import open_memmap
            def mm(suffix, dtype, shape):
                return open_memmap(pref + suffix, mode="w+", dtype=dtype, shape=shape) if create else open_memmap(pref + suffix, mode="r+")
            self.npy = dict(pos=mm(".pos.npy", np.float32, (npn, ntn, nrec, natoms, 3)), box=mm(".box.npy", np.float32, (npn * ntn * nrec,)),
                            natoms=mm(".natoms.npy", np.uint16, (npn, ntn, nrec)))
            for name in THERMO_NAMES:
                self.npy[name] = mm(".%s.npy" % name, np.float32, (npn, ntn, nrec))
            self.ntn = ntn
        self.count = 0
        self.err = None
        self.q = queue.Queue(maxsize=depth)
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            if self.err is not None:
                continue
            try:
                self._write(*item)
            except Exception as e:          # surfaced by the next put() / close()
                self.err = e

    def _write(self, s, th17, box, x):
        want = self.npy is not None
        parsed = nm.append_thrm_batch(th17, self.thrm_paths, parse_back=want)
        pos, bx = nm.append_traj_batch(self.natoms, box, x, self.traj_paths, nthreads=self.nthreads, parse_back=want)
        if want:
            i, j = np.divmod(self.slots, self.ntn)
            self.npy["pos"][i, j, s] = pos
            self.npy["box"][self.slots * self.nrec + s] = bx
            self.npy["natoms"][i, j, s] = self.natoms
            for c, name in enumerate(THERMO_NAMES):
                self.npy[name][i, j, s] = parsed[:, c]

    def put(self, thermo, box, x):
        """thermo: (nloc, 18) records of the cycle; box (nloc,), x (nloc, 3 natoms). The arrays are owned by the writer afterwards."""
        if self.err is not None:
            raise self.err
        self.q.put((self.count, np.ascontiguousarray(thermo[:, _THRM_COLS]), box, x))
        self.count += 1

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.npy is not None:
            for a in self.npy.values():
                a.flush()
            self.npy = None
        if self.err is not None:
            raise self.err


def consolidate_outputs(args, pref, npn, ntn):
    """consolidate_outputs (lammps_remcmc.py:289-316): the per-replica files concatenated in pressure-major, temperature order
    (streamed, no whole-file reads), then removed"""
    for ext in (".thrm", ".traj"):
        with open(pref + ext, "wb") as out:
            for i in range(npn):
                for j in range(ntn):
                    with open(replica_prefix(args.name, args.element, i, j) + ext, "rb") as fh:
                        shutil.copyfileobj(fh, out, 1 << 22)
    for i in range(npn):
        for j in range(ntn):
            for ext in (".thrm", ".traj"):
                os.remove(replica_prefix(args.name, args.element, i, j) + ext)


# ----------------------------------------------------------------------------- restart files (N4)
def dump_restart(path, natoms, state, thermo):
    """dump_samples_restart (lammps_remcmc.py:821-828): object array (NS, 21) in the reference's slot layout"""
    ns = state["box"].size
    rows = []
    for k in range(ns):
        th = thermo[k]
        rows.append([natoms, state["x"][k].copy(), state["v"][k].copy(), th[0], th[1], th[2], th[3], state["box"][k],
                     state["box"][k] ** 3, state["dx"][k], state["dv"][k], state["dt"][k]] + [0.0] * 9)
    arr = np.empty((ns, 21), dtype=object)
    for k, row in enumerate(rows):
        for c, v in enumerate(row):
            arr[k, c] = v
    np.save(path, arr)


def load_restart(path):
    """load_samples_restart (lammps_remcmc.py:810-818); numpy >= 1.16.3 needs allow_pickle"""
    arr = np.load(path, allow_pickle=True)
    x = np.stack([np.asarray(r[1], dtype=np.float64) for r in arr])
    v = np.stack([np.asarray(r[2], dtype=np.float64) for r in arr])
    box = np.array([float(r[7]) for r in arr])
    dx, dv, dt = (np.array([float(r[c]) for r in arr]) for c in (9, 10, 11))
    return int(arr[0][0]), x, v, box, dx, dv, dt


# ----------------------------------------------------------------------------- distributed plumbing
class Comm:
    """one process per GPU; torch.distributed (NCCL over NVLink on GPUs, gloo on CPU tests) is plumbing only"""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if not dist.is_initialized():
                # NM_DIST_BACKEND=gloo: several ranks on ONE GPU (tests); NCCL needs one device per rank
                backend = os.environ.get("NM_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
                if backend == "nccl":
                    torch.cuda.set_device(self.local_rank)
                dist.init_process_group(backend=backend)
            self.dist = dist
        self.backend = self.dist.get_backend() if self.dist is not None else None

    def row_shard(self, npn):
        """pressure row u lives on rank u mod G: every rank holds an even mix of low- and high-pressure rows
        and all temperatures of each; exchanges never cross rows (lammps_remcmc.py:782-789). Returns the global rows of this rank."""
        if npn % self.world:
            raise ValueError("pressure_number (%d) must be a multiple of the number of GPUs (%d)" % (npn, self.world))
        return list(range(self.rank, npn, self.world))

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


class TableGather:
    """(pe + ke, vol) of every slot in the job -- what the reference's client holds in STATE -- kept current OFF the
    critical path: every rank packs its slots (nm_exchange_pack) and one asynchronous all-gather (16 bytes per replica,
    NCCL over NVLink) runs beside the next cycle; the swaps themselves are decided locally. latest() returns the most
    recent completed job-wide table in global slot order."""

    def __init__(self, comm, eng, torch):
        self.comm, self.eng, self.torch = comm, eng, torch
        on_gpu = torch.cuda.is_available()
        coll = "cuda" if on_gpu and getattr(comm, "backend", None) != "gloo" else "cpu"      # where the collective runs
        self.stage = [torch.empty((eng.n_rep, 2), dtype=torch.float64, device="cuda") for _ in range(2)] if on_gpu and coll == "cpu" else None
        self.local = [torch.empty((eng.n_rep, 2), dtype=torch.float64, device=coll) for _ in range(2)]
        self.full = [torch.empty((comm.world, eng.n_rep, 2), dtype=torch.float64, device=coll) for _ in range(2)]
        self.work = [None, None]
        self.cur = 0
        nrow = eng.n_rep // eng.nt
        # rank r, local row lr -> global row r + lr * world
        self.order = np.array([(r + lr * comm.world) * eng.nt + j for r in range(comm.world) for lr in range(nrow) for j in range(eng.nt)])

    def start(self):
        b = self.cur = self.cur ^ 1
        if self.work[b] is not None:
            self.work[b].wait()
        if self.stage is not None:              # gloo with a GPU engine: pack on the device, collective on the host
            self.eng.exchange_pack(self.stage[b].data_ptr())
            self.local[b].copy_(self.stage[b])
        else:
            self.eng.exchange_pack(self.local[b].data_ptr())
        if self.comm.world == 1:
            self.full[b][0].copy_(self.local[b], non_blocking=True)
            self.work[b] = None
        else:
            self.work[b] = self.comm.dist.all_gather_into_tensor(self.full[b].view(-1, 2), self.local[b], async_op=True)

    def latest(self):
        b = self.cur
        if self.work[b] is not None:
            self.work[b].wait(); self.work[b] = None
        flat = self.full[b].view(-1, 2).cpu().numpy()
        out = np.empty_like(flat)
        out[self.order] = flat
        return out

    def finish(self):
        for b in (0, 1):
            if self.work[b] is not None:
                self.work[b].wait(); self.work[b] = None


def initial_thermo(eng, natoms, mass, st):
    """thermo record of a freshly uploaded state (what lammps_extract returns after init_sample's last 'run 0'): pe and the
    pair virial from nm_eval, ke / temp from the velocities, total pressure as compute pressure forms it"""
    pe, w, _, _ = eng.eval(want_forces=False)
    ke = 0.5 * mass * (st["v"] ** 2).sum(1)
    dof = 3.0 * natoms - 3.0
    temp = 2.0 * ke / dof
    vol = st["box"] ** 3
    th = np.zeros((eng.n_rep, nm.THERMO_WIDTH))
    th[:, 0], th[:, 1], th[:, 2], th[:, 3], th[:, 4], th[:, 5] = temp, pe, ke, (dof * temp + w) / 3.0 / vol, st["box"], vol
    th[:, 6], th[:, 7], th[:, 8] = st["dx"], st["dv"], st["dt"]
    return th


def run(args, comm=None, log=print):
    """the main loop of lammps_remcmc.py:959-1001 with the per-replica work on the GPU"""
    import torch
    comm = comm or Comm()
    el = args.element
    if UNITS[el] != "lj":
        raise SystemExit("element %s needs a MEAM potential that the reference does not ship; only LJ runs" % el)
    npn, ntn, sz = args.pressure_number, args.temperature_number, args.supercell_size
    nsmpl, cutoff, mod = args.sample_number, args.sample_cutoff, args.sample_mod
    P, T = grids(args.pressure_range[0], args.pressure_range[1], npn, args.temperature_range[0], args.temperature_range[1], ntn)
    dt0 = TIMESTEP[UNITS[el]]
    pref = file_prefix(args.name, el)
    rows = comm.row_shard(npn)
    ns, nloc = npn * ntn, len(rows) * ntn
    natoms = 4 * sz ** 3 if LAT[el][0] == "fcc" else 2 * sz ** 3
    if comm.rank == 0:
        np.save(pref + ".virial.trgt.npy", P)
        np.save(pref + ".temp.trgt.npy", T)
    et, pf = init_constants(P, T)
    temp = np.tile(T.astype(np.float64), npn)
    np.random.seed(SEED)
    device = comm.local_rank if torch.cuda.is_available() else 0
    stream = torch.cuda.current_stream().cuda_stream
    eng = nm.Engine(natoms=natoms, n_rep=nloc, nt=ntn, n_rep_global=ns, rep_offset=comm.rank * ntn, row_stride=comm.world, device=device,
                    nstps=args.timesteps, mod=mod, bulk_move=args.bulk_move, ppos=args.position_move,
                    pvol=args.volume_move, lat_scale=LAT[el][1], mass=MASS[el], rc=RC, seed=SEED, stream=stream)
    gs = eng.global_slots()                       # global slot k = i*NT + j of every local slot
    eng.set_labels(et[gs], pf[gs], temp[gs])
    if args.restart:
        rf = os.path.join(os.getcwd(), "%s.%s.%s.lammps.rstrt.%04d.npy" % (args.restart_name, el.lower(), LAT[el][0], args.restart_step))
        _, x, v, box, dx, dv, dt = load_restart(rf)
        box = np.array([text6(b) for b in box])
        eng.set_state(x=x[gs], v=v[gs], box=box[gs], dx=dx[gs], dv=dv[gs], dt=dt[gs])
        eng.exchange(-1, uniforms=np.random.rand(npn * ntn * (ntn - 1) // 2), want_perm=False)
    else:
        x, v, box = init_samples(P, T, sz, args.pos_displace, SEED, slots=gs, interpolate=args.interpolate_states, device=device)
        box = np.array([text6(b) for b in box])          # init_lammps: 'change_box ... %f'
        eng.set_state(x=x, v=v, box=box, dx=np.full(nloc, args.pos_displace), dv=np.full(nloc, args.vol_displace),
                      dt=np.full(nloc, dt0))
        if args.interpolate_states:
            eng.velocity_create(0)                       # the 'velocity create T[j]' draw stays in STATE (:420-425)
    record = cutoff < nsmpl
    writer = None
    if record:
        headers = [header_text(args, P[k // ntn], T[k % ntn], nsmpl, cutoff, mod, dt0).encode() for k in gs.tolist()]
        if args.direct_npy and comm.world > 1:
            if comm.rank == 0:
                StreamWriter(args, pref, [], npn, ntn, natoms, nsmpl - cutoff, [], direct_npy=True, create=True).close()   # creates the memmaps
            comm.barrier()
        writer = StreamWriter(args, pref, gs, npn, ntn, natoms, nsmpl - cutoff, headers, nthreads=max(1, args.threads),
                              direct_npy=args.direct_npy, create=comm.world == 1)
    gather = TableGather(comm, eng, torch)
    # STEP = -1: dump_samples_restart() before the loop (lammps_remcmc.py:975-976)
    st = eng.get_state()
    _gather_and_dump(comm, pref, 0, natoms, st, initial_thermo(eng, natoms, MASS[el], st), gs, ns)
    swaps_total = 0
    for step in range(nsmpl):
        eng.run_cycle(step)
        th = eng.get_thermo()
        if (step + 1) > cutoff:
            st = eng.get_state(want_v=False)
            writer.put(th, st["box"], st["x"])
        eng.adapt()
        if (step + 1) % args.restart_dump == 0:
            st = eng.get_state()
            _gather_and_dump(comm, pref, step + 1, natoms, st, th, gs, ns)
        if (step + 1) != nsmpl:
            _, swaps = eng.exchange(step)
            swaps_total += swaps
            gather.start()                               # job-wide (pe + ke, vol) table: asynchronous, for the log
            if args.verbose:
                tot = _sum_over_ranks(comm, torch, swaps)
                if comm.rank == 0:
                    log("%d replica exchanges performed" % tot)
    gather.finish()
    if writer is not None:
        writer.close()
        comm.barrier()
        if comm.rank == 0:
            consolidate_outputs(args, pref, npn, ntn)
    counters = eng.counters()
    eng.close()
    return counters, _sum_over_ranks(comm, torch, swaps_total)


def _sum_over_ranks(comm, torch, value):
    if comm.world == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device="cuda" if torch.cuda.is_available() else "cpu")
    comm.dist.all_reduce(t)
    return int(t.item())


def _gather_and_dump(comm, pref, step, natoms, st, th, gs, ns):
    """dump_samples_restart (lammps_remcmc.py:821-828) in global slot order. Multi-rank: every rank writes its part next to the
    target, rank 0 assembles (the restart file is one object array by format)."""
    path = pref + ".rstrt.%04d.npy" % step
    if comm.world == 1:
        dump_restart(path, natoms, st, th)
        return
    part = "%s.part%03d.npz" % (path, comm.rank)
    np.savez(part, gs=gs, th=th, **st)
    comm.barrier()
    if comm.rank == 0:
        full = {k: None for k in ("x", "v", "box", "dx", "dv", "dt")}
        th_full = np.empty((ns, th.shape[1]))
        for r in range(comm.world):
            name = "%s.part%03d.npz" % (path, r)
            with np.load(name) as z:
                for k in full:
                    if full[k] is None:
                        full[k] = np.empty((ns,) + z[k].shape[1:])
                    full[k][z["gs"]] = z[k]
                th_full[z["gs"]] = z["th"]
            os.remove(name)
        dump_restart(path, natoms, full, th_full)
    comm.barrier()


def main(argv=None):
    args = parse_args(argv)
    counters, swaps = run(args)
    if args.verbose and int(os.environ.get("RANK", "0")) == 0:
        print("hmc atom-steps: %d, sweeps: %d, exchanges: %d" % (counters["hmc_atom_steps"], counters["sweeps"], swaps))


if __name__ == "__main__":
    main(sys.argv[1:])
