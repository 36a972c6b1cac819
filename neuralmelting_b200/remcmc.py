"""Host driver of the replica-exchange NPT Monte Carlo run: the B200 drop-in for
/root/reference/scripts/lammps_remcmc.py.

Same command line (every flag of lammps_remcmc.py:25-85 is accepted, the Dask/joblib/PBS ones are
kept for compatibility and ignored: the replica grid runs on the GPUs), same element dictionaries
(:873-893), same (P, T) grids (:895-897), same output files: <prefix>.virial.trgt.npy,
<prefix>.temp.trgt.npy, <prefix>.thrm, <prefix>.traj (text formats of :176-256, consumed unchanged
by lammps_parse.py) and <prefix>.rstrt.%04d.npy restart dumps (:821-828). The per-replica LAMMPS
work (:394-691) happens inside the CUDA engine (neuralmelting_b200.engine); no LAMMPS, Dask, numba
or CPU fallback is involved.

Multi-GPU: one process per GPU (torchrun); each rank owns whole pressure rows. The exchange
all-gathers (pe + ke, vol) of every replica over NCCL and every rank replays the same sweep.
"""
import argparse
import os
import sys

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


def build_parser():
    parser = argparse.ArgumentParser(description="replica-exchange NPT Monte Carlo on B200 (lammps_remcmc.py drop-in)")
    for short, long_, kw in _FLAGS:
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


def init_samples(P, T, sz, dx, rng, interpolate=False, device=0):
    """init_sample for every slot (lammps_remcmc.py:394-456): relaxed fcc at P[i], then 'displace_atoms all random'
    by +-DX*LAT (text-rounded) per axis; velocities zero. With -is the log-volume is raised by 0.75 (j+1)/NT
    (:412); the reference's subsequent 'run 1024' has no integrator defined and leaves positions unchanged."""
    frac = fcc_fractional(sz)
    n = frac.shape[0]
    boxes_p = relaxed_boxes(P.astype(np.float64), sz, device=device)
    nt = T.size
    ns = P.size * nt
    x = np.empty((ns, 3 * n))
    box = np.empty(ns)
    d = text6(dx * LAT["LJ"][1])
    for k in range(ns):
        i, j = divmod(k, nt)
        L = boxes_p[i]
        pos = frac * L + d * 2.0 * (rng.random((n, 3)) - 0.5)
        pos -= np.floor(pos / L) * L
        if interpolate:
            Lnew = np.cbrt(np.exp(np.log(L ** 3) + 0.75 * (j + 1) / nt))
            pos *= Lnew / L
            L = text6(Lnew)
        x[k] = pos.reshape(-1)
        box[k] = L
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
    """write_traj (lammps_remcmc.py:248-256) for a batch of replicas: list of bytes, one record per replica"""
    import ctypes as C
    L = nm.load_library()
    nrep = box.size
    x = np.ascontiguousarray(x, dtype=np.float64)
    box = np.ascontiguousarray(box, dtype=np.float64)
    off = np.zeros(nrep + 1, dtype=np.int64)
    L.nm_format_traj_batch.restype = C.c_int64
    L.nm_format_traj_batch.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]
    total = L.nm_format_traj_batch(nrep, natoms, box.ctypes.data, x.ctypes.data, None, 0, off.ctypes.data, nthreads)
    if total < 0:
        raise nm.NmError(int(total), L.nm_last_error().decode())
    buf = C.create_string_buffer(int(total) + 1)
    L.nm_format_traj_batch(nrep, natoms, box.ctypes.data, x.ctypes.data, buf, int(total) + 1, off.ctypes.data, nthreads)
    raw = buf.raw
    return [raw[off[k]:off[k + 1]] for k in range(nrep)]


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
                backend = "nccl" if torch.cuda.is_available() else "gloo"
                if backend == "nccl":
                    torch.cuda.set_device(self.local_rank)
                dist.init_process_group(backend=backend)
            self.dist = dist

    def row_shard(self, npn):
        """contiguous blocks of whole pressure rows; exchanges never cross rows (lammps_remcmc.py:782-789)"""
        if npn % self.world:
            raise ValueError("pressure_number (%d) must be a multiple of the number of GPUs (%d)" % (npn, self.world))
        per = npn // self.world
        return self.rank * per, per

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


def allgather_table(comm, eng, torch):
    """(pe + ke, vol) of every slot in the job: nm_exchange_pack on each rank + one all-gather (16 bytes per replica)"""
    local = torch.empty((eng.n_rep, 2), dtype=torch.float64, device="cuda")
    eng.exchange_pack(local.data_ptr())
    if comm.world == 1:
        return local
    full = torch.empty((eng.n_rep_global, 2), dtype=torch.float64, device="cuda")
    comm.dist.all_gather_into_tensor(full, local)
    return full


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
    row0, nrow = comm.row_shard(npn)
    ns, nloc, off = npn * ntn, nrow * ntn, row0 * ntn
    natoms = 4 * sz ** 3 if LAT[el][0] == "fcc" else 2 * sz ** 3
    if comm.rank == 0:
        np.save(pref + ".virial.trgt.npy", P)
        np.save(pref + ".temp.trgt.npy", T)
    et, pf = init_constants(P, T)
    temp = np.tile(T.astype(np.float64), npn)
    np.random.seed(SEED)
    rng = np.random.default_rng(SEED)
    device = comm.local_rank if torch.cuda.is_available() else 0
    stream = torch.cuda.current_stream().cuda_stream
    eng = nm.Engine(natoms=natoms, n_rep=nloc, nt=ntn, n_rep_global=ns, rep_offset=off, device=device,
                    nstps=args.timesteps, mod=mod, bulk_move=args.bulk_move, ppos=args.position_move,
                    pvol=args.volume_move, lat_scale=LAT[el][1], mass=MASS[el], rc=RC, seed=SEED, stream=stream)
    sl = slice(off, off + nloc)
    eng.set_labels(et[sl], pf[sl], temp[sl])
    if args.restart:
        rf = os.path.join(os.getcwd(), "%s.%s.%s.lammps.rstrt.%04d.npy" % (args.restart_name, el.lower(), LAT[el][0], args.restart_step))
        _, x, v, box, dx, dv, dt = load_restart(rf)
        box = np.array([text6(b) for b in box])
        eng.set_state(x=x[sl], v=v[sl], box=box[sl], dx=dx[sl], dv=dv[sl], dt=dt[sl])
        table = allgather_table(comm, eng, torch)
        eng.exchange_apply(table.data_ptr(), et, pf, -1, uniforms=np.random.rand(npn * ntn * (ntn - 1) // 2), want_perm=False)
    else:
        x, v, box = init_samples(P[row0:row0 + nrow], T, sz, args.pos_displace, np.random.default_rng(SEED + 1 + comm.rank),
                                 interpolate=args.interpolate_states, device=device)
        box = np.array([text6(b) for b in box])          # init_lammps: 'change_box ... %f'
        eng.set_state(x=x, v=v, box=box, dx=np.full(nloc, args.pos_displace), dv=np.full(nloc, args.vol_displace),
                      dt=np.full(nloc, dt0))
    record = cutoff < nsmpl
    thrm_parts = [[] for _ in range(nloc)]
    traj_parts = [[] for _ in range(nloc)]
    if record:
        for k in range(nloc):
            i, j = divmod(off + k, ntn)
            thrm_parts[k].append(header_text(args, P[i], T[j], nsmpl, cutoff, mod, dt0).encode())
    swaps_total = 0
    for step in range(nsmpl):
        eng.run_cycle(step)
        th = eng.get_thermo()
        if (step + 1) > cutoff:
            st = eng.get_state(want_v=False)
            recs = traj_records(natoms, st["box"], st["x"], nthreads=max(1, args.threads))
            for k in range(nloc):
                thrm_parts[k].append(thrm_line(th[k]))
                traj_parts[k].append(recs[k])
        eng.adapt()
        if (step + 1) % args.restart_dump == 0:
            st = eng.get_state()
            _gather_and_dump(comm, torch, pref, step + 1, natoms, st, th, ns, off, nloc)
        if (step + 1) != nsmpl:
            table = allgather_table(comm, eng, torch)
            _, swaps = eng.exchange_apply(table.data_ptr(), et, pf, step)
            swaps_total += swaps
            if args.verbose and comm.rank == 0:
                log("%d replica exchanges performed" % swaps)
    if record:
        _consolidate(comm, pref, thrm_parts, traj_parts)
    counters = eng.counters()
    eng.close()
    return counters, swaps_total


def _gather_and_dump(comm, torch, pref, step, natoms, st, th, ns, off, nloc):
    if comm.world == 1:
        dump_restart(pref + ".rstrt.%04d.npy" % step, natoms, st, th)
        return
    objs = [None] * comm.world if comm.rank == 0 else None
    comm.dist.gather_object((st, th), objs, dst=0)
    if comm.rank == 0:
        full = {k: np.concatenate([o[0][k] for o in objs]) for k in ("x", "v", "box", "dx", "dv", "dt")}
        dump_restart(pref + ".rstrt.%04d.npy" % step, natoms, full, np.concatenate([o[1] for o in objs]))


def _consolidate(comm, pref, thrm_parts, traj_parts):
    """consolidate_outputs (lammps_remcmc.py:289-316): pressure-major, temperature, sample order"""
    if comm.world > 1:
        objs = [None] * comm.world if comm.rank == 0 else None
        comm.dist.gather_object((thrm_parts, traj_parts), objs, dst=0)
        if comm.rank != 0:
            return
        thrm_parts = [p for o in objs for p in o[0]]
        traj_parts = [p for o in objs for p in o[1]]
    with open(pref + ".thrm", "wb") as fh:
        for parts in thrm_parts:
            fh.write(b"".join(parts))
    with open(pref + ".traj", "wb") as fh:
        for parts in traj_parts:
            fh.write(b"".join(parts))


def main(argv=None):
    args = parse_args(argv)
    counters, swaps = run(args)
    if args.verbose and int(os.environ.get("RANK", "0")) == 0:
        print("hmc atom-steps: %d, sweeps: %d, exchanges: %d" % (counters["hmc_atom_steps"], counters["sweeps"], swaps))


if __name__ == "__main__":
    main(sys.argv[1:])
