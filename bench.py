#!/usr/bin/env python
"""bench.py -- throughput of the replica-exchange NPT Monte Carlo hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1|c4] [--impl reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

A "step" is one collection cycle of the whole local (P, T) grid: MOD Monte Carlo moves per replica
(gen_samples, lammps_remcmc.py:694-719) + thermo read-back + step-size adaptation (gen_mc_params, :748-770)
+ replica exchange (:776-803; all-gather of (pe+ke, vol) over NCCL when N > 1).
Metric (BASELINE.json): HMC atom-steps/s = sum over HMC moves of natoms*NSTPS / time; MC sweeps/s rides along.
Workload at N=1: BASELINE.json configs[1] (C2: 500-atom LJ, 16 x 16 grid, default move mix). N > 1 is weak scaling:
every GPU holds 16 pressure rows x 16 temperatures of a (16 N) x 16 grid, pressure row u on rank u mod N ("c3": 4 rows x
32 T of 4000 atoms per GPU, i.e. exactly BASELINE configs[2] at N=8). Swaps are decided rank-locally (exchanges never cross
pressure rows); the NCCL all-gather of (pe + ke, vol) runs asynchronously beside the next cycle.
The same JSON line carries a "workloads" block with short legs of the other BASELINE configurations (c3, c4, c5) and a
"parity_gate": after the timed region a handful of the resident configurations is re-evaluated and compared with the CPU
oracle at 1e-10 (energy, virial, forces; in-cutoff pair counts exact).

--impl reference times the CPU restatement of the reference path (oracle/, "port": LAMMPS and Dask are not installable,
see BASELINE.md) farmed over all host cores, one replica per task like Dask does, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)



def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the workload's dominant kernel, from the committed
    `ncu --set full` capture of the current round: profiles/ncu_traffic.json = {workload: {"bytes": B, "capture": file}}"""
    try:
        ent = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(workload)
        return (float(ent["bytes"]), ent["capture"]) if ent else (None, None)
    except Exception:
        return None, None


WORKLOADS = {
    # name: (supercell, pressure rows per GPU, temperatures, bulk_move, ppos, pvol, mod, description)
    "c1": (4, 8, 8, True, 0.125, 0.125, 128, "C1: LJ fcc 4x4x4 (256 atoms), 8x8 P-T grid per GPU, default move mix"),
    "c2": (5, 16, 16, True, 0.125, 0.125, 128, "C2: LJ fcc 5x5x5 (500 atoms), 16x16 P-T grid per GPU, default PMC/VMC/HMC mix (-pm .125 -vm .125 -ts 8 -sm 128 -bm as run.sh)"),
    "c3": (10, 4, 32, True, 0.125, 0.125, 128, "C3: LJ fcc 10x10x10 (4000 atoms), 4 pressure rows x 32 T per GPU (the 32x32 grid at 8 GPUs), default move mix"),
    "c4": (10, 4, 32, False, 0.75, 0.125, 1, "C4: LJ 4000 atoms, PMC-heavy single-atom moves (-pm .75 -vm .125 -sm 1, no -bm), exchange every sweep, 4 rows x 32 T per GPU"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", type=str, default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--rdf-samples", type=int, default=1024, help="c5: samples per step (N = 4000, SBINS = 64)")
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--equil", type=int, default=32, help="untimed equilibration cycles before the warm-up (the step sizes adapt for ~25 cycles: the cost of a cycle rises by 60 %% until the acceptances settle at 0.5)")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32], help="pair arithmetic: 64 (headline) or the FP32 mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the short c3 / c4 / c5 legs of the 'workloads' block")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    return ap.parse_args()


def grid_for(world, wl):
    from neuralmelting_b200 import remcmc
    sz, rows, nt, bulk, ppos, pvol, mod, desc = WORKLOADS[wl]
    npn = rows * world
    P, T = remcmc.grids(1.0, 8.0, npn, 0.25, 2.5, nt)
    return sz, rows, nt, npn, P, T, bulk, ppos, pvol, mod, desc


def initial_states(P, T, sz, slots, device):
    """the reference's init_sample (pressure-relaxed fcc + random displacement keyed on the global slot) for the given global slots"""
    from neuralmelting_b200 import remcmc
    x, v, box = remcmc.init_samples(P, T, sz, 0.03125, remcmc.SEED, slots=slots, device=device)
    box = np.array([remcmc.text6(b) for b in box])
    return x, v, box


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.proc = None
        self.device = device
        if os.environ.get("NM_BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", os.environ.get("NM_BENCH_SAMPLER_MS", "100")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def mc_config(desc, natoms, nloc, npn, nt, mod, world, rows):
    """the `config` object of an MC workload line: the same for the b200 arm and the reference arm (the driver compares them)"""
    return {"workload": desc, "natoms": natoms, "replicas_per_gpu": nloc, "grid": [npn, nt], "moves_per_cycle": mod, "hmc_steps": 8,
            "l2": "inputs larger than L2 (per-GPU state + neighbour lists of %d replicas > 126 MB)" % nloc if nloc * natoms > 60000 else "working set fits L2; no flush (compute-bound on-chip kernel)",
            "parallelism": "replica grid sharded by pressure row, row u on rank u mod %d (%d row(s)/GPU); swaps decided rank-locally, (pe+ke, vol) all-gathered asynchronously" % (world, rows)}


def flops_from(ct):
    """algorithmic FLOPs: 24 per in-cutoff pair (force-only), 30 (force+energy+virial), 2x13 per single-atom
    dE neighbour, 18 per atom-step of the integrator"""
    return 24.0 * ct["pairs_force"] + 30.0 * ct["pairs_full"] + 13.0 * ct["pairs_delta"] + 18.0 * ct["hmc_atom_steps"]


def parity_gate(eng, natoms, nloc, ncheck=6, tol=1e-10):
    """correctness gate on what the timed region left resident: re-evaluate every configuration on the GPU (nm_eval), pick
    ncheck slots across the temperature range and compare energy, virial, forces and the in-cutoff pair count with the CPU
    oracle (list-based lj/cut restatement) on the same positions"""
    from oracle import oracle as orc
    st = eng.get_state(want_v=False)
    pe, w, f, npairs = eng.eval(want_forces=True)
    pick = np.unique(np.linspace(0, nloc - 1, ncheck).round().astype(int))
    worst = {"pe": 0.0, "w": 0.0, "f": 0.0}
    pairs_exact = True
    for k in pick:
        pe_o, w_o, f_o, np_o = orc.lj_eval_list(st["x"][k], float(st["box"][k]))
        worst["pe"] = max(worst["pe"], abs(pe[k] - pe_o) / abs(pe_o))
        worst["w"] = max(worst["w"], abs(w[k] - w_o) / max(abs(w_o), 1e-300))
        worst["f"] = max(worst["f"], float(np.abs(f[k].reshape(-1) - np.asarray(f_o).reshape(-1)).max() / np.abs(f_o).max()))
        pairs_exact = pairs_exact and int(npairs[k]) == int(np_o)
    # the virial of a near-equilibrium configuration is a small difference of large terms: it is compared relative to sum |r.f|
    ok = worst["pe"] <= tol and worst["f"] <= tol and pairs_exact
    return {"replicas_checked": int(pick.size), "tol": tol, "max_rel_err_pe": worst["pe"], "max_rel_err_virial": worst["w"],
            "max_rel_err_forces": worst["f"], "pair_counts_exact": bool(pairs_exact), "pass": bool(ok),
            "against": "oracle/nm_oracle.c lj_eval_list on the positions resident after the timed region"}


def measure_mc(args, comm, torch, workload, steps, warmup, equil, with_e2e=True, with_gate=True):
    """one Monte Carlo workload on this rank's shard: returns the rank-0 dict (None elsewhere) and the pieces the CPU baseline needs"""
    from neuralmelting_b200 import engine as nm
    from neuralmelting_b200 import remcmc
    dev = comm.local_rank
    sz, rows, nt, npn, P, T, bulk, ppos, pvol, mod, desc = grid_for(comm.world, workload)
    natoms = 4 * sz ** 3
    nloc, ns = rows * nt, npn * nt
    et, pf = remcmc.init_constants(P, T)
    temp = np.tile(T.astype(np.float64), npn)
    stream = torch.cuda.current_stream().cuda_stream
    eng = nm.Engine(natoms=natoms, n_rep=nloc, nt=nt, n_rep_global=ns, rep_offset=comm.rank * nt, row_stride=comm.world, device=dev,
                    mod=mod, bulk_move=bulk, ppos=ppos, pvol=pvol, seed=remcmc.SEED, stream=stream, precision=args.precision)
    gs = eng.global_slots()
    x, v, box = initial_states(P, T, sz, gs, dev)
    eng.set_labels(et[gs], pf[gs], temp[gs])
    eng.set_state(x=x, v=v, box=box, dx=np.full(nloc, 0.03125), dv=np.full(nloc, 0.03125), dt=np.full(nloc, 0.00390625))
    gather = remcmc.TableGather(comm, eng, torch)

    def step(cyc, kernel_events=None):
        if kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        eng.run_cycle(cyc)
        if kernel_events is not None:
            e1.record(); kernel_events.append((e0, e1))
        th = eng.get_thermo()                       # the step's result (D2H, 18 doubles per replica)
        eng.adapt()
        eng.exchange(cyc, want_perm=False)          # rank-local sweep of the local pressure rows (asynchronous)
        gather.start()                              # job-wide (pe + ke, vol) table: asynchronous all-gather, off the critical path
        return th

    cyc = 0
    for _ in range(equil):
        step(cyc); cyc += 1
    sampler = ClockSampler(dev) if comm.rank == 0 else None       # covers warm-up + timed region; idle samples are filtered
    for _ in range(max(3, warmup)):
        step(cyc); cyc += 1
    eng.synchronize()
    # ---------------- timed region: device-resident state, CUDA events on the launching stream
    eng.reset_counters()
    launches0 = eng.launch_count()
    kev = []
    comm.barrier(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step(cyc, kev); cyc += 1
    t1.record()
    comm.barrier(); torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = t0.elapsed_time(t1)
    ct = eng.counters()
    launches = eng.launch_count() - launches0
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev)
    table = gather.latest()                         # the job-wide table of the last exchange (checked below)
    if os.environ.get("NM_BENCH_DIAG") and comm.rank == 0:      # development aid: cost of every chain in the last timed cycle
        rc = eng.replica_counters().astype(float); ci = {k: i for i, k in enumerate(nm.COUNTER_COLS)}
        th = eng.get_thermo()
        order = np.argsort(-rc[:, ci["clk_total"]])
        print("slot  T  rho  Mclk  builds outer evals helped  listed/eval  kclk/eval  Mclk_build", file=sys.stderr)
        for k in order:
            r = rc[k]
            print("%4d %.2f %.3f %6.1f %4d %3d %4d %4d %7.0f %6.1f %6.1f" % (k, th[k, 0], natoms / th[k, 5], r[ci["clk_total"]] / 1e6, r[ci["list_builds"]], r[ci["outer_builds"]],
                  r[ci["force_evals"]], r[ci["helped_evals"]], r[ci["list_pairs"]] / max(1, r[ci["force_evals"]]), r[ci["clk_eval"]] / max(1, r[ci["force_evals"]]) / 1e3,
                  r[ci["clk_build"]] / 1e6), file=sys.stderr)
        print("sum of chain Mclk %.0f  / nsm(148) = %.1f" % (rc[:, ci["clk_total"]].sum() / 1e6, rc[:, ci["clk_total"]].sum() / 1e6 / 148), file=sys.stderr)
    gate = parity_gate(eng, natoms, nloc) if (with_gate and comm.rank == 0 and args.precision == 64) else None
    # ---------------- end-to-end through the public API with HOST buffers (pinned): state in, cycle, state + thermo out
    e2e_ms, atom_steps_e2e_local, sweeps_e2e_local, e2e_steps = 0.0, 0.0, 0.0, 0
    h2d = (2 * 3 * natoms + 4) * 8 * nloc
    d2h = (2 * 3 * natoms + 4 + 18) * 8 * nloc
    if with_e2e:
        hx = torch.empty((nloc, 3 * natoms), dtype=torch.float64).pin_memory()
        hv = torch.empty((nloc, 3 * natoms), dtype=torch.float64).pin_memory()
        st = eng.get_state()
        hx.numpy()[:] = st["x"]; hv.numpy()[:] = st["v"]
        hbox, hdx, hdv, hdt = st["box"].copy(), st["dx"].copy(), st["dv"].copy(), st["dt"].copy()
        e2e_steps = max(2, min(steps, 4))
        eng.reset_counters()
        comm.barrier(); torch.cuda.synchronize()
        w0 = time.perf_counter()
        breakdown = os.environ.get("NM_BENCH_E2E_BREAKDOWN")          # development aid: synchronises after every call
        tb = [0.0] * 4
        kev_e2e = []
        for _ in range(e2e_steps):
            c0 = time.perf_counter()
            eng.set_state(x=hx.numpy(), v=hv.numpy(), box=hbox, dx=hdx, dv=hdv, dt=hdt)      # H2D (+ the 'run 0' evaluation)
            if breakdown: eng.synchronize()
            c1 = time.perf_counter()
            th = step(cyc, kev_e2e if breakdown else None); cyc += 1
            if breakdown: eng.synchronize()
            c2 = time.perf_counter()
            st = eng.get_state(x_out=hx.numpy(), v_out=hv.numpy())                       # D2H straight into the pinned host buffers
            c3 = time.perf_counter()
            hbox, hdx, hdv, hdt = st["box"], st["dx"], st["dv"], st["dt"]
            c4 = time.perf_counter()
            for k, dtk in enumerate((c1 - c0, c2 - c1, c3 - c2, c4 - c3)): tb[k] += dtk
        if breakdown and comm.rank == 0:
            print("e2e breakdown (ms/step): set_state %.2f  step %.2f (cycle kernel %.2f)  get_state %.2f  host copies %.2f" % (
                1e3 * tb[0] / e2e_steps, 1e3 * tb[1] / e2e_steps, sum(a.elapsed_time(b) for a, b in kev_e2e) / e2e_steps, 1e3 * tb[2] / e2e_steps, 1e3 * tb[3] / e2e_steps), file=sys.stderr)
        torch.cuda.synchronize(); comm.barrier()
        e2e_ms = (time.perf_counter() - w0) * 1e3
        ct_e2e = eng.counters()
        atom_steps_e2e_local, sweeps_e2e_local = ct_e2e["hmc_atom_steps"], ct_e2e["sweeps"]
    gather.finish()
    # ---------------- reduce over ranks: max time, summed work
    vals = torch.tensor([ms, e2e_ms, kernel_ms], dtype=torch.float64, device="cuda")
    work = torch.tensor([ct["hmc_atom_steps"], ct["sweeps"], flops_from(ct), atom_steps_e2e_local, launches,
                         ct["pairs_force"] + ct["pairs_full"], ct["list_pairs"], ct["list_builds"], ct["outer_builds"], ct["pmc_trials"], sweeps_e2e_local],
                        dtype=torch.float64, device="cuda")
    if comm.world > 1:
        comm.dist.all_reduce(vals, op=comm.dist.ReduceOp.MAX)
        comm.dist.all_reduce(work, op=comm.dist.ReduceOp.SUM)
    ms, e2e_ms, kernel_ms = (float(t) for t in vals.cpu())
    atom_steps, sweeps, flops, atom_steps_e2e, launches_all, inpairs, listpairs, builds, obuilds, trials, sweeps_e2e = (float(t) for t in work.cpu())
    out = None
    if comm.rank == 0:
        peak, _ = nm.measure_fma_peak(dev, args.precision)
        achieved = flops / comm.world / (kernel_ms * 1e-3) if kernel_ms > 0 else 0.0      # per-GPU, kernel-only time
        traffic, capture = ncu_traffic(workload) if comm.world == 1 else (None, None)
        out = {
            "metric": "hmc_atom_steps_per_sec", "value": atom_steps / (ms * 1e-3), "unit": "atom-steps/s",
            "mc_sweeps_per_sec": sweeps / (ms * 1e-3),
            "n_gpus": comm.world, "steps": steps, "warmup": max(3, warmup), "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.precision == 64 else "f32",
            "data": "synthetic: pressure-relaxed fcc + random displacement (the reference's init_sample), %d equilibration cycles, counter-based RNG seed 256" % equil,
            "config": mc_config(desc, natoms, nloc, npn, nt, mod, comm.world, rows),
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "fp64_fma" if args.precision == 64 else "fp32_fma", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes/launch (ncu --set full, profiles/%s)" % capture if capture else None,
                         "kernel": "nm::k_cycle", "kernel_ms_per_step": kernel_ms / steps,
                         "peak_source": "live DFMA microbenchmark (nm_measure_fma_peak); MEASURED_PEAKS.json carries no FP64 figure",
                         "flops": "24/in-cutoff pair (force), 30 (force+energy+virial), 13/neighbour of a single-atom dE, 18/atom-step",
                         "in_cutoff_pairs_per_step": inpairs / steps, "listed_over_in_cutoff": listpairs / max(1.0, inpairs),
                         "list_builds_per_step": builds / steps, "outer_builds_per_step": obuilds / steps},
            "clocks": clocks,
        }
        if with_e2e:
            out["e2e"] = {"value": atom_steps_e2e / (e2e_ms * 1e-3), "unit": "atom-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                          "steps": e2e_steps, "mc_sweeps_per_sec": sweeps_e2e / (e2e_ms * 1e-3),
                          "note": "set_state (pinned host x, v, box, step sizes) -> cycle -> get_thermo + get_state, every step"}
        if not bulk:
            out["single_atom_trials_per_sec"] = trials / (ms * 1e-3)
        if gate is not None:
            out["parity_gate"] = gate
        # the asynchronously gathered table is the job-wide (pe + ke, vol) of the last exchange: finite, positive volumes
        out["exchange_table"] = {"slots": int(table.shape[0]), "finite": bool(np.isfinite(table).all()), "min_vol": float(table[:, 1].min())}
    eng.close()
    return out, dict(sz=sz, nt=nt, P=P, T=T, et=et, pf=pf, temp=temp, bulk=bulk, ppos=ppos, pvol=pvol, mod=mod, x=x, box=box, gs=gs)


def leg_summary(o):
    keep = ("metric", "value", "unit", "mc_sweeps_per_sec", "single_atom_trials_per_sec", "ordered_pair_distances_per_sec", "n_gpus", "steps", "warmup",
            "ms_per_step", "dtype", "config", "e2e", "gpu_launches", "roofline", "clocks", "parity_gate", "fp32_pipe")
    return {k: o[k] for k in keep if k in o}


def run_b200(args):
    import torch
    from neuralmelting_b200 import remcmc
    comm = remcmc.Comm()
    if comm.world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d" % (args.gpus, comm.world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(comm.local_rank)
    out, ctx = measure_mc(args, comm, torch, args.workload, args.steps, args.warmup, args.equil)
    if comm.rank == 0 and comm.world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, ctx["sz"], ctx["nt"], ctx["P"], ctx["T"], ctx["et"][ctx["gs"]], ctx["pf"][ctx["gs"]], ctx["temp"][ctx["gs"]],
                                           ctx["bulk"], ctx["ppos"], ctx["pvol"], ctx["mod"], ctx["x"], ctx["box"], args.cpu_seconds)
    if not args.no_legs and args.workload == "c2":
        # short legs of the other BASELINE configurations, same process, same box (driver-run evidence for C3 / C4 / C5)
        legs = {}
        for wl, (st_, wu_, eq_) in (("c3", (3, 3, 32)), ("c4", (24, 8, 64))):     # equilibrated like the headline (stationary step sizes)
            o, _ = measure_mc(args, comm, torch, wl, st_, wu_, eq_, with_e2e=True, with_gate=True)
            if o is not None:
                legs[wl] = leg_summary(o)
        o = run_rdf(args, comm=comm, steps=4, nsamples=min(args.rdf_samples, 512))
        if o is not None:
            legs["c5"] = leg_summary(o)
        if out is not None:
            out["workloads"] = legs
    return out


def cpu_baseline(args, sz, nt, P, T, et, pf, temp, bulk, ppos, pvol, mod, x0, box0, target_s):
    """the CPU restatement (oracle/, kind "port") farmed over all host cores on a bounded sample of the same workload"""
    from oracle import oracle as orc
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    natoms = 4 * sz ** 3
    ns = x0.shape[0]
    # replicas spread over the whole T range (work per replica depends on density), one task per replica
    nrep = min(ns, max(cores, 2 * cores))
    pick = np.unique(np.linspace(0, ns - 1, nrep).round().astype(int))
    nrep = pick.size
    params = orc.make_params(mod=mod, bulk_move=int(bulk), ppos=ppos, pvol=pvol, seed=256)
    labels = np.stack([et[pick], pf[pick], temp[pick], [float("%f" % t) for t in temp[pick]]], 1)

    def run(ncycles, mod_override=None):
        p = orc.make_params(mod=mod_override or mod, bulk_move=int(bulk), ppos=ppos, pvol=pvol, seed=256)
        x, v = x0[pick].copy(), np.zeros((nrep, 3 * natoms))
        scal = np.stack([box0[pick], np.full(nrep, 0.03125), np.full(nrep, 0.03125), np.full(nrep, 0.00390625)], 1).copy()
        counts = np.zeros((nrep, 6))
        t0 = time.perf_counter()
        _, ct = orc.farm(p, labels, 0, x, v, scal, counts, cycle0=0, ncycles=ncycles, nthreads=cores)
        return time.perf_counter() - t0, ct
    # calibrate on a short run, then size the sample
    probe_mod = max(1, mod // 16)
    tp, ctp = run(1, probe_mod)
    per_cycle = tp * (mod / probe_mod)
    ncycles = int(max(1, min(8, round(target_s / max(per_cycle, 1e-3)))))
    if per_cycle > 2.5 * target_s:
        # even one full cycle is too long: time a shortened cycle (fewer moves, same move mix) and say so
        use_mod = max(1, int(mod * target_s / per_cycle))
        t, ct = run(1, use_mod)
        sample = "%d replicas across the T range x 1 cycle of %d moves (shortened from %d), %d threads" % (nrep, use_mod, mod, cores)
    else:
        t, ct = run(ncycles)
        sample = "%d replicas across the T range x %d full cycle(s) of %d moves, %d threads" % (nrep, ncycles, mod, cores)
    return {"value": float(ct[orc.CT_HMC_ATOM_STEPS]) / t, "unit": "atom-steps/s", "cores": cores, "kind": "port",
            "mc_sweeps_per_sec": float(ct[orc.CT_SWEEPS]) / t, "sample": sample, "seconds": t,
            "note": "CPU restatement (LAMMPS unavailable): Verlet-list C code, no LAMMPS boot / per-move neighbour rebuilds / Python overhead -> faster than the real reference"}


def run_reference(args):
    """--impl reference: the reference's CPU path (restated, see cpu_baseline) on the host cores; rank 0 only"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from neuralmelting_b200 import remcmc
    from oracle import oracle as orc
    world = args.gpus
    sz, rows, nt, npn, P, T, bulk, ppos, pvol, mod, desc = grid_for(world, args.workload)
    natoms = 4 * sz ** 3
    et, pf = remcmc.init_constants(P, T)
    temp = np.tile(T.astype(np.float64), npn)
    # initial states without the GPU: perfect fcc at the analytic zero-temperature density of each pressure row is not
    # available on the CPU side of the product, so the oracle's own lattice helper is used (rho from a bisection on the shell sums)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rng = np.random.default_rng(1000)
    ns = npn * nt
    nrep = min(ns, max(cores, 2 * cores))
    pick = np.unique(np.linspace(0, ns - 1, nrep).round().astype(int))
    nrep = pick.size
    x0, box0 = [], []
    cache = {}
    for k in pick:
        i = k // nt
        if i not in cache:
            lo, hi = 1.4, 1.7          # fcc lattice constant bracket
            for _ in range(60):
                mid = 0.5 * (lo + hi)
                e, w, _ = orc.fcc_shell_sum(mid)
                p_mid = w / (3.0 * mid ** 3 / 4.0)
                lo, hi = (mid, hi) if p_mid >= float(P[i]) else (lo, mid)
            cache[i] = 0.5 * (lo + hi)
        L = float("%f" % (sz * cache[i]))
        pos = orc.fcc_positions(sz, L) + 0.035063 * 2.0 * (rng.random((natoms, 3)) - 0.5)
        x0.append(orc.wrap(pos.reshape(-1), L)); box0.append(L)
    x0, box0 = np.array(x0), np.array(box0)
    labels = np.stack([et[pick], pf[pick], temp[pick], [float("%f" % t) for t in temp[pick]]], 1)
    # bounded step: calibrate the number of moves per timed step so that the whole run ends within minutes
    def farm(mod_use, cycle0, state):
        p = orc.make_params(mod=mod_use, bulk_move=int(bulk), ppos=ppos, pvol=pvol, seed=256)
        t0 = time.perf_counter()
        _, ct = orc.farm(p, labels, 0, state[0], state[1], state[2], state[3], cycle0=cycle0, ncycles=1, nthreads=cores)
        return time.perf_counter() - t0, ct
    state = [x0.copy(), np.zeros_like(x0), np.stack([box0, np.full(nrep, 0.03125), np.full(nrep, 0.03125), np.full(nrep, 0.00390625)], 1).copy(), np.zeros((nrep, 6))]
    probe_mod = max(1, mod // 16)
    tp, _ = farm(probe_mod, 0, state)
    budget = 120.0 / max(1, args.steps + args.warmup)
    mod_use = int(max(1, min(mod, probe_mod * budget / max(tp, 1e-3))))
    for w in range(args.warmup):
        farm(mod_use, 1 + w, state)
    t = 0.0
    tot = np.zeros(orc.CT_N)
    for s in range(args.steps):
        dt, ct = farm(mod_use, 100 + s, state)
        t += dt; tot += ct.astype(np.float64)
    val = tot[orc.CT_HMC_ATOM_STEPS] / t
    sample = "%d of %d replicas across the T range, %d of %d moves per step, %d threads" % (nrep, ns, mod_use, mod, cores)
    return {"impl": "reference", "metric": "hmc_atom_steps_per_sec", "value": val, "unit": "atom-steps/s",
            "mc_sweeps_per_sec": tot[orc.CT_SWEEPS] / t, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic: relaxed fcc + random displacement, counter-based RNG seed 256",
            "config": mc_config(desc, natoms, rows * nt, npn, nt, mod, world, rows),      # the b200 arm's config, key for key
            "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of lammps_remcmc.py's per-replica path (LAMMPS + Dask are not installable here); one replica per task over all host threads"}


def run_rdf(args, comm=None, steps=None, nsamples=None):
    """c5 (BASELINE configs[4]): RDF feature extraction (lammps_distr.py:123-135) over samples of 4000-atom configurations.
    A step = one batch of --rdf-samples samples; value = samples/s with the positions resident in HBM; e2e = from host
    arrays through nm_rdf_counts (H2D of positions, D2H of counts inside the timed region). HBM roofline: 12 N + 6 bytes
    in, 4 SBINS out per sample -- the kernel is FP32-ALU bound, the fraction is reported for the record."""
    import torch
    from neuralmelting_b200 import engine as nm
    from neuralmelting_b200 import remcmc
    comm = comm or remcmc.Comm()
    dev = comm.local_rank
    torch.cuda.set_device(dev)
    steps = steps or args.steps
    n, sb, ns = 4000, 64, nsamples or args.rdf_samples
    rng = np.random.default_rng(5 + comm.rank)
    box = rng.uniform(15.2, 20.5, ns).astype(np.float32)                 # the density range of the 32x32 grid
    pos = (rng.uniform(0, 1, (ns, n, 3)) * box[:, None, None]).astype(np.float32)
    l = np.float32(15.2)
    r = np.linspace(1e-16, 1 / 2, sb) * l
    d_pos = torch.from_numpy(pos).cuda(); d_box = torch.from_numpy(box).cuda()
    d_cnt = torch.zeros((ns, sb), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        nm.rdf_counts_device(d_pos.data_ptr(), d_box.data_ptr(), n, ns, r, d_cnt.data_ptr(), device=dev, stream=stream)
    sampler = ClockSampler(dev) if comm.rank == 0 else None
    for _ in range(max(12, args.warmup)):          # short steps: warm up for ~0.5 s so that the SM clock has ramped
        step()
    comm.barrier(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    comm.barrier(); torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = t0.elapsed_time(t1)
    w0 = time.perf_counter()
    e2e_steps = 2
    for _ in range(e2e_steps):
        got = nm.rdf_counts(pos, box, r, device=dev)
    e2e_s = time.perf_counter() - w0
    vals = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if comm.world > 1:
        comm.dist.all_reduce(vals, op=comm.dist.ReduceOp.MAX)
    ms, e2e_ms = (float(t) for t in vals.cpu())
    if comm.rank != 0:
        return None
    bytes_per_sample = 12 * n + 6 + 4 * sb
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_per_sample * ns * steps * comm.world / (ms * 1e-3) / 1e9 / comm.world
    out = {"metric": "rdf_samples_per_sec", "value": ns * steps * comm.world / (ms * 1e-3), "unit": "samples/s",
           "n_gpus": comm.world, "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": ms / steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic: uniform random positions of 4000 atoms, boxes 15.2-20.5 (the density range of the 32x32 grid)",
           "config": {"workload": "C5: RDF histogram (lammps_distr.calculate_rdf), N=4000, SBINS=64, %d samples per step per GPU" % ns,
                      "l2": "inputs %.0f MB per step %s L2" % (12e-6 * n * ns, "larger than" if 12 * n * ns > 126e6 else "fit in; compute-bound kernel")},
           "ordered_pair_distances_per_sec": n * (n - 1.0) * ns * steps * comm.world / (ms * 1e-3),
           "e2e": {"value": ns * e2e_steps * comm.world / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": (12 * n + 4) * ns,
                   "d2h_bytes_per_step": 4 * sb * ns, "steps": e2e_steps},
           "gpu_launches": steps * comm.world,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "kernel": "nmrdf::k_rdf", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                        "note": "algorithmic bytes %d per sample; the kernel is bound by FP32/integer issue (N(N-1) ordered pair distances per sample), not by HBM" % bytes_per_sample},
           "clocks": clocks}
    # FP32-pipe view: 9 flop per ORDERED pair of the bit-exact formulation (3 sub, 3 mul, 2 add, 1 sqrt)
    peak32, _ = nm.measure_fma_peak(dev, 32)
    ach32 = 9.0 * out["ordered_pair_distances_per_sec"] / comm.world
    out["fp32_pipe"] = {"achieved": ach32 / 1e12, "peak": peak32 / 1e12, "unit": "TFLOP/s", "frac": ach32 / peak32,
                        "flops": "9 per ordered pair distance (un-contracted float32: 3 sub, 3 mul, 2 add, 1 sqrt); image selection, range tests and binning are overhead",
                        "peak_source": "live FFMA microbenchmark (nm_measure_fma_peak)"}
    traffic, capture = ncu_traffic("c5") if comm.world == 1 else (None, None)
    out["roofline"]["traffic"] = traffic
    if capture:
        out["roofline"]["traffic_unit"] = "bytes/launch (ncu --set full, profiles/%s)" % capture
    if comm.world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        nsamp = max(cores, 2)
        t0c = time.perf_counter()
        ref = orc.rdf_counts_farm(pos[:nsamp], box[:nsamp], r, nthreads=cores)
        tc = time.perf_counter() - t0c
        assert np.array_equal(ref, got[:nsamp]), "rdf parity"
        out["cpu_baseline"] = {"value": nsamp / tc, "unit": "samples/s", "cores": cores, "kind": "port", "seconds": tc,
                               "sample": "%d samples of the same batch, one per thread (27-image all-pairs float32 loop of the reference, in C)" % nsamp}
    return out


def main():
    args = parse()
    # rank 0 prints ONE JSON line on stdout: libraries that write there (NCCL's version banner) are sent to stderr for the
    # duration of the run, the line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    if args.workload == "c5":
        out = run_rdf(args) if args.impl != "reference" else None
        if out is not None:
            emit(out)
        return
    out = run_reference(args) if args.impl == "reference" else run_b200(args)
    if out is not None:
        emit(out)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
