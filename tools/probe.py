"""ad-hoc timing probe (development aid, not the benchmark)"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from neuralmelting_b200 import engine as nm
from oracle import oracle as orc

def grid(n_side, np_, nt):
    n = 4 * n_side ** 3
    P = np.linspace(1, 8, np_, dtype=np.float32).astype(float)
    import os
    tlo, thi = float(os.environ.get('TLO', 0.25)), float(os.environ.get('THI', 2.5))
    T = np.linspace(tlo, thi, nt, dtype=np.float32).astype(float)
    rng = np.random.default_rng(0)
    xs, boxes, et, pf, tt = [], [], [], [], []
    for i in range(np_):
        for j in range(nt):
            rho = min(1.1, 1.05 - 0.35 * (T[j] - 0.25) / 2.25 + 0.02 * (P[i] - 1))
            box = n_side * (4 / rho) ** (1 / 3)
            x = orc.fcc_positions(n_side, box) + rng.normal(0, 0.03, (n, 3))
            xs.append(orc.wrap(x.reshape(-1), box)); boxes.append(box); et.append(T[j]); pf.append(P[i] / T[j]); tt.append(T[j])
    return n, np.array(xs), np.array(boxes), np.array(et), np.array(pf), np.array(tt)

def main():
    n_side, np_, nt = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    ncyc = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    bulk = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    skin = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
    oskin = float(sys.argv[7]) if len(sys.argv) > 7 else 0.0
    n, x, box, et, pf, tt = grid(n_side, np_, nt)
    ns = np_ * nt
    for prec in (64, 32):
        fl, ms = nm.measure_fma_peak(0, prec)
        print("fma peak fp%d: %.2f TFLOP/s (%.3f ms)" % (prec, fl / 1e12, ms))
    eng = nm.Engine(natoms=n, n_rep=ns, nt=nt, bulk_move=bool(bulk), skin=skin, skin_outer=oskin, precision=int(os.environ.get('PREC', 64)),
                    mod=int(os.environ.get('MOD', 128)), ppos=float(os.environ.get('PPOS', 0.125)), pvol=float(os.environ.get('PVOL', 0.125)))
    eng.set_labels(et, pf, tt)
    t0 = time.time()
    eng.set_state(x=x, v=np.zeros_like(x), box=box, dx=np.full(ns, .03125), dv=np.full(ns, .03125), dt=np.full(ns, .00390625))
    print("set_state %.3f s" % (time.time() - t0))
    for cyc in range(ncyc):
        eng.reset_counters()
        t0 = time.time()
        eng.run_cycle(cyc); eng.synchronize()
        dt = time.time() - t0
        th = eng.get_thermo(); eng.adapt(); perm, sw = eng.exchange(cyc)
        ct = eng.counters()
        flops = 24.0 * ct["pairs_force"] + 30.0 * ct["pairs_full"]
        print("cycle %d: %.1f ms  atom-steps/s %.3e  sweeps/s %.3e  pair-flops %.2f TF/s  builds %d evals %d  listpairs/inpairs %.2f swaps %d  <ah> %.2f <av> %.2f <ap> %.2f" % (
            cyc, dt * 1e3, ct["hmc_atom_steps"] / dt, ct["sweeps"] / dt, flops / dt / 1e12, ct["list_builds"], ct["force_evals"],
            ct["list_pairs"] / max(1, ct["pairs_force"] + ct["pairs_full"]), sw, th[:, 17].mean(), th[:, 16].mean(), th[:, 15].mean()))
        print("   per-call clk: outer %.0f inner %.0f vel %.0f | share outer %.2f inner %.2f vel %.2f" % (ct["clk_outer"] / max(1, ct["outer_builds"] or ct["list_builds"]), ct["clk_inner"] / max(1, ct["list_builds"]), ct["clk_vel"] / max(1, ct["hmc_moves"]), ct["clk_outer"] / ct["clk_total"], ct["clk_inner"] / ct["clk_total"], ct["clk_vel"] / ct["clk_total"]))
        print("   build phases (debug build): tiles %.0f  tiles+ghost table %.0f  row walk %.0f  clk per build" % (ct["clk_outer"] / max(1, ct["list_builds"]), ct["dbg_loopit"] / max(1, ct["list_builds"]), ct["dbg_loopclk"] / max(1, ct["list_builds"])))
        if not bulk:
            print("   iterative PMC (debug build): rounds %d  trials/round %.1f  clk/round: evaluate %.0f commit %.0f  acc %.2f  (of the evaluate: proposals %.0f)" % (ct["helped_evals"], ct["pmc_trials"] / max(1, ct["helped_evals"]), ct["dbg_loopclk"] / max(1, ct["helped_evals"]), ct["dbg_loopit"] / max(1, ct["helped_evals"]), th[:, 15].mean(), ct["clk_vel"] / max(1, ct["helped_evals"])))
        print("   outer builds %d | clk share: eval %.2f build %.2f  | clk/eval %.0f clk/build %.0f  | total Mclk/CTA %.1f" % (ct["outer_builds"], ct["clk_eval"] / ct["clk_total"], ct["clk_build"] / ct["clk_total"], ct["clk_eval"] / max(1, ct["force_evals"]), ct["clk_build"] / max(1, ct["list_builds"]), ct["clk_total"] / ns / 1e6))
    clk = eng.cta_clocks().astype(float).reshape(np_, nt) / 1e6
    print("per-slot Mclk by temperature (mean over P):", np.round(clk.mean(0), 1))
    print("per-slot Mclk min %.1f mean %.1f max %.1f" % (clk.min(), clk.mean(), clk.max()))
    rc = eng.replica_counters().astype(float); ci = {k: i for i, k in enumerate(nm.COUNTER_COLS)}
    order = np.argsort(-rc[:, ci["clk_total"]])
    print("slot   T     rho   Mclk  sweeps hmc builds evals  Mclk_build Mclk_eval  pairs/eval")
    for k in list(order[:12]) + list(order[-4:]):
        r = rc[k]
        print("%4d  %.2f  %.3f  %5.1f  %4d %4d  %4d  %5d   %6.1f  %6.1f   %7.0f" % (k, tt[k], n / th[k, 5], r[ci["clk_total"]] / 1e6, r[ci["sweeps"]], r[ci["hmc_moves"]],
              r[ci["list_builds"]], r[ci["force_evals"]], r[ci["clk_build"]] / 1e6, r[ci["clk_eval"]] / 1e6, r[ci["list_pairs"]] / max(1, r[ci["force_evals"]])))
    if os.environ.get("SMID"):
        smid = rc[:, ci["helped_evals"]].astype(int)
        cnt = np.bincount(smid, minlength=160)
        print("CTAs per SM histogram:", np.bincount(cnt[:int(smid.max()) + 1]), " distinct SMs:", (cnt > 0).sum(), " max smid:", smid.max())
        clkm = rc[:, ci["clk_total"]] / 1e6
        for nshare in (1, 2):
            sel = cnt[smid] == nshare
            if sel.any():
                print("  CTAs on SMs with %d CTA(s): n=%d mean Mclk %.1f max %.1f" % (nshare, sel.sum(), clkm[sel].mean(), clkm[sel].max()))
        pair_sum = np.zeros(160)
        np.add.at(pair_sum, smid, clkm)
        print("  per-SM sum of CTA Mclk: min %.1f mean %.1f max %.1f" % (pair_sum[cnt > 0].min(), pair_sum[cnt > 0].mean(), pair_sum[cnt > 0].max()))
    ctot = np.sort(rc[:, ci["clk_total"]] / 1e6)[::-1]
    print("sorted clk_total (Mclk) deciles:", np.round(ctot[::max(1, len(ctot) // 16)], 1), " kernel %.1f Mclk at 1.965 GHz; sum/148 = %.1f" % (dt * 1965e6 / 1e6, ctot.sum() / 148))
    print("T col0 thermo:", np.round(th[:nt, :6], 3)[::max(1, nt // 4)])

main()
