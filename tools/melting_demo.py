"""end-to-end sanity run of the drop-in pipeline on a small grid: equilibrate, collect, parse (the lammps_parse.py
logic), RDF + CDF stage; prints per-temperature averages of one pressure row (the melting transition shows as the
jump in energy / density between the solid and the liquid branch)"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, ".")
from neuralmelting_b200 import distr, remcmc

def main():
    td = tempfile.mkdtemp()
    os.chdir(td)
    t0 = time.time()
    a = remcmc.parse_args("-n demo -ss 4 -pn 4 -tn 16 -sn 120 -sc 60 -sm 64 -bm".split())
    counters, swaps = remcmc.run(a, log=lambda *x: None)
    t1 = time.time()
    pref = remcmc.file_prefix("demo", "LJ")
    P, T = np.load(pref + ".virial.trgt.npy"), np.load(pref + ".temp.trgt.npy")
    th = np.loadtxt(pref + ".thrm", dtype=np.float32).reshape(P.size, T.size, -1, 17)
    data = [l.split() for l in open(pref + ".traj")]
    hdr = np.array([v for v in data if len(v) == 2])
    natoms = hdr[:, 0].astype(np.uint16).reshape(P.size, T.size, -1)
    box = hdr[:, 1].astype(np.float32)
    x = np.concatenate([np.array(v).astype(np.float32) for v in data if len(v) == 3], 0).reshape(P.size, T.size, natoms.shape[2], 256, 3)
    np.save(pref + ".natoms.npy", natoms); np.save(pref + ".box.npy", box); np.save(pref + ".pos.npy", x)
    g = distr.run(distr.build_parser().parse_args("-n demo -sb 64 -cb 11".split()))
    t2 = time.time()
    print("run: %d cycles x 64 moves, 64 replicas of 256 atoms: %.1f s (%.3g HMC atom-steps, %d exchanges); parse+RDF+CDF of %d samples: %.1f s" % (
        120, t1 - t0, counters["hmc_atom_steps"], swaps, natoms.size, t2 - t1))
    r = np.load(pref + ".r.npy")
    for i in (0, P.size - 1):
        print("pressure %.2f:   T      <temp>   <pe>/N   <rho>    <press>   ah     g(r) peak" % P[i])
        for j in range(T.size):
            m = th[i, j].mean(0)
            print("              %5.2f   %6.3f  %7.3f  %6.3f  %7.3f  %5.2f   %5.2f at r=%.2f" % (T[j], m[0], m[1] / 256, 256 / m[4], m[3], m[16],
                  g[i, j].mean(0).max(), r[g[i, j].mean(0).argmax()]))

main()
