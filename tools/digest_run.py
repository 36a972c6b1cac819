"""tiny runs that print a digest of the results (development aid: run-to-run / helper-schedule determinism checks):
python tools/sanitize_small.py small|large"""
import sys
import numpy as np
sys.path.insert(0, ".")
from neuralmelting_b200 import engine as nm
from oracle import oracle as orc

def main():
    large = len(sys.argv) > 1 and sys.argv[1] == "large"
    n_side = 10 if large else 4
    n = 4 * n_side ** 3
    nrep = 2 if large else 3
    rng = np.random.default_rng(0)
    xs, boxes = [], []
    for rho in ([1.0, 0.7] if large else [1.05, 0.9, 0.7]):
        box = orc.round6(n_side * (4 / rho) ** (1 / 3))
        xs.append(orc.wrap((orc.fcc_positions(n_side, box) + rng.normal(0, 0.04, (n, 3))).reshape(-1), box)); boxes.append(box)
    x, box = np.array(xs), np.array(boxes)
    T = np.linspace(0.5, 2.0, nrep); P = np.full(nrep, 2.0)
    with nm.Engine(natoms=n, n_rep=nrep, nt=nrep, mod=4 if large else 10, bulk_move=True, seed=5, ppos=0.25, pvol=0.25, skin_outer=0.25 if large else 0.0) as eng:
        eng.set_labels(T, P / T, T, T)
        eng.set_state(x=x, v=np.zeros_like(x), box=box, dx=np.full(nrep, 0.004), dv=np.full(nrep, 0.01), dt=np.full(nrep, 0.005))
        for cyc in range(2):
            eng.run_cycle(cyc); th = eng.get_thermo(); eng.adapt(); eng.exchange(cyc)
        ct = eng.counters(); st = eng.get_state()
    import hashlib
    h = hashlib.sha256(th.tobytes() + st["x"].tobytes() + st["v"].tobytes()).hexdigest()[:16]
    print("digest", h, "helped_evals", ct["helped_evals"], "list_builds", ct["list_builds"], "outer_builds", ct["outer_builds"])

main()
