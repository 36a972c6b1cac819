import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from neuralmelting_b200 import engine as nm
from oracle import oracle as orc
import test_gpu_parity as t
for n_side in (4, 5, 10):
    rho = [1.122, 1.1, 0.9, 0.6]; sig = [0.0, 0.05, 0.15, 0.3]
    x, box = t._configs(orc, n_side, rho, sig, seed=50 + n_side)
    n = 4 * n_side ** 3
    with nm.Engine(natoms=n, n_rep=4, nt=4, precision=32) as eng:
        eng.set_state(x=x, box=box)
        pe, w, f, npairs = eng.eval()
    for k in range(4):
        pe_o, w_o, f_o, np_o = orc.lj_eval_list(x[k], box[k])
        print(n_side, k, "dnp", npairs[k] - np_o, "pe rel %.2e" % (abs(pe[k] - pe_o) / abs(pe_o)), "w rel %.2e" % (abs(w[k] - w_o) / max(abs(w_o), abs(pe_o))),
              "f abs %.2e fmax %.2e w/n %.2e" % (np.abs(f[k] - f_o).max(), np.abs(f_o).max(), abs(w_o) / n))
