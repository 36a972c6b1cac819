import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from neuralmelting_b200 import engine as nm
from oracle import oracle as orc
import test_gpu_parity as t
np.set_printoptions(linewidth=200, precision=12)
bulk = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
th_o, th_g, (xo, vo, scal), st, ct = t._run_both(nm, orc, 4, bulk, mod=1, ncycles=72, rho=[1.1, 1.0, 0.85, 0.6], temps=[0.4, 0.9, 1.6, 2.5], press=[1, 3, 5, 8])
rel = np.abs(th_g - th_o) / np.maximum(np.abs(th_o), 1e-300)
for cyc in range(72):
    bad = rel[cyc, :, :9].max(1)
    kinds = ["P" if th_o[cyc, k, 9] else ("V" if th_o[cyc, k, 11] else "H") for k in range(4)]
    accs = [int(th_o[cyc, k, 10] + th_o[cyc, k, 12] + th_o[cyc, k, 14]) for k in range(4)]
    print(cyc, " ".join("%s%d:%.1e" % (kinds[k], accs[k], bad[k]) for k in range(4)), "cnt_eq", np.array_equal(th_g[cyc, :, 9:], th_o[cyc, :, 9:]))
    if bad.max() > 1e-6:
        k = int(bad.argmax())
        print("  gpu", th_g[cyc, k, :9]); print("  cpu", th_o[cyc, k, :9]); break
