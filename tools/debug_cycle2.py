import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from neuralmelting_b200 import engine as nm
from oracle import oracle as orc
import test_gpu_parity as t
np.set_printoptions(linewidth=200, precision=3)
for bulk in (True, False):
    th_o, th_g, (xo, vo, scal), st, ct = t._run_both(nm, orc, 4, bulk, mod=24, ncycles=4, rho=[1.1, 1.0, 0.85, 0.6], temps=[0.4, 0.9, 1.6, 2.5], press=[1, 3, 5, 8])
    rel = np.abs(th_g - th_o) / np.maximum(np.abs(th_o), 1e-300)
    print("bulk", bulk, "counters equal", np.array_equal(th_g[..., 9:], th_o[..., 9:]))
    print(rel[..., :6].max(2))
    print(ct)
