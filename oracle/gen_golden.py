"""Generate tests/golden/*.npz|json by RUNNING THE REFERENCE'S OWN FUNCTIONS.

Run in the build container only (needs /root/reference, numba):
    python oracle/gen_golden.py
The GPU box has no /root/reference; tests read the committed fixtures instead.

What is pinned here (reference = /root/reference/scripts):
  * calculate_rdf            lammps_distr.py:123-135   (numba njit, as shipped)
  * calculate_spatial edges  lammps_distr.py:82-98     (restated inline: R, DNI)
  * calculate_cdf            lammps_distr.py:161-171   (.py_func: pure NumPy histogramdd)
  * replica_exchange         lammps_remcmc.py:776-803  (.py_func, module globals injected)
  * gen_mc_param             lammps_remcmc.py:726-745
  * write_thrm / write_traj / init_header  lammps_remcmc.py:176-256
  * parse_args defaults      lammps_remcmc.py:22-100, lammps_distr.py:15-53
The LJ / MD physics cannot be pinned this way (LAMMPS absent) -- see oracle/nm_oracle.c header.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/scripts"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def text_roundtrip(x):
    """positions as lammps_parse.py sees them: '%.4E' text -> float32 (lammps_remcmc.py:256, lammps_parse.py:93)"""
    return np.array([float("%.4E" % v) for v in np.asarray(x).reshape(-1)], dtype=np.float32).reshape(np.shape(x))


def fcc(sz, box):
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    cells = np.array([[i, j, k] for k in range(sz) for j in range(sz) for i in range(sz)], dtype=float)
    return (cells[:, None, :] + basis[None]).reshape(-1, 3) / sz * box


def gen_rdf():
    sys.path.insert(0, REF)
    import lammps_distr as ld
    rng = np.random.default_rng(20261018)
    cases = {}
    sbins = 64
    # three samples per case sharing one minimum box (edges depend on min(box) over ALL samples)
    for name, sz, sigma in (("n108", 3, 0.08), ("n256", 4, 0.15), ("n500", 5, 0.4)):
        n = 4 * sz ** 3
        boxes = np.array([float("%.4E" % (sz * a)) for a in (1.55, 1.61, 1.75)], dtype=np.float32)
        pos = np.zeros((3, n, 3), dtype=np.float32)
        for s in range(3):
            x = fcc(sz, float(boxes[s])) + rng.normal(0, sigma, (n, 3))
            x -= np.floor(x / float(boxes[s])) * float(boxes[s])
            if s == 2:   # a few atoms slightly outside the box, as LAMMPS hands them back after 'run N'
                x[:5] += float(boxes[s]) * 1.0
                x[5:9] -= 0.01
            pos[s] = text_roundtrip(x)
        natoms = np.full(3, n, dtype=np.uint16)
        l = np.min(boxes)
        r = np.linspace(1e-16, 1 / 2, sbins)
        dr = r[1] - r[0]
        dv = 4 * np.pi * np.square(r) * dr
        r = r * l
        dv = dv * l ** 3
        nrho = np.divide(natoms, np.power(boxes, 3))
        dni = np.multiply(nrho[:, np.newaxis], dv[np.newaxis, :])
        b = [-1, 0, 1]
        br = np.array([[b[i], b[j], b[k]] for i in range(3) for j in range(3) for k in range(3)], dtype=np.int8)
        rd = np.zeros(sbins, dtype=np.float32)
        g = []
        for s in range(3):
            out = ld.calculate_rdf(natoms[s], boxes[s], br, pos[s], r, rd)
            g.append(np.array(out))
        g = np.array(g)
        print("rdf", name, "dtype", g.dtype, "edge dtype", r.dtype, "sum counts", (g * n).sum(1))
        cases[name] = dict(pos=pos, box=boxes, natoms=natoms, r=r, dni=dni, g=g)
    np.savez_compressed(os.path.join(OUT, "rdf_reference.npz"),
                        **{"%s_%s" % (k, f): v for k, d in cases.items() for f, v in d.items()})


def gen_cdf():
    """calculate_cdf.py_func (the jitted wrapper raises TypingError under numba 0.65: np.histogramdd unsupported)"""
    sys.path.insert(0, REF)
    import lammps_distr as ld
    rng = np.random.default_rng(77)
    out = {}
    for name, sz, cb in (("n108_cb8", 3, 8), ("n256_cb11", 4, 11)):
        n = 4 * sz ** 3
        boxes = np.array([float("%.4E" % (sz * a)) for a in (1.56, 1.7)], dtype=np.float32)
        pos = np.zeros((2, n, 3), dtype=np.float32)
        for s in range(2):
            x = fcc(sz, float(boxes[s])) + rng.normal(0, 0.2, (n, 3))
            x -= np.floor(x / float(boxes[s])) * float(boxes[s])
            if s == 1:
                x[:4] += float(boxes[s])
                x[0] = 0.0
            pos[s] = text_roundtrip(x)
        l = float(np.min(boxes))
        rv = np.array([np.linspace(0, l, cb + 1) for _ in range(3)], dtype=np.float64)
        rv -= l / 2
        b = [-1, 0, 1]
        br = np.array([[b[i], b[j], b[k]] for i in range(3) for j in range(3) for k in range(3)], dtype=np.int8)
        cd = np.zeros((cb, cb, cb), dtype=np.float32)
        res = []
        for s in range(2):
            res.append(np.array(ld.calculate_cdf.py_func(np.uint16(n), boxes[s], br, pos[s], rv, cd)))
        res = np.array(res)
        print("cdf", name, res.dtype, "counts", (res * n).sum(axis=(1, 2, 3)))
        out.update({name + "_pos": pos, name + "_box": boxes, name + "_rv": rv, name + "_c": res, name + "_natoms": np.full(2, n, np.uint16)})
    np.savez_compressed(os.path.join(OUT, "cdf_reference.npz"), **out)


def import_remcmc():
    stub = types.ModuleType("lammps")
    stub.lammps = object
    sys.modules["lammps"] = stub
    sys.path.insert(0, REF)
    import lammps_remcmc as lr
    return lr


def gen_exchange(lr):
    out = {}
    rng = np.random.default_rng(7)
    for name, np_, nt, mode in (("g2x4", 2, 4, "random"), ("g4x8", 4, 8, "random"),
                                 ("g3x6_anti", 3, 6, "anti"), ("g2x5_inf", 2, 5, "inf")):
        ns = np_ * nt
        P = np.linspace(1, 8, np_, dtype=np.float32)
        T = np.linspace(0.25, 2.5, nt, dtype=np.float32)
        # CONST as init_constant computes it (lj branch, lammps_remcmc.py:128-131); the script was written
        # for numpy-1 scalar promotion (float32 scalars promote to float64 in 1.0*T) -- keep that.
        const = [(1.0 * float(T[k % nt]), float(P[k // nt]) / (1.0 * float(T[k % nt]))) for k in range(ns)]
        state = []
        for k in range(ns):
            n = 32
            pe = float(rng.normal(-6.0 * n, 20.0))
            if mode == "anti":
                pe = -200.0 + 30.0 * (nt - 1 - k % nt)     # hot slots hold the low energies -> every swap accepted
            ke = float(abs(rng.normal(1.5 * n * T[k % nt], 3.0)))
            vol = float(rng.normal(n / 0.9, 2.0))
            if mode == "inf" and k % nt == 1:
                pe = 1e308
            if mode == "inf" and k % nt == 3:
                pe = float("nan")
            st = [n, rng.normal(size=3 * n), rng.normal(size=3 * n), 2 * ke / (3 * n - 3), pe, ke, 1.0,
                  vol ** (1 / 3), vol, 0.03 + 0.001 * k, 0.04 + 0.001 * k, 0.004 + 0.0001 * k] + list(np.zeros(9))
            st[0] = 1000 + k     # tag: natoms slot doubles as the configuration id (it is swapped with [:12])
            state.append(st)
        seed = 256
        np.random.seed(seed)
        uniforms = np.random.rand(np_ * nt * (nt - 1) // 2)
        lr.NP, lr.NT, lr.STATE, lr.CONST, lr.VERBOSE, lr.PARALLEL = np_, nt, state, const, False, False
        np.random.seed(seed)
        with np.errstate(all="ignore"):
            lr.replica_exchange.py_func()
        after = np.random.rand()
        np.random.seed(seed)
        np.random.rand(uniforms.size)
        assert after == np.random.rand(), "replica_exchange did not draw exactly one uniform per pair"
        perm = np.array([st[0] - 1000 for st in state], dtype=np.int32)
        pe0 = np.zeros(ns)
        ke0 = np.zeros(ns)
        vol0 = np.zeros(ns)
        dx0 = np.zeros(ns)
        for k in range(ns):      # pre-exchange values by ORIGINAL slot, recovered through the tags
            src = perm[k]
            pe0[src], ke0[src], vol0[src], dx0[src] = state[k][4], state[k][5], state[k][8], state[k][9]
        out[name + "_shape"] = np.array([np_, nt])
        out[name + "_pe"] = pe0
        out[name + "_ke"] = ke0
        out[name + "_vol"] = vol0
        out[name + "_dx"] = dx0
        out[name + "_et"] = np.array([c[0] for c in const])
        out[name + "_pf"] = np.array([c[1] for c in const])
        out[name + "_uniforms"] = uniforms
        out[name + "_perm"] = perm
        print("exchange", name, "perm", perm.tolist())
    np.savez_compressed(os.path.join(OUT, "exchange_reference.npz"), **out)


def gen_adapt_and_text(lr):
    res = {}
    # gen_mc_param over the ratio cases that matter (float32 ratios as gen_sample makes them)
    cases = []
    for nap, ntp in ((0, 0), (1, 2), (3, 7), (5, 9), (0, 5), (16, 16), (64, 128), (63, 128), (65, 128)):
        with np.errstate(invalid="ignore"):
            a = np.nan_to_num(np.float32(nap) / np.float32(ntp))
        st = [0] * 9 + [0.03125, 0.0625, 0.00390625] + [ntp, nap, ntp, nap, ntp, nap, a, a, a]
        new = lr.gen_mc_param(st)
        cases.append(dict(nap=nap, ntp=ntp, ratio=float(a), dx=float(new[9]), dv=float(new[10]), dt=float(new[11]),
                          tail=[float(v) for v in new[12:]]))
    res["adapt"] = cases
    # text writers
    rng = np.random.default_rng(3)
    n = 7
    x = rng.uniform(-0.5, 9.5, 3 * n)
    x[0], x[1], x[2] = 0.0, 1.23456e-7, 12345.678
    state = [n, x, rng.normal(size=3 * n), 1.01234567, -1234.56789, 385.123456, 2.3456789, 6.5432109, 280.123456,
             0.031250, 0.0332031, 0.00390625, 13.0, 4.0, 18.0, 9.0, 97.0, 85.0,
             float(np.float32(4) / np.float32(13)), 0.5, float(np.float32(85) / np.float32(97))]
    with tempfile.TemporaryDirectory() as td:
        thrm, traj = os.path.join(td, "a.thrm"), os.path.join(td, "a.traj")
        lr.NSMPL, lr.CUTOFF, lr.MOD, lr.NSWPS = 1024, 0, 128, 1024 * 128
        lr.PPOS, lr.PVOL, lr.PHMC, lr.NSTPS, lr.SEED = 0.125, 0.125, 0.75, 8, 256
        lr.EL, lr.SZ, lr.DX, lr.DV, lr.DT = "LJ", 5, 0.03125, 0.03125, 0.00390625
        lr.UNITS = {"LJ": "lj"}
        lr.LAT = {"LJ": ("fcc", 1.122)}
        lr.MASS = {"LJ": 1.0}
        lr.NP, lr.NT = 4, 8
        lr.P = np.linspace(1, 8, 4, dtype=np.float32)
        lr.T = np.linspace(0.25, 2.5, 8, dtype=np.float32)
        # init_header uses np.unravel_index(dims=...) which numpy >= 1.21 rejects: shim the keyword only
        orig = np.unravel_index
        np.unravel_index = lambda k, dims=None, order="C", shape=None: orig(k, dims if shape is None else shape, order=order)
        try:
            lr.init_header(13, (thrm, traj))
        finally:
            np.unravel_index = orig
        lr.write_thrm((thrm, traj), state)
        lr.write_traj((thrm, traj), state)
        res["thrm_text"] = open(thrm).read()
        res["traj_text"] = open(traj).read()
    res["state_scalars"] = [float(v) if not isinstance(v, np.ndarray) else None for v in state]
    res["state_x"] = x.tolist()
    res["header_k"] = 13
    # argparse defaults
    argv = sys.argv
    sys.argv = ["x"]
    res["remcmc_defaults"] = [v if not isinstance(v, (np.floating, np.integer)) else float(v) for v in lr.parse_args()]
    import lammps_distr as ld
    res["distr_defaults"] = list(ld.parse_args())
    sys.argv = argv
    with open(os.path.join(OUT, "host_reference.json"), "w") as fh:
        json.dump(res, fh, indent=1)
    print("adapt/text/defaults written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_rdf()
    gen_cdf()
    lr = import_remcmc()
    gen_exchange(lr)
    gen_adapt_and_text(lr)
