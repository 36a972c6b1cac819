"""GPU tests of the host drivers end to end: lammps_remcmc.py-compatible run -> files -> parse -> RDF stage."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _parse_like_reference(pref):
    """what lammps_parse.py:37-103 does with the consolidated files"""
    P, T = np.load(pref + ".virial.trgt.npy"), np.load(pref + ".temp.trgt.npy")
    pn, tn = P.size, T.size
    cols = np.loadtxt(pref + ".thrm", dtype=np.float32)
    assert cols.shape[1] == 17
    data = [line.split() for line in open(pref + ".traj")]
    hdr = np.array([v for v in data if len(v) == 2])
    natoms = hdr[:, 0].astype(np.uint16).reshape(pn, tn, -1)
    box = hdr[:, 1].astype(np.float32)
    x = np.concatenate([np.array(v).astype(np.float32) for v in data if len(v) == 3], 0)
    x = x.reshape(pn, tn, natoms.shape[2], natoms[0, 0, 0], 3)
    return cols.reshape(pn, tn, -1, 17), natoms, box, x


def test_remcmc_run_writes_reference_compatible_files(nm, orc, tmp_path):
    from neuralmelting_b200 import distr, remcmc
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        args = remcmc.parse_args("-n t1 -ss 4 -pn 2 -tn 3 -sn 4 -sc 1 -sm 6 -bm -rd 2".split())
        counters, swaps = remcmc.run(args, log=lambda *a: None)
        assert counters["sweeps"] == 2 * 3 * 4 * 6
        pref = remcmc.file_prefix("t1", "LJ")
        th, natoms, box, x = _parse_like_reference(pref)
        assert th.shape == (2, 3, 3, 17) and x.shape == (2, 3, 3, 256, 3) and (natoms == 256).all()
        # recorded thermo is physical: temperature columns positive, volume = box^3 within the text precision
        assert (th[..., 0] >= 0).all() and (th[..., 4] > 0).all()
        np.testing.assert_allclose(th[..., 4].reshape(-1), box.astype(np.float64) ** 3, rtol=5e-4)
        assert (x >= 0).all() and (x.reshape(2, 3, 3, -1).max(-1) <= box.reshape(2, 3, 3) * (1 + 1e-4)).all()
        assert os.path.exists(pref + ".rstrt.0002.npy") and os.path.exists(pref + ".rstrt.0004.npy")
        rst = np.load(pref + ".rstrt.0004.npy", allow_pickle=True)
        assert rst.shape == (6, 21)
        # energies in the files match a fresh oracle evaluation of the dumped configuration (text precision aside)
        st = rst[4]
        pe_o = orc.lj_eval_list(np.asarray(st[1]), float(st[7]))[0]
        assert abs(pe_o - float(st[4])) <= 1e-9 * abs(pe_o)
        # restart run continues from the dump
        args2 = remcmc.parse_args("-n t2 -r -rn t1 -rs 4 -ss 4 -pn 2 -tn 3 -sn 2 -sm 4 -bm".split())
        c2, _ = remcmc.run(args2, log=lambda *a: None)
        assert c2["sweeps"] == 2 * 3 * 2 * 4
        # RDF stage on the parsed arrays (what lammps_parse.py would have written)
        np.save(pref + ".natoms.npy", natoms)
        np.save(pref + ".box.npy", box)
        np.save(pref + ".pos.npy", x)
        g = distr.run(distr.build_parser().parse_args("-n t1 -sb 32 -cb 6".split()))
        cdf = np.load(pref + ".cdf.npy")
        assert cdf.shape == (2, 3, 3, 6, 6, 6) and np.load(pref + ".rv.npy").shape == (3, 7) and np.load(pref + ".dn.npy").shape == (18,)
        assert abs(cdf.mean() - 1.0) < 0.35          # pair density relative to the ideal gas, averaged over the cell
        assert g.shape == (2, 3, 3, 32) and g.dtype == np.float64
        r = np.load(pref + ".r.npy")
        flat_x, flat_n = x.reshape(-1, 256, 3), natoms.reshape(-1)
        cnt = orc.rdf_counts(flat_x[5], box[5], r)
        dni = np.load(pref + ".dni.npy").reshape(-1, 32)
        np.testing.assert_array_equal(g.reshape(-1, 32)[5], (cnt.astype(np.float32) / np.float32(flat_n[5])) / dni[5])
    finally:
        os.chdir(cwd)


def test_iterative_pmc_run_and_ensemble_sanity(nm):
    """-pm heavy single-atom sweeps (config 4 style): acceptance adapts towards 0.5 and energies stay finite"""
    from neuralmelting_b200 import remcmc
    P, T = remcmc.grids(1, 8, 1, 0.6, 2.0, 4)
    x, v, box = remcmc.init_samples(P, T, 4, 0.03125, np.random.default_rng(0))
    et, pf = remcmc.init_constants(P, T)
    with nm.Engine(natoms=256, n_rep=4, nt=4, mod=4, bulk_move=False, ppos=0.75, pvol=0.125) as eng:
        eng.set_labels(et, pf, np.tile(T.astype(np.float64), 1))
        eng.set_state(x=x, v=v, box=np.array([remcmc.text6(b) for b in box]), dx=np.full(4, .03125), dv=np.full(4, .03125),
                      dt=np.full(4, .00390625))
        aps = []
        for c in range(30):
            eng.run_cycle(c)
            th = eng.get_thermo()
            eng.adapt()
            eng.exchange(c)
            aps.append(th[:, 15].copy())
        assert np.isfinite(th).all()
        late = np.array(aps[-10:])
        assert 0.3 < late[late > 0].mean() < 0.7
