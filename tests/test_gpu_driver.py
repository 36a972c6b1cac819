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
        # STEP = -1 dump before the loop (lammps_remcmc.py:975-976): the initial state, v = 0, pe of the start configuration
        r0 = np.load(pref + ".rstrt.0000.npy", allow_pickle=True)
        assert r0.shape == (6, 21) and float(r0[3][5]) == 0.0 and not np.asarray(r0[3][2]).any()
        pe0 = orc.lj_eval_list(np.asarray(r0[3][1]), float(r0[3][7]))[0]
        assert abs(pe0 - float(r0[3][4])) <= 1e-9 * abs(pe0)
        assert not [f for f in os.listdir(tmp_path) if f.endswith(".lammps.thrm") and ".0" in f.split("fcc")[1]]   # per-replica parts consolidated away
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
    x, v, box = remcmc.init_samples(P, T, 4, 0.03125, 0)
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


def test_direct_npy_emission_equals_parsing_the_text(nm, orc, tmp_path):
    """N2: -dn writes .pos/.box/.natoms/.<thermo>.npy during the run; they equal what lammps_parse.py:37-103 reads back from
    the consolidated text (the same 5-significant-digit round trip)"""
    from neuralmelting_b200 import remcmc
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        args = remcmc.parse_args("-n d1 -ss 4 -pn 2 -tn 2 -sn 3 -sc 1 -sm 4 -bm -dn -nt 2".split())
        remcmc.run(args, log=lambda *a: None)
        pref = remcmc.file_prefix("d1", "LJ")
        th, natoms, box, x = _parse_like_reference(pref)
        np.testing.assert_array_equal(np.load(pref + ".pos.npy"), x)
        np.testing.assert_array_equal(np.load(pref + ".box.npy"), box)
        np.testing.assert_array_equal(np.load(pref + ".natoms.npy"), natoms)
        for c, name in enumerate(remcmc.THERMO_NAMES):
            np.testing.assert_array_equal(np.load(pref + ".%s.npy" % name), th[..., c], err_msg=name)
    finally:
        os.chdir(cwd)


def test_relaxed_boxes_hit_the_survey_anchors(nm):
    """N3: zero-temperature pressure-relaxed fcc densities: rho0(P=1) = 1.08969, rho0(P=8) = 1.16740"""
    from neuralmelting_b200 import remcmc
    for sz in (4, 5):
        L = remcmc.relaxed_boxes(np.array([1.0, 8.0]), sz)
        rho = 4 * sz ** 3 / L ** 3
        np.testing.assert_allclose(rho, [1.08969, 1.16740], atol=6e-6)


def test_interpolated_start_keeps_the_velocity_draw(nm, tmp_path):
    """-is: the 'velocity all create T[j]' draw of init_sample stays in STATE (lammps_remcmc.py:420-425): KE = (3N-3)/2 T
    before 'zero angular' takes a little out, zero total momentum, and it is what rstrt.0000 holds"""
    from neuralmelting_b200 import remcmc
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        args = remcmc.parse_args("-n i1 -ss 4 -pn 1 -tn 3 -sn 1 -sc 1 -sm 2 -bm -is".split())
        remcmc.run(args, log=lambda *a: None)
        r0 = np.load(remcmc.file_prefix("i1", "LJ") + ".rstrt.0000.npy", allow_pickle=True)
        _, T = remcmc.grids(1, 8, 1, 0.25, 2.5, 3)
        for k in range(3):
            v = np.asarray(r0[k][2]).reshape(-1, 3)
            ke = 0.5 * (v ** 2).sum()
            assert 0.97 * 1.5 * 255 * float("%f" % T[k]) < ke <= 1.5 * 255 * float("%f" % T[k]) * (1 + 1e-12)
            assert np.abs(v.sum(0)).max() < 1e-10
            assert float(r0[k][5]) == pytest.approx(ke, rel=1e-12)
            # volumes interpolated upwards with the temperature index (:412)
        vols = [float(r0[k][8]) for k in range(3)]
        assert vols[0] < vols[1] < vols[2]
    finally:
        os.chdir(cwd)


def _two_rank_worker(rank, port, tmp, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NM_DIST_BACKEND="gloo")
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    try:
        import torch.distributed as dist
        from neuralmelting_b200 import remcmc
        os.chdir(tmp)
        args = remcmc.parse_args("-n s2 -ss 4 -pn 2 -tn 3 -sn 4 -sc 1 -sm 6 -bm -rd 2 -dn".split())
        counters, swaps = remcmc.run(args, log=lambda *a: None)
        args2 = remcmc.parse_args("-n s3 -r -rn s2 -rs 2 -ss 4 -pn 2 -tn 3 -sn 2 -sm 4 -bm".split())
        remcmc.run(args2, log=lambda *a: None)
        dist.destroy_process_group()
        q.put((rank, "ok", swaps))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc()), -1))


def test_two_rank_run_writes_the_same_files_as_one_rank(nm, tmp_path):
    """driver level: remcmc.run on two ranks (rows dealt cyclically; both ranks share this GPU, gloo collectives) writes
    byte-identical .thrm / .traj / .npy / restart files to a single-rank run, and a two-rank restart continues the same way"""
    import torch.multiprocessing as mp
    from neuralmelting_b200 import remcmc
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir(); two.mkdir()
    cwd = os.getcwd()
    os.chdir(one)
    try:
        _, swaps1 = remcmc.run(remcmc.parse_args("-n s2 -ss 4 -pn 2 -tn 3 -sn 4 -sc 1 -sm 6 -bm -rd 2 -dn".split()), log=lambda *a: None)
        remcmc.run(remcmc.parse_args("-n s3 -r -rn s2 -rs 2 -ss 4 -pn 2 -tn 3 -sn 2 -sm 4 -bm".split()), log=lambda *a: None)
    finally:
        os.chdir(cwd)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_two_rank_worker, args=(r, port, str(two), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    assert res[0][2] == swaps1
    names = sorted(f for f in os.listdir(one))
    assert names == sorted(os.listdir(two))
    for f in names:
        a, b = open(one / f, "rb").read(), open(two / f, "rb").read()
        if ".rstrt." in f:
            ra, rb = np.load(one / f, allow_pickle=True), np.load(two / f, allow_pickle=True)
            assert ra.shape == rb.shape
            for k in range(ra.shape[0]):
                for c in range(ra.shape[1]):
                    assert np.array_equal(np.asarray(ra[k][c]), np.asarray(rb[k][c])), (f, k, c)
        else:
            assert a == b, f
