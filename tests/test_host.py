"""CPU tests of the host side: flag surface, grids, constants, text formats (native formatter), restart files,
and that the C-ABI library loads and exports every symbol include/nm_b200.h declares (no GPU work)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLDEN, "host_reference.json")))


def test_library_exports_every_declared_symbol(nm):
    header = open(os.path.join(ROOT, "include", "nm_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)       # prototypes only, not the prose
    names = set(re.findall(r"\b(nm_[a-z_0-9]+)\s*\(", body))
    assert len(names) >= 25
    lib = ctypes.CDLL(nm.LIB_PATH)
    for name in sorted(names):
        assert hasattr(lib, name), "libnm_b200.so does not export %s" % name
    assert lib.nm_abi_version() == 2


def test_no_gpu_is_a_loud_error(nm):
    """the product never falls back to the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nm.NmError) as ei:
        nm.Engine(natoms=256, n_rep=1, nt=1)
    assert ei.value.code == nm.NM_ENODEV
    with pytest.raises(nm.NmError):
        nm.rdf_counts(np.zeros((1, 4, 3), np.float32), np.ones(1, np.float32), np.linspace(0.1, 0.5, 8))


def test_product_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "neuralmelting_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.replace("oracle/nm_oracle.c", "").lower() or f == "nm_device.cuh", f


def test_flag_surface_matches_reference(gold):
    from neuralmelting_b200 import distr, remcmc
    a = remcmc.parse_args([])
    order = ["verbose", "restart", "parallel", "client", "distributed", "interpolate_states", "bulk_move", "restart_dump",
             "restart_name", "restart_step", "queue", "allocation", "nodes", "procs_per_node", "walltime", "memory", "workers",
             "threads", "method", "name", "element", "supercell_size", "pressure_number", "pressure_range", "temperature_number",
             "temperature_range", "sample_cutoff", "sample_number", "sample_mod", "position_move", "volume_move", "timesteps",
             "pos_displace", "vol_displace"]
    flat = []
    for k in order:
        v = getattr(a, k)
        flat += list(v) if isinstance(v, list) else [v]
    assert flat == gold["remcmc_defaults"]
    d = distr.build_parser().parse_args([])
    dflat = [d.verbose, d.parallel, d.client, d.distributed, d.queue, d.allocation, d.nodes, d.procs_per_node, d.walltime,
             d.memory, d.workers, d.threads, d.method, d.name, d.element, d.spherical_bins, d.cartesian_bins]
    assert dflat == gold["distr_defaults"]
    b = remcmc.parse_args("-v -r -bm -is -n run -ss 10 -pn 32 -tn 32 -pr 2 9 -tr .5 3 -sc 7 -sn 9 -sm 3 -pm .75 -vm .1 -ts 4 -dx .01 -dv .02 -rd 5 -rn a -rs 6 -c -nw 16 -nt 1 -mt fork".split())
    assert (b.bulk_move, b.supercell_size, b.pressure_range, b.sample_mod, b.position_move) == (True, 10, [2.0, 9.0], 3, 0.75)


def test_grids_and_constants():
    from neuralmelting_b200 import remcmc
    P, T = remcmc.grids(1, 8, 4, 0.25, 2.5, 8)
    assert P.dtype == np.float32 and T.dtype == np.float32
    et, pf = remcmc.init_constants(P, T)
    assert et.dtype == np.float64 and et.shape == (32,)
    k = 13
    i, j = divmod(k, 8)
    assert et[k] == float(T[j]) and pf[k] == float(P[i]) / float(T[j])


def test_text_formats_byte_identical_to_reference(nm, gold):
    """write_thrm / write_traj / init_header of the reference (golden text) against the native formatter"""
    from neuralmelting_b200 import remcmc
    st = gold["state_scalars"]
    x = np.array(gold["state_x"])
    thermo = np.array([st[3], st[4], st[5], st[6], st[7], st[8], st[9], st[10], st[11]] + st[12:21])
    args = remcmc.parse_args([])
    P, T = remcmc.grids(1, 8, 4, 0.25, 2.5, 8)
    i, j = divmod(gold["header_k"], 8)
    text = remcmc.header_text(args, P[i], T[j], 1024, 0, 128, 0.00390625).encode() + remcmc.thrm_line(thermo)
    assert text.decode() == gold["thrm_text"]
    rec = remcmc.traj_records(int(st[0]), np.array([st[7]]), x.reshape(1, -1), nthreads=2)[0]
    assert rec.decode() == gold["traj_text"]
    assert nm.format_traj(int(st[0]), st[7], x).decode() == gold["traj_text"]
    # python's '%.4E' and the native formatter agree on awkward values
    vals = np.array([0.0, -0.0, 1e-300, 9.99995e4, 9.99994999e4, 1.00005, 0.99995, -1234.56789, 5e-5, 1e100, 2.5e-5, 1.23455e-7, 3.0])
    for v in vals:
        assert nm.format_thrm(np.full(17, v)).decode() == 17 * " %.4E" % tuple([v] * 17) + "\n"
    rng = np.random.default_rng(0)
    big = rng.normal(0, 1, 17 * 200) * 10.0 ** rng.integers(-8, 8, 17 * 200)
    for row in big.reshape(-1, 17):
        assert nm.format_thrm(row).decode() == 17 * " %.4E" % tuple(row) + "\n"


def test_restart_round_trip(tmp_path):
    from neuralmelting_b200 import remcmc
    rng = np.random.default_rng(1)
    ns, n = 6, 8
    state = dict(x=rng.normal(size=(ns, 3 * n)), v=rng.normal(size=(ns, 3 * n)), box=rng.uniform(6, 7, ns),
                 dx=rng.uniform(.01, .05, ns), dv=rng.uniform(.01, .05, ns), dt=rng.uniform(.003, .005, ns))
    th = rng.normal(size=(ns, 18))
    path = str(tmp_path / "a.rstrt.0003.npy")
    remcmc.dump_restart(path, n, state, th)
    arr = np.load(path, allow_pickle=True)
    assert arr.shape == (ns, 21) and arr.dtype == object           # the reference's np.array(STATE, dtype=object)
    assert arr[2][0] == n and arr[2][8] == state["box"][2] ** 3 and all(v == 0 for v in arr[2][12:])
    natoms, x, v, box, dx, dv, dt = remcmc.load_restart(path)
    assert natoms == n
    for a, b in ((x, state["x"]), (v, state["v"]), (box, state["box"]), (dx, state["dx"]), (dt, state["dt"])):
        np.testing.assert_array_equal(a, b)


def test_reference_parse_script_consumes_our_files(nm, tmp_path):
    """acceptance consumer: the UNMODIFIED lammps_parse.py runs on the files the streaming writer produced (build container
    only), and the directly emitted .npy files (N2) are bit-identical to what the parser derives from the text"""
    ref = "/root/reference/scripts/lammps_parse.py"
    if not os.path.exists(ref):
        pytest.skip("reference not present (GPU box)")
    from neuralmelting_b200 import remcmc
    rng = np.random.default_rng(2)
    pn, tn, s, n = 2, 3, 4, 32
    args = remcmc.parse_args(["-n", "t", "-pn", str(pn), "-tn", str(tn), "-ss", "2"])
    P, T = remcmc.grids(1, 8, pn, 0.25, 2.5, tn)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        pref = remcmc.file_prefix("t", "LJ")
        np.save(pref + ".virial.trgt.npy", P)
        np.save(pref + ".temp.trgt.npy", T)
        ns = pn * tn
        heads = [remcmc.header_text(args, P[k // tn], T[k % tn], s, 0, 128, 0.00390625).encode() for k in range(ns)]
        w = remcmc.StreamWriter(args, pref, np.arange(ns), pn, tn, n, s, heads, nthreads=3, direct_npy=True)
        xs, boxes = np.empty((s, ns, 3 * n)), np.empty((s, ns))
        for q in range(s):
            th = np.abs(rng.normal(size=(ns, 18))) * 10.0 ** rng.integers(-6, 6, (ns, 18)) + 1e-7
            boxes[q] = rng.uniform(5, 6, ns)
            xs[q] = rng.uniform(0, 1, (ns, 3 * n)) * boxes[q][:, None] * 10.0 ** rng.integers(-4, 1, (ns, 3 * n))
            th[:, 4], th[:, 5] = boxes[q], boxes[q] ** 3
            w.put(th, boxes[q].copy(), xs[q].copy())
        w.close()
        remcmc.consolidate_outputs(args, pref, pn, tn)
        names = ("pos", "box", "natoms") + remcmc.THERMO_NAMES
        ours = {name: np.load(pref + ".%s.npy" % name) for name in names}
        for name in names:
            os.remove(pref + ".%s.npy" % name)
        subprocess.check_call([sys.executable, ref, "-n", "t"], cwd=str(tmp_path))
        for name in names:
            theirs = np.load(pref + ".%s.npy" % name)
            assert theirs.dtype == ours[name].dtype and theirs.shape == ours[name].shape, name
            np.testing.assert_array_equal(theirs, ours[name], err_msg=name)
        pos, natoms, box = ours["pos"], ours["natoms"], ours["box"]
        assert pos.shape == (pn, tn, s, n, 3) and pos.dtype == np.float32
        assert natoms.shape == (pn, tn, s) and natoms.dtype == np.uint16 and (natoms == n).all()
        assert box.shape == (pn * tn * s,) and ours["vol"].shape == (pn, tn, s)
        want = np.array([[float("%.4E" % v) for v in row] for row in xs.reshape(-1, 3 * n)], dtype=np.float32).reshape(s, pn, tn, n, 3)
        np.testing.assert_array_equal(pos, want.transpose(1, 2, 0, 3, 4))
    finally:
        os.chdir(cwd)


def test_streaming_writer_text_equals_the_one_shot_formatter(nm, tmp_path):
    """the appended per-replica records are byte-identical to format_thrm / format_traj (golden-pinned above); runs on the GPU box too"""
    from neuralmelting_b200 import remcmc
    rng = np.random.default_rng(4)
    ns, n = 5, 7
    box = rng.uniform(5, 9, ns)
    x = rng.normal(0, 3, (ns, 3 * n))
    vals = rng.normal(0, 1, (ns, 17)) * 10.0 ** rng.integers(-9, 9, (ns, 17))
    paths = [str(tmp_path / ("r%d.traj" % k)) for k in range(ns)]
    tpaths = [str(tmp_path / ("r%d.thrm" % k)) for k in range(ns)]
    for rep in range(2):
        pos, bx = nm.append_traj_batch(n, box, x, paths, nthreads=2, parse_back=True)
        parsed = nm.append_thrm_batch(vals, tpaths, parse_back=True)
    for k in range(ns):
        assert open(paths[k], "rb").read() == 2 * nm.format_traj(n, box[k], x[k])
        assert open(tpaths[k], "rb").read() == 2 * nm.format_thrm(vals[k])
    np.testing.assert_array_equal(pos.reshape(ns, -1), np.array([[np.float32(float("%.4E" % v)) for v in row] for row in x]))
    np.testing.assert_array_equal(bx, np.array([np.float32(float("%.4E" % b)) for b in box]))
    np.testing.assert_array_equal(parsed, np.array([[np.float32(float("%.4E" % v)) for v in row] for row in vals]))
    with pytest.raises(nm.NmError):
        nm.append_traj_batch(n, box, x, [str(tmp_path / "no_such_dir" / "a.traj")] * ns)


def test_distr_setup_matches_reference_golden():
    from neuralmelting_b200 import distr
    g = np.load(os.path.join(GOLDEN, "rdf_reference.npz"))
    r, dni = distr.spatial_setup(g["n256_natoms"], g["n256_box"], 64)
    np.testing.assert_array_equal(r, g["n256_r"])
    np.testing.assert_array_equal(dni, g["n256_dni"])
