"""CPU tests that pin the oracle's LJ restatement (parity unpinned by the reference: LAMMPS is absent) with
self-derived anchors: analytic fcc shell sums, three independent implementations, finite-difference forces."""
import numpy as np
import pytest

A0 = (4 / 1.122) ** (1 / 3)


def _liquid(orc, sz, rho, sigma, seed):
    rng = np.random.default_rng(seed)
    box = sz * (4 / rho) ** (1 / 3)
    x = orc.fcc_positions(sz, box) + rng.normal(0, sigma, (4 * sz ** 3, 3))
    return orc.wrap(x.reshape(-1), box).reshape(-1, 3), box


def test_fcc_anchors(orc):
    """E/N, W/(3V) of the perfect fcc crystal at rho* = 1.122, rc = 2.5 (SURVEY section 4 anchors, re-derived from shells)"""
    e, w, nn = orc.fcc_shell_sum(A0)
    assert nn == 78
    assert abs(e - (-8.034879297835)) < 1e-11
    for sz in (4, 5):
        box = sz * A0
        pe, W, f, npairs = orc.lj_eval_n2(orc.fcc_positions(sz, box), box)
        n = 4 * sz ** 3
        assert npairs == 39 * n
        assert abs(pe / n - e) < 1e-11 and abs(W / n - w) < 1e-10
        assert abs(W / (3 * box ** 3) - 3.540508883950) < 1e-10
        assert np.abs(f).max() < 1e-12


@pytest.mark.parametrize("sz,rho,sigma", [(4, 1.1, 0.05), (5, 0.9, 0.2), (5, 0.5, 0.5), (7, 0.8, 0.3)])
def test_three_implementations_agree(orc, sz, rho, sigma):
    x, box = _liquid(orc, sz, rho, sigma, seed=sz)
    a = orc.lj_eval_n2(x, box)
    b = orc.lj_eval_list(x, box)
    c = orc.lj_eval_numpy(x, box)
    assert a[3] == b[3] == c[3]
    scale = max(abs(a[0]), 1.0)
    assert abs(a[0] - b[0]) < 1e-13 * scale and abs(a[0] - c[0]) < 1e-12 * scale
    assert abs(a[1] - b[1]) < 1e-12 * max(abs(a[1]), scale) and abs(a[1] - c[1]) < 1e-11 * max(abs(a[1]), scale)
    fs = np.abs(a[2]).max()
    assert np.abs(a[2] - b[2]).max() < 1e-13 * fs and np.abs(a[2] - c[2]).max() < 1e-12 * fs


def test_forces_are_minus_gradient(orc):
    x, box = _liquid(orc, 4, 0.95, 0.1, seed=2)
    pe, w, f, _ = orc.lj_eval_n2(x, box)
    h = 1e-6
    rng = np.random.default_rng(0)
    for _ in range(6):
        i, c = rng.integers(256), rng.integers(3)
        xp, xm = x.copy(), x.copy()
        xp[i, c] += h
        xm[i, c] -= h
        fd = -(orc.lj_eval_n2(xp, box)[0] - orc.lj_eval_n2(xm, box)[0]) / (2 * h)
        assert abs(fd - f[i, c]) < 1e-5 * max(1.0, abs(f[i, c]))
    # virial = -dE/dlnV * 3 for a homogeneous scaling (no pair crosses the cutoff for a tiny strain is not guaranteed -> loose)
    assert np.abs(f.sum(0)).max() < 1e-10 * np.abs(f).max()


def test_single_atom_delta_equals_total_difference(orc):
    x, box = _liquid(orc, 4, 1.0, 0.08, seed=4)
    rng = np.random.default_rng(1)
    e0 = orc.lj_eval_n2(x, box)[0]
    for _ in range(8):
        k = int(rng.integers(256))
        xn = x[k] + rng.uniform(-0.1, 0.1, 3)
        xn -= np.floor(xn / box) * box
        y = x.copy()
        y[k] = xn
        assert abs(orc.lj_delta_atom(x, k, xn, box) - (orc.lj_eval_n2(y, box)[0] - e0)) < 1e-10


def test_cycle_invariants(orc):
    """one oracle cycle: counters add up, KE after a rejected HMC equals (3N-3)/2 T minus the rotation part, box is '%f'-rounded"""
    x, box = _liquid(orc, 4, 1.0, 0.03, seed=7)
    box = orc.round6(box)
    p = orc.make_params(mod=40, bulk_move=1, seed=3)
    xx, vv = x.reshape(-1).copy(), np.zeros(768)
    scal, counts = np.array([box, .03125, .03125, .00390625]), np.zeros(6)
    th, ct, tr = orc.cycle(p, [1.0, 2.0, 1.0, 1.0], 0, 0, xx, vv, scal, counts, trace=True)
    assert counts[0] + counts[2] + counts[4] == 40 == ct[orc.CT_SWEEPS]
    assert ct[orc.CT_HMC_ATOM_STEPS] == 256 * 8 * counts[4]
    assert abs(scal[0] * 1e6 - round(scal[0] * 1e6)) < 1e-6
    assert th[4] == scal[0] and abs(th[5] - scal[0] ** 3) < 1e-9
    assert 0 <= xx.min() and xx.max() < scal[0]
    dof = 3 * 256 - 3
    assert abs(th[0] - 2 * th[2] / dof) < 1e-12
    # pe of the trace's last row equals the reported pe
    assert tr[-1, 0] == th[1]


# LAMMPS examples/melt (in.melt: `lattice fcc 0.8442`, `region box block 0 10 0 10 0 10`, `velocity all create 3.0 87287`,
# `pair_style lj/cut 2.5`, `pair_coeff 1 1 1.0 1.0 2.5`), step-0 thermo line of the logs LAMMPS ships with the example
# (log.*.melt.g++.*): "Step Temp E_pair E_mol TotEng Press" = 0 3 -6.7733681 0 -2.2744931 -3.7033504 (per-atom energies,
# lj units). It pins energy, virial and the 3N-3 dof convention of thermo_temp / thermo_press to LAMMPS's own output.
MELT_RHO, MELT_T = 0.8442, 3.0
MELT_EPAIR, MELT_TOTENG, MELT_PRESS = -6.7733681, -2.2744931, -3.7033504


def melt_thermo(n, pe, w, box, temp):
    """thermo_pe / N, (pe + ke) / N and thermo_press with ke = dof/2 k_B T, dof = 3N - 3 (what `velocity create` leaves)"""
    dof = 3 * n - 3
    ke = 0.5 * dof * temp
    return pe / n, (pe + ke) / n, (dof * temp + w) / (3.0 * box ** 3)


def test_lammps_melt_example_step0_pins_the_oracle(orc):
    """third-party pin: the oracle reproduces every printed digit of LAMMPS's published step-0 line for this potential"""
    a = (4 / MELT_RHO) ** (1 / 3)
    n, box = 4000, 10 * a
    x = orc.fcc_positions(10, box)
    for ev in (orc.lj_eval_list, orc.lj_eval_n2):
        pe, w, f, npairs = ev(x, box)
        e_pair, toteng, press = melt_thermo(n, pe, w, box, MELT_T)
        assert abs(e_pair - MELT_EPAIR) < 5e-8
        assert abs(toteng - MELT_TOTENG) < 5e-8
        assert abs(press - MELT_PRESS) < 5e-8
        assert np.abs(f).max() < 1e-11
    e, wn, nn = orc.fcc_shell_sum(a)
    assert abs(e - MELT_EPAIR) < 5e-8 and npairs == n * nn // 2


def test_iterative_sweep_candidate_lists_are_bitwise_identical_to_the_all_atom_sum(orc):
    """the per-sweep candidate lists of the oracle's iter_position_mc (what the timed CPU baseline of C4 runs) visit the
    same non-zero terms in the same order as the O(N) loop: identical positions, energies and counters, bit for bit"""
    n_side, n = 6, 864
    rng = np.random.default_rng(3)
    box = orc.round6(n_side * (4 / 0.85) ** (1 / 3))
    x0 = orc.wrap((orc.fcc_positions(n_side, box) + rng.normal(0, 0.08, (n, 3))).reshape(-1), box)
    params = orc.make_params(mod=3, bulk_move=0, ppos=1.0, pvol=0.0, seed=99)
    label = np.array([1.2, 2.0 / 1.2, 1.2, 1.2])
    out = []
    for lists in (1, 0):
        orc.set_delta_lists(lists)
        x, v = x0.copy(), np.zeros(3 * n)
        scal = np.array([box, 0.09, 0.03125, 0.00390625])
        th, ct = orc.cycle(params, label, 5, 0, x, v, scal, np.zeros(6))
        out.append((x, th, ct))
    orc.set_delta_lists(1)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert 0 < out[0][1][10] < out[0][1][9]          # accepted and rejected trials
