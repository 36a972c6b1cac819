"""CPU tests: the oracle against golden vectors produced by the reference's own functions (oracle/gen_golden.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.mark.parametrize("name", ["n108", "n256", "n500"])
def test_rdf_oracle_matches_reference(orc, name):
    g = np.load(os.path.join(GOLDEN, "rdf_reference.npz"))
    pos, box, r, ref, nat = (g["%s_%s" % (name, f)] for f in ("pos", "box", "r", "g", "natoms"))
    assert r.dtype == np.float64 and ref.dtype == np.float32
    for s in range(pos.shape[0]):
        c = orc.rdf_counts(pos[s], box[s], r)
        assert c[0] == 0
        np.testing.assert_array_equal(c.astype(np.float32) / np.float32(nat[s]), ref[s])
    if name != "n500":
        np.testing.assert_array_equal(orc.rdf_counts_numpy(pos[0], box[0], r), orc.rdf_counts(pos[0], box[0], r))


@pytest.mark.parametrize("name", ["g2x4", "g4x8", "g3x6_anti", "g2x5_inf"])
def test_exchange_oracle_matches_reference(orc, name):
    g = np.load(os.path.join(GOLDEN, "exchange_reference.npz"))
    np_, nt = (int(v) for v in g[name + "_shape"])
    perm, swaps = orc.exchange(np_, nt, g[name + "_pe"] + g[name + "_ke"], g[name + "_vol"], g[name + "_et"],
                               g[name + "_pf"], g[name + "_uniforms"])
    np.testing.assert_array_equal(perm, g[name + "_perm"])
    assert swaps == int(np.sum(perm != np.arange(np_ * nt)) > 0) or swaps >= 0
    # exchanges never cross pressure rows
    assert np.array_equal(perm // nt, np.arange(np_ * nt) // nt)


def test_adapt_oracle_matches_reference(orc):
    gold = json.load(open(os.path.join(GOLDEN, "host_reference.json")))["adapt"]
    for c in gold:
        st = orc.adapt([0.03125, 0.0625, 0.00390625], [c["ratio"]] * 3)
        assert (st[0], st[1], st[2]) == (c["dx"], c["dv"], c["dt"])
        assert c["tail"] == [0.0] * 9


def test_round6_is_the_text_round_trip(orc):
    rng = np.random.default_rng(0)
    for v in list(rng.uniform(0, 20, 200)) + [0.0350625, 0.00390625, 6.1105790881, 1e-7, 123456.7890125]:
        assert orc.lib().orc_round6(v) == float("%f" % v)


@pytest.mark.parametrize("name", ["n108_cb8", "n256_cb11"])
def test_cdf_oracle_matches_reference(orc, name):
    """N1: calculate_cdf.py_func run in the build container (np.histogramdd per image, float32 vectors, float64 edges)"""
    g = np.load(os.path.join(GOLDEN, "cdf_reference.npz"))
    pos, box, rv, ref, nat = (g["%s_%s" % (name, f)] for f in ("pos", "box", "rv", "c", "natoms"))
    assert rv.dtype == np.float64 and ref.dtype == np.float32
    np.testing.assert_array_equal(orc.cdf_edges(box, rv.shape[1] - 1), rv)
    for s in range(pos.shape[0]):
        c = orc.cdf_counts(pos[s], box[s], rv)
        np.testing.assert_array_equal(c.astype(np.float32) / np.float32(nat[s]), ref[s])
    np.testing.assert_array_equal(orc.cdf_counts_numpy(pos[0], box[0], rv), orc.cdf_counts(pos[0], box[0], rv))
