"""Host driver of the structural-feature stage: the B200 drop-in for the RDF half of
/root/reference/scripts/lammps_distr.py.

Same flags (lammps_distr.py:18-47; the Dask/joblib/PBS ones are accepted and ignored), same inputs
(<prefix>.natoms/.box/.pos.npy written by lammps_parse.py, :63-70), same outputs <prefix>.dni.npy,
<prefix>.r.npy, <prefix>.rdf.npy (:336-338) with the same dtypes. The 27-image all-pairs float32 distance
histogram of calculate_rdf (:123-135) runs in the CUDA kernel of csrc/nm_rdf.cu, bit-exact bin counts.
The Cartesian pair-vector density (calculate_cdf, :161-171, -cb) follows in the same way (<prefix>.dn/.rv/.cdf.npy, :368-370).
"""
import argparse
import os
import sys

import numpy as np

from . import engine as nm

LAT = {"Ti": "bcc", "Al": "fcc", "Ni": "fcc", "Cu": "fcc", "LJ": "fcc"}

_FLAGS = [
    ("-v", "--verbose", dict(action="store_true", help="verbose output")),
    ("-p", "--parallel", dict(action="store_true", help="parallel run (accepted, ignored: GPU kernel)")),
    ("-c", "--client", dict(action="store_true", help="dask client run mode (accepted, ignored)")),
    ("-d", "--distributed", dict(action="store_true", help="distributed run (accepted, ignored)")),
    ("-q", "--queue", dict(type=str, default="jobqueue", help="submission queue (ignored)")),
    ("-a", "--allocation", dict(type=str, default="startup", help="submission allocation (ignored)")),
    ("-nn", "--nodes", dict(type=int, default=1, help="number of nodes (ignored)")),
    ("-np", "--procs_per_node", dict(type=int, default=16, help="number of processors per node (ignored)")),
    ("-w", "--walltime", dict(type=int, default=2, help="job walltime (ignored)")),
    ("-m", "--memory", dict(type=int, default=32, help="total job memory (ignored)")),
    ("-nw", "--workers", dict(type=int, default=16, help="total job worker count (ignored)")),
    ("-nt", "--threads", dict(type=int, default=1, help="threads per worker (ignored)")),
    ("-mt", "--method", dict(type=str, default="fork", help="parallelization method (ignored)")),
    ("-n", "--name", dict(type=str, default="remcmc_init", help="name of simulation")),
    ("-e", "--element", dict(type=str, default="LJ", help="element choice")),
    ("-sb", "--spherical_bins", dict(type=int, default=64, help="number of bins for spherical distributions")),
    ("-cb", "--cartesian_bins", dict(type=int, default=16, help="number of bins for cartesian distributions")),
]


def build_parser():
    parser = argparse.ArgumentParser(description="RDF structural features on B200 (lammps_distr.py drop-in)")
    for short, long_, kw in _FLAGS:
        parser.add_argument(short, long_, **kw)
    return parser


def spatial_setup(natoms, box, sbins):
    """the RDF part of calculate_spatial (lammps_distr.py:73-98): edges R (float64, scaled by the minimum box over
    ALL samples) and the ideal-gas shell populations DNI"""
    nrho = np.divide(natoms, np.power(box, 3))
    l = np.min(box)
    r = np.linspace(1e-16, 1 / 2, sbins)
    dr = r[1] - r[0]
    dv = 4 * np.pi * np.square(r) * dr
    r = r * l
    dv = dv * l ** 3
    dni = np.multiply(nrho[:, np.newaxis], dv[np.newaxis, :])
    return r, dni


def calculate_rdfs(natoms, box, pos, r, device=0, batch=4096, verbose=False):
    """calculate_rdf for every sample (lammps_distr.py:174-192): float32 array (S, SBINS) = counts / natoms"""
    ns = natoms.size
    out = np.empty((ns, r.size), dtype=np.float32)
    for s0 in range(0, ns, batch):
        s1 = min(ns, s0 + batch)
        counts = nm.rdf_counts(pos[s0:s1], box[s0:s1], r, device=device)
        out[s0:s1] = counts.astype(np.float32) / natoms[s0:s1, None].astype(np.float32)
        if verbose:
            print("rdf: %d / %d samples" % (s1, ns))
    return out


def cartesian_setup(natoms, box, cbins):
    """the CDF part of calculate_spatial (lammps_distr.py:108-116): edges RV (3, CBINS+1) float64 centred on zero and
    the ideal-gas voxel population DN. RV is built in float64 as under the numpy-1 rules the script was written for."""
    nrho = np.divide(natoms, np.power(box, 3))
    l = float(np.min(box))
    rv = np.array([np.linspace(0, l, cbins + 1) for _ in range(3)], dtype=np.float64)
    rv -= l / 2
    drv = rv[0, 1] - rv[0, 0]
    dn = nrho * drv ** 3
    return rv, dn


def calculate_cdfs(natoms, box, pos, rv, device=0, batch=1024, verbose=False):
    """calculate_cdf for every sample (lammps_distr.py:216-234): float32 array (S, CB, CB, CB) = counts / natoms"""
    ns, cb = natoms.size, rv.shape[1] - 1
    out = np.empty((ns, cb, cb, cb), dtype=np.float32)
    for s0 in range(0, ns, batch):
        s1 = min(ns, s0 + batch)
        counts = nm.cdf_counts(pos[s0:s1], box[s0:s1], rv, device=device)
        out[s0:s1] = counts.astype(np.float32) / natoms[s0:s1, None, None, None].astype(np.float32)
        if verbose:
            print("cdf: %d / %d samples" % (s1, ns))
    return out


def run(args):
    prefix = os.path.join(os.getcwd(), "%s.%s.%s.lammps" % (args.name, args.element.lower(), LAT[args.element]))
    P = np.load(prefix + ".virial.trgt.npy")
    T = np.load(prefix + ".temp.trgt.npy")
    pn, tn = P.size, T.size
    natoms = np.load(prefix + ".natoms.npy").reshape(-1)
    box = np.load(prefix + ".box.npy").reshape(-1)
    pos = np.load(prefix + ".pos.npy").reshape(-1, natoms[0], 3)
    ns = natoms.size
    rns = np.int32(ns / (pn * tn))
    r, dni = spatial_setup(natoms, box, args.spherical_bins)
    if args.verbose:
        print("computing %s %s samples" % (ns, args.element.lower()))
    g = calculate_rdfs(natoms, box, pos, r, verbose=args.verbose)
    g = np.divide(g, dni)                                     # lammps_distr.py:319
    dni = dni.reshape(pn, tn, rns, r.size)
    g = g.reshape(pn, tn, rns, r.size)
    np.save(prefix + ".dni.npy", dni)
    np.save(prefix + ".r.npy", r)
    np.save(prefix + ".rdf.npy", g)
    rv, dn = cartesian_setup(natoms, box, args.cartesian_bins)
    c = calculate_cdfs(natoms, box, pos, rv, verbose=args.verbose)
    c = np.divide(c, dn[:, np.newaxis, np.newaxis, np.newaxis])      # lammps_distr.py:365
    c = c.reshape(pn, tn, rns, *(3 * (rv.shape[1] - 1,)))
    np.save(prefix + ".dn.npy", dn)
    np.save(prefix + ".rv.npy", rv)
    np.save(prefix + ".cdf.npy", c)
    if args.verbose:
        print("calculations finalized")
    return g


def main(argv=None):
    run(build_parser().parse_args(argv))


if __name__ == "__main__":
    main(sys.argv[1:])
