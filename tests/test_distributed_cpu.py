"""world_size-2 gloo tests (CPU) of the multi-rank host logic: pressure rows dealt out cyclically (row u -> rank u mod G),
the asynchronous all-gather of the (pe + ke, vol) table into global slot order, per-replica streamed output written by both
ranks and consolidated by rank 0 in (P, T) order, and the multi-rank restart dump."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


class _StubEngine:
    """what TableGather needs of an engine, on the CPU: n_rep, nt and exchange_pack into a host tensor"""

    def __init__(self, n_rep, nt, table):
        self.n_rep, self.nt, self.table = n_rep, nt, np.ascontiguousarray(table)

    def exchange_pack(self, ptr):
        import ctypes
        ctypes.memmove(ptr, self.table.ctypes.data, self.table.nbytes)


def _worker(rank, world, port, tmp, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from neuralmelting_b200 import remcmc
    from oracle import oracle as orc
    comm = remcmc.Comm()
    try:
        npn, nt, natoms = 4, 5, 6
        rows = comm.row_shard(npn)
        assert rows == [rank, rank + 2]
        ns, nloc = npn * nt, len(rows) * nt
        gs = np.array([u * nt + j for u in rows for j in range(nt)])
        rng = np.random.default_rng(5)
        etot, vol = rng.normal(-1500, 30, ns), rng.normal(280, 5, ns)          # the job-wide truth
        P, T = remcmc.grids(1, 8, npn, 0.25, 2.5, nt)
        et, pf = remcmc.init_constants(P, T)
        # every rank packs its local slots; the asynchronous all-gather delivers the job-wide table in global slot order
        tg = remcmc.TableGather(comm, _StubEngine(nloc, nt, np.stack([etot[gs], vol[gs]], 1)), torch)
        tg.start()
        table = tg.latest()
        tg.finish()
        assert np.array_equal(table[:, 0], etot) and np.array_equal(table[:, 1], vol)
        # the sweep of a row needs only that row: the rows of this rank, swept from the job-wide uniforms, agree with the
        # job-wide sweep restricted to them
        u = orc.exchange_uniforms(256, 7, npn * nt * (nt - 1) // 2)
        perm, swaps = orc.exchange(npn, nt, table[:, 0], table[:, 1], et, pf, u)
        assert all(p // nt == k // nt for k, p in enumerate(perm))              # swaps stay inside a pressure row
        per = nt * (nt - 1) // 2
        for r in rows:
            sl = slice(r * nt, (r + 1) * nt)
            perm_r, _ = orc.exchange(1, nt, etot[sl], vol[sl], et[sl], pf[sl], u[r * per:(r + 1) * per])
            assert np.array_equal(perm_r + r * nt, perm[sl])
        # streamed per-replica output from both ranks, consolidation on rank 0 in (pressure, temperature) order
        os.chdir(tmp)
        args = remcmc.parse_args(["-n", "c", "-pn", str(npn), "-tn", str(nt)])
        pref = remcmc.file_prefix("c", "LJ")
        nrec = 2
        if rank == 0:
            remcmc.StreamWriter(args, pref, [], npn, nt, natoms, nrec, [], direct_npy=True, create=True).close()
        comm.barrier()
        w = remcmc.StreamWriter(args, pref, gs, npn, nt, natoms, nrec, [("# h%d\n" % k).encode() for k in gs], direct_npy=True, create=False)
        for s in range(nrec):
            th = np.zeros((nloc, 18)); th[:, 0] = gs + 100 * s
            w.put(th, np.full(nloc, 7.0 + s), np.repeat((gs + 100.0 * s)[:, None], 3 * natoms, 1))
        w.close()
        comm.barrier()
        if rank == 0:
            remcmc.consolidate_outputs(args, pref, npn, nt)
            lines = open(pref + ".thrm").read().split("\n")
            assert [l for l in lines if l.startswith("#")] == ["# h%d" % k for k in range(ns)]
            cols = np.loadtxt(pref + ".thrm", dtype=np.float32).reshape(npn, nt, nrec, 17)
            assert np.array_equal(cols[..., 0], np.arange(ns).reshape(npn, nt, 1) + 100 * np.arange(nrec))
            data = [l.split() for l in open(pref + ".traj")]
            assert len([v for v in data if len(v) == 2]) == ns * nrec and len([v for v in data if len(v) == 3]) == ns * nrec * natoms
            pos = np.load(pref + ".pos.npy")
            assert pos.shape == (npn, nt, nrec, natoms, 3) and np.array_equal(pos[..., 0, 0], cols[..., 0])
            assert np.array_equal(np.load(pref + ".box.npy").reshape(npn, nt, nrec)[0, 0], [7.0, 8.0])
            assert not [f for f in os.listdir(tmp) if ".0" in f and f.endswith((".thrm", ".traj"))]     # per-replica parts removed
        # multi-rank restart dump: one object array in global slot order
        st = dict(x=np.repeat(gs[:, None] * 1.0, 3 * natoms, 1), v=np.zeros((nloc, 3 * natoms)), box=gs + 0.5, dx=gs * 1.0, dv=gs * 2.0, dt=gs * 3.0)
        remcmc._gather_and_dump(comm, pref, 3, natoms, st, np.zeros((nloc, 18)), gs, ns)
        if rank == 0:
            _, x, v, box, dx, dv, dt = remcmc.load_restart(pref + ".rstrt.0003.npy")
            assert np.array_equal(box, np.arange(ns) + 0.5) and np.array_equal(x[:, 0], np.arange(ns)) and np.array_equal(dt, 3.0 * np.arange(ns))
        q.put((rank, "ok", swaps))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc()), -1))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_and_consolidation(orc, nm, tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    assert res[0][2] == res[1][2]


def test_row_shard_is_cyclic_and_rejects_uneven_split():
    sys.path.insert(0, ROOT)
    from neuralmelting_b200 import remcmc
    c = remcmc.Comm.__new__(remcmc.Comm)
    c.rank, c.world, c.dist = 1, 3, None
    with pytest.raises(ValueError):
        c.row_shard(32)
    c.world = 4
    assert c.row_shard(32) == [1, 5, 9, 13, 17, 21, 25, 29]
