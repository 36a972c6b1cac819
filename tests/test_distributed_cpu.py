"""world_size-2 gloo tests (CPU) of the multi-rank host logic: pressure-row sharding, the all-gathered exchange table
replayed identically on every rank, and rank-0 consolidation of the per-replica output in (P, T) order."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, tmp, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from neuralmelting_b200 import remcmc
    from oracle import oracle as orc
    comm = remcmc.Comm()
    try:
        npn, nt = 4, 5
        row0, nrow = comm.row_shard(npn)
        assert (row0, nrow) == (rank * 2, 2)
        ns, nloc, off = npn * nt, nrow * nt, row0 * nt
        rng = np.random.default_rng(5)
        etot, vol = rng.normal(-1500, 30, ns), rng.normal(280, 5, ns)          # the job-wide truth
        P, T = remcmc.grids(1, 8, npn, 0.25, 2.5, nt)
        et, pf = remcmc.init_constants(P, T)
        # each rank packs its local slots, one all-gather builds the job-wide table (16 bytes per replica)
        local = torch.tensor(np.stack([etot[off:off + nloc], vol[off:off + nloc]], 1))
        full = torch.empty((ns, 2), dtype=torch.float64)
        dist.all_gather_into_tensor(full, local)
        table = full.numpy()
        assert np.array_equal(table[:, 0], etot) and np.array_equal(table[:, 1], vol)
        # every rank replays the same sweep from the same counter-based uniforms -> identical permutation
        u = orc.exchange_uniforms(256, 7, npn * nt * (nt - 1) // 2)
        perm, swaps = orc.exchange(npn, nt, table[:, 0], table[:, 1], et, pf, u)
        perms = [None, None]
        dist.all_gather_object(perms, perm.tolist())
        assert perms[0] == perms[1]
        assert all(p // nt == k // nt for k, p in enumerate(perm))              # swaps stay inside a pressure row
        # consolidation on rank 0 in (pressure, temperature) order
        os.chdir(tmp)
        thrm = [[("h%d\n" % (off + k)).encode(), b"x\n"] for k in range(nloc)]
        traj = [[("t%d\n" % (off + k)).encode()] for k in range(nloc)]
        remcmc._consolidate(comm, os.path.join(tmp, "c"), thrm, traj)
        comm.barrier()
        if rank == 0:
            text = open(os.path.join(tmp, "c.thrm")).read().split()
            assert [w for w in text if w.startswith("h")] == ["h%d" % k for k in range(ns)]
            assert open(os.path.join(tmp, "c.traj")).read().split() == ["t%d" % k for k in range(ns)]
        q.put((rank, "ok", swaps))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc()), -1))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_and_consolidation(orc, tmp_path):
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


def test_row_shard_rejects_uneven_split():
    sys.path.insert(0, ROOT)
    from neuralmelting_b200 import remcmc
    c = remcmc.Comm.__new__(remcmc.Comm)
    c.rank, c.world, c.dist = 1, 3, None
    with pytest.raises(ValueError):
        c.row_shard(32)
    c.world = 4
    assert c.row_shard(32) == (8, 8)
