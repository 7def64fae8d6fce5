"""N>1 host logic on CPU: two and three gloo ranks (even and ragged view ownership) run the multi-GPU expansion driver
(densepoints_b200/distributed.py) over an oracle-backed level backend; both must end with
the store and grids of the single-process 1-thread FIFO."""
import os
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from densepoints_b200 import distributed as dd
from densepoints_b200 import scenes

CELL = 5


def _scene(only_views=None):
    sc = scenes.make_plane_scene(seed=5, n_views=4, width=160, height=120, yaw_spread_deg=14.0,
                                 only_views=only_views)
    seeds = scenes.make_seeds(sc, 40, seed=6, depth_noise=0.004, tilt_deg=4.0)
    return sc, seeds


class OracleLevelBackend:
    """Same three steps as CudaLevelBackend, computed with the CPU oracle."""

    def __init__(self, orc, V, prm, seeds, nvis, vis):
        self.orc, self.V, self.prm = orc, V, prm
        self.org = orc.Organizer(V, prm)
        self.org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
        self.begin = 0
        self.words = 9 + V.n

    def frontier(self):
        return self.begin, self.org.size()

    def local(self, cell_size, rank, world, rov, max_records):
        st = self.org.export()
        rows = []
        own = None
        if world > 1 and rov is None:        # equal-work contiguous ranges (dp_parent_flags_kernel)
            w = np.where(st["nvis"][self.begin:] >= 2, st["nvis"][self.begin:], 0).astype(np.int64)
            scan = np.cumsum(w) - w
            own = ((scan * world * 8) // max(int(w.sum()), 1)) % world      # DP_RANGE_INTERLEAVE = 8
        for i in range(self.begin, self.org.size()):
            if st["nvis"][i] < 2:
                continue
            if world > 1 and (own[i - self.begin] != rank if own is not None
                              else rov[st["ref"][i]] != rank):
                continue
            kids = self.orc.expand_patch(self.V, self.prm, cell_size, st["pos"][i], st["nrm"][i],
                                         st["ref"][i], st["vis"][i, :st["nvis"][i]])
            for d, p, n, v in kids:
                row = np.full(self.words, -1, np.int32)
                row[0] = (i - self.begin) * 4 + d
                row[1] = st["ref"][i]
                row[2] = len(v)
                row[3:6] = p.view(np.int32)
                row[6:9] = n.view(np.int32)
                row[9:9 + len(v)] = v
                rows.append(row)
        buf = torch.from_numpy(np.array(rows, np.int32).reshape(-1, self.words))
        return buf, len(rows)

    def frontier_weights(self):
        st = self.org.export()
        w = np.zeros(self.V.n, np.int64)
        for i in range(self.begin, self.org.size()):
            if st["nvis"][i] >= 2:
                w[st["ref"][i]] += st["nvis"][i]
        return w

    def commit_gathered(self, recv, world, capacity, counts):
        parts = [recv[r * capacity: r * capacity + c] for r, c in enumerate(counts) if c > 0]
        total = int(sum(counts))
        records = torch.cat(parts) if parts else recv[:0]
        return self.commit(records, total)

    def commit(self, records, total):
        n0 = self.org.size()
        rec = records.numpy()[:total]
        for row in rec[np.argsort(rec[:, 0], kind="stable")]:
            self.org.try_insert(row[3:6].copy().view(np.float32), row[6:9].copy().view(np.float32),
                                int(row[1]), row[9:9 + row[2]])
        self.begin = n0
        return self.org.size() - n0


def _worker(rank, world, port, outdir, balance, m=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    # every rank renders its slice of the views, then the images are exchanged
    sc, seeds = _scene(only_views=dd.views_of_rank(4, rank, world))
    assert not sc.images[(rank + 1) % world].any()       # a view this rank did not render
    dd.share_images(sc.images, rank, world)
    full, _ = _scene()
    assert all(np.array_equal(a, b) for a, b in zip(sc.images, full.images))
    V = orc.Views(sc.P, sc.images)
    prm = orc.default_params(minimum_visible_image=2, max_patches_per_cell=m)
    nvis, vis, _, _ = orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
    be = OracleLevelBackend(orc, V, prm, seeds, nvis, vis)
    # fixed contiguous ownership, or re-balanced every level from the frontier (assign_views)
    # (equal-work ranges of the frontier: ownership="ranges", the table stays None)
    rov = None if balance else dd.partition_views(seeds["ref"], sc.n_views, world)
    stats = dd.expand_distributed(be, CELL, -1, rank, world, rov,
                                  ownership="ranges" if balance == 2 else "views")
    ex = be.org.export()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), **ex,
             grids=np.concatenate([be.org.grid(v).ravel() for v in range(sc.n_views)]),
             local=stats["local_records"], passed=stats["passed"])
    dist.destroy_process_group()


def test_partition_views_is_contiguous_and_balanced():
    ref = np.repeat(np.arange(8), [10, 10, 10, 10, 40, 40, 40, 40])
    rov = dd.partition_views(ref, 8, 2)
    assert (np.diff(rov) >= 0).all() and set(rov) == {0, 1}
    load = [np.isin(ref, np.where(rov == r)[0]).sum() for r in range(2)]
    assert max(load) <= 0.65 * len(ref)
    assert (dd.partition_views(ref, 8, 1) == 0).all()


import pytest


def test_assign_views_is_deterministic_and_balanced():
    w = np.array([50, 0, 7, 7, 30, 30, 1, 25], np.int64)
    rov = dd.assign_views(w, 3)
    assert np.array_equal(rov, dd.assign_views(w.copy(), 3)) and set(rov) <= {0, 1, 2}
    load = np.array([w[rov == r].sum() for r in range(3)])
    assert load.max() <= 1.34 * w.sum() / 3            # LPT bound 4/3 - 1/(3m)
    assert (dd.assign_views(w, 1) == 0).all()


@pytest.mark.parametrize("world,balance,m", [(2, 0, 1), (3, 0, 1), (2, 1, 1), (3, 1, 1), (2, 2, 1), (3, 2, 1),
                                             (2, 2, 2)])     # m: max_patches_per_cell
def test_multi_rank_expansion_matches_single_process_fifo(orc, world, balance, m):
    sc, seeds = _scene()
    V = orc.Views(sc.P, sc.images)
    prm = orc.default_params(minimum_visible_image=2, max_patches_per_cell=m)
    nvis, vis, _, _ = orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
    ref_org = orc.Organizer(V, prm)
    ref_org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    n_seed = ref_org.size()
    ref_org.expand_fifo(CELL)
    want = ref_org.export()
    assert ref_org.size() > n_seed
    with tempfile.TemporaryDirectory() as d:
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_worker, args=(world, port + world + 10 * int(balance) + 40 * (m - 1), d, balance, m), nprocs=world,
                 join=True)
        got = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
    grids = np.concatenate([ref_org.grid(v).ravel() for v in range(sc.n_views)])
    for g in got:
        for k in want:
            assert np.array_equal(g[k], want[k]), k
        assert np.array_equal(g["grids"], grids)
    # the work really was split: no rank produced all the records, together they produced all
    assert all(int(g["local"]) < int(g["passed"]) for g in got)
    assert sum(int(g["local"]) for g in got) == int(got[0]["passed"])
    assert sum(int(g["local"]) > 0 for g in got) >= 2
