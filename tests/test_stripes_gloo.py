"""CPU, world_size 2, gloo: the row-stripe sharding logic (broadcast of x,
stripe ownership, ragged all-gather of the kNN lists) with the compute
operators replaced by the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleOps(object):
    """Stand-in operators (CPU tensors) with the device operators' contract."""

    def prepare(self, x, metric):
        from oracle import sdm_oracle as orc
        return {"d": orc.pdist(x.numpy(), metric)}

    def pdist_stripe(self, cells, r0, r1, out=None):
        return torch.from_numpy(np.ascontiguousarray(cells["d"][r0:r1]))

    def knn_from_d(self, d, k, exclude, row_begin):
        d = d.numpy()
        order = np.argsort(d, kind="mergesort", axis=1)[:, :k + 1]
        rows = np.arange(d.shape[0])[:, None]
        if exclude == 0:
            idx = order[:, 1:]
        else:
            keep = order != (rows + row_begin)
            keep[keep.all(axis=1), 0] = False
            idx = order[keep].reshape(d.shape[0], k)
        return (torch.from_numpy(idx.astype(np.int64)),
                torch.from_numpy(np.take_along_axis(d, idx, axis=1)))

    def knn_stripe(self, cells, k, exclude, r0, r1):
        return self.knn_from_d(self.pdist_stripe(cells, r0, r1), k, exclude,
                               r0)


    # -- consumers: NumPy stand-ins with the device operators' contract -----
    def squareform_rows(self, d, n, row_begin, out=None):
        a = d.numpy()
        runs = [a[r, row_begin + r + 1:] for r in range(a.shape[0])]
        v = np.concatenate(runs) if runs else np.zeros(0)
        flags = torch.zeros(1, dtype=torch.int32)
        if np.isnan(v).any():
            flags |= 1
        if any(not (a[r, row_begin + r] == 0) for r in range(a.shape[0])):
            flags |= 2
        seg = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
        if out is not None:
            out.copy_(seg)
            seg = out
        return seg, flags

    def new_mask(self, n, device):
        return torch.ones(n, dtype=torch.uint8)

    def masked_kth(self, d, alive, rank, row_begin):
        d, a = d.numpy(), alive.numpy().astype(bool)
        out = np.full(d.shape[0], np.inf)
        for r in range(d.shape[0]):
            if a[row_begin + r]:
                out[r] = np.sort(d[r][a])[rank]
        return torch.from_numpy(out)

    def rare_update(self, kth, cutoff, it, alive, last_kept):
        a = alive.numpy().astype(bool)
        keep = a & (kth.numpy() <= cutoff)
        last_kept.numpy()[keep] = it
        alive.numpy()[a & ~keep] = 0
        return int(keep.sum())

    def vec_max(self, v):
        return float(v.max())

    def impute_iter(self, x_curr, knn, n_do_i, mpv, it, x_next, idc, stats,
                    r0, r1):
        x, kn = x_curr.numpy(), knn.numpy()
        present = x >= mpv
        for s in range(r0, r1):
            mask = (present[kn[s]].sum(axis=0) >= n_do_i) & ~present[s]
            x_next.numpy()[s] = x[s]
            if mask.any():
                x_next.numpy()[s, mask] = np.median(x[kn[s]][:, mask], axis=0)
                idc.numpy()[s, mask] = it
        sub = idc.numpy()[r0:r1] > 0
        stats[0] = int(sub.sum())
        stats[1] = int(sub.any(axis=1).sum())


    def prepare_csr(self, indptr, indices, data, shape, metric):
        import scipy.sparse as sps
        from oracle import sdm_oracle as orc
        m = sps.csr_matrix((data.numpy(), indices.numpy(), indptr.numpy()), shape=shape)
        return {"d": orc.pdist(m, metric)}


def _input_worker(rank, world, port, n, p, k, q):
    """gather_x (row shards of x -> full x on every rank) and the CSR entry."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sps
        from oracle import sdm_oracle as orc
        from scedar_b200.stripes import (StripedDistanceMatrix, gather_x,
                                         gather_x_from_host, shard_rows)
        rng = np.random.default_rng(9)
        xn = rng.random((n, p))
        b, e = shard_rows(n, world)[rank]
        full = gather_x(torch.from_numpy(xn[b:e].copy()), n)
        ok_gather = bool(np.array_equal(full.numpy(), xn))
        # the chunked form (upload range by range, all-gather behind it)
        for chunks in (1, 3, 1000):
            full2 = gather_x_from_host(torch.from_numpy(xn[b:e].copy()), n,
                                       device="cpu", chunks=chunks)
            ok_gather = ok_gather and bool(np.array_equal(full2.numpy(), xn))
        xs = sps.random(n, p, density=0.2, format="csr", random_state=3).astype(np.float64)
        csr = None
        if rank == 0:
            csr = (torch.from_numpy(xs.indptr.astype(np.int64)),
                   torch.from_numpy(xs.indices.astype(np.int32)),
                   torch.from_numpy(xs.data.copy()))
        sdm = StripedDistanceMatrix.from_csr(csr, "cosine", shape=(n, p), nnz=xs.nnz,
                                             device="cpu", ops=OracleOps())
        idx, dd = sdm.knn(k, exclude=1)
        ridx, rdist = orc.knns_precomputed(orc.pdist(xs, "cosine"), k)
        ok_csr = bool(np.array_equal(idx.numpy(), ridx) and np.array_equal(dd.numpy(), rdist))
        q.put((rank, ok_gather, ok_csr))
    finally:
        dist.destroy_process_group()


def test_two_rank_inputs():
    world, n, p, k = 2, 301, 9, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_input_worker, args=(r, world, port, n, p, k, q))
             for r in range(world)]
    for pr in procs:
        pr.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(g[1] is True and g[2] is True for g in got), got


def _consumer_worker(rank, world, port, n, p, k, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import knn_oracle as korc
        from oracle import sdm_oracle as orc
        from scedar_b200.stripes import StripedDistanceMatrix
        rng = np.random.default_rng(5)
        xn = np.concatenate([rng.normal(0, 1, (n - 12, p)),
                             rng.uniform(-9, 9, (12, p))])
        x = torch.from_numpy(xn) if rank == 0 else None
        sdm = StripedDistanceMatrix(x, "euclidean", shape=(n, p),
                                    device="cpu", ops=OracleOps())
        d = orc.pdist(xn, "euclidean")
        cut = float(np.median(np.sort(d, axis=0)[k]) * 1.1)
        lk = sdm.rare_samples(k, cut, 4).numpy()
        res, prog = korc.rare_pdist(d, k, cut, 4)
        want = np.full(n, -1)
        for i, kept in enumerate(prog):
            want[kept] = i
        ok_rare = bool(np.array_equal(lk, want)) and 0 < len(res) < n
        # imputation on thresholded data, lists from the striped kNN
        xi = np.where(xn > 0.3, xn, 0.0)
        idx, _ = sdm.knn(k, exclude=0)
        got, idc, (entries, cells) = sdm.impute(torch.from_numpy(xi), idx, k,
                                                2, 0.3, 3)
        wx, widc, wstats = korc.impute(xi, idx.numpy(), k, 2, 0.3, 3)
        ok_imp = (np.array_equal(got.numpy(), wx)
                  and np.array_equal(idc.numpy(), widc)
                  and (entries, cells) == wstats and entries > 0)
        q.put((rank, ok_rare, ok_imp))
    finally:
        dist.destroy_process_group()


def test_two_rank_consumers():
    """Row-striped rare-sample detection and imputation (world 2, gloo)
    against the single-process NumPy restatement."""
    world, n, p, k = 2, 300, 10, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_consumer_worker,
                         args=(r, world, port, n, p, k, q))
             for r in range(world)]
    for pr in procs:
        pr.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(g[1] is True and g[2] is True for g in got), got


def _worker(rank, world, port, n, p, k, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sdm_oracle as orc
        from scedar_b200.stripes import StripedDistanceMatrix
        x = None
        if rank == 0:
            x = torch.from_numpy(
                np.random.default_rng(1).poisson(1.0, (n, p)).astype(float))
        sdm = StripedDistanceMatrix(x, "euclidean", shape=(n, p),
                                    device="cpu", ops=OracleOps())
        stripe = sdm.compute_d()
        assert stripe.shape == (sdm.row_end - sdm.row_begin, n)
        idx, dd = sdm.knn(k, exclude=1)
        idx2, dd2 = sdm.knn(k, exclude=0, from_d=False)
        ok = None
        if rank == 0:
            d = orc.pdist(x.numpy(), "euclidean")
            ridx, rdist = orc.knns_precomputed(d, k)
            fidx, _ = orc.knn_drop_first(d, k)
            ok = (np.array_equal(idx.numpy(), ridx)
                  and np.array_equal(dd.numpy(), rdist)
                  and np.array_equal(idx2.numpy(), fidx))
        q.put((rank, sdm.stripes, tuple(idx.shape), ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [300, 129])
def test_two_rank_stripes(n):
    world, p, k = 2, 12, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, p, k, q))
             for r in range(world)]
    for pr in procs:
        pr.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    got.sort()
    assert got[0][1] == got[1][1]                    # same partition
    assert got[0][2] == got[1][2] == (n, k)          # every rank: all lists
    assert got[0][3] is True


def _hclust_worker(rank, world, port, n, p, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.cluster.hierarchy as sch
        from oracle import hclust_oracle as horc
        from oracle import sdm_oracle as orc
        from scedar_b200.stripes import StripedDistanceMatrix
        xn = np.random.default_rng(4).normal(size=(n, p))
        x = torch.from_numpy(xn) if rank == 0 else None
        sdm = StripedDistanceMatrix(x, "euclidean", shape=(n, p),
                                    device="cpu", ops=OracleOps())
        d = orc.pdist(xn, "euclidean")
        want = horc.squareform(d)
        ok = True
        for dst in range(world):
            sf = sdm.condensed(dst=dst)
            ok &= (sf is None) == (rank != dst)
            if rank == dst:
                ok &= bool(np.array_equal(sf.numpy(), want))
        z = sdm.hclust_linkage(linkage="average")
        ok &= bool(np.array_equal(z, sch.linkage(want, method="average")))
        za = sdm.hclust_linkage(linkage="auto", is_euc_dist=True, dst=world - 1)
        zo, _ = horc.hclust_linkage(d, linkage="auto", is_euc_dist=True)
        ok &= bool(np.array_equal(za, zo))
        # what scipy's validity check rejects is raised on every rank, even
        # though only one rank's stripe holds the NaN
        sdm.d_stripe = sdm.d_stripe.clone()
        if rank == world - 1:
            sdm.d_stripe[-2, -1] = float("nan")
        try:
            sdm.condensed()
            ok = False
        except ValueError as e:
            ok &= "symmetric" in str(e)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(260, 2), (130, 2), (390, 3)])
def test_condensed_and_linkage_over_ranks(n, world):
    """Row-striped condensed form (exact-size send / recv of the ragged
    segments) and the linkage on top of it (gloo, world 2 and 3; n = 130
    leaves the last rank two rows)."""
    p = 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_hclust_worker, args=(r, world, port, n, p, q))
             for r in range(world)]
    for pr in procs:
        pr.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(g[1] is True for g in got), got
