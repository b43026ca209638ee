"""Block minima of D -- the by-product of the INT8 split Gram launches -- and
the selection that reads them instead of the whole matrix.

Two things are held bit for bit: (1) every launch form writes, for every row
of every output stripe, exactly the minimum of each aligned run of 32 entries of
that row of D; (2) the lists of ``scedar_knn_from_d_blockmin`` equal those of
``scedar_knn_from_d`` on the same matrix -- indices and distances, both
exclusion rules, exact ties (integer counts) included, k from 1 to 127.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    import scedar_b200
    from scedar_b200 import _capi
    _capi.load()
    return scedar_b200


def run_minima(d, ldb):
    """Reference block minima of a stripe d [rows, n] (torch, device)."""
    import torch
    rows, n = d.shape
    nb = (n + 31) // 32
    pad = torch.full((rows, nb * 32), float("inf"), dtype=d.dtype, device=d.device)
    pad[:, :n] = d
    return pad.view(rows, nb, 32).amin(dim=2)


def case(seed, n, p):
    rng = np.random.default_rng(seed)
    if seed % 2:
        x = rng.poisson(0.7, size=(n, p)).astype(np.float64)      # exact ties
        x[3] = x[2]
        x[n - 1] = x[0]
    else:
        x = np.log1p(rng.poisson(rng.gamma(0.4, 2.0, size=p), size=(n, p))) \
            + 1e-3 * rng.random((n, p))
    return x


@pytest.mark.parametrize("n,p,metric,seed", [
    (33, 5, "euclidean", 1), (130, 17, "cosine", 2), (1000, 64, "correlation", 3),
    (1031, 200, "euclidean", 4), (2050, 300, "cosine", 5),
    (4129, 130, "correlation", 6), (3000, 4500, "correlation", 7)])
def test_minima_and_lists_whole_matrix(b200, n, p, metric, seed):
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
    x = case(seed, n, p)
    cells = dev.Cells.from_dense(torch.from_numpy(x).cuda(), metric)
    bmin = dev.new_blockmin(n, n)
    bmin.fill_(-3.0)
    cells.set_blockmin([bmin])
    for prec in ("fp64_i8", "fp64_i8s6"):
        d = cells.pdist_symmetric(precision=prec)
        nb = (n + 31) // 32
        assert torch.equal(bmin[:, :nb], run_minima(d, bmin.stride(0))), prec
        for k in sorted({1, 2, min(15, n - 1), min(40, n - 1), min(100, n - 1),
                         min(127, n - 1)}):
            for ex in (EXCLUDE_FIRST, EXCLUDE_SELF):
                i0, v0 = dev.knn_from_d(d, k, ex)
                i1, v1 = cells.knn_from_d_blockmin(d, bmin, k, ex)
                assert torch.equal(i0, i1) and torch.equal(v0, v1), (prec, k, ex)
    # a sub-range of rows selects from the rows' own minima
    r0, r1 = n // 3, n - 2
    i0, v0 = dev.knn_from_d(d[r0:r1], 5, EXCLUDE_SELF, row_begin=r0)
    i1, v1 = cells.knn_from_d_blockmin(d[r0:r1], bmin[r0:r1], 5, EXCLUDE_SELF,
                                       row_begin=r0)
    assert torch.equal(i0, i1) and torch.equal(v0, v1)
    cells.close()


def test_stripe_and_row_range_forms(b200):
    """Row stripes (the multi-GPU ownership of D on one device) and the
    row-range form behind the upload write the same minima."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_FIRST
    from scedar_b200.stripes import StripedDistanceMatrix
    n, p = 2700, 150
    x = case(11, n, p)
    xt = torch.from_numpy(x).cuda()
    cells = dev.Cells.from_dense(xt, "correlation")
    full = cells.pdist_symmetric(precision="fp64_i8")
    want = run_minima(full, 0)
    nb = want.shape[1]
    for r0, r1 in ((0, 1280), (1280, 2700), (512, 640), (2688, 2700)):
        bmin = dev.new_blockmin(r1 - r0, n)
        bmin.fill_(-1.0)
        cells.set_blockmin([bmin])
        d = cells.pdist_stripe(r0, r1, precision="fp64_i8")
        assert torch.equal(d, full[r0:r1])
        assert torch.equal(bmin[:, :nb], want[r0:r1]), (r0, r1)
        i0, v0 = dev.knn_from_d(d, 15, EXCLUDE_FIRST, row_begin=r0)
        i1, v1 = cells.knn_from_d_blockmin(d, bmin, 15, EXCLUDE_FIRST, row_begin=r0)
        assert torch.equal(i0, i1) and torch.equal(v0, v1)
    cells.close()
    # the form bench.py's e2e runs (upload hidden behind row ranges), and the
    # one its `value` runs, through StripedDistanceMatrix
    xh = torch.from_numpy(x).pin_memory()
    s1 = StripedDistanceMatrix.from_host(xh, "correlation", precision="fp64_i8")
    assert s1.bmin is not None
    assert torch.equal(s1.d_stripe, full)
    assert torch.equal(s1.bmin[:, :nb], want)
    a_i, a_d = s1.knn(15, EXCLUDE_FIRST)
    s2 = StripedDistanceMatrix(xt, "correlation", precision="fp64_i8")
    s2.compute_d()
    assert s2.bmin is not None and torch.equal(s2.bmin[:, :nb], want)
    b_i, b_d = s2.knn(15, EXCLUDE_FIRST)
    r_i, r_d = dev.knn_from_d(full, 15, EXCLUDE_FIRST)
    assert torch.equal(a_i, r_i) and torch.equal(a_d, r_d)
    assert torch.equal(b_i, r_i) and torch.equal(b_d, r_d)
    # the DMMA kernel leaves no minima: the selection reads D
    s3 = StripedDistanceMatrix(xt, "correlation", precision="fp64")
    s3.compute_d()
    assert s3.bmin is None
    for s in (s1, s2, s3):
        s.close()


def test_peer_store_forms_single_gpu_emulation(b200):
    """The multi-GPU peer-store launches, ranks emulated one after the other on
    one GPU: every stripe's minima buffer (allocated behind the stripe, where
    a peer reaches it through the same mapping) is complete and exact."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_SELF
    from scedar_b200.stripes import partition, piece_layout, stripe_pieces
    n, p = 1900, 90
    x = torch.from_numpy(case(21, n, p)).cuda()
    for metric in ("correlation", "euclidean"):
        cells = dev.Cells.from_dense(x, metric)
        full = cells.pdist_symmetric(precision="fp64_i8")
        want = run_minima(full, 0)
        nb = want.shape[1]
        for world in (2, 3, 8):
            parts = partition(n, world)
            for form in ("stripe", "rows", "pieces1", "pieces3", "pieces8", "pieces0"):
                peers = dev.PeerStripes(parts, n, blockmin=True)
                for r in range(world):
                    peers.tensor(r).fill_(-7.0)
                    peers.bmin_tensor(r).fill_(-7.0)
                cells.set_blockmin(peers.bmin_ptrs(), peers.ldb)
                for r in range(world):
                    if form == "stripe":
                        dev.pdist_stripe_p2p(cells, peers, r, precision="fp64_i8")
                    elif form.startswith("pieces"):
                        continue
                    else:
                        for b, e, _ in piece_layout(n, world, 3):
                            dev.pdist_p2p_rows(cells, peers, r, b, e,
                                               precision="fp64_i8")
                if form.startswith("pieces"):
                    # pieces of every stripe at once, launch q on every rank
                    nq = int(form[6:])
                    if nq:
                        table, _ = stripe_pieces(parts, n, nq)
                    else:           # the growing default of from_host_stripes
                        from scedar_b200.stripes import GROWING_PIECES
                        table, rg = stripe_pieces(parts, n, cuts=GROWING_PIECES)
                        nq = len(rg)
                    for q in range(nq):
                        for r in range(world):
                            dev.pdist_p2p_pieces(cells, peers, r, table, q, nq,
                                                 precision="fp64_i8")
                torch.cuda.synchronize()
                for r, (b, e) in enumerate(parts):
                    assert torch.equal(peers.tensor(r), full[b:e]), (metric, world, form, r)
                    assert torch.equal(peers.bmin_tensor(r)[:, :nb], want[b:e]), \
                        (metric, world, form, r)
                    i0, v0 = dev.knn_from_d(full[b:e], 15, EXCLUDE_SELF, row_begin=b)
                    i1, v1 = cells.knn_from_d_blockmin(
                        peers.tensor(r), peers.bmin_tensor(r), 15, EXCLUDE_SELF,
                        row_begin=b)
                    assert torch.equal(i0, i1) and torch.equal(v0, v1)
                cells.set_blockmin(None)
                peers.close()
        cells.close()


def test_blockmin_guards(b200):
    """No minima without an INT8 split launch: the call fails instead of
    selecting from stale memory; detaching stops the writes."""
    import torch
    from scedar_b200 import device as dev
    x = torch.from_numpy(case(5, 300, 20)).cuda()
    cells = dev.Cells.from_dense(x, "cosine")
    bmin = dev.new_blockmin(300, 300)
    cells.set_blockmin([bmin])
    d = cells.pdist_symmetric(precision="fp64")           # DMMA: writes none
    with pytest.raises(ValueError):
        cells.knn_from_d_blockmin(d, bmin, 5)
    cells.pdist_symmetric(precision="fp64_i8", out=d)
    cells.knn_from_d_blockmin(d, bmin, 5)
    cells.set_blockmin(None)
    bmin.fill_(-2.0)
    cells.pdist_symmetric(precision="fp64_i8", out=d)
    torch.cuda.synchronize()
    assert float(bmin.max()) == -2.0
    with pytest.raises(ValueError):
        cells.knn_from_d_blockmin(d, bmin, 5)
    with pytest.raises(ValueError):                        # ldb too small
        cells.set_blockmin([bmin], ldb=3)
    cells.close()


def test_dropin_uses_minima_and_overlapped_upload(b200, monkeypatch):
    """Through the drop-in class: a dense host matrix large enough for the auto
    policy's INT8 split kernel goes up range by range behind the contraction
    (pageable source), D comes with its block minima, and the neighbour lists
    -- selected through them -- equal the oracle's stable order on that D."""
    import torch
    from oracle import sdm_oracle as orc
    from scedar_b200 import device as dev, synth
    from scedar_b200 import sdm as sdm_mod
    monkeypatch.setattr(sdm_mod, "UPLOAD_OVERLAP_MIN_CELLS", 0)   # take that route here
    n, p, k = 6016, 1500, 15
    x = synth.log_counts(n, p, seed=5)
    sdm = b200.SampleDistanceMatrix(x, metric="correlation")
    assert sdm._gram_precision() == "fp64_i8"
    d = sdm._device_d()
    assert sdm._bmin_dev is not None and tuple(d.shape) == (n, n)
    # same matrix as the one-copy route
    cells = dev.Cells.from_dense(dev.to_device(x), "correlation")
    assert torch.equal(d, cells.pdist_symmetric(precision="fp64_i8"))
    cells.close()
    lut = sdm.s_knn_ind_lut(k)
    want, _ = orc.knn_drop_first(d.cpu().numpy(), k)
    assert np.array_equal(np.array([lut[i] for i in range(n)]), want)
    ki, kd = sdm.s_knns(k)
    ri, rd = orc.knns_precomputed(d.cpu().numpy(), k)
    assert np.array_equal(np.asarray(ki), ri) and np.array_equal(np.asarray(kd), rd)
    sdm.release_device()
    assert sdm._bmin_dev is None
