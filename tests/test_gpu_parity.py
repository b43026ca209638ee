"""Parity of the CUDA path (through the C ABI) against the oracle and the
golden vectors.  All tests need a B200: ``pytest -m gpu``."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import sdm_oracle as orc

pytestmark = pytest.mark.gpu

# float64 path tolerance stated by BASELINE.json: 1e-10 relative
RTOL = 1e-10


@pytest.fixture(scope="module")
def b200():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    import scedar_b200
    from scedar_b200 import _capi
    _capi.load()  # fails loudly if the .so is missing
    return scedar_b200


def assert_d_close(got, want):
    assert got.dtype == np.float64 and got.shape == want.shape
    scale = max(want.max(), 1e-300) if want.size else 1.0
    assert np.abs(got - want).max() / scale <= RTOL
    big = want > 1e-6
    if big.any():
        assert (np.abs(got - want)[big] / want[big]).max() <= RTOL
    assert np.array_equal(got, got.T)
    assert not got.diagonal().any()
    assert (got >= 0).all()


def near_tie_rows(idx, ref_idx, d_ref, rel=1e-13):
    """Rows where indices differ; each must be a near tie in the reference
    distances (the two candidates within `rel`, three orders below the 1e-10
    distance tolerance -- SURVEY.md section 8d).  Returns the count."""
    bad = np.where((idx != ref_idx).any(axis=1))[0]
    for r in bad:
        a = d_ref[r, idx[r]]
        b = d_ref[r, ref_idx[r]]
        assert np.all(np.abs(a - b) <= rel * np.maximum(a, b)), r
    return len(bad)


DENSE = [("ranf222", False), ("counts_ties", True), ("edge_rows", True),
         ("logcounts", False)]


@pytest.mark.parametrize("name,full", DENSE)
def test_pdist_golden(b200, golden, name, full):
    g = golden(name)
    x = g["x"]
    for m in sorted({k.split("_")[0] for k in g.files if k.endswith("_lut")}):
        got = b200.SampleDistanceMatrix(x, metric=m)._d
        assert_d_close(got, orc.pdist(x, m))
        if full:
            want = g[m + "_d"]
            if name == "counts_ties":
                # integer counts: Gram exact, epilogue in reference order
                if m == "correlation":
                    assert_d_close(got, want)
                else:
                    assert np.array_equal(got, want), m
            else:
                assert_d_close(got, want) if m != "euclidean" else \
                    np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-7)
        else:
            np.testing.assert_allclose(got[:8], g[m + "_d_rows"], rtol=1e-9,
                                       atol=1e-13)
            np.testing.assert_allclose(got[::25, ::25], g[m + "_d_grid"],
                                       rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("name,full", DENSE)
def test_knn_lut_golden(b200, golden, name, full):
    """s_knn_ind_lut == ranks 1..k of the reference's stable argsort."""
    g = golden(name)
    x, k = g["x"], int(g["k"])
    for m in sorted({k_.split("_")[0] for k_ in g.files
                     if k_.endswith("_lut")}):
        sdm = b200.SampleDistanceMatrix(x, metric=m)
        lut = sdm.s_knn_ind_lut(k)
        got = np.array([lut[i] for i in range(x.shape[0])])
        want = g[m + "_lut"]
        if name == "counts_ties" and m != "correlation":
            assert np.array_equal(got, want), m      # thousands of exact ties
        elif name == "edge_rows":
            # duplicates / zero rows: compare on the reference's own D
            sdm2 = b200.SampleDistanceMatrix(x, d=g[m + "_d"], metric=m)
            lut2 = sdm2.s_knn_ind_lut(k)
            assert np.array_equal(
                np.array([lut2[i] for i in range(x.shape[0])]), want), m
        elif name == "counts_ties":
            # centred integer counts: mathematically exact ties whose order
            # is decided by rounding noise; every mismatch must be such a tie
            near_tie_rows(got, want, orc.pdist(x, m))
        else:
            assert near_tie_rows(got, want, orc.pdist(x, m)) == 0, m


@pytest.mark.parametrize("name", ["ranf222", "logcounts"])
def test_s_knns_golden(b200, golden, name):
    g = golden(name)
    x, k = g["x"], int(g["k"])
    for m in sorted({k_.split("_")[0] for k_ in g.files
                     if k_.endswith("_lut")}):
        for use_pdist in (True, False):
            sdm = b200.SampleDistanceMatrix(x, metric=m, use_pdist=use_pdist)
            targets, distances = sdm.s_knns(k)
            assert isinstance(targets, list) and len(targets) == x.shape[0]
            assert targets[0].dtype == np.int64
            assert distances[0].dtype == np.float64
            key = "_knn_" if use_pdist else "_raw_knn_"
            assert np.array_equal(np.array(targets), g[m + key + "idx"])
            np.testing.assert_allclose(np.array(distances),
                                       g[m + key + "dist"], rtol=1e-9,
                                       atol=1e-12)
            assert sdm._last_k == k and sdm._last_knns[0] is targets


def test_precomputed_d_paths(b200, golden):
    """User-supplied d: top-k, both exclusion rules, on the reference's own D
    with exact ties -- indices bit-exact against the stable argsort."""
    g = golden("counts_ties")
    x = g["x"]
    for m in ("cosine", "correlation", "euclidean"):
        d = g[m + "_d"]
        sdm = b200.SampleDistanceMatrix(x, d=d.copy(), metric=m)
        for k in (1, 10, 31, 40, 100, 140):
            lut = sdm.s_knn_ind_lut(k)
            got = np.array([lut[i] for i in range(x.shape[0])])
            assert np.array_equal(got, g[m + "_arg"][1:k + 1].T), (m, k)
            idx, dist = sdm.s_knns(k)
            ridx, rdist = orc.knns_precomputed(d, k)
            assert np.array_equal(np.array(idx), ridx), (m, k)
            assert np.array_equal(np.array(dist), rdist), (m, k)


def test_col_sort_golden(b200, golden):
    g = golden("counts_ties")
    x = g["x"]
    for m in ("cosine", "euclidean"):
        sdm = b200.SampleDistanceMatrix(x, d=g[m + "_d"].copy(), metric=m)
        assert np.array_equal(sdm._col_argsorted_d, g[m + "_arg"])
        assert sdm._col_argsorted_d.dtype == np.int64
        assert np.array_equal(sdm._col_sorted_d,
                              np.sort(g[m + "_d"], axis=0))
        assert np.array_equal(sdm.s_ith_nn_ind(3), g[m + "_arg"][3])
        assert sdm._col_argsorted_d is sdm._col_argsorted_d  # memoised


def test_col_sort_large(b200):
    """n > 4096 exercises the global bitonic passes."""
    import torch
    from scedar_b200 import device as dev
    rng = np.random.default_rng(5)
    d = rng.integers(0, 50, size=(6, 5000)).astype(np.float64)  # many ties
    srt, arg = dev.col_sort(torch.from_numpy(d).cuda())
    assert np.array_equal(arg.cpu().numpy(),
                          np.argsort(d, kind="mergesort", axis=1))
    assert np.array_equal(srt.cpu().numpy(), np.sort(d, axis=1))


def test_csr_golden(b200, golden):
    g = golden("csr_lognormal")
    xs = sps.csr_matrix((g["data"], g["indices"], g["indptr"]),
                        shape=tuple(g["shape"]))
    k = int(g["k"])
    for m in ("cosine", "euclidean"):
        sdm = b200.SampleDistanceMatrix(xs, metric=m)
        got = sdm._d
        assert_d_close(got, g[m + "_d"])
        dense = b200.SampleDistanceMatrix(xs.toarray(), metric=m)._d
        assert np.array_equal(got, dense)        # same kernels after densify
        idx, dist = sdm.s_knns(k)
        np.testing.assert_allclose(np.array(dist), g[m + "_knn_dist"],
                                   rtol=1e-9, atol=1e-12)
        srt = np.sort(g[m + "_d"], axis=1)[:, :k + 2]
        untied = (np.diff(srt, axis=1) != 0).all(axis=1)
        assert np.array_equal(np.array(idx)[untied],
                              g[m + "_knn_idx"][untied])
        conn = sdm.s_knn_connectivity_matrix(5)
        want = sps.csr_matrix((g[m + "_conn_data"], g[m + "_conn_indices"],
                               g[m + "_conn_indptr"]), shape=got.shape)
        untied5 = (np.diff(np.sort(g[m + "_d"], axis=1)[:, :7], axis=1)
                   != 0).all(axis=1)
        np.testing.assert_allclose(conn.toarray()[untied5],
                                   want.toarray()[untied5], rtol=1e-9)
    with pytest.raises(ValueError):
        b200.SampleDistanceMatrix(xs, metric="correlation")._d
    with pytest.raises(ValueError):
        b200.SampleDistanceMatrix(xs, metric="correlation").d


def test_num_correct_dist_mat_golden(b200, golden):
    g = golden("ncdm")
    ncdm = b200.SampleDistanceMatrix.num_correct_dist_mat
    for src, want, ub, nw in (("raw", "fixed", None, g["n_warn"][0]),
                              ("raw", "fixed_ub", 0.7, g["n_warn"][1]),
                              ("sym", "fixed_sym", None, g["n_warn"][2])):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            m = g[src].copy()
            out = ncdm(m, ub)
        assert out is m
        assert np.array_equal(out, g[want])
        assert len(w) == int(nw)


# ---- the reference's own known-answer tests, against the CUDA path --------
def test_reference_known_answers(b200):
    SDM = b200.SampleDistanceMatrix
    # tests/test_eda/test_dense_sdm.py:23-52
    r2, r8 = np.sqrt(2), np.sqrt(8)
    np.testing.assert_allclose(
        SDM([[0, 0], [1, 1], [2, 2]], metric="euclidean").d,
        [[0, r2, r8], [r2, 0, r2], [r8, r2, 0]])
    x = np.array([[0, 1, 2, 3], [1, 2, 0, 6]])
    e = np.sqrt(np.power(x[0] - x[1], 2).sum())
    np.testing.assert_allclose(SDM(x, metric="euclidean", nprocs=5).d,
                               [[0, e], [e, 0]])
    np.testing.assert_allclose(SDM(x, metric="correlation", nprocs=5).d,
                               [[0, 0.3618551], [0.3618551, 0]], rtol=1e-6)
    # :503-551  i-th NN, LUT and the tie-break vector
    nn = SDM([[0], [1], [5], [6], [10], [20]], metric="euclidean")
    np.testing.assert_allclose(nn.s_ith_nn_d(0), 0)
    np.testing.assert_allclose(nn.s_ith_nn_d(1), [1, 1, 1, 1, 4, 10])
    np.testing.assert_allclose(nn.s_ith_nn_d(2), [5, 4, 4, 4, 5, 14])
    x3 = [[v] * 3 for v in (0, 1, 5, 6, 10, 20)]
    for use_pdist in (True, False):
        nn = SDM(x3, metric="euclidean", use_pdist=use_pdist)
        if use_pdist:
            assert nn.s_ith_nn_ind(0).tolist() == [0, 1, 2, 3, 4, 5]
            assert nn.s_ith_nn_ind(1).tolist() == [1, 0, 3, 2, 3, 4]
            assert nn.s_ith_nn_ind(2).tolist() == [2, 2, 1, 4, 2, 3]
        assert nn.s_knn_ind_lut(0) == dict(zip(range(6), [[]] * 6))
        assert nn.s_knn_ind_lut(1) == dict(zip(range(6), [[1], [0], [3], [2],
                                                          [3], [4]]))
        assert nn.s_knn_ind_lut(3) == dict(zip(range(6), [
            [1, 2, 3], [0, 2, 3], [3, 1, 0], [2, 4, 1], [3, 2, 1],
            [4, 3, 2]]))
        nn.s_knn_ind_lut(5)
        for bad in (-1, -0.5, 6, 6.5, 7):
            with pytest.raises(ValueError):
                nn.s_knn_ind_lut(bad)
    # :1041-1061 connectivity
    nn = SDM([[1], [2], [6]], metric="euclidean")
    np.testing.assert_allclose(nn.s_knn_connectivity_matrix(1).toarray(),
                               [[0, 1, 0], [1, 0, 0], [0, 4, 0]])
    with pytest.raises(ValueError):
        nn.s_knn_connectivity_matrix(0)
    with pytest.raises(ValueError):
        nn.s_knn_connectivity_matrix(1, index_params={})
    with pytest.raises(ValueError):
        nn.s_knn_connectivity_matrix(1, query_params={})
    # :472-501 num_correct_dist_mat
    t = np.array([[0, 1, 2], [0.5, 0, 1.5], [1, 1.6, 0.5]])
    with pytest.warns(UserWarning):
        c = SDM.num_correct_dist_mat(t.copy())
    assert np.array_equal(c, [[0, 0.5, 1], [0.5, 0, 1.6], [1, 1.6, 0]])
    with pytest.warns(UserWarning):
        c = SDM.num_correct_dist_mat(t.copy(), 1)
    assert np.array_equal(c, [[0, 0.5, 1], [0.5, 0, 1], [1, 1, 0]])
    # 20 collinear points, k=10 (tests/test_eda/test_no_pdist_sdm.py:1099-1145)
    xl = np.arange(20, dtype=float).reshape(-1, 1) ** 1.5
    for use_pdist in (True, False):
        nn = SDM(xl, metric="euclidean", use_pdist=use_pdist)
        ref = orc.knns_precomputed(orc.pdist(xl, "euclidean"), 10)[0]
        lut = nn.s_knn_ind_lut(10)
        assert np.array_equal(np.array([lut[i] for i in range(20)]), ref)


def test_sklearn_crosscheck(b200):
    # tests/test_eda/test_dense_sdm.py:1231-1250
    import sklearn.metrics.pairwise as skp
    np.random.seed(222)
    x = np.random.ranf(10000).reshape(500, -1)
    SDM = b200.SampleDistanceMatrix
    for m, fn in (("cosine", SDM.cosine_pdist),
                  ("correlation", SDM.correlation_pdist)):
        skd = skp.pairwise_distances(x, metric=m)
        np.testing.assert_allclose(fn(x), skd, atol=1e-12)
        np.testing.assert_allclose(SDM(x, metric=m)._d, skd, atol=1e-12)


def test_duplicates_and_zero_rows(b200):
    SDM = b200.SampleDistanceMatrix
    x = np.array([[0, 0], [1, 1], [200, 200], [200, 200], [200, 200],
                  [100, 100]], dtype=float)
    sdm = SDM(x, metric="euclidean")
    assert sdm.s_knn_ind_lut(2)[3] == [3, 4]          # appendix B
    assert sdm.s_knns(2)[0][3].tolist() == [2, 4]
    z = np.array([[0, 0, 0], [1, 2, 3], [0, 0, 0], [5, 5, 5], [2, 4, 6.5]])
    d = SDM(z, metric="cosine")._d
    assert d[0, 1] == 1.0 and d[0, 2] == 1.0 and d[0, 0] == 0.0
    d = SDM(z, metric="correlation")._d
    assert d[3, 1] == 1.0 and d[3, 3] == 0.0 and not np.isnan(d).any()


def test_empty_and_lazy(b200):
    SDM = b200.SampleDistanceMatrix
    with pytest.raises(ValueError):
        SDM(np.empty(0), metric="euclidean")
    sdm = SDM(np.empty((0, 0)), metric="euclidean")
    assert sdm._d.shape == (0, 0)
    assert sdm._col_sorted_d.shape == (0, 0)
    assert sdm._col_argsorted_d.shape == (0, 0)
    assert SDM(np.empty((3, 0)), metric="cosine")._d.shape == (3, 3)


@pytest.mark.parametrize("n,p", [(1, 7), (2, 1), (127, 33), (129, 17),
                                 (300, 1), (257, 640)])
def test_ragged_shapes(b200, n, p):
    """Tile-edge sizes: n, p not multiples of the 128 x 16 tiling, odd row
    strides (no 128-bit loads possible)."""
    rng = np.random.default_rng(n * 1000 + p)
    x = rng.gamma(0.5, 2.0, size=(n, p))
    for m in ("cosine", "correlation", "euclidean"):
        got = b200.SampleDistanceMatrix(x, metric=m)._d
        assert_d_close(got, orc.pdist(x, m))


def test_stripes_and_fused_topk(b200):
    """Row stripes of D and the fused (no D) kNN agree with the symmetric
    full-matrix kernel bit for bit; fused top-k == top-k from D."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
    x = torch.from_numpy(np.random.default_rng(3).poisson(
        0.7, size=(1000, 90)).astype(np.float64)).cuda()
    for m in ("cosine", "correlation", "euclidean"):
        cells = dev.Cells.from_dense(x, m)
        full = cells.pdist_symmetric()
        assert torch.equal(full, cells.pdist_stripe(0, 1000))
        assert torch.equal(full[100:333], cells.pdist_stripe(100, 333))
        assert torch.equal(full[997:], cells.pdist_stripe(997, 1000))
        # the three routes of the FP64 kNN (capi.cu: knn_stripe_fp64) agree
        import os
        for mode in ("fused", "chunked", "symmetric", None):
            if mode is None:
                os.environ.pop("SCEDAR_B200_KNN_FP64", None)
            else:
                os.environ["SCEDAR_B200_KNN_FP64"] = mode
            for k in (1, 15, 31):
                for ex in (EXCLUDE_FIRST, EXCLUDE_SELF):
                    i0, d0 = dev.knn_from_d(full, k, ex)
                    i1, d1 = cells.knn_stripe(k, ex)
                    assert torch.equal(i0, i1) and torch.equal(d0, d1), (m, k, ex)
                    i2, d2 = cells.knn_stripe(k, ex, 130, 700)
                    assert torch.equal(i0[130:700], i2)
        for k in (40, 100):   # unfused wide-list path
            i0, d0 = dev.knn_from_d(full, k, EXCLUDE_SELF)
            i1, d1 = cells.knn_stripe(k, EXCLUDE_SELF)
            assert torch.equal(i0, i1) and torch.equal(d0, d1)
        ref = orc.knns_precomputed(full.cpu().numpy(), 15)
        i0, d0 = dev.knn_from_d(full, 15, EXCLUDE_SELF)
        assert np.array_equal(i0.cpu().numpy(), ref[0])
        cells.close()


def test_c1_scale_properties(b200):
    """BASELINE config 1 shape (3005 x 19972): size-independent properties --
    symmetry, zero diagonal, range, row subsample against the oracle, kNN
    self-consistency."""
    from scedar_b200 import synth
    x = synth.log_counts(3005, 19972, seed=3005)
    sdm = b200.SampleDistanceMatrix(x, metric="correlation")
    d = sdm._d
    assert np.array_equal(d, d.T) and not d.diagonal().any()
    assert d.min() >= 0 and d.max() <= 2
    rows = np.arange(0, 3005, 97)
    xc = x - x.mean(axis=1, keepdims=True)
    nrm = np.sqrt((xc * xc).sum(axis=1))
    want = 1 - (xc[rows] @ xc.T) / nrm[rows, None] / nrm[None, :]
    want[np.arange(len(rows)), rows] = 0
    assert np.abs(d[rows] - want).max() <= 1e-10
    lut = sdm.s_knn_ind_lut(15)
    ref = np.argsort(d, kind="mergesort", axis=1)[:, 1:16]
    assert np.array_equal(np.array([lut[i] for i in range(3005)]), ref)


def test_host_abi(b200):
    """The host-buffer C entry points (plain pointers, no torch)."""
    import ctypes
    from scedar_b200 import _capi
    lib = _capi.load()
    x = np.random.default_rng(9).gamma(0.4, 2.0, size=(700, 130))
    d = np.empty((700, 700))
    _capi.check(lib.scedar_pdist_dense_host(
        x.ctypes.data_as(ctypes.c_void_p), 700, 130, 1, 0,
        d.ctypes.data_as(ctypes.c_void_p), 0))
    assert_d_close(d, orc.pdist(x, "correlation"))
    idx = np.empty((700, 15), dtype=np.int64)
    dist = np.empty((700, 15))
    _capi.check(lib.scedar_knn_dense_host(
        x.ctypes.data_as(ctypes.c_void_p), 700, 130, 1, 0, 15, 1,
        idx.ctypes.data_as(ctypes.c_void_p),
        dist.ctypes.data_as(ctypes.c_void_p), 0))
    ridx, rdist = orc.knns_precomputed(d, 15)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    with pytest.raises(ValueError):
        _capi.check(lib.scedar_knn_dense_host(
            x.ctypes.data_as(ctypes.c_void_p), 700, 130, 1, 0, 700, 1,
            idx.ctypes.data_as(ctypes.c_void_p),
            dist.ctypes.data_as(ctypes.c_void_p), 0))


def test_p2p_balanced_symmetric_single_gpu_emulation(b200):
    """The multi-GPU peer-store kernel, with the N ranks emulated one after
    the other on one GPU (all stripes local): every stripe must equal the
    rows of the whole-matrix result bit for bit."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200.stripes import partition
    x = torch.from_numpy(np.random.default_rng(8).gamma(
        0.5, 1.0, size=(1100, 70))).cuda()
    for m in ("correlation", "euclidean"):
        cells = dev.Cells.from_dense(x, m)
        full = cells.pdist_symmetric()
        for world in (1, 2, 3, 8, 11):
            parts = partition(1100, world)
            peers = dev.PeerStripes(parts, 1100)
            for r in range(world):
                peers.tensor(r).fill_(-7.0)
            for r in range(world):
                dev.pdist_stripe_p2p(cells, peers, r)
            torch.cuda.synchronize()
            for r, (b, e) in enumerate(parts):
                assert torch.equal(peers.tensor(r), full[b:e]), (m, world, r)
            peers.close()
        cells.close()


@pytest.mark.parametrize("metric", ["correlation", "cosine", "euclidean"])
def test_fast_path_tcgen05(b200, metric):
    """precision='fast' (tcgen05 + TMA, 3xFP16 split): the distance matrix is
    within the stated tolerance 1e-4 * max(D) of the FP64 path, and the kNN is
    EXACT (same indices, same float64 distances) after the FP64 re-rank."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
    from scedar_b200 import synth
    for n, p in ((700, 300), (1500, 130), (2100, 2000)):
        x = torch.from_numpy(synth.log_counts(n, p, seed=n + p)).cuda()
        cells = dev.Cells.from_dense(x, metric)
        d64 = cells.pdist_symmetric()
        dfast = cells.pdist_stripe(0, n, precision="fast")
        err = (dfast - d64).abs().max().item() / d64.max().item()
        assert err <= 1e-4, (metric, n, p, err)   # stated tolerance
        assert torch.equal(dfast.diagonal(), torch.zeros_like(dfast.diagonal()))
        part = cells.pdist_stripe(200, 555, precision="fast")
        assert torch.equal(part, dfast[200:555])
        # whole matrix: only the tiles touching the lower triangle are
        # contracted, the rest is written as transposes -> bit-symmetric
        dsym = cells.pdist_symmetric(precision="fast")
        assert torch.equal(dsym, dsym.T)
        assert torch.equal(dsym.diagonal(), torch.zeros_like(dsym.diagonal()))
        assert (dsym - d64).abs().max().item() / d64.max().item() <= 1e-4
        assert torch.equal(torch.tril(dsym), torch.tril(dfast))
        # k + 9 <= 32: one 32-candidate list per cell; larger k re-ranks 64
        # candidates drawn from >= 4 column splits (k <= 55)
        for k, ex in ((15, EXCLUDE_FIRST), (15, EXCLUDE_SELF), (1, EXCLUDE_SELF),
                      (23, EXCLUDE_FIRST), (24, EXCLUDE_SELF), (30, EXCLUDE_SELF),
                      (47, EXCLUDE_FIRST), (55, EXCLUDE_SELF)):
            i64, v64 = cells.knn_stripe(k, ex)
            ifa, vfa = cells.knn_stripe(k, ex, precision="fast")
            assert torch.equal(i64, ifa), (metric, n, p, k, ex)
            assert torch.equal(v64, vfa), "re-rank must reproduce the FP64 path bit for bit"
        ifa, vfa = cells.knn_stripe(15, EXCLUDE_SELF, 300, 650, precision="fast")
        i64, v64 = cells.knn_stripe(15, EXCLUDE_SELF, 300, 650)
        assert torch.equal(i64, ifa)
        cells.close()


def test_fast_path_ties_and_fallback(b200):
    """Integer counts: thousands of exact ties defeat any margin, so the fast
    kNN must fall back to the FP64 kernel for those rows and still be exact."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_FIRST
    from scedar_b200 import synth
    x = torch.from_numpy(synth.raw_counts(900, 60, seed=4)).cuda()
    for metric in ("euclidean", "cosine"):
        cells = dev.Cells.from_dense(x, metric)
        i64, v64 = cells.knn_stripe(10, EXCLUDE_FIRST)
        ifa, vfa = cells.knn_stripe(10, EXCLUDE_FIRST, precision="fast")
        assert torch.equal(i64, ifa) and torch.equal(v64, vfa)
        cells.close()


@pytest.mark.parametrize("n,p", [(2, 1), (5, 3), (33, 70), (129, 31), (257, 64),
                                 (300, 256), (513, 257)])
def test_fast_path_small_and_ragged(b200, n, p):
    """Tiny and ragged shapes through the tcgen05 kernels (all three tile
    variants: single CTA, CTA pair from 256 K columns, A-resident up to 64):
    D within tolerance and bit-symmetric, kNN identical to the FP64 path."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
    rng = np.random.default_rng(n * 1000 + p)
    x = torch.from_numpy(rng.normal(0.0, 1.0, size=(n, p)) + 0.3).cuda()
    for metric in ("euclidean", "cosine", "correlation"):
        if metric == "correlation" and p < 2:
            continue
        cells = dev.Cells.from_dense(x, metric)
        d64 = cells.pdist_symmetric()
        scale = max(d64.max().item(), 1e-300)
        for d in (cells.pdist_stripe(0, n, precision="fast"),
                  cells.pdist_symmetric(precision="fast")):
            assert (d - d64).abs().max().item() / scale <= 1e-4, (metric, n, p)
            assert float(d.diagonal().abs().max()) == 0.0
        assert torch.equal(d, d.T)
        for k in sorted({1, min(15, n - 1), min(40, n - 1)}):
            for ex in (EXCLUDE_FIRST, EXCLUDE_SELF):
                i64, v64 = cells.knn_stripe(k, ex)
                ifa, vfa = cells.knn_stripe(k, ex, precision="fast")
                assert torch.equal(i64, ifa) and torch.equal(v64, vfa), (metric, n, p, k)
        cells.close()


@pytest.mark.parametrize("n,p", [(130, 20), (1000, 90), (5000, 301)])
def test_incremental_cells_and_lower_rows(b200, n, p):
    """Cells filled range by range + the lower triangle contracted range by
    range (the upload-hiding pipeline) give the whole-matrix result bit for
    bit, and so does the pinned-host helper."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200.stripes import StripedDistanceMatrix
    rng = np.random.default_rng(n + p)
    xh = torch.from_numpy(rng.poisson(1.3, size=(n, p)).astype(np.float64)
                          + rng.random((n, p))).pin_memory()
    x = xh.cuda()
    for metric in ("correlation", "cosine", "euclidean"):
        ref_cells = dev.Cells.from_dense(x, metric)
        want = ref_cells.pdist_symmetric()
        cells = dev.Cells.alloc(n, p, metric)
        d = torch.full((n, n), -7.0, dtype=torch.float64, device="cuda")
        for b, e in dev.upload_ranges(n, first=256):
            cells.fill_rows(x[b:e], b)
            cells.pdist_lower_rows(b, e, d)
        assert torch.equal(d, want), metric
        i0, v0 = ref_cells.knn_stripe(7, 1)
        i1, v1 = cells.knn_stripe(7, 1)           # the filled cells are complete
        assert torch.equal(i0, i1) and torch.equal(v0, v1)
        cells.close()
        sdm = StripedDistanceMatrix.from_host(xh, metric)
        assert torch.equal(sdm.d_stripe, want)
        i2, _ = sdm.knn(7, 1)
        assert torch.equal(i2, i0)
        sdm.close()
        ref_cells.close()
    with pytest.raises(ValueError):
        dev.Cells.alloc(n, p, "cosine").fill_rows(x[5:100], 5)   # not 128-aligned


def test_fast_path_row_granular_fallback(b200):
    """A few cells whose candidate margin cannot be verified (40 + 31 exact
    duplicates: more than 32 candidates at distance 0) are recomputed by the
    FP64 kernel row by row (gathered rows), the rest keep the tensor-core
    result: bit-identical lists either way, also through the tile-granular
    fallback."""
    import os
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
    rng = np.random.default_rng(12)
    x = rng.normal(0.0, 1.0, size=(3000, 40))
    x[100:140] = x[100]
    x[2000:2031] = x[2950]
    x[2950] = x[2000]
    xt = torch.from_numpy(x).cuda()
    for metric in ("euclidean", "cosine"):
        cells = dev.Cells.from_dense(xt, metric)
        for k, ex in ((15, EXCLUDE_SELF), (15, EXCLUDE_FIRST), (30, EXCLUDE_SELF)):
            i64, v64 = cells.knn_stripe(k, ex)
            for forced in (None, "1"):
                if forced:
                    os.environ["SCEDAR_B200_TC_TILE_FALLBACK"] = forced
                else:
                    os.environ.pop("SCEDAR_B200_TC_TILE_FALLBACK", None)
                ifa, vfa = cells.knn_stripe(k, ex, precision="fast")
                assert torch.equal(i64, ifa) and torch.equal(v64, vfa), (metric, k, ex, forced)
                sfa, _ = cells.knn_stripe(k, ex, 64, 2500, precision="fast")
                assert torch.equal(sfa, i64[64:2500])
            os.environ.pop("SCEDAR_B200_TC_TILE_FALLBACK", None)
        cells.close()
