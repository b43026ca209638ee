"""Pin the CPU oracle: reference known answers + golden vectors produced by
the live reference (tests/golden/make_golden.py).  CPU only."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import sdm_oracle as orc

DENSE_CASES = [("ranf222", False), ("counts_ties", True), ("edge_rows", True),
               ("logcounts", False)]


def _metrics(g):
    return sorted({k.split("_")[0] for k in g.files
                   if k.endswith("_lut")})


@pytest.mark.parametrize("name,full", DENSE_CASES)
def test_dense_golden(golden, name, full):
    g = golden(name)
    x, k = g["x"], int(g["k"])
    for m in _metrics(g):
        d = orc.pdist(x, m)
        assert d.dtype == np.float64 and d.shape == (x.shape[0],) * 2
        assert np.array_equal(d, d.T) and not d.diagonal().any()
        if full:
            np.testing.assert_allclose(d, g[m + "_d"], rtol=1e-12, atol=1e-15)
        else:
            np.testing.assert_allclose(d[:8], g[m + "_d_rows"], rtol=1e-12,
                                       atol=1e-15)
            np.testing.assert_allclose(d[::25, ::25], g[m + "_d_grid"],
                                       rtol=1e-12, atol=1e-15)
        # index parity is asserted on the reference's own D when it is stored
        # (integer inputs => exact ties must resolve identically)
        dd = g[m + "_d"] if full else d
        arg = orc.col_argsorted_d(dd)
        if full:
            assert np.array_equal(arg, g[m + "_arg"])
        else:
            assert np.array_equal(arg[:k + 2], g[m + "_arg_top"])
            np.testing.assert_allclose(orc.col_sorted_d(dd)[:k + 2],
                                       g[m + "_sorted_top"], rtol=1e-12,
                                       atol=1e-15)
        lut = orc.knn_ind_lut(dd, k)
        assert np.array_equal(np.array([lut[i] for i in range(len(lut))]),
                              g[m + "_lut"])
        idx, dist = orc.knn_drop_first(dd, k)
        assert np.array_equal(idx, g[m + "_lut"])


@pytest.mark.parametrize("name", ["ranf222", "logcounts"])
def test_knns_golden_continuous(golden, name):
    """sklearn's kneighbors: no ties on continuous data, so the reference's
    output is fully determined and the oracle must match it exactly."""
    g = golden(name)
    x, k = g["x"], int(g["k"])
    for m in _metrics(g):
        idx, dist = orc.knns_precomputed(orc.pdist(x, m), k)
        assert np.array_equal(idx, g[m + "_knn_idx"])
        np.testing.assert_allclose(dist, g[m + "_knn_dist"], rtol=1e-12)
        ridx, rdist = orc.knns_raw(x, m, k)
        assert np.array_equal(ridx, g[m + "_raw_knn_idx"])
        # the no-pdist path uses sklearn's own distance formulas
        np.testing.assert_allclose(rdist, g[m + "_raw_knn_dist"], rtol=1e-9,
                                   atol=1e-12)


@pytest.mark.parametrize("name", ["counts_ties", "edge_rows"])
def test_knns_golden_ties(golden, name):
    """With exact ties sklearn's order is unspecified: distances must agree,
    and indices must agree wherever the reference's row has no tie."""
    g = golden(name)
    k = int(g["k"])
    for m in _metrics(g):
        d = g[m + "_d"]
        idx, dist = orc.knns_precomputed(d, k)
        np.testing.assert_array_equal(dist, g[m + "_knn_dist"])
        srt = np.sort(d, axis=1)[:, :k + 2]
        untied = (np.diff(srt, axis=1) != 0).all(axis=1)
        assert np.array_equal(idx[untied], g[m + "_knn_idx"][untied])
        # every returned neighbour really is at the returned distance
        assert np.array_equal(np.take_along_axis(d, idx, axis=1), dist)


def test_csr_golden(golden):
    g = golden("csr_lognormal")
    xs = sps.csr_matrix((g["data"], g["indices"], g["indptr"]),
                        shape=tuple(g["shape"]))
    k = int(g["k"])
    for m in ("cosine", "euclidean"):
        d = orc.pdist(xs, m)
        np.testing.assert_allclose(d, g[m + "_d"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(d, orc.pdist(xs.toarray(), m), rtol=1e-12,
                                   atol=1e-14)
        assert np.array_equal(orc.col_argsorted_d(d)[:17], g[m + "_arg_top"])
        # disjoint supports give exact cosine ties (distance 1): sklearn's tie
        # order is unspecified, so compare indices on untied rows only
        idx, dist = orc.knns_precomputed(d, k)
        np.testing.assert_allclose(dist, g[m + "_knn_dist"], rtol=1e-12)
        srt = np.sort(g[m + "_d"], axis=1)[:, :k + 2]
        untied = (np.diff(srt, axis=1) != 0).all(axis=1)
        assert untied.sum() > 100
        assert np.array_equal(idx[untied], g[m + "_knn_idx"][untied])
        conn = orc.knn_connectivity(*orc.knns_precomputed(d, 5), d.shape[0])
        want = sps.csr_matrix((g[m + "_conn_data"], g[m + "_conn_indices"],
                               g[m + "_conn_indptr"]), shape=d.shape)
        assert conn.nnz == want.nnz == 5 * d.shape[0]
        np.testing.assert_allclose(conn.toarray()[untied],
                                   want.toarray()[untied], rtol=1e-12)
    with pytest.raises(ValueError):
        orc.pdist(xs, "correlation")     # test_sparse_sdm.py:39-41


def test_num_correct_dist_mat_golden(golden):
    g = golden("ncdm")
    for src, want, ub, nw in (("raw", "fixed", None, g["n_warn"][0]),
                              ("raw", "fixed_ub", 0.7, g["n_warn"][1]),
                              ("sym", "fixed_sym", None, g["n_warn"][2])):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            m = g[src].copy()
            out = orc.num_correct_dist_mat(m, ub)
        assert out is m
        assert np.array_equal(out, g[want])
        assert len(w) == int(nw)


# ---- the reference's own known answers -------------
def test_known_answers_d():
    # tests/test_eda/test_dense_sdm.py:23-52
    d = orc.pdist([[0, 0], [1, 1], [2, 2]], "euclidean")
    r2, r8 = np.sqrt(2), np.sqrt(8)
    np.testing.assert_allclose(d, [[0, r2, r8], [r2, 0, r2], [r8, r2, 0]])
    x = np.array([[0, 1, 2, 3], [1, 2, 0, 6]])
    np.testing.assert_allclose(orc.pdist(x, "correlation"),
                               [[0, 0.3618551], [0.3618551, 0]], rtol=1e-6)
    e = np.sqrt(np.power(x[0] - x[1], 2).sum())
    np.testing.assert_allclose(orc.pdist(x, "euclidean"), [[0, e], [e, 0]])


def test_known_answers_num_correct():
    # tests/test_eda/test_dense_sdm.py:472-501
    t = np.array([[0, 1, 2], [0.5, 0, 1.5], [1, 1.6, 0.5]])
    with pytest.warns(UserWarning):
        c = orc.num_correct_dist_mat(t.copy())
    assert np.array_equal(c, [[0, 0.5, 1], [0.5, 0, 1.6], [1, 1.6, 0]])
    with pytest.warns(UserWarning):
        c = orc.num_correct_dist_mat(t.copy(), 1)
    assert np.array_equal(c, [[0, 0.5, 1], [0.5, 0, 1], [1, 1, 0]])
    with pytest.raises(ValueError):
        orc.num_correct_dist_mat(np.zeros((3, 2)))


def test_known_answers_nn():
    # tests/test_eda/test_dense_sdm.py:503-551
    d = orc.pdist([[0], [1], [5], [6], [10], [20]], "euclidean")
    s = orc.col_sorted_d(d)
    np.testing.assert_allclose(s[0], 0)
    np.testing.assert_allclose(s[1], [1, 1, 1, 1, 4, 10])
    np.testing.assert_allclose(s[2], [5, 4, 4, 4, 5, 14])
    x3 = [[v] * 3 for v in (0, 1, 5, 6, 10, 20)]
    d = orc.pdist(x3, "euclidean")
    a = orc.col_argsorted_d(d)
    assert a[0].tolist() == [0, 1, 2, 3, 4, 5]
    assert a[1].tolist() == [1, 0, 3, 2, 3, 4]
    assert a[2].tolist() == [2, 2, 1, 4, 2, 3]
    assert orc.knn_ind_lut(d, 0) == dict(zip(range(6), [[]] * 6))
    # sample 2 is equidistant from 0 and 4: the lower index wins
    assert orc.knn_ind_lut(d, 3) == dict(zip(range(6), [
        [1, 2, 3], [0, 2, 3], [3, 1, 0], [2, 4, 1], [3, 2, 1], [4, 3, 2]]))
    idx, _ = orc.knns_precomputed(d, 3)
    assert idx.tolist() == [[1, 2, 3], [0, 2, 3], [3, 1, 0], [2, 4, 1],
                            [3, 2, 1], [4, 3, 2]]
    for bad in (-1, -0.5, 6, 6.5, 7):
        with pytest.raises(ValueError):
            orc.knn_ind_lut(d, bad)


def test_known_answers_connectivity():
    # tests/test_eda/test_dense_sdm.py:1041-1046
    d = orc.pdist([[1], [2], [6]], "euclidean")
    conn = orc.knn_connectivity(*orc.knns_precomputed(d, 1), 3)
    np.testing.assert_allclose(conn.toarray(),
                               [[0, 1, 0], [1, 0, 0], [0, 4, 0]])
    with pytest.raises(ValueError):
        orc.knns_precomputed(d, 0)


def test_sklearn_crosscheck():
    # tests/test_eda/test_dense_sdm.py:1231-1250 (assert_allclose, rtol 1e-7)
    import sklearn.metrics.pairwise as skp
    np.random.seed(222)
    x = np.random.ranf(10000).reshape(500, -1)
    for m in ("cosine", "correlation", "euclidean"):
        np.testing.assert_allclose(orc.pdist(x, m),
                                   skp.pairwise_distances(x, metric=m),
                                   atol=1e-12)


def test_duplicate_rows_accessors_disagree():
    # the LUT accessor drops rank 0, kneighbors drops self
    x = np.array([[0, 0], [1, 1], [200, 200], [200, 200], [200, 200],
                  [100, 100]], dtype=float)
    d = orc.pdist(x, "euclidean")
    assert orc.knn_ind_lut(d, 2)[3] == [3, 4]
    assert orc.knns_precomputed(d, 2)[0][3].tolist() == [2, 4]


def test_empty():
    # tests/test_eda/test_dense_sdm.py:60-70
    d = orc.pdist(np.empty((0, 0)), "euclidean")
    assert d.shape == (0, 0)
    assert orc.col_sorted_d(d).shape == (0, 0)
    assert orc.col_argsorted_d(d).shape == (0, 0)
