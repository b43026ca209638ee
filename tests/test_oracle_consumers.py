"""Pin oracle/knn_oracle.py (rare-sample detection, imputation, affinity
edges): the reference's known-answer tests + golden vectors produced by the
live reference (tests/golden/make_golden_consumers.py).  CPU only."""
import numpy as np
import pytest

from oracle import knn_oracle as korc
from oracle import sdm_oracle as orc


def last_kept(progress, n):
    lk = np.full(n, -1, dtype=np.int64)
    for i, kept in enumerate(progress):
        lk[np.asarray(kept, dtype=np.int64)] = i
    return lk


# reference: tests/test_knn/test_detection_dense.py:9-24, 170-181
RARE_X = np.array([[0, 0], [1, 1], [200, 200], [200, 200], [200, 200],
                   [100, 100], [101, 101], [99, 99], [100, 100], [102, 102]],
                  dtype=np.float64)
RARE_CASES = [((3, 10, 5), list(range(5, 10))), ((4, 10, 6), list(range(5, 10))),
              ((5, 10, 7), [])]


@pytest.mark.parametrize("params,want", RARE_CASES)
def test_rare_known_answers(params, want):
    d = orc.pdist(RARE_X, "euclidean")
    res, prog = korc.rare_pdist(d, *params)
    assert res == want and prog[-1] == want and len(prog) == params[2] + 1
    assert prog[0] == list(range(10))
    res, prog = korc.rare_no_pdist(RARE_X, "euclidean", *params)
    assert res == want and prog[-1] == want


def test_rare_golden(golden):
    g = golden("consumers")
    x = g["rare_x"]
    d = orc.pdist(x, "euclidean")
    for j, (k, cut, it) in enumerate(g["rare_params"]):
        res, prog = korc.rare_pdist(d, int(k), float(cut), int(it))
        assert np.array_equal(last_kept(prog, x.shape[0]),
                              g["rare_pdist_lk{}".format(j)])
        assert 0 < len(res) < x.shape[0]       # the case is not degenerate
        res, prog = korc.rare_no_pdist(x, "euclidean", int(k), float(cut),
                                       int(it))
        assert np.array_equal(last_kept(prog, x.shape[0]),
                              g["rare_nopdist_lk{}".format(j)])
    k, cut, it = g["rare_cos_params"]
    res, prog = korc.rare_pdist(orc.pdist(x, "cosine"), int(k), float(cut),
                                int(it))
    assert np.array_equal(last_kept(prog, x.shape[0]), g["rare_cos_lk"])


# reference: tests/test_knn/test_imputation.py:12-43
IMP_X = np.array([[0, 0], [1, 1], [1, 2], [2, 3], [2, 5], [0, 100], [0, 100],
                  [0, 101], [10, 100]], dtype=np.float64)


def _lut_arr(x, k):
    lut = orc.knn_ind_lut(orc.pdist(x, "euclidean"), k)
    return np.array([lut[i] for i in range(x.shape[0])])


def test_impute_known_answers():
    got, idc, _ = korc.impute(IMP_X, _lut_arr(IMP_X, 1), 1, 1, 0.5, 1)
    want = np.array([[1, 1], [1, 1], [1, 2], [2, 3], [2, 5], [0, 100],
                     [0, 100], [0, 101], [10, 100]], dtype=np.float64)
    np.testing.assert_equal(got, want)
    got, idc, _ = korc.impute(IMP_X, _lut_arr(IMP_X, 3), 3, 1, 1.5, 1)
    want = np.array([[np.median([1, 1, 2]), np.median([1, 2, 3])],
                     [np.median([1, 1, 2]), np.median([1, 2, 3])],
                     [np.median([1, 1, 2]), 2], [2, 3], [2, 5],
                     [np.median([0, 0, 10]), 100],
                     [np.median([0, 0, 10]), 100],
                     [np.median([0, 0, 10]), 101], [10, 100]])
    np.testing.assert_equal(got, want)
    assert idc[0, 0] == 1 and idc[3, 0] == 0
    got, _, _ = korc.impute(IMP_X, _lut_arr(IMP_X, 5), 5, 3, 0.5, 5)
    assert got[0, 0] > 0 and got[0, 1] > 0 and (got[5:8, 0] > 0).all()


def test_impute_golden(golden):
    g = golden("consumers")
    x = g["imp_x"]
    for j, (k, n_do, mpv, it) in enumerate(g["imp_params"]):
        knn = g["imp_knn{}".format(j)]
        assert np.array_equal(_lut_arr(x, int(k)), knn)
        got, idc, (entries, cells) = korc.impute(x, knn, int(k), int(n_do),
                                                 float(mpv), int(it))
        np.testing.assert_array_equal(got, g["imp_x{}".format(j)])
        np.testing.assert_array_equal(idc, g["imp_idc{}".format(j)])
        assert entries == (idc > 0).sum() > 0


def test_connectivity_and_affinity_golden(golden):
    g = golden("consumers")
    x = g["conn_x"]
    idx, dist = orc.knns_precomputed(orc.pdist(x, "euclidean"), 6)
    conn = orc.knn_connectivity(idx, dist, x.shape[0])
    assert conn.has_sorted_indices and int(g["conn_sorted"][0]) == 1
    # sklearn leaves exact ties unordered, so WHICH of several cells tied at
    # the k-th distance is kept is unspecified: rows without a tie at that
    # boundary have a determined neighbour set (CSR columns are sorted) and
    # must match exactly, the others as multisets of distances
    gi, gp, gd = g["conn_indices"], g["conn_indptr"], g["conn_data"]
    assert np.array_equal(conn.indptr, gp)
    assert np.isneginf(gd).any()               # exact zero distances occur
    d = orc.pdist(x, "euclidean")
    exact_rows = 0
    for r in range(x.shape[0]):
        a, b = slice(gp[r], gp[r + 1]), slice(conn.indptr[r], conn.indptr[r + 1])
        np.testing.assert_array_equal(np.sort(conn.data[b]), np.sort(gd[a]))
        others = np.sort(np.delete(d[r], r))
        if others[5] != others[6]:
            assert np.array_equal(conn.indices[b], gi[a]), r
            np.testing.assert_array_equal(conn.data[b], gd[a])
            exact_rows += 1
    assert exact_rows >= 10
    # affinity arithmetic on the reference's own matrix
    import scipy.sparse as sps
    ref_conn = sps.csr_matrix((gd, gi, gp), shape=(x.shape[0],) * 2)
    s, t, w = korc.affinity_weights(ref_conn, aff_scale=2.5)
    assert np.array_equal(np.stack([s, t], axis=1), g["aff_edges"])
    np.testing.assert_array_equal(w, g["aff_weights"])
