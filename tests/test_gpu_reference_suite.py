"""The reference's own hot-path tests, restated against the drop-in class.

Every case below is one of the reference's pytest cases for the distance
matrix / kNN path (tests/test_eda/test_dense_sdm.py, test_sparse_sdm.py and
test_no_pdist_sdm.py of scedar v0.2.0; the lines are cited per test), with the
same inputs and the same expected values, run through
``scedar_b200.SampleDistanceMatrix`` on the B200.  The reference keeps three
copies of the suite (dense input, CSR input, ``use_pdist=False``); here the
three are one parametrised suite.  Cases of those files that exercise
plotting, t-SNE / UMAP / PCA, HNSW or the classified-sample subclasses are
outside the path and are not restated.
"""
import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu

R2, R8 = np.sqrt(2), np.sqrt(8)
X_3X2 = [[0, 0], [1, 1], [2, 2]]
X_2X4 = np.array([[0, 1, 2, 3], [1, 2, 0, 6]])
LINE3 = [[v] * 3 for v in (0, 1, 5, 6, 10, 20)]
D_3X3 = np.array([[0, R2, R8], [R2, 0, R2], [R8, R2, 0]])


@pytest.fixture(scope="module")
def SDM():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    import scedar_b200
    from scedar_b200 import _capi
    _capi.load()
    return scedar_b200.SampleDistanceMatrix


def as_form(x, form):
    """dense: what the reference's dense suite passes (lists / arrays);
    sparse: the CSR twin of test_sparse_sdm.py."""
    if form == "sparse":
        return sps.csr_matrix(np.asarray(x, dtype=float))
    return x


FORMS = ["dense", "sparse"]


# ---- construction ----------------------------------------------------------
@pytest.mark.parametrize("form", FORMS)
def test_valid_init(SDM, form):
    """test_dense_sdm.py:23-58, test_sparse_sdm.py:25-47."""
    np.testing.assert_allclose(
        SDM(as_form(X_3X2, form), metric="euclidean").d, D_3X3)
    e = np.sqrt(np.power(X_2X4[0] - X_2X4[1], 2).sum())
    np.testing.assert_allclose(
        SDM(as_form(X_2X4, form), metric="euclidean", nprocs=5).d,
        [[0, e], [e, 0]])
    if form == "dense":
        a, b = X_2X4[0] - X_2X4[0].mean(), X_2X4[1] - X_2X4[1].mean()
        c = 1 - np.dot(a, b) / (np.linalg.norm(a, 2) * np.linalg.norm(b, 2))
        d3 = SDM(X_2X4, metric="correlation", nprocs=5).d
        np.testing.assert_allclose(d3, [[0, 0.3618551], [0.3618551, 0]])
        np.testing.assert_allclose(d3, [[0, c], [c, 0]])
    else:
        # test_sparse_sdm.py:39-41 -- correlation on CSR must raise
        with pytest.raises(ValueError):
            SDM(as_form(X_2X4, form), metric="correlation", nprocs=5).d
    SDM(as_form(X_3X2, form), D_3X3)
    SDM(as_form(X_3X2, form), D_3X3, metric="euclidean")
    SDM([[1, 2]], metric="euclidean")


def test_valid_init_no_pdist(SDM):
    """test_no_pdist_sdm.py:25-45."""
    sdm = SDM(sps.csr_matrix(X_3X2), metric="euclidean", use_pdist=False)
    with pytest.raises(ValueError):
        sdm.d
    with pytest.raises(ValueError):
        SDM(sps.csr_matrix(X_2X4), metric="correlation", use_pdist=False,
            nprocs=5).d
    with pytest.raises(ValueError):
        SDM(sps.csr_matrix(X_3X2), D_3X3, use_pdist=False)
    SDM([[1, 2]], metric="euclidean", use_pdist=False)


@pytest.mark.parametrize("use_pdist", [True, False])
def test_empty_init(SDM, use_pdist):
    """test_dense_sdm.py:60-70, test_sparse_sdm.py:49-62,
    test_no_pdist_sdm.py:47-61."""
    with pytest.raises(ValueError):
        SDM(np.empty(0), metric="euclidean")
    sdm = SDM(np.empty((0, 0)), metric="euclidean", use_pdist=use_pdist)
    assert len(sdm.sids) == 0 and len(sdm.fids) == 0
    assert sdm._x.shape == (0, 0)
    if use_pdist:
        assert sdm._d.shape == (0, 0)
        assert sdm._col_sorted_d.shape == (0, 0)
        assert sdm._col_argsorted_d.shape == (0, 0)
    else:
        for attr in ("_d", "_col_sorted_d", "_col_argsorted_d"):
            with pytest.raises(ValueError):
                getattr(sdm, attr)


@pytest.mark.parametrize("form", FORMS)
def test_init_wrong_metric(SDM, form):
    """test_dense_sdm.py:72-96: the constructor is lazy, ``.d`` raises."""
    x = as_form(X_3X2, form)
    with pytest.raises(Exception):
        SDM(x, metric="precomputed")
    for bad in ("unknown", 1, 1., ("euclidean",), ["euclidean"]):
        SDM(x, metric=bad)
        with pytest.raises(Exception):
            SDM(x, metric=bad).d


@pytest.mark.parametrize("form", FORMS)
def test_init_wrong_d(SDM, form):
    """test_dense_sdm.py:98-129: non-numeric d, wrong shapes."""
    x = as_form(X_3X2, form)
    with pytest.raises(Exception):
        SDM(x, np.array([[0, R2, R8], ["1a1", 0, R2], [R8, R2, 0]]))
    for bad in (np.array([[0, R2], [R2, 0]]),
                np.array([[0, R2], [R2, 0], [1, 2]]), np.arange(6)):
        with pytest.raises(Exception):
            SDM(x, bad)


def test_cache_slots_are_memoised(SDM):
    """test_dense_sdm.py:131-168 checks that ``to_classified`` hands the
    SAME cached objects on (``_lazy_load_d``, ``_lazy_load_col_sorted_d``,
    ``_lazy_load_col_argsorted_d``): the slots must exist under these names and
    hold the arrays the properties return."""
    sdm = SDM(np.arange(100).reshape(50, -1), metric="euclidean")
    sdm.s_ith_nn_d(1)
    sdm.s_ith_nn_ind(1)
    assert sdm._lazy_load_d is not None and sdm._lazy_load_d is sdm._d
    assert sdm._lazy_load_col_sorted_d is sdm._col_sorted_d
    assert sdm._lazy_load_col_argsorted_d is sdm._col_argsorted_d
    assert sdm._lazy_load_col_argsorted_d.dtype == np.int64
    assert sdm._lazy_load_col_sorted_d.shape == (50, 50)


@pytest.mark.parametrize("form", FORMS)
def test_ind_x(SDM, form):
    """test_dense_sdm.py:340-372: subsetting slices D (bitwise)."""
    sids, fids = list("abcdef"), list(range(10, 20))
    np.random.seed(340)
    sdm = SDM(as_form(np.random.ranf(60).reshape(6, -1), form), sids=sids,
              fids=fids)
    ss = sdm.ind_x([0, 5], list(range(9)))
    assert ss._x.shape == (2, 9)
    assert ss.sids == ["a", "f"] and ss.fids == list(range(10, 19))
    np.testing.assert_equal(ss.d, sdm._d[np.ix_((0, 5), (0, 5))])
    for args in ((), (None, None)):
        ss = sdm.ind_x(*args)
        assert ss._x.shape == (6, 10)
        assert ss.sids == sids and ss.fids == fids
        np.testing.assert_equal(ss.d, sdm._d)
    with pytest.raises(IndexError):
        sdm.ind_x([6])
    with pytest.raises(IndexError):
        sdm.ind_x(None, ["a"])


# ---- num_correct_dist_mat ----------------------------------------------------
def test_num_correct_dist_mat(SDM):
    """test_dense_sdm.py:472-501."""
    t = np.array([[0, 1, 2], [0.5, 0, 1.5], [1, 1.6, 0.5]])
    with pytest.warns(UserWarning):
        c = SDM.num_correct_dist_mat(t)
    np.testing.assert_equal(c, [[0, 0.5, 1], [0.5, 0, 1.6], [1, 1.6, 0]])
    assert c is t                                  # in place, same object
    c2 = SDM.num_correct_dist_mat(t, 1)
    np.testing.assert_equal(c2, [[0, 0.5, 1], [0.5, 0, 1], [1, 1, 0]])
    bad = np.array([[0, 0.5], [0.5, 0], [1, 1]])
    with pytest.raises(Exception):
        SDM.num_correct_dist_mat(bad, 1)
    with pytest.raises(Exception):
        SDM.num_correct_dist_mat(bad)


# ---- i-th nearest neighbour ----------------------------------------------------
@pytest.mark.parametrize("form", FORMS)
def test_s_ith_nn_d(SDM, form):
    """test_dense_sdm.py:503-511, test_sparse_sdm.py:470-482."""
    nn = SDM(as_form([[0], [1], [5], [6], [10], [20]], form),
             metric="euclidean")
    np.testing.assert_allclose([0, 0, 0, 0, 0, 0], nn.s_ith_nn_d(0))
    np.testing.assert_allclose([1, 1, 1, 1, 4, 10], nn.s_ith_nn_d(1))
    np.testing.assert_allclose([5, 4, 4, 4, 5, 14], nn.s_ith_nn_d(2))


@pytest.mark.parametrize("form", FORMS)
def test_s_ith_nn_ind(SDM, form):
    """test_dense_sdm.py:513-523, test_sparse_sdm.py:484-496."""
    nn = SDM(as_form(LINE3, form), metric="euclidean")
    np.testing.assert_allclose([0, 1, 2, 3, 4, 5], nn.s_ith_nn_ind(0))
    np.testing.assert_allclose([1, 0, 3, 2, 3, 4], nn.s_ith_nn_ind(1))
    np.testing.assert_allclose([2, 2, 1, 4, 2, 3], nn.s_ith_nn_ind(2))


def test_s_ith_nn_no_pdist_raises(SDM):
    """test_no_pdist_sdm.py:398-411."""
    with pytest.raises(ValueError):
        SDM([[0], [1], [5], [6], [10], [20]], metric="euclidean",
            use_pdist=False).s_ith_nn_d(0)
    with pytest.raises(ValueError):
        SDM(LINE3, metric="euclidean", use_pdist=False).s_ith_nn_ind(0)


# ---- kNN look-up table ----------------------------------------------------------
@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("use_pdist", [True, False])
def test_knn_ind_lut(SDM, form, use_pdist):
    """test_dense_sdm.py:537-551, test_no_pdist_sdm.py:1099-1122 -- including
    the tie-break of sample 2 (equidistant from 0 and 4: 0 wins)."""
    nn = SDM(as_form(LINE3, form), metric="euclidean", use_pdist=use_pdist)
    assert nn.s_knn_ind_lut(0) == dict(zip(range(6), [[]] * 6))
    assert nn.s_knn_ind_lut(1) == dict(zip(range(6), [[1], [0], [3], [2],
                                                      [3], [4]]))
    assert nn.s_knn_ind_lut(2) == dict(zip(range(6), [
        [1, 2], [0, 2], [3, 1], [2, 4], [3, 2], [4, 3]]))
    assert nn.s_knn_ind_lut(3) == dict(zip(range(6), [
        [1, 2, 3], [0, 2, 3], [3, 1, 0], [2, 4, 1], [3, 2, 1], [4, 3, 2]]))
    nn.s_knn_ind_lut(5)
    if not use_pdist:
        line = [[v] for v in range(20)]
        a = SDM(as_form(line, form), metric="euclidean", use_pdist=False)
        b = SDM(as_form(line, form), metric="euclidean", use_pdist=False)
        assert a.s_knn_ind_lut(0) == dict(zip(range(20), [[]] * 20))
        assert a.s_knn_ind_lut(10) == b.s_knn_ind_lut(10)


@pytest.mark.parametrize("use_pdist", [True, False])
def test_knn_ind_lut_wrong_args(SDM, use_pdist):
    """test_dense_sdm.py:553-574, test_no_pdist_sdm.py:1124-1145."""
    nn = SDM(LINE3, metric="euclidean", use_pdist=use_pdist)
    for bad in (-1, -0.5, 6, 6.5, 7):
        with pytest.raises(ValueError):
            nn.s_knn_ind_lut(bad)


# ---- connectivity matrix -------------------------------------------------------
@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("use_pdist,x", [(True, [[1], [2], [6]]),
                                         (False, [[0], [1], [5]])])
def test_s_knn_connectivity_matrix(SDM, form, use_pdist, x):
    """test_dense_sdm.py:1041-1061, test_no_pdist_sdm.py:885-907 (the HNSW
    line is nmslib's and not restated)."""
    nn = SDM(as_form(x, form), metric="euclidean", use_pdist=use_pdist)
    np.testing.assert_allclose([[0, 1, 0], [1, 0, 0], [0, 4, 0]],
                               nn.s_knn_connectivity_matrix(1).toarray())
    assert nn.s_knn_connectivity_matrix(
        1, use_hnsw=False, use_pca=False).shape == (3, 3)
    with pytest.raises(ValueError):
        nn.s_knn_connectivity_matrix(0)
    with pytest.raises(ValueError):
        nn.s_knn_connectivity_matrix(1, use_hnsw=False, use_pca=False,
                                     index_params={})
    with pytest.raises(ValueError):
        nn.s_knn_connectivity_matrix(1, use_hnsw=False, use_pca=False,
                                     index_params=None, query_params={})


# ---- cosine / correlation against sklearn --------------------------------------
@pytest.mark.parametrize("form", FORMS)
def test_cosine_pdist(SDM, form):
    """test_dense_sdm.py:1231-1239, test_sparse_sdm.py:1229-1238
    (assert_allclose, default rtol 1e-7)."""
    import sklearn.metrics.pairwise as skp
    np.random.seed(222)
    x = np.random.ranf(10000).reshape(500, -1)
    skd = skp.pairwise_distances(x, metric="cosine")
    np.testing.assert_allclose(SDM.cosine_pdist(as_form(x, form)), skd)
    np.testing.assert_allclose(SDM(as_form(x, form), metric="cosine")._d, skd)


def test_correlation_pdist(SDM):
    """test_dense_sdm.py:1241-1250; CSR must raise (test_sparse_sdm.py:
    1240-1246)."""
    import sklearn.metrics.pairwise as skp
    np.random.seed(222)
    x = np.random.ranf(10000).reshape(500, -1)
    skd = skp.pairwise_distances(x, metric="correlation")
    np.testing.assert_allclose(SDM.correlation_pdist(x), skd)
    np.testing.assert_allclose(SDM(x, metric="correlation")._d, skd)
    with pytest.raises(ValueError):
        SDM(sps.csr_matrix(x), metric="correlation")._d
