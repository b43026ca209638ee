"""Parity of the consumers behind the path (kNN connectivity /
affinity, RareSampleDetection, FeatureImputation) against the oracle, the
reference's known-answer tests and golden vectors from the live reference.
All through the C ABI; needs a B200: ``pytest -m gpu``."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import knn_oracle as korc
from oracle import sdm_oracle as orc

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


def last_kept(progress, n):
    lk = np.full(n, -1, dtype=np.int64)
    for i, kept in enumerate(progress):
        lk[np.asarray(kept, dtype=np.int64)] = i
    return lk


# ---------------------------------------------------------------- connectivity
def test_connectivity_csr(b200, golden):
    g = golden("consumers")
    x = g["conn_x"]
    n = x.shape[0]
    for use_pdist in (True, False):
        sdm = b200.SampleDistanceMatrix(x, metric="euclidean",
                                        use_pdist=use_pdist)
        conn = sdm.s_knn_connectivity_matrix(6)
        assert sps.isspmatrix_csr(conn) and conn.shape == (n, n)
        assert conn.has_sorted_indices and conn.indices.dtype == np.int32
        idx, dist = orc.knns_precomputed(orc.pdist(x, "euclidean"), 6)
        want = orc.knn_connectivity(idx, dist, n)
        assert np.array_equal(conn.indptr, want.indptr)
        assert np.array_equal(conn.indices, want.indices)
        np.testing.assert_array_equal(conn.data, want.data)
        assert np.isneginf(conn.data).any()
        # the live reference, on the rows whose neighbour set is determined
        gi, gp, gd = g["conn_indices"], g["conn_indptr"], g["conn_data"]
        d = orc.pdist(x, "euclidean")
        for r in range(n):
            others = np.sort(np.delete(d[r], r))
            a = slice(gp[r], gp[r + 1])
            np.testing.assert_array_equal(np.sort(conn.data[a]), np.sort(gd[a]))
            if others[5] != others[6]:
                assert np.array_equal(conn.indices[a], gi[a])


def test_connectivity_reference_vector(b200):
    # tests/test_eda/test_dense_sdm.py:1041-1046
    sdm = b200.SampleDistanceMatrix([[1], [2], [6]], metric="euclidean")
    np.testing.assert_equal(sdm.s_knn_connectivity_matrix(1).toarray(),
                            [[0, 1, 0], [1, 0, 0], [0, 4, 0]])


def test_affinity_edges(b200, golden):
    g = golden("consumers")
    n = g["conn_x"].shape[0]
    ref_conn = sps.csr_matrix((g["conn_data"], g["conn_indices"],
                               g["conn_indptr"]), shape=(n, n))
    s, t, w = b200.SampleDistanceMatrix.knn_conn_mat_to_aff_edges(
        ref_conn, aff_scale=2.5)
    assert np.array_equal(np.stack([s, t], axis=1), g["aff_edges"])
    np.testing.assert_array_equal(w, g["aff_weights"])


# ------------------------------------------------------------- rare detection
RARE_X = [[0, 0], [1, 1], [200, 200], [200, 200], [200, 200],
          [100, 100], [101, 101], [99, 99], [100, 100], [102, 102]]


@pytest.mark.parametrize("use_pdist", [True, False])
def test_rare_reference_vectors(b200, use_pdist):
    # tests/test_knn/test_detection_dense.py:9-24, 170-181
    from scedar_b200 import knn
    sdm = b200.SampleDistanceMatrix(RARE_X, metric="euclidean",
                                    use_pdist=use_pdist)
    rsd = knn.RareSampleDetection(sdm)
    d0 = sdm._d.copy() if use_pdist else None
    resl = rsd.detect_rare_samples([3, 4, 5], [10] * 3, [5, 6, 7])
    assert resl == [list(range(5, 10)), list(range(5, 10)), []]
    assert len(rsd._res_lut) == 3
    assert rsd._res_lut[(3, 10, 5)][1][-1] == resl[0]
    assert rsd._res_lut[(4, 10, 6)][1][-1] == resl[1]
    assert rsd._res_lut[(5, 10, 7)][1][-1] == resl[2]
    if use_pdist:
        np.testing.assert_equal(sdm._d, d0)      # d must not change
    # parameter validation (test_detection_dense.py:140-168)
    for bad in ((0, 1, 1), (10, 1, 1), (1, 0, 1), (1, -0.1, 1), (1, 1, 0),
                ([1, 2], 1, 1), (1, [1, 2], 1), (1, 1, [1, 2])):
        with pytest.raises(ValueError):
            rsd.detect_rare_samples(*bad)


def test_rare_golden(b200, golden):
    from scedar_b200 import knn
    g = golden("consumers")
    x = g["rare_x"]
    n = x.shape[0]
    ks = [int(p[0]) for p in g["rare_params"]]
    cs = [float(p[1]) for p in g["rare_params"]]
    its = [int(p[2]) for p in g["rare_params"]]
    for use_pdist, tag in ((True, "pdist"), (False, "nopdist")):
        sdm = b200.SampleDistanceMatrix(x, metric="euclidean",
                                        use_pdist=use_pdist)
        rsd = knn.RareSampleDetection(sdm)
        resl = rsd.detect_rare_samples(ks, cs, its)
        for j in range(len(ks)):
            res, prog = rsd._res_lut[(ks[j], cs[j], its[j])]
            want = g["rare_{}_lk{}".format(tag, j)]
            assert np.array_equal(last_kept(prog, n), want)
            assert res == resl[j] == np.nonzero(want == its[j])[0].tolist()
    k, cut, it = g["rare_cos_params"]
    rsd = knn.RareSampleDetection(b200.SampleDistanceMatrix(x, metric="cosine"))
    rsd.detect_rare_samples(int(k), float(cut), int(it))
    prog = rsd._res_lut[(int(k), float(cut), int(it))][1]
    assert np.array_equal(last_kept(prog, n), g["rare_cos_lk"])


def test_rare_larger_vs_oracle(b200):
    """4,000 blob cells + outliers: device loop == NumPy restatement run on
    the same (device-computed) distance matrix; pdist and no-pdist agree."""
    from scedar_b200 import knn, synth
    x = synth.blobs(4000, 40, seed=9, n_blobs=6, outlier_frac=0.02)
    sdm = b200.SampleDistanceMatrix(x, metric="euclidean")
    d = sdm._d
    cut = float(np.median(np.sort(d, axis=0)[15]))
    res = knn.RareSampleDetection(sdm).detect_rare_samples(15, cut, 5)[0]
    want, prog = korc.rare_pdist(d, 15, cut, 5)
    assert res == want and 0 < len(res) < 4000
    nd = b200.SampleDistanceMatrix(x, metric="euclidean", use_pdist=False)
    res2 = knn.RareSampleDetection(nd).detect_rare_samples(15, cut, 5)[0]
    assert res2 == want


def test_masked_kth_abi(b200):
    import torch
    from scedar_b200 import device as dev
    rng = np.random.default_rng(3)
    n = 1500
    d = rng.random((n, n))
    d = (d + d.T) / 2
    np.fill_diagonal(d, 0)
    alive = rng.random(n) < 0.6
    dd = torch.from_numpy(d).cuda()
    aa = torch.from_numpy(alive.astype(np.uint8)).cuda()
    sub = np.sort(d[np.ix_(alive, alive)], axis=0)
    for rank in (0, 1, 17, 40, 100):
        got = dev.masked_kth(dd, aa, rank).cpu().numpy()
        assert np.isposinf(got[~alive]).all()
        np.testing.assert_array_equal(got[alive], sub[rank])
    # stripe form
    got = dev.masked_kth(dd[256:700], aa, 5, row_begin=256).cpu().numpy()
    full = dev.masked_kth(dd, aa, 5).cpu().numpy()
    np.testing.assert_array_equal(got, full[256:700])
    # compaction + gather
    sel = dev.compact_alive(aa, int(alive.sum())).cpu().numpy()
    assert np.array_equal(sel, np.nonzero(alive)[0])
    rows = dev.gather_rows(dd, torch.from_numpy(sel).cuda()).cpu().numpy()
    np.testing.assert_array_equal(rows, d[alive])
    big = torch.zeros(300001, dtype=torch.uint8, device="cuda")
    pick = np.sort(rng.choice(300001, size=12345, replace=False))
    big[torch.from_numpy(pick).cuda()] = 1
    assert np.array_equal(dev.compact_alive(big, 12345).cpu().numpy(), pick)


# ----------------------------------------------------------------- imputation
IMP_X = [[0, 0], [1, 1], [1, 2], [2, 3], [2, 5], [0, 100], [0, 100],
         [0, 101], [10, 100]]


def test_impute_reference_vectors(b200):
    # tests/test_knn/test_imputation.py:12-76
    from scedar_b200 import knn
    tsdm = b200.SampleDistanceMatrix(IMP_X, metric="euclidean")
    fmp = knn.FeatureImputation(tsdm)
    d = fmp._sdm._d.copy()
    res = fmp.impute_features([1, 3, 5], [1, 1, 3], [0.5, 1.5, 0.5], [1, 1, 5])
    assert len(fmp._res_lut) == 3
    assert all(r._x.shape == (9, 2) for r in res)
    assert all(not np.array_equal(r._d, d) for r in res)
    np.testing.assert_equal(res[0]._x, np.array(
        [[1, 1], [1, 1], [1, 2], [2, 3], [2, 5], [0, 100], [0, 100],
         [0, 101], [10, 100]]))
    np.testing.assert_equal(res[1]._x, np.array(
        [[np.median([1, 1, 2]), np.median([1, 2, 3])],
         [np.median([1, 1, 2]), np.median([1, 2, 3])],
         [np.median([1, 1, 2]), 2], [2, 3], [2, 5],
         [np.median([0, 0, 10]), 100], [np.median([0, 0, 10]), 100],
         [np.median([0, 0, 10]), 101], [10, 100]]))
    assert res[2]._x[0, 0] > 0 and res[2]._x[0, 1] > 0
    assert (res[2]._x[5:8, 0] > 0).all()
    np.testing.assert_equal(fmp._res_lut[(3, 1, 1.5, 1, np.median)][0],
                            res[1]._x)
    np.testing.assert_equal(fmp._sdm._d, d)
    np.testing.assert_equal(fmp._sdm.x, IMP_X)
    stats = fmp._res_lut[(3, 1, 1.5, 1, np.median)][2]
    assert "picked up 8 total features in 6 (66.7%) cells" in stats
    assert fmp._res_lut[(3, 1, 1.5, 1, np.median)][1].toarray().tolist() == [
        [1, 1], [1, 1], [1, 0], [0, 0], [0, 0], [1, 0], [1, 0], [1, 0], [0, 0]]
    for bad in ((0, 1, 1, 1), (9, 1, 1, 1), (3, 4, 1, 1), (3, 0, 1, 1),
                (3, 1, 1, 0), ([1, 2], 1, 1, 1)):
        with pytest.raises(ValueError):
            fmp.impute_features(*bad)
    with pytest.raises(ValueError):
        fmp.impute_features(3, 1, 1, 1, statistic_fun=1)


def test_impute_golden(b200, golden):
    from scedar_b200 import knn
    g = golden("consumers")
    x = g["imp_x"]
    sdm = b200.SampleDistanceMatrix(x, metric="euclidean")
    fmp = knn.FeatureImputation(sdm)
    P = g["imp_params"]
    res = fmp.impute_features([int(p[0]) for p in P], [int(p[1]) for p in P],
                              [float(p[2]) for p in P], [int(p[3]) for p in P])
    for j, p in enumerate(P):
        np.testing.assert_array_equal(res[j]._x, g["imp_x{}".format(j)])
        key = (int(p[0]), int(p[1]), float(p[2]), int(p[3]), np.median)
        idc = fmp._res_lut[key][1]
        assert sps.issparse(idc)
        np.testing.assert_array_equal(idc.toarray(), g["imp_idc{}".format(j)])
    # sparse x in, sparse x out (imputation.py:44-47, 101-102)
    sp = knn.FeatureImputation(b200.SampleDistanceMatrix(
        sps.csr_matrix(x), metric="euclidean"))
    r = sp.impute_features(int(P[0][0]), int(P[0][1]), float(P[0][2]),
                           int(P[0][3]))[0]
    assert sps.issparse(r._x)
    np.testing.assert_array_equal(r._x.toarray(), g["imp_x0"])


def test_impute_larger_vs_oracle(b200):
    from scedar_b200 import knn, synth
    x = synth.log_counts(1200, 333, seed=77)
    sdm = b200.SampleDistanceMatrix(x, metric="correlation")
    for k, n_do, mpv, it in ((15, 5, 1.0, 2), (16, 4, 0.7, 3)):
        lut = sdm.s_knn_ind_lut(k)
        arr = np.array([lut[i] for i in range(x.shape[0])])
        got = knn.FeatureImputation._impute_features_runner(
            x, lut, k, n_do, mpv, it)
        want, idc, (entries, cells) = korc.impute(x, arr, k, n_do, mpv, it)
        np.testing.assert_array_equal(got[0], want)
        np.testing.assert_array_equal(got[1].toarray(), idc)
        assert "picked up {} total features in {} ".format(entries, cells) \
            in got[2].splitlines()[-2]


def test_striped_consumers_single_rank(b200):
    """StripedDistanceMatrix.rare_samples / impute (world 1, CUDA operators)
    agree with the drop-in classes and the oracle."""
    import torch
    from scedar_b200 import device as dev, knn, synth
    from scedar_b200._capi import EXCLUDE_FIRST
    from scedar_b200.stripes import StripedDistanceMatrix
    x = synth.blobs(1500, 25, seed=2, n_blobs=3, outlier_frac=0.03)
    sdm = StripedDistanceMatrix(dev.to_device(x), "euclidean")
    d = sdm.compute_d().cpu().numpy()
    cut = float(np.median(np.sort(d, axis=0)[8]) * 1.1)
    lk = sdm.rare_samples(8, cut, 5).cpu().numpy()
    res, prog = korc.rare_pdist(d, 8, cut, 5)
    assert np.array_equal(lk, last_kept(prog, 1500)) and 0 < len(res) < 1500
    xi = synth.log_counts(1500, 90, seed=6)
    s2 = StripedDistanceMatrix(dev.to_device(xi), "correlation")
    idx, _ = s2.knn(9, EXCLUDE_FIRST)
    got, idc, stats = s2.impute(dev.to_device(xi), idx, 9, 3, 0.6, 2)
    wx, widc, wstats = korc.impute(xi, idx.cpu().numpy(), 9, 3, 0.6, 2)
    np.testing.assert_array_equal(got.cpu().numpy(), wx)
    np.testing.assert_array_equal(idc.cpu().numpy(), widc)
    assert stats == wstats
    sdm.close()
    s2.close()
