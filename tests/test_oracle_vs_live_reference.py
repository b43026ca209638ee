"""The oracle against the LIVE reference on seeded random inputs.  Runs only
where /root/reference exists (the build container); on the GPU box the
committed golden vectors stand in for it."""
import warnings

import numpy as np
import pytest

from oracle import knn_oracle as korc
from oracle import sdm_oracle as orc
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(),
                                reason="live reference only in the build container")


@pytest.fixture(scope="module")
def ref():
    warnings.simplefilter("ignore")
    sdm = load_reference()
    import scedar.knn as ref_knn
    return sdm, ref_knn


def last_kept(progress, n):
    lk = np.full(n, -1, dtype=np.int64)
    for i, kept in enumerate(progress):
        lk[np.asarray(kept, dtype=np.int64)] = i
    return lk


@pytest.mark.parametrize("seed", range(8))
def test_pdist_and_knn(ref, seed):
    SDM, _ = ref
    rng = np.random.default_rng(seed)
    n, p = int(rng.integers(5, 120)), int(rng.integers(2, 60))
    x = rng.normal(size=(n, p)) if seed % 2 else np.log1p(rng.poisson(1.0, (n, p)) * 1.0)
    x = x + rng.random((n, p)) * 1e-3          # no exact ties: sklearn's order is then defined
    k = int(rng.integers(1, n - 1))
    for metric in ("cosine", "correlation", "euclidean"):
        s = SDM(x, metric=metric)
        d = orc.pdist(x, metric)
        np.testing.assert_allclose(d, s._d, rtol=1e-10, atol=1e-13)
        lut = s.s_knn_ind_lut(k)
        want = orc.knn_ind_lut(s._d, k)
        assert all(lut[i] == want[i] for i in range(n))
        t, dist = s.s_knns(k)
        oi, od = orc.knns_precomputed(s._d, k)
        assert np.array_equal(np.array(t), oi)
        np.testing.assert_allclose(np.array(dist), od, rtol=1e-12)
        t2, _ = SDM(x, metric=metric, use_pdist=False).s_knns(k)
        assert np.array_equal(np.array(t2), orc.knns_raw(x, metric, k)[0])
        conn = s.s_knn_connectivity_matrix(min(k, 5))
        oc = orc.knn_connectivity(*orc.knns_precomputed(s._d, min(k, 5)), n)
        assert (conn != oc).nnz == 0


@pytest.mark.parametrize("seed", range(6))
def test_rare_and_impute(ref, seed):
    SDM, ref_knn = ref
    rng = np.random.default_rng(100 + seed)
    n, p = int(rng.integers(30, 150)), int(rng.integers(3, 40))
    x = np.concatenate([rng.normal(0, 1, (n - 6, p)), rng.uniform(-8, 8, (6, p))])
    x = x + rng.random((n, p)) * 1e-3
    k = int(rng.integers(2, 9))
    n_iter = int(rng.integers(1, 6))
    s = SDM(x, metric="euclidean")
    cut = float(np.median(np.sort(s._d, axis=0)[k]) * rng.uniform(0.9, 1.4))
    rsd = ref_knn.RareSampleDetection(s)
    res = rsd.detect_rare_samples(k, cut, n_iter)[0]
    ores, oprog = korc.rare_pdist(s._d, k, cut, n_iter)
    assert res == ores
    assert np.array_equal(last_kept(rsd._res_lut[(k, cut, n_iter)][1], n),
                          last_kept(oprog, n))
    nd = ref_knn.RareSampleDetection(SDM(x, metric="euclidean", use_pdist=False))
    try:
        res2 = nd.detect_rare_samples(k, cut, n_iter)[0]
    except TypeError:            # the reference's parmap swallowed sklearn's ValueError
        with pytest.raises(ValueError):
            korc.rare_no_pdist(x, "euclidean", k, cut, n_iter)
    else:
        assert res2 == korc.rare_no_pdist(x, "euclidean", k, cut, n_iter)[0]
    # imputation on a thresholded copy
    xi = np.where(x > 0.2, x, 0.0)
    si = SDM(xi, metric="euclidean")
    n_do = int(rng.integers(1, k + 1))
    fi = ref_knn.FeatureImputation(si)
    out = fi.impute_features(k, n_do, 0.2, n_iter)[0]
    lut = si.s_knn_ind_lut(k)
    arr = np.array([lut[i] for i in range(n)])
    ox, oidc, _ = korc.impute(xi, arr, k, n_do, 0.2, n_iter)
    np.testing.assert_array_equal(np.asarray(out._x), ox)
    key = (k, n_do, 0.2, n_iter, np.median)
    np.testing.assert_array_equal(fi._res_lut[key][1].toarray(), oidc)


@pytest.mark.parametrize("seed", range(6))
def test_hclust(ref, seed):
    """oracle/hclust_oracle.py against the live HClustTree (sdm.py:1369-1693)
    on the reference's own D (bit-identical input, so the linkage matrices
    must be identical too)."""
    from oracle import hclust_oracle as horc
    import scedar.eda as ref_eda
    SDM, _ = ref
    H = ref_eda.HClustTree
    rng = np.random.default_rng(200 + seed)
    n, p = int(rng.integers(6, 90)), int(rng.integers(2, 30))
    x = rng.normal(size=(n, p)) + rng.integers(0, 3, size=(n, 1)) * 2.0
    metric = ("euclidean", "cosine", "correlation")[seed % 3]
    if metric == "correlation" and p < 3:
        p = 3
        x = rng.normal(size=(n, p))
    d = SDM(x, metric=metric)._d
    euc = metric == "euclidean"
    for lt in ("single", "complete", "average", "weighted") + \
            (("centroid", "median", "ward") if euc else ()):
        z, _ = horc.hclust_linkage(d, linkage=lt)
        np.testing.assert_array_equal(z, H.hclust_linkage(d, linkage=lt))
    rounds = int(rng.integers(-2, 12))
    for oo in (False, True):
        z, _ = horc.hclust_linkage(d, linkage="auto", n_eval_rounds=rounds,
                                   is_euc_dist=euc, optimal_ordering=oo)
        want = H.hclust_linkage(d, linkage="auto", n_eval_rounds=rounds,
                                is_euc_dist=euc, optimal_ordering=oo)
        np.testing.assert_array_equal(z, want)
    tree = H.hclust_tree(d, linkage="average")
    z, _ = horc.hclust_linkage(d, linkage="average")
    assert tree.leaf_ids() == horc.leaf_ids(z)
    r = int(rng.integers(1, 8))
    import scipy.cluster.hierarchy as sch
    assert tree.n_round_bipar_cnt(r) == horc.n_round_bipar_cnt(sch.to_tree(z), r)
    assert H.sort_x_by_d(x, metric=metric, optimal_ordering=True) == \
        horc.sort_x_by_d(x, metric=metric, optimal_ordering=True)
    # the product's host logic (no device involved) on the same linkage
    from scedar_b200 import hclust as ph
    assert ph.bipartition_counts(z, r) == tree.n_round_bipar_cnt(r)
    sf = horc.squareform(d)
    for oo in (False, True):
        got = ph.linkage_from_condensed(sf, n, linkage="auto",
                                        n_eval_rounds=rounds, is_euc_dist=euc,
                                        optimal_ordering=oo)
        want = H.hclust_linkage(d, linkage="auto", n_eval_rounds=rounds,
                                is_euc_dist=euc, optimal_ordering=oo)
        np.testing.assert_array_equal(got, want)
    # MultinomialMdl.mdl on count vectors (mdl.py:86-106)
    from scedar.eda.mdl import MultinomialMdl
    for c in tree.n_round_bipar_cnt(r) + [[], [4], [2, 2, 2]]:
        assert horc.multinomial_mdl(c) == MultinomialMdl(np.array(c)).mdl
        assert ph._multinomial_mdl(np.array(c)) == MultinomialMdl(np.array(c)).mdl
