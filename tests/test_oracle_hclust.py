"""Pin oracle/hclust_oracle.py: the reference's known-answer tests
(tests/test_eda/test_dense_sdm.py:170-193, 1255-1304) + golden linkage
matrices produced by the live reference (tests/golden/make_golden_hclust.py),
and the host-side logic of scedar_b200.hclust that needs no device.  CPU only."""
import numpy as np
import pytest
import scipy.cluster.hierarchy as sch
import scipy.spatial.distance as spdist

from oracle import hclust_oracle as horc
from oracle import sdm_oracle as orc

METRICS = ("euclidean", "cosine", "correlation")
LINKAGES = ("single", "complete", "average", "weighted", "ward")


def assert_same_linkage(z, want):
    assert z.shape == want.shape
    np.testing.assert_array_equal(z[:, [0, 1, 3]], want[:, [0, 1, 3]])
    np.testing.assert_allclose(z[:, 2], want[:, 2], rtol=1e-9, atol=1e-12)


# reference: tests/test_eda/test_dense_sdm.py:1255-1304
X_5X2 = np.array([[0, 0], [100, 100], [1, 1], [101, 101], [80, 80]],
                 dtype=np.float64)


def test_known_tree_5x2():
    d = orc.pdist(X_5X2, "euclidean")
    z, _ = horc.hclust_linkage(d, linkage="auto")
    root = sch.to_tree(z)
    assert horc.leaf_ids(z) == [0, 2, 4, 1, 3]
    assert root.get_left().get_count() == 2
    assert root.get_right().get_count() == 3
    assert root.get_left().pre_order(lambda nd: nd.get_id()) == [0, 2]
    assert root.get_right().pre_order(lambda nd: nd.get_id()) == [4, 1, 3]
    # is_euc_dist widens the candidate set, n_eval_rounds=-1 is clamped
    z2, _ = horc.hclust_linkage(d, linkage="auto", n_eval_rounds=-1,
                                is_euc_dist=True)
    assert sorted(horc.leaf_ids(z2)) == [0, 1, 2, 3, 4]


# reference: tests/test_eda/test_dense_sdm.py:170-193
@pytest.mark.parametrize("x", [
    [[0, 5, 30, 10], [1, 5, 30, 10], [0, 5, 33, 10], [2, 5, 30, 7],
     [2, 5, 30, 9]],
    [[0, 0, 30, 10], [1, 2, 30, 10], [0, 3, 33, 10], [2, 4, 30, 7],
     [2, 5, 30, 9]]])
def test_known_sort_x_by_d(x):
    x = np.array(x, dtype=np.float64)
    assert horc.sort_x_by_d(x.T, metric="euclidean",
                            optimal_ordering=True) == [2, 3, 1, 0]


def test_invalid_dmat():
    # reference: tests/test_eda/test_dense_sdm.py:1306-1311
    with pytest.raises(ValueError):
        horc.hclust_linkage(np.arange(5))
    with pytest.raises(ValueError):
        horc.hclust_linkage(np.arange(10).reshape(2, 5))


def test_squareform_is_scipys():
    horc._self_check()
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 3, 17, 130):
        d = spdist.squareform(spdist.pdist(rng.normal(size=(n, 4)))) \
            if n > 1 else np.zeros((n, n))
        got = horc.squareform(d)
        np.testing.assert_array_equal(got, spdist.squareform(d))
    bad = np.zeros((3, 3))
    bad[0, 1] = np.nan
    bad[1, 0] = np.nan
    with pytest.raises(ValueError, match="symmetric"):
        horc.squareform(bad)
    bad = np.zeros((3, 3))
    bad[1, 1] = 1e-300
    with pytest.raises(ValueError, match="diagonal"):
        horc.squareform(bad)


@pytest.mark.parametrize("metric", METRICS)
def test_golden_linkages(golden, metric):
    g = golden("hclust")
    d = orc.pdist(g["x"], metric)
    for lt in LINKAGES:
        z, _ = horc.hclust_linkage(d, linkage=lt)
        assert_same_linkage(z, g["z_{}_{}".format(metric, lt)])
    z, _ = horc.hclust_linkage(d, linkage="auto",
                               is_euc_dist=(metric == "euclidean"))
    assert_same_linkage(z, g["z_{}_auto".format(metric)])
    z, _ = horc.hclust_linkage(d, linkage="auto", n_eval_rounds=11,
                               is_euc_dist=(metric == "euclidean"),
                               optimal_ordering=True)
    assert_same_linkage(z, g["z_{}_auto_oo".format(metric)])
    assert horc.sort_x_by_d(g["x"], metric=metric) == \
        g["leaves_{}".format(metric)].tolist()
    z, _ = horc.hclust_linkage(d, linkage="average")
    cnt = horc.n_round_bipar_cnt(sch.to_tree(z), 5)
    want = [[v for v in row if v >= 0] for row in
            g["bipar_{}".format(metric)].tolist()]
    assert cnt == want


def test_golden_dirty_matrix(golden):
    g = golden("hclust")
    z, _ = horc.hclust_linkage(g["dirty_d"].copy(), linkage="complete")
    assert_same_linkage(z, g["z_dirty_complete"])


def test_golden_sort_features(golden):
    g = golden("hclust")
    order = horc.sort_x_by_d(g["sf_x"].T, metric="euclidean",
                             optimal_ordering=True)
    assert order == g["sf_fids"].tolist()


def test_golden_ind_x(golden):
    g = golden("hclust")
    sel = g["indx_sel"]
    d = orc.pdist(g["x"], "cosine")
    np.testing.assert_allclose(d[sel][:, sel], g["indx_d"], rtol=1e-10,
                               atol=1e-14)


# ---- host logic of the product that runs without a device -------------------
def test_product_bipartition_counts_and_mdl(golden):
    from scedar_b200 import hclust as ph
    g = golden("hclust")
    for metric in METRICS:
        z = g["z_{}_average".format(metric)]
        got = ph.bipartition_counts(z, 5)
        want = [[v for v in row if v >= 0] for row in
                g["bipar_{}".format(metric)].tolist()]
        assert got == want
        tree = ph.HClustTree.hct_from_lkg(z)
        assert tree.n_round_bipar_cnt(5) == want
        assert tree.prev is None and tree.left().prev is tree
        assert tree.count() == 150
        assert tree.left_count() + tree.right_count() == 150
        assert tree.leaf_ids() == horc.leaf_ids(z)
        assert tree.left_leaf_ids() + tree.right_leaf_ids() == tree.leaf_ids()
    leaf = tree
    while leaf.count() > 1:
        leaf = leaf.left()
    assert leaf.left().count() == 0 and leaf.left().leaf_ids() == []
    assert leaf.left().left_leaf_ids() == [] and leaf.right().count() == 0
    for c in ([], [3], [3, 3, 3], [1, 2, 2, 0, 0, 7]):
        assert ph._multinomial_mdl(np.array(c)) == pytest.approx(
            horc.multinomial_mdl(c), rel=0, abs=0)


def test_product_condensed_offsets():
    from scedar_b200 import device as dev
    for n in (0, 1, 2, 5, 129, 1000):
        assert dev.condensed_len(n) == len(spdist.squareform(np.zeros((n, n))))
        if n > 1:
            assert dev.condensed_base(0, n) == 0
            assert dev.condensed_base(n, n) == dev.condensed_len(n)
            for i in range(min(n, 7)):
                assert dev.condensed_base(i + 1, n) - dev.condensed_base(i, n) \
                    == n - 1 - i
