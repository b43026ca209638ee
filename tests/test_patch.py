"""scedar_b200.patch: re-pointing a reference-like class at the B200 path."""
import numpy as np
import pytest


class RefLike(object):
    """The slots the reference's SampleDistanceMatrix.__init__ fills
    (scedar/eda/sfm.py:43-75, scedar/eda/sdm.py:93-146), nothing else."""

    def __init__(self, x, d=None, metric="cosine", use_pdist=True, nprocs=None):
        self._x = np.asarray(x, dtype="float64")
        self._sids = np.arange(self._x.shape[0])
        self._fids = np.arange(self._x.shape[1])
        self._nprocs = 1
        self._lazy_load_d = d
        self._metric = metric
        self._use_pdist = use_pdist
        self._lazy_load_col_sorted_d = None
        self._lazy_load_col_argsorted_d = None
        self._last_k = None
        self._last_knns = None

    def untouched(self):
        return "reference method"


def test_install_wiring_cpu():
    from scedar_b200 import patch
    cls = type("RefCopy", (RefLike,), {})
    patch.install_sdm(cls)
    tree = patch.install_hclust(type("T", (object,), {}))
    with pytest.raises(ValueError):
        tree.hclust_linkage(np.arange(5))        # rejected before any device use
    patch.install_sdm(cls)                       # idempotent
    s = cls(np.zeros((0, 3)), metric="euclidean")
    assert s._cells_lut == {} and s._d_dev is None and s._d_user is False
    assert s._d.shape == (0, 0)                  # empty input needs no device
    assert s.untouched() == "reference method"
    with pytest.raises(ValueError):
        cls(np.ones((4, 3)), use_pdist=False)._d
    with pytest.raises(ValueError):
        cls(np.ones((4, 3))).s_knn_ind_lut(4)


def test_install_on_live_reference_cpu():
    from oracle.ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("live reference only exists in the build container")
    import warnings
    warnings.simplefilter("ignore")
    ref = load_reference()
    cls = type("RefSub", (ref,), {})             # leave the shared class alone
    from scedar_b200 import patch
    patch.install_sdm(cls)
    s = cls(np.zeros((0, 4)), metric="euclidean")
    assert s._d.shape == (0, 0) and s.sids == []
    with pytest.raises(ValueError):
        cls(np.arange(12.0).reshape(4, 3), use_pdist=False)._d
    # HClustTree: the reference's own hclust_tree reaches the patched
    # hclust_linkage by name (sdm.py:1671-1678) and keeps its error behaviour
    # (tests/test_eda/test_dense_sdm.py:1306-1311) -- no device involved
    import scedar.eda.sdm as ref_sdm_mod
    tree_cls = ref_sdm_mod.HClustTree
    orig = tree_cls.__dict__["hclust_linkage"]
    try:
        patch.install_hclust(tree_cls)
        with pytest.raises(ValueError):
            tree_cls.hclust_tree(np.arange(5))
        with pytest.raises(ValueError):
            tree_cls.hclust_tree(np.arange(10).reshape(2, 5))
    finally:
        tree_cls.hclust_linkage = orig


@pytest.mark.gpu
def test_patched_class_matches_dropin():
    import scedar_b200
    from scedar_b200 import patch, synth
    cls = patch.install_sdm(type("RefCopy", (RefLike,), {}))
    x = synth.log_counts(400, 120, seed=8)
    for metric in ("correlation", "euclidean"):
        a = cls(x, metric=metric)
        b = scedar_b200.SampleDistanceMatrix(x, metric=metric)
        assert np.array_equal(a._d, b._d)
        assert a.s_knn_ind_lut(15) == b.s_knn_ind_lut(15)
        ta, da = a.s_knns(15)
        tb, db = b.s_knns(15)
        assert np.array_equal(np.array(ta), np.array(tb))
        assert np.array_equal(a._col_argsorted_d, b._col_argsorted_d)
        ca, cb = a.s_knn_connectivity_matrix(5), b.s_knn_connectivity_matrix(5)
        assert (ca != cb).nnz == 0
        assert np.array_equal(a._condensed_d(), b._condensed_d())
        sel = [7, 399, 0]
        assert np.array_equal(a.d_block(sel, sel), b._d[sel][:, sel])


@pytest.mark.gpu
def test_patched_hclust_matches_dropin():
    import scipy.cluster.hierarchy as sch
    import scedar_b200
    from scedar_b200 import patch, synth

    class RefTree(object):
        """The call structure of the reference (sdm.py:1671-1682)."""
        @staticmethod
        def hclust_linkage(dmat, linkage="complete", **kw):
            raise AssertionError("not patched")

        @staticmethod
        def hclust_tree(dmat, linkage="complete", **kw):
            return sch.to_tree(RefTree.hclust_linkage(dmat, linkage=linkage,
                                                      **kw))

    patch.install_hclust(RefTree)
    sdm_cls = patch.install_sdm(type("RefCopy", (RefLike,), {}))
    x = synth.blobs(300, 20, seed=2, n_blobs=3)
    a = sdm_cls(x, metric="euclidean")
    want = scedar_b200.HClustTree.hclust_linkage(
        scedar_b200.SampleDistanceMatrix(x, metric="euclidean")._d,
        linkage="ward")
    # the patched SDM itself (D stays in HBM) and its host matrix
    got = RefTree.hclust_tree(a, linkage="ward")
    assert got.pre_order(lambda nd: nd.get_id()) == \
        sch.to_tree(want).pre_order(lambda nd: nd.get_id())
    assert a._lazy_load_d is None
    assert np.array_equal(RefTree.hclust_linkage(a._d, linkage="ward"), want)


def test_patched_live_reference_on_fake_device(monkeypatch):
    """The reference's OWN class with the B200 hooks patched in (device
    operators replaced by the CPU stand-ins of tests/fake_device.py) against
    the unpatched reference: the wiring of every hook, on the real class, in
    the build container only."""
    from oracle.ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("live reference only exists in the build container")
    import warnings
    warnings.simplefilter("ignore")
    import fake_device
    fake_device.install(monkeypatch)
    ref = load_reference()
    import scedar.eda.sdm as ref_sdm_mod
    from scedar_b200 import patch, synth
    cls = patch.install_sdm(type("RefSub", (ref,), {}))
    tree = patch.install_hclust(type("TreeSub", (ref_sdm_mod.HClustTree,), {}))
    x = synth.blobs(90, 14, seed=6, n_blobs=3)
    for metric in ("cosine", "correlation", "euclidean"):
        a, b = cls(x, metric=metric), ref(x, metric=metric)
        np.testing.assert_allclose(a._d, b._d, rtol=1e-10, atol=1e-13)
        assert a._d is a._lazy_load_d
        assert a.s_knn_ind_lut(7) == b.s_knn_ind_lut(7)
        ta, da = a.s_knns(7)
        tb, db = b.s_knns(7)
        assert np.array_equal(np.array(ta), np.array(tb))
        np.testing.assert_allclose(np.array(da), np.array(db), rtol=1e-10)
        assert np.array_equal(a._col_argsorted_d, b._col_argsorted_d)
        np.testing.assert_allclose(a.s_ith_nn_d(3), b.s_ith_nn_d(3),
                                   rtol=1e-10)
        ca, cb = a.s_knn_connectivity_matrix(4), b.s_knn_connectivity_matrix(4)
        assert np.array_equal(ca.indices, cb.indices)
        np.testing.assert_allclose(ca.data, cb.data, rtol=1e-10)
        # reference methods that were not touched keep working on the patched
        # class: ind_x slices the D the hooks produced
        sub = a.ind_x([5, 1, 60])
        np.testing.assert_allclose(sub._d, b.ind_x([5, 1, 60])._d,
                                   rtol=1e-10, atol=1e-13)
        za = tree.hclust_linkage(a, linkage="average")
        zb = ref_sdm_mod.HClustTree.hclust_linkage(b._d, linkage="average")
        assert np.array_equal(za[:, [0, 1, 3]], zb[:, [0, 1, 3]])
        np.testing.assert_allclose(za[:, 2], zb[:, 2], rtol=1e-9)
    nd_a = cls(x, metric="euclidean", use_pdist=False)
    nd_b = ref(x, metric="euclidean", use_pdist=False)
    assert nd_a.s_knn_ind_lut(5) == nd_b.s_knn_ind_lut(5)
    with pytest.raises(ValueError):
        nd_a._d


def test_patch_keeps_what_the_device_path_does_not_own(monkeypatch):
    """Installing the patch must not remove anything from a working scedar:
    metrics outside cosine / correlation / euclidean keep the reference's own
    sklearn route (sdm.py:1215-1217, 1064-1079), the sorted-column accessors
    then work on the matrix it leaves behind, ``use_pca=True`` searches the
    reference's own PCA coordinates, and ``use_hnsw=True`` reaches the
    reference's HNSW code (cluster.Community's defaults,
    scedar/cluster/community.py:117)."""
    from oracle.ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("live reference only exists in the build container")
    import warnings
    warnings.simplefilter("ignore")
    import fake_device
    fake_device.install(monkeypatch)
    ref = load_reference()
    import scedar.knn as ref_knn
    from scedar_b200 import patch, synth
    cls = patch.install_sdm(type("RefSub2", (ref,), {}))
    x = synth.blobs(60, 9, seed=11, n_blobs=3)
    for metric in ("cityblock", "chebyshev"):
        a, b = cls(x, metric=metric), ref(x, metric=metric)
        assert np.array_equal(a._d, b._d)
        assert a.s_knn_ind_lut(4) == b.s_knn_ind_lut(4)
        assert np.array_equal(a._col_argsorted_d, b._col_argsorted_d)
        ta, _ = a.s_knns(4)
        tb, _ = b.s_knns(4)
        assert np.array_equal(np.array(ta), np.array(tb))
        na = cls(x, metric=metric, use_pdist=False)
        nb = ref(x, metric=metric, use_pdist=False)
        assert na.s_knn_ind_lut(4) == nb.s_knn_ind_lut(4)
    # PCA-space search: the reference's PCA, the path's kNN
    a, b = cls(x, metric="euclidean"), ref(x, metric="euclidean")
    ta, da = a.s_knns(5, use_pca=True)
    tb, db = b.s_knns(5, use_pca=True)
    assert np.array_equal(np.array(ta), np.array(tb))
    np.testing.assert_allclose(np.array(da), np.array(db), rtol=1e-9)
    # HNSW goes to the reference's own method (nmslib is a stub in this image:
    # reaching it is what is checked)
    seen = {}

    def fake_hnsw(self, k, **kw):
        seen["k"] = k
        return [np.zeros(k, dtype=int)] * 60, [np.zeros(k)] * 60

    monkeypatch.setattr(cls, "_s_knns_hnsw", fake_hnsw, raising=False)
    a.s_knns(3, use_hnsw=True)
    assert seen == {"k": 3}
    # the kNN consumers: the reference's own wrappers, serial, device runners
    rare_cls = type("RareSub", (ref_knn.RareSampleDetection,), {})
    imp_cls = type("ImpSub", (ref_knn.FeatureImputation,), {})
    patch.install_knn(rare_cls, imp_cls)
    a, b = cls(x, metric="euclidean"), ref(x, metric="euclidean")
    ra = rare_cls(a).detect_rare_samples(3, 6.0, 2, nprocs=4)
    rb = ref_knn.RareSampleDetection(b).detect_rare_samples(3, 6.0, 2)
    assert ra == rb
    xi = np.where(x > np.median(x), x, 0.0)
    ia = imp_cls(cls(xi, metric="euclidean")).impute_features(4, 2, 1, 1,
                                                              nprocs=3)
    ib = ref_knn.FeatureImputation(ref(xi, metric="euclidean")).impute_features(
        4, 2, 1, 1)
    assert isinstance(ia[0], ref)          # the reference's class comes back
    assert np.array_equal(ia[0]._x, ib[0]._x)
    # a statistic other than the median is the reference's own loop
    ia = imp_cls(cls(xi, metric="euclidean")).impute_features(
        4, 2, 1, 1, statistic_fun=np.mean)
    ib = ref_knn.FeatureImputation(ref(xi, metric="euclidean")).impute_features(
        4, 2, 1, 1, statistic_fun=np.mean)
    np.testing.assert_allclose(ia[0]._x, ib[0]._x)
