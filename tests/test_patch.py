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
