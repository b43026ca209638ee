"""Re-point an installed scedar at the B200 path.

``install()`` patches the compute hooks of the reference classes in place:
``scedar.eda.SampleDistanceMatrix`` keeps its constructor, ids, plotting,
t-SNE / UMAP / PCA, ``to_classified`` ... and gets ``_d``, the sorted-column
accessors and the kNN accessors from :class:`scedar_b200.SampleDistanceMatrix`
(scedar/eda/sdm.py:187-222, 751-947, 1053-1083, 1193-1305);
``scedar.eda.HClustTree.hclust_linkage`` (sdm.py:1629-1669) gets the device
clean-up + condensed form of :mod:`scedar_b200.hclust` and accepts the SDM
itself in place of its ``_d``;
``scedar.knn.RareSampleDetection`` / ``FeatureImputation`` get the device
loops of :mod:`scedar_b200.knn` (scedar/knn/detection.py:33-123,
scedar/knn/imputation.py:35-135).
"""
from . import knn as _knn
from . import sdm as _sdm

# compute hooks and the helpers they call (all defined on the drop-in class)
SDM_HOOKS = (
    "_d", "_col_sorted_d", "_col_argsorted_d", "_s_knns_skl", "s_knns",
    "s_knn_ind_lut", "s_knn_connectivity_matrix", "s_ith_nn_d", "s_ith_nn_ind",
    "_validate_pdist", "_check_metric", "_cells", "_device_d", "_d_stripes",
    "_col_sort", "_knn_from_pdist", "release_device", "_is_sparse",
    "d_block", "_condensed_d",
)
SDM_STATIC_HOOKS = (
    "cosine_pdist", "correlation_pdist", "num_correct_dist_mat", "_static_pdist",
    "knn_conn_mat_to_aff_edges", "knn_conn_mat_to_aff_graph", "_warn_ncdm",
)


def install_sdm(cls):
    """Patch a ``SampleDistanceMatrix``-like class (the reference's) in place
    and return it."""
    fast = _sdm.SampleDistanceMatrix
    for name in SDM_HOOKS:
        setattr(cls, name, fast.__dict__[name])
    for name in SDM_STATIC_HOOKS:
        setattr(cls, name, staticmethod(getattr(fast, name)))
    if not hasattr(cls, "precision"):
        cls.precision = fast.precision
    orig_init = cls.__init__
    if getattr(orig_init, "_scedar_b200", False):
        return cls

    def __init__(self, *args, **kwargs):
        orig_init(self, *args, **kwargs)
        # HBM-side caches of the B200 path (not part of the reference's state)
        self._cells_lut = {}
        self._d_dev = None
        self._knn_dev = None
        self._d_user = self._lazy_load_d is not None

    __init__._scedar_b200 = True
    cls.__init__ = __init__
    return cls


def install_knn(rare_cls=None, impute_cls=None):
    """Patch ``scedar.knn.RareSampleDetection`` / ``FeatureImputation``."""
    if rare_cls is not None:
        for name in ("_pdist_rare_s_detect", "_no_pdist_rare_s_detect",
                     "detect_rare_samples"):
            setattr(rare_cls, name, _knn.RareSampleDetection.__dict__[name])
    if impute_cls is not None:
        impute_cls._impute_features_runner = staticmethod(
            _knn.FeatureImputation._impute_features_runner)
        impute_cls.impute_features = _knn.FeatureImputation.__dict__[
            "impute_features"]
    return rare_cls, impute_cls


def install_hclust(cls):
    """Patch an ``HClustTree``-like class: ``hclust_linkage`` only -- the
    reference's ``hclust_tree`` / ``sort_x_by_d`` / MIRAC call it by name."""
    from . import hclust as _hclust
    cls.hclust_linkage = staticmethod(_hclust.HClustTree.hclust_linkage)
    return cls


def install():
    """Patch the installed ``scedar`` package."""
    import scedar.eda.sdm as ref_sdm
    import scedar.knn as ref_knn
    install_sdm(ref_sdm.SampleDistanceMatrix)
    install_hclust(ref_sdm.HClustTree)
    install_knn(ref_knn.RareSampleDetection, ref_knn.FeatureImputation)
    return ref_sdm.SampleDistanceMatrix
