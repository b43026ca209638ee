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
    "d_block", "_condensed_d", "_ref_hook", "_knn_by_full_sort", "_stripes_of",
    "_gram_precision", "_compute_device_d", "_knn_of_stripe",
)
SDM_STATIC_HOOKS = (
    "cosine_pdist", "correlation_pdist", "num_correct_dist_mat", "_static_pdist",
    "knn_conn_mat_to_aff_edges", "knn_conn_mat_to_aff_graph", "_warn_ncdm",
    "_on_b200_path",
)
# The reference's own versions of these stay reachable as ``_ref_<name>``: the
# patched hooks hand over to them for everything the device path does not own
# -- any metric other than cosine / correlation / euclidean (sklearn's
# ``pairwise_distances`` / ``NearestNeighbors``, sdm.py:1215-1217, 1064-1079)
# and approximate HNSW search (sdm.py:949-1051).  Installing the patch therefore
# removes nothing from a working scedar: ``Community`` / ``CommunityMIRAC`` with
# their defaults (use_pca=True, use_hnsw=True) run as before.
SDM_KEEP_REF = ("_d", "s_knns", "_s_knns_skl")


def install_sdm(cls):
    """Patch a ``SampleDistanceMatrix``-like class (the reference's) in place
    and return it."""
    fast = _sdm.SampleDistanceMatrix
    orig_init = cls.__init__
    if getattr(orig_init, "_scedar_b200", False):
        return cls
    for name in SDM_KEEP_REF:
        if name in cls.__dict__ or hasattr(cls, name):
            setattr(cls, "_ref_" + name, getattr(cls, name))
    for name in SDM_HOOKS:
        setattr(cls, name, fast.__dict__[name])
    for name in SDM_STATIC_HOOKS:
        setattr(cls, name, staticmethod(getattr(fast, name)))
    if not hasattr(cls, "precision"):
        cls.precision = fast.precision
    cls.I8_MIN_CELLS = fast.I8_MIN_CELLS
    cls.I8_MIN_WORK = fast.I8_MIN_WORK
    cls.I8_MAX_FEATURES = fast.I8_MAX_FEATURES
    if not hasattr(cls, "_pca_x"):
        cls._pca_x = fast.__dict__["_pca_x"]

    def __init__(self, *args, **kwargs):
        orig_init(self, *args, **kwargs)
        # HBM-side caches of the B200 path (not part of the reference's state)
        self._cells_lut = {}
        self._d_dev = None
        self._bmin_dev = None
        self._knn_dev = None
        self._d_user = self._lazy_load_d is not None

    __init__._scedar_b200 = True
    cls.__init__ = __init__
    return cls


def _serial(method):
    """The reference's wrapper with ``nprocs`` pinned to 1: ``utils.parmap``
    then maps serially in this process (scedar/utils.py:41-43) instead of
    forking children that would inherit a CUDA context."""
    def wrapper(self, *args, **kwargs):
        import inspect
        names = list(inspect.signature(method).parameters)
        pos = names.index("nprocs") - 1 if "nprocs" in names else None
        args = list(args)
        if pos is not None and pos < len(args):
            if int(args[pos]) < 1:
                raise ValueError("nprocs should be >= 1. nprocs: {}".format(
                    args[pos]))
            args[pos] = 1
        elif "nprocs" in kwargs:
            if int(kwargs["nprocs"]) < 1:
                raise ValueError("nprocs should be >= 1. nprocs: {}".format(
                    kwargs["nprocs"]))
            kwargs["nprocs"] = 1
        return method(self, *args, **kwargs)
    wrapper.__doc__ = method.__doc__
    wrapper._scedar_b200 = True
    return wrapper


def install_knn(rare_cls=None, impute_cls=None):
    """Patch ``scedar.knn.RareSampleDetection`` / ``FeatureImputation``: only
    the per-parameter-tuple runners are replaced (the device loops); the
    reference's own ``detect_rare_samples`` / ``impute_features`` keep their
    validation, caching and result wrapping and are merely made serial."""
    if rare_cls is not None and not getattr(
            rare_cls.detect_rare_samples, "_scedar_b200", False):
        for name in ("_pdist_rare_s_detect", "_no_pdist_rare_s_detect"):
            setattr(rare_cls, "_ref_" + name, getattr(rare_cls, name))
            setattr(rare_cls, name, _knn.RareSampleDetection.__dict__[name])
        rare_cls.detect_rare_samples = _serial(rare_cls.detect_rare_samples)
    if impute_cls is not None and not getattr(
            impute_cls.impute_features, "_scedar_b200", False):
        impute_cls._ref__impute_features_runner = staticmethod(
            impute_cls.__dict__["_impute_features_runner"].__func__
            if isinstance(impute_cls.__dict__.get("_impute_features_runner"),
                          staticmethod)
            else impute_cls._impute_features_runner)
        ref_runner = impute_cls._ref__impute_features_runner

        def runner(pb_x, knn_ordered_ind_dict, k, n_do, min_present_val,
                   n_iter, statistic_fun=_knn.np.median, verbose=False):
            if statistic_fun is not _knn.np.median:
                # only the median has a device kernel: any other statistic is
                # the reference's own loop
                return ref_runner(pb_x, knn_ordered_ind_dict, k, n_do,
                                  min_present_val, n_iter,
                                  statistic_fun=statistic_fun, verbose=verbose)
            return _knn.FeatureImputation._impute_features_runner(
                pb_x, knn_ordered_ind_dict, k, n_do, min_present_val, n_iter,
                statistic_fun=statistic_fun, verbose=verbose)

        impute_cls._impute_features_runner = staticmethod(runner)
        impute_cls.impute_features = _serial(impute_cls.impute_features)
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
