"""Drop-in ``SampleDistanceMatrix`` for the distance-matrix / kNN hot path.

Mirrors the reference interface for this path (scedar/eda/sdm.py:42-146,
187-222, 729-947, 1053-1083, 1189-1305 and the input contract of
scedar/eda/sfm.py:43-75): same constructor, same lazy ``_d`` / ``d`` property,
same kNN accessors, same cache slot names, same error conventions.  The
arithmetic the reference delegates to NumPy / SciPy / scikit-learn runs in the
CUDA library behind ``include/scedar_b200.h``; there is no CPU fallback, and
metrics outside cosine / correlation / euclidean raise.

Plotting, t-SNE / UMAP / PCA, HNSW and the classified-sample subclasses are
not part of this path and are not provided here; see
INTEGRATION.md for patching these hooks into an installed scedar instead.
"""
import os
import warnings
from collections import defaultdict

import numpy as np
import scipy.sparse as spsparse
import torch

from . import device as dev
from ._capi import (EXCLUDE_FIRST, EXCLUDE_SELF, METRIC_IDS, NCDM_ASYMMETRIC,
                    NCDM_BAD_DIAG)

# D up to this many bytes is also kept resident in HBM between accessors
DEVICE_D_CACHE_BYTES = 48 << 30
# stripe size used when D is streamed to the host in pieces
STRIPE_BYTES = 2 << 30
# a dense host matrix of at least this size AND this many cells is uploaded range
# by range behind the contraction instead of in one blocking copy (with fewer
# cells the contraction is too short to hide anything: config 1, 3,005 x 19,972,
# contracts in 2 ms and uploads in 70)
UPLOAD_OVERLAP_BYTES = 64 << 20
UPLOAD_OVERLAP_MIN_CELLS = 20000
MAX_TOPK = 127
FAST_MAX_K = 55     # the tensor-core pass re-ranks up to 64 candidates


def _check_is_valid_sfids(sfids):
    """scedar/eda/mtype.py:52-71."""
    if sfids is None:
        raise ValueError("[sf]ids cannot be None")
    if type(sfids) != list:
        raise ValueError("[sf]ids must be a homogenous list of int or str")
    kinds = set(map(type, sfids))
    if len(kinds) > 1 or (len(kinds) == 1 and
                          type(sfids[0]) not in (str, int)):
        raise ValueError("[sf]ids must be a homogenous list of int or str")
    arr = np.array(sfids)
    if arr.ndim != 1 or np.unique(arr).shape[0] != arr.shape[0]:
        raise ValueError("[sf]ids must not contain duplicated values")


class SampleDistanceMatrix(object):
    """Cells x genes matrix with a lazily computed cell x cell distance matrix
    and k-nearest-neighbour accessors, computed on one B200.

    Parameters are the reference's (sdm.py:93-94).  ``nprocs`` is accepted and
    stored for compatibility; the device is the current CUDA device.

    ``precision`` (class or instance attribute, default from the environment
    variable ``SCEDAR_B200_PRECISION``, else ``"fp64"``) selects the Gram
    kernel: ``"fp64"`` matches the reference's float64 to 1e-10 -- with the
    DMMA kernel, or, from 1,024 cells and n^2 p = 2e10 up, with the
    INT8 split on the tcgen05 tensor cores (``"fp64_i8"``, 2.7 x faster, same
    bar; ``SCEDAR_B200_FP64_KERNEL=dmma|i8`` forces one; both can also be
    named directly as ``"fp64_dmma"`` / ``"fp64_i8"`` / ``"fp64_i8s6"``);
    ``"fast"`` runs the contraction on the tcgen05 tensor cores -- the
    distance matrix is then within 1e-5 (cosine, correlation) or, for
    euclidean, within 2e-6 * max|x|^2 on the squared distance (1e-4 * max(D)
    unless two cells nearly coincide), while kNN results stay exact (float64
    re-rank of the tensor-core candidates).
    """
    precision = os.environ.get("SCEDAR_B200_PRECISION", "fp64")
    # "fp64" picks the INT8 split kernel where it pays: enough cells to fill the
    # SMs with 128 x 64 tiles and enough work (n^2 p) to amortise the slicing
    I8_MIN_CELLS = 1024
    I8_MIN_WORK = 2e10
    I8_MAX_FEATURES = 1 << 16

    def _gram_precision(self, n=None, p=None):
        """The kernel name handed to the device layer for ``self.precision``."""
        prec = self.precision
        if prec == "fp64_dmma":
            return "fp64"
        if prec != "fp64":
            return prec
        forced = os.environ.get("SCEDAR_B200_FP64_KERNEL", "auto")
        if forced == "dmma":
            return "fp64"
        n = self._x.shape[0] if n is None else n
        p = self._x.shape[1] if p is None else p
        fits = 0 < p <= self.I8_MAX_FEATURES
        if forced == "i8":
            return "fp64_i8" if fits else "fp64"
        big = n >= self.I8_MIN_CELLS and float(n) * n * p >= self.I8_MIN_WORK
        return "fp64_i8" if (fits and big) else "fp64"

    def __init__(self, x, d=None, metric="cosine", use_pdist=True,
                 sids=None, fids=None, nprocs=None):
        super(SampleDistanceMatrix, self).__init__()
        # -- input contract, sfm.py:43-75 (np.asarray instead of
        #    np.array(copy=False), which NumPy 2 rejects for lists) ----------
        if x is None:
            raise ValueError("x cannot be None")
        if spsparse.issparse(x):
            x = spsparse.csr_matrix(x, dtype="float64")
        else:
            try:
                x = np.asarray(x, dtype="float64")
            except ValueError as e:
                raise ValueError("Features must be float. {}".format(e))
        if x.ndim != 2:
            raise ValueError("x has shape (n_samples, n_features)")
        if sids is None:
            sids = list(range(x.shape[0]))
        else:
            _check_is_valid_sfids(sids)
            if len(sids) != x.shape[0]:
                raise ValueError("x has shape (n_samples, n_features)")
        if fids is None:
            fids = list(range(x.shape[1]))
        else:
            _check_is_valid_sfids(fids)
            if len(fids) != x.shape[1]:
                raise ValueError("x has shape (n_samples, n_features)")
        self._x = x
        self._sids = np.array(sids)
        self._fids = np.array(fids)

        # -- sdm.py:97-146 ----------------------------------------------------
        if d is not None:
            if not use_pdist:
                raise ValueError("pdist cannot be provided when "
                                 "use pdist = False")
            try:
                d = np.array(d, dtype="float64")
            except ValueError as e:
                raise ValueError("d must be float. {}".format(e))
            if (d.ndim != 2 or d.shape[0] != d.shape[1]
                    or d.shape[0] != self._x.shape[0]):
                raise ValueError("d should have shape (n_samples, n_samples)")
            d = self.num_correct_dist_mat(d)
        else:
            if metric == "precomputed":
                raise ValueError("metric cannot be precomputed when "
                                 "d is None.")
        self._nprocs = 1 if nprocs is None else max(int(nprocs), 1)
        self._lazy_load_d = d
        self._metric = metric
        self._use_pdist = use_pdist
        self._lazy_load_col_sorted_d = None
        self._lazy_load_col_argsorted_d = None
        self._knn_ng_lut = {}
        self._last_k = None
        self._last_knns = None
        # device-side caches (not part of the reference's state)
        self._cells_lut = {}
        self._d_dev = None
        self._bmin_dev = None   # block minima of _d_dev (INT8 split kernel only)
        self._d_user = d is not None
        self._knn_dev = None    # (k, exclude, metric, idx, dist) last lists in HBM

    # ------------------------------------------------------------------ ids
    @property
    def sids(self):
        return self._sids.tolist()

    @property
    def fids(self):
        return self._fids.tolist()

    @property
    def x(self):
        return self._x.copy() if self._is_sparse else self._x.tolist()

    @property
    def _is_sparse(self):
        return spsparse.issparse(self._x)

    def s_id_to_ind(self, selected_sids):
        sid_list = self.sids
        return [sid_list.index(i) for i in selected_sids]

    def f_id_to_ind(self, selected_fids):
        fid_list = self.fids
        return [fid_list.index(i) for i in selected_fids]

    def ind_x(self, selected_s_inds=None, selected_f_inds=None):
        """sdm.py:662-699 -- subsetting slices the cached D, no recompute."""
        if selected_s_inds is None:
            selected_s_inds = slice(None, None)
        if selected_f_inds is None:
            selected_f_inds = slice(None, None)
        for sel in (selected_s_inds, selected_f_inds):
            # NumPy's rule for a dense x (IndexError); scipy.sparse answers a
            # non-integer index with whatever its version happens to raise
            if not isinstance(sel, slice) and \
                    np.asarray(sel).dtype.kind not in "iub":
                raise IndexError("only integers, slices and integer or "
                                 "boolean arrays are valid indices")
        sub_d = None
        if self._use_pdist:
            sub_d = self.d_block(selected_s_inds, selected_s_inds)
        return SampleDistanceMatrix(
            x=self._x[selected_s_inds, :][:, selected_f_inds].copy(),
            d=sub_d, metric=self._metric, use_pdist=self._use_pdist,
            sids=self._sids[selected_s_inds].tolist(),
            fids=self._fids[selected_f_inds].tolist(), nprocs=self._nprocs)

    def d_block(self, row_inds, col_inds):
        """``self._d[row_inds][:, col_inds]`` as a fresh array -- the slicing
        of ``ind_x`` (sdm.py:684-692) and of MIRAC's cluster encoders
        (scedar/cluster/mirac.py:316-331).  A D that so far lives only in HBM
        is sliced there (``scedar_d_gather_block``) and only the block comes
        to the host; a D already on the host is sliced where it is."""
        self._validate_pdist()
        n = self._x.shape[0]
        if self._lazy_load_d is None and self._device_d() is not None:
            pos = np.arange(n, dtype=np.int64)
            rows = np.atleast_1d(pos[row_inds])     # NumPy's own index rules
            cols = np.atleast_1d(pos[col_inds])
            blk = dev.gather_block(self._d_dev, dev.to_device(rows),
                                   dev.to_device(cols))
            return blk.cpu().numpy()
        return self._d[row_inds][:, col_inds].copy()

    def _condensed_d(self):
        """``scipy.spatial.distance.squareform(self._d)`` (the input of
        ``sch.linkage``, sdm.py:1666) without the n x n host matrix: the
        strict upper triangle is packed on the device, stripe by stripe when D
        is not resident, and only n(n-1)/2 doubles cross to the host."""
        from .hclust import _condensed_flags_check
        self._validate_pdist()
        n = self._x.shape[0]
        out = np.empty(dev.condensed_len(n), dtype=np.float64)
        if n == 0:
            return out
        flags = dev.zeros(1, torch.int32)
        for r0, r1, stripe in self._d_stripes():
            seg, _ = dev.squareform_rows(stripe, n, row_begin=r0, flags=flags)
            b0, b1 = dev.condensed_base(r0, n), dev.condensed_base(r1, n)
            if b1 > b0:
                dev.to_host(seg, out[b0:b1])
        _condensed_flags_check(int(flags.item()))
        return out

    def sort_features(self, fdist_metric="cosine", optimal_ordering=False):
        """sdm.py:178-184: reorder the features by the leaf order of the
        auto-linkage tree of their pairwise distances."""
        from .hclust import HClustTree
        order = HClustTree.sort_x_by_d(
            self._x.T, metric=fdist_metric, optimal_ordering=optimal_ordering,
            nprocs=self._nprocs)
        self._x = self._x[:, order]
        self._fids = self._fids[order]
        for c in self._cells_lut.values():   # prepared from the old order
            c.close()
        self._cells_lut = {}

    def id_x(self, selected_sids=None, selected_fids=None):
        s = None if selected_sids is None else self.s_id_to_ind(selected_sids)
        f = None if selected_fids is None else self.f_id_to_ind(selected_fids)
        return self.ind_x(s, f)

    # ------------------------------------------------------- device helpers
    @staticmethod
    def _on_b200_path(metric):
        """cosine / correlation / euclidean run on the device; every other
        metric is sklearn's (sdm.py:1215-1217) and is none of this path's
        business."""
        return isinstance(metric, str) and metric in METRIC_IDS

    def _ref_hook(self, name):
        """The reference's own implementation of a patched hook, kept by
        ``patch.install_sdm`` as ``_ref_<name>`` (None on the stand-alone
        drop-in class, which has nothing to fall back to)."""
        return getattr(type(self), "_ref_" + name, None)

    def _check_metric(self, metric):
        if not isinstance(metric, str) or metric not in METRIC_IDS:
            raise ValueError(
                "metric {!r} is not on the B200 path; supported: {}".format(
                    metric, sorted(METRIC_IDS)))
        if self._is_sparse and metric == "correlation":
            raise ValueError("Only the following metrics are supportted for "
                             "sparse matrices, ['cosine', 'euclidean']")

    def _cells(self, metric):
        """Prepared cells in HBM for ``metric`` (cached per metric)."""
        self._check_metric(metric)
        if metric not in self._cells_lut:
            if self._is_sparse:
                x = self._x
                if not x.has_canonical_format:
                    x = x.copy()
                    x.sum_duplicates()
                ip = x.indptr.astype(np.int64, copy=False)
                cells = dev.Cells.from_csr(
                    dev.to_device(ip),
                    dev.to_device(x.indices.astype(np.int32, copy=False)),
                    dev.to_device(x.data.astype(np.float64, copy=False)),
                    x.shape, metric)
            else:
                cells = dev.Cells.from_dense(dev.to_device(self._x), metric)
            self._cells_lut[metric] = cells
        return self._cells_lut[metric]

    def _device_d(self):
        """D resident in HBM, or None when it is too large to keep."""
        n = self._x.shape[0]
        limit = DEVICE_D_CACHE_BYTES
        if self._d_dev is None and n * n * 8 > limit:
            # a larger D is still kept (and computed symmetrically, half the
            # work) when it takes less than half of the free device memory
            free, _ = torch.cuda.mem_get_info()
            limit = max(limit, free // 2)
        if self._d_dev is None and n * n * 8 <= limit:
            self._bmin_dev = None
            if self._lazy_load_d is not None:
                self._d_dev = dev.to_device(self._lazy_load_d)
            else:
                self._d_dev = self._compute_device_d()
        return self._d_dev

    def _compute_device_d(self):
        """The whole D of ``self._x`` into HBM.  With the INT8 split kernel the
        block minima of D come with it (the selection then reads n/32 minima
        per row instead of the row); a large dense host matrix that is not on
        the device yet goes up range by range behind the contraction of the
        rows already there (``device.pdist_symmetric_from_host``)."""
        metric = self._metric
        self._check_metric(metric)
        prec = self._gram_precision()
        n, p = self._x.shape
        bmin = None
        if prec in ("fp64_i8", "fp64_i8s6") and n >= 4096:
            bmin = dev.new_blockmin(n, n)
        if (metric not in self._cells_lut and not self._is_sparse
                and prec != "fast" and n * p * 8 >= UPLOAD_OVERLAP_BYTES
                and n >= UPLOAD_OVERLAP_MIN_CELLS):
            xh = np.ascontiguousarray(self._x, dtype=np.float64)
            cells, d = dev.pdist_symmetric_from_host(xh, metric, precision=prec,
                                                     bmin=bmin)
            self._cells_lut[metric] = cells
        else:
            cells = self._cells(metric)
            if bmin is not None:
                cells.set_blockmin([bmin])
            d = cells.pdist_symmetric(precision=prec)
        self._bmin_dev = bmin
        return d

    def _knn_of_stripe(self, stripe, r0, k, exclude):
        """Top-k of the rows of a stripe yielded by ``_d_stripes``."""
        bmin = getattr(self, "_bmin_dev", None)
        if bmin is not None and stripe is self._d_dev:
            return self._cells_lut[self._metric].knn_from_d_blockmin(
                stripe, bmin, k, exclude, row_begin=r0)
        return dev.knn_from_d(stripe, k, exclude, row_begin=r0)

    def _d_stripes(self):
        """Yield (row_begin, row_end, device stripe of D) covering all rows."""
        n = self._x.shape[0]
        whole = self._device_d()
        if whole is not None:
            yield 0, n, whole
            return
        rows = max(128, (STRIPE_BYTES // (8 * n)) // 128 * 128)
        if self._lazy_load_d is not None:
            for r0 in range(0, n, rows):
                r1 = min(n, r0 + rows)
                yield r0, r1, dev.to_device(self._lazy_load_d[r0:r1])
            return
        cells = self._cells(self._metric)
        for r0 in range(0, n, rows):
            r1 = min(n, r0 + rows)
            yield r0, r1, cells.pdist_stripe(
                r0, r1, precision=self._gram_precision())

    # ----------------------------------------------------- distance matrix
    @staticmethod
    def num_correct_dist_mat(dmat, upper_bound=None):
        """sdm.py:187-222, in place on the caller's array (round trip through
        HBM); emits the reference's two ``UserWarning``s."""
        if (not isinstance(dmat, np.ndarray) or dmat.ndim != 2
                or dmat.shape[0] != dmat.shape[1]):
            raise ValueError("dmat must be a 2D (n_samples, n_samples)"
                             " np array")
        if dmat.shape[0] == 0:
            return dmat
        t = dev.to_device(dmat.astype(np.float64, copy=False))
        SampleDistanceMatrix._warn_ncdm(
            dev.num_correct_dist_mat_(t, upper_bound))
        dmat[...] = t.cpu().numpy()
        return dmat

    @staticmethod
    def _warn_ncdm(flags):
        """The two ``UserWarning``s of sdm.py:196-213 from the kernel's flag
        bits."""
        if flags & NCDM_BAD_DIAG:
            warnings.warn("distance matrix might not be numerically "
                          "correct. diag vals should be close to 0.")
        if flags & NCDM_ASYMMETRIC:
            warnings.warn("distance matrix might not be numerically "
                          "correct. should be approximately symmetric.")

    @staticmethod
    def _static_pdist(x, metric):
        if spsparse.issparse(x):
            return SampleDistanceMatrix(x, metric=metric)._d
        x = np.asarray(x, dtype="float64")
        if x.ndim != 2:
            raise ValueError("x has shape (n_samples, n_features)")
        return SampleDistanceMatrix(x, metric=metric)._d

    @staticmethod
    def cosine_pdist(x):
        """sdm.py:1223-1266.  The result already carries the clean-up of
        ``num_correct_dist_mat`` (it is fused into the kernel epilogue)."""
        return SampleDistanceMatrix._static_pdist(x, "cosine")

    @staticmethod
    def correlation_pdist(x):
        """sdm.py:1268-1287."""
        return SampleDistanceMatrix._static_pdist(x, "correlation")

    @property
    def d(self):
        return self._d.tolist()

    def _validate_pdist(self):
        """The checks the reference's ``_d`` makes before computing
        (sdm.py:1194-1209, 1219-1220)."""
        if not self._use_pdist:
            raise ValueError("This SDM instance does not use pdist.")
        if self._is_sparse and self._metric not in (
                "cityblock", "cosine", "euclidean", "l1", "l2", "manhattan"):
            raise ValueError("Only the following metrics are supportted for "
                             "sparse matrices, ['cityblock', 'cosine', "
                             "'euclidean', 'l1', 'l2', 'manhattan']")
        if self._lazy_load_d is None:
            n = self._x.shape[0]
            if self._x.shape[0] * self._x.shape[1] == 0:
                self._lazy_load_d = np.zeros((n, n))
            else:
                self._check_metric(self._metric)

    @property
    def _d(self):
        """sdm.py:1193-1221: lazy, memoised in ``_lazy_load_d``.  D may
        already be resident in HBM (an earlier kNN accessor); the host copy is
        made on first access."""
        if (self._lazy_load_d is None and self._use_pdist
                and not self._on_b200_path(self._metric)
                and self._ref_hook("_d") is not None):
            # a metric the device path does not own: the reference's own
            # (sklearn) route, untouched; the accessors below then work on
            # the matrix it leaves in _lazy_load_d
            return self._ref_hook("_d").fget(self)
        self._validate_pdist()
        if self._lazy_load_d is None:
            n = self._x.shape[0]
            out = np.empty((n, n), dtype=np.float64)
            for r0, r1, stripe in self._d_stripes():
                dev.to_host(stripe, out[r0:r1])
            self._lazy_load_d = out
            self._d_own = out
        return self._lazy_load_d

    # ------------------------------------------------------ sorted columns
    def _col_sort(self, want_sorted, want_arg):
        if self._use_pdist:
            # np.sort(self._d, ...) (sdm.py:1293, 1302): the reference fills
            # _lazy_load_d on the way, and to_classified hands that slot on
            self._d
        self._validate_pdist()
        n = self._x.shape[0]
        srt = np.empty((n, n), dtype=np.float64) if want_sorted else None
        arg = np.empty((n, n), dtype=np.int64) if want_arg else None
        if n:
            # the sort writes up to two more [rows, n] arrays: a resident D is
            # sorted in row blocks, not in one piece
            block = max(128, (STRIPE_BYTES // (8 * n)) // 128 * 128)
            for r0, r1, stripe in self._d_stripes():
                for b0 in range(r0, r1, block):
                    b1 = min(r1, b0 + block)
                    s, a = dev.col_sort(stripe[b0 - r0:b1 - r0], want_sorted,
                                        want_arg)
                    if want_sorted:
                        dev.to_host(s, srt[b0:b1])
                    if want_arg:
                        dev.to_host(a, arg[b0:b1])
        # cell-major -> the reference's (rank, cell) orientation (a view)
        return (srt.T if want_sorted else None, arg.T if want_arg else None)

    @property
    def _col_sorted_d(self):
        """sdm.py:1289-1296."""
        if self._lazy_load_col_sorted_d is None:
            self._lazy_load_col_sorted_d = self._col_sort(True, False)[0]
        return self._lazy_load_col_sorted_d

    @property
    def _col_argsorted_d(self):
        """sdm.py:1298-1305."""
        if self._lazy_load_col_argsorted_d is None:
            self._lazy_load_col_argsorted_d = self._col_sort(False, True)[1]
        return self._lazy_load_col_argsorted_d

    def s_ith_nn_d(self, i):
        return self._col_sorted_d[i, :]

    def s_ith_nn_ind(self, i):
        return self._col_argsorted_d[i, :]

    # ------------------------------------------------------------------ kNN
    def _knn_from_pdist(self, k, exclude):
        """Top-k of every column of D (== row, D is symmetric)."""
        self._validate_pdist()
        n = self._x.shape[0]
        idx = np.empty((n, k), dtype=np.int64)
        dist = np.empty((n, k), dtype=np.float64)
        # no D anywhere yet and too large to be worth keeping: kNN without D
        # (the library still routes through a scratch D when that is cheaper)
        fused = (self._lazy_load_d is None and self._d_dev is None
                 and k <= 31 and n * n * 8 > DEVICE_D_CACHE_BYTES)
        # a tensor-core D is approximate: exact neighbours come from the
        # candidate + float64 re-rank kernel (k <= 55) or the FP64 kernels
        # (larger k), never from that matrix.  Only a D this object computed
        # itself may be bypassed like this -- one that was handed in (the
        # constructor, to_classified, a restored cache) is the caller's truth.
        own_d = (self._lazy_load_d is None
                 or self._lazy_load_d is getattr(self, "_d_own", None))
        exact_fast = (self.precision == "fast" and own_d
                      and not getattr(self, "_d_user", False)
                      and self._on_b200_path(self._metric))
        if fused or exact_fast:
            # D is not needed (or would not fit): fused Gram -> top-k
            i, dd = self._cells(self._metric).knn_stripe(
                k, exclude, precision=(
                    self._gram_precision() if k <= FAST_MAX_K or not exact_fast
                    else "fp64"))
            self._knn_dev = (k, exclude, self._metric, i, dd)
            return i.cpu().numpy(), dd.cpu().numpy()
        parts = []
        for r0, r1, stripe in self._d_stripes():
            i, dd = self._knn_of_stripe(stripe, r0, k, exclude)
            parts.append((i, dd))
            idx[r0:r1] = i.cpu().numpy()
            dist[r0:r1] = dd.cpu().numpy()
        if parts:
            self._knn_dev = (k, exclude, self._metric,
                             torch.cat([q[0] for q in parts]),
                             torch.cat([q[1] for q in parts]))
        return idx, dist

    def s_knn_ind_lut(self, k, metric=None, use_pca=False, use_hnsw=False,
                      index_params=None, query_params=None, verbose=False):
        """sdm.py:751-790: ``{i: [1st_NN_ind, ..., kth_NN_ind]}``."""
        n = self._x.shape[0]
        if k < 0 or k > n - 1:
            raise ValueError("k ({}) should >= 0 and <= n_samples-1".format(k))
        k = int(k)
        if self._use_pdist:
            if k == 0:
                return dict(zip(range(n), [[] for _ in range(n)]))
            if self._lazy_load_col_argsorted_d is not None or k > MAX_TOPK:
                arr = self._col_argsorted_d[1:k + 1, :].T
            else:
                # ranks 1..k of the stable column argsort, without sorting
                # the other n-k-1 ranks
                arr = self._knn_from_pdist(k, EXCLUDE_FIRST)[0]
            return dict(zip(range(n), arr.tolist()))
        if self._last_k is None or self._last_k < k:
            compute_k = min(10, max(n - 1, 0)) if k < 10 else k
            targets, distances = self.s_knns(
                compute_k, metric=metric, use_pca=use_pca, use_hnsw=use_hnsw,
                index_params=index_params, query_params=query_params,
                verbose=verbose)
        else:
            targets, distances = self._last_knns
        lut = defaultdict(list)
        for i in range(len(targets)):
            lut[i]
            lut[i].extend(int(t) for t in targets[i][:k])
        return lut

    def s_knns(self, k, metric=None, use_pca=False, use_hnsw=False,
               index_params=None, query_params=None, verbose=False):
        """sdm.py:792-870: exact kNN, ``(list of index arrays, list of
        distance arrays)``, self excluded, ascending distance."""
        if k < 1:
            raise ValueError("k should >= 1")
        if use_hnsw:
            # approximate search (nmslib, sdm.py:949-1051) is the reference's
            # own: a patched class keeps it, the stand-alone class has none
            ref = self._ref_hook("s_knns")
            if ref is None:
                raise NotImplementedError(
                    "approximate HNSW search (nmslib) is outside the B200 "
                    "path; the exact brute-force path is use_hnsw=False")
            return ref(self, k, metric=metric, use_pca=use_pca,
                       use_hnsw=use_hnsw, index_params=index_params,
                       query_params=query_params, verbose=verbose)
        if index_params is not None:
            raise ValueError("index_params are not used with use_hnsw=False.")
        if query_params is not None:
            raise ValueError("query_params are not used with use_hnsw=False.")
        targets, distances = self._s_knns_skl(k, metric=metric,
                                              use_pca=use_pca, verbose=verbose)
        self._last_k = k
        self._last_knns = (targets, distances)
        return targets, distances

    def _knn_by_full_sort(self, stripes, n, k):
        """k + 1 <= n nearest of every cell by the full stable sort of its row
        of D (any k; the list kernels keep at most 128 entries), self dropped
        by sklearn's rule.  ``stripes`` yields (r0, r1, device stripe)."""
        idx = np.empty((n, k), dtype=np.int64)
        dist = np.empty((n, k), dtype=np.float64)
        for r0, r1, stripe in stripes:
            block = max(128, (STRIPE_BYTES // (8 * n)) // 128 * 128)
            for b0 in range(r0, r1, block):
                b1 = min(r1, b0 + block)
                srt, arg = dev.col_sort(stripe[b0 - r0:b1 - r0], True, True)
                a = arg[:, :k + 1].cpu().numpy()
                v = srt[:, :k + 1].cpu().numpy()
                keep = a != np.arange(b0, b1)[:, None]
                keep[keep.all(axis=1), 0] = False
                idx[b0:b1] = a[keep].reshape(b1 - b0, k)
                dist[b0:b1] = v[keep].reshape(b1 - b0, k)
        return idx, dist

    def _s_knns_skl(self, k, metric=None, use_pca=False, verbose=False):
        """sdm.py:1053-1083 (the slot sklearn's NearestNeighbors filled)."""
        if metric is None:
            metric = self._metric
        if not self._on_b200_path(metric) or (
                self._is_sparse and metric == "correlation"):
            ref = self._ref_hook("_s_knns_skl")
            if ref is not None:
                return ref(self, k, metric=metric, use_pca=use_pca,
                           verbose=verbose)
            self._check_metric(metric)          # raises
        n = self._x.shape[0]
        k = int(k)
        if k + 1 > n:
            raise ValueError("Expected n_neighbors < n_samples_fit, but "
                             "n_neighbors = {}, n_samples_fit = {}".format(
                                 k, n))
        if use_pca:
            # sdm.py:1060-1065: the same search over the PCA coordinates
            # (the reference's own _pca_x on a patched class)
            if verbose:
                print("Construct {} distance KNN graph on PCA.".format(metric))
            pcs = np.ascontiguousarray(self._pca_x, dtype=np.float64)
            cells = dev.Cells.from_dense(dev.to_device(pcs), metric)
            try:
                if k > MAX_TOPK:
                    idx, dist = self._knn_by_full_sort(
                        self._stripes_of(cells), n, k)
                else:
                    i, dd = cells.knn_stripe(
                        k, EXCLUDE_SELF,
                        precision=self._gram_precision(n, pcs.shape[1]))
                    idx, dist = i.cpu().numpy(), dd.cpu().numpy()
            finally:
                cells.close()
            return list(idx), list(dist)
        if verbose:
            print("Construct {} distance KNN graph on raw data.".format(
                metric))
        if self._use_pdist:
            if k > MAX_TOPK:
                arg = self._col_argsorted_d[:k + 1, :].T
                srt = self._col_sorted_d[:k + 1, :].T
                keep = arg != np.arange(n)[:, None]
                keep[keep.all(axis=1), 0] = False
                idx = arg[keep].reshape(n, k)
                dist = srt[keep].reshape(n, k)
            else:
                idx, dist = self._knn_from_pdist(k, EXCLUDE_SELF)
        elif k > MAX_TOPK:
            # more neighbours than the list kernels keep: D row stripes are
            # formed in HBM and sorted there; the matrix never exists whole
            idx, dist = self._knn_by_full_sort(
                self._stripes_of(self._cells(metric)), n, k)
        else:
            i, dd = self._cells(metric).knn_stripe(
                k, EXCLUDE_SELF, precision=self._gram_precision())
            self._knn_dev = (k, EXCLUDE_SELF, metric, i, dd)
            idx, dist = i.cpu().numpy(), dd.cpu().numpy()
        return list(idx), list(dist)

    def _stripes_of(self, cells):
        """(r0, r1, D stripe in HBM) of prepared cells, a few GB at a time;
        exact kernels only (these stripes are searched, not returned)."""
        n = cells.n
        rows = max(128, (STRIPE_BYTES // (8 * max(n, 1))) // 128 * 128)
        prec = self._gram_precision(cells.n, cells.p)
        if prec == "fast":
            prec = "fp64"
        for r0 in range(0, n, rows):
            r1 = min(n, r0 + rows)
            yield r0, r1, cells.pdist_stripe(r0, r1, precision=prec)

    def s_knn_connectivity_matrix(self, k, metric=None, use_pca=False,
                                  use_hnsw=False, index_params=None,
                                  query_params=None, verbose=False):
        """sdm.py:872-947: CSR, entry (i, j) = distance when j is one of i's
        kNN; exact-zero distances stored as -inf."""
        self._knn_dev = None
        targets, distances = self.s_knns(
            k=k, metric=metric, use_pca=use_pca, use_hnsw=use_hnsw,
            index_params=index_params, query_params=query_params,
            verbose=verbose)
        n = self._x.shape[0]
        kd = self._knn_dev
        if kd is not None and kd[0] == k and kd[1] == EXCLUDE_SELF:
            idx_dev, dist_dev = kd[3], kd[4]       # still resident in HBM
        else:
            idx_dev = dev.to_device(np.asarray(targets, dtype=np.int64))
            dist_dev = dev.to_device(np.asarray(distances, dtype=np.float64))
        indptr, indices, data = dev.knn_connectivity_csr(idx_dev, dist_dev)
        return spsparse.csr_matrix(
            (data.cpu().numpy(), indices.cpu().numpy(),
             indptr.cpu().numpy().astype(np.int32 if n * k < 2 ** 31
                                         else np.int64)), shape=(n, n))

    @staticmethod
    def knn_conn_mat_to_aff_edges(knn_conn_mat, aff_scale=1):
        """The edge list of ``knn_conn_mat_to_aff_graph`` (sdm.py:1085-1093)
        without igraph: ``(sources, targets, weights)`` with ``-inf -> 0`` and
        ``weights = (max - weights) * aff_scale``, computed on the device."""
        m = spsparse.csr_matrix(knn_conn_mat)
        sources, targets = m.nonzero()
        data = np.asarray(m[sources, targets]).ravel().astype(np.float64)
        w = dev.knn_affinity(dev.to_device(data), aff_scale)
        return sources, targets, w.cpu().numpy()

    @staticmethod
    def knn_conn_mat_to_aff_graph(knn_conn_mat, aff_scale=1):
        """sdm.py:1085-1093; needs python-igraph for the Graph object (the
        edge arithmetic itself is ``knn_conn_mat_to_aff_edges``)."""
        try:
            import igraph as ig
        except ImportError:
            raise ImportError("python-igraph is required for the Graph "
                              "object; use knn_conn_mat_to_aff_edges for the "
                              "edge list")
        s_, t_, w = SampleDistanceMatrix.knn_conn_mat_to_aff_edges(
            knn_conn_mat, aff_scale)
        return ig.Graph(edges=list(zip(s_.tolist(), t_.tolist())),
                        directed=False, edge_attrs={"weight": w})

    # ------------------------------------------------------------------ PCA
    @property
    def _pca_x(self):
        """sdm.py:1313-1334: the samples in the space of the first
        ``min(100, n, p)`` principal components (dense) or singular vectors
        (sparse, ``min(100, p - 1)``), memoised in ``_lazy_load_pca_x``."""
        if getattr(self, "_lazy_load_pca_x", None) is None:
            from . import pca as _pca
            self._lazy_load_pca_x = _pca.pca_x(self._x)
        return self._lazy_load_pca_x

    # ------------------------------------------------------------- read-only
    @property
    def metric(self):
        return self._metric

    def release_device(self):
        """Drop every HBM-resident cache (prepared cells, D)."""
        for c in self._cells_lut.values():
            c.close()
        self._cells_lut = {}
        self._d_dev = None
        self._bmin_dev = None
        self._knn_dev = None
        torch.cuda.empty_cache()
