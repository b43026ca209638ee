"""Drop-in ``scedar.knn``: rare-sample detection and kNN feature imputation on
top of the B200 distance-matrix / kNN path.

Same classes, arguments, return values and error behaviour as
scedar/knn/detection.py and scedar/knn/imputation.py.  The reference runs one
parameter tuple per forked ``parmap`` child (scedar/utils.py:17-60); a CUDA
context does not survive ``fork``, so the tuples run one after another on the
device here and ``nprocs`` is accepted for compatibility only.  An exception
inside one tuple propagates (the reference's ``parmap`` swallows it into the
result list and then fails on ``res[0]``).
"""
import time

import numpy as np
import scipy.sparse as spsp
import torch

from . import device as dev
from ._capi import EXCLUDE_SELF
from .sdm import MAX_TOPK, SampleDistanceMatrix


def _as_list(v):
    return [v] if np.isscalar(v) else list(v)


class RareSampleDetection(object):
    """K nearest neighbour detection of rare samples
    (scedar/knn/detection.py:7-31).

    The pdist variant keeps D resident in HBM and, per iteration, takes the
    ``i_k``-th smallest entry of every live column restricted to the live
    rows (``scedar_masked_kth``) instead of re-slicing and re-sorting the
    shrinking matrix; the no-pdist variant gathers the kept cells and reruns
    the fused Gram -> top-k kernel.
    """

    def __init__(self, sdm):
        super(RareSampleDetection, self).__init__()
        self._sdm = sdm
        self._res_lut = {}

    # detection.py:33-79
    def _no_pdist_rare_s_detect(self, k, d_cutoff, n_iter, metric=None,
                                use_pca=False, use_hnsw=False,
                                index_params=None, query_params=None):
        param_key = (k, d_cutoff, n_iter)
        if param_key in self._res_lut:
            return self._res_lut[param_key]
        sdm = self._sdm
        metric = sdm._metric if metric is None else metric
        if (use_pca or use_hnsw or index_params or query_params
                or not SampleDistanceMatrix._on_b200_path(metric)):
            # PCA-space / HNSW / foreign-metric searches: the reference's own
            # loop when this is a patched scedar class, else the refusals of
            # SampleDistanceMatrix.s_knns
            ref = getattr(type(self), "_ref__no_pdist_rare_s_detect", None)
            if ref is not None:
                return ref(self, k, d_cutoff, n_iter, metric=metric,
                           use_pca=use_pca, use_hnsw=use_hnsw,
                           index_params=index_params,
                           query_params=query_params)
            sdm.s_knns(k, metric=metric, use_pca=use_pca, use_hnsw=use_hnsw,
                       index_params=index_params, query_params=query_params)
            if use_pca:
                raise NotImplementedError(
                    "rare-sample detection over PCA coordinates: build the "
                    "SampleDistanceMatrix on the PCs instead")
        sdm._check_metric(metric)
        x = sdm._x.toarray() if spsp.issparse(sdm._x) else sdm._x
        x_dev = dev.to_device(np.ascontiguousarray(x, dtype=np.float64))
        def kth_nn(xd):
            """Distance to the k-th neighbour of every cell: tensor [m, 1]."""
            m = xd.shape[0]
            if k + 1 > m:
                raise ValueError(
                    "Expected n_neighbors < n_samples_fit, but n_neighbors = "
                    "{}, n_samples_fit = {}".format(k, m))
            cells = dev.Cells.from_dense(xd, metric)
            try:
                if k > MAX_TOPK:
                    # beyond the list kernels: sorted D stripes in HBM
                    _, dist = sdm._knn_by_full_sort(sdm._stripes_of(cells),
                                                    m, k)
                    return dev.to_device(
                        np.ascontiguousarray(dist[:, k - 1:k]))
                _, dist = cells.knn_stripe(
                    k, EXCLUDE_SELF,
                    precision=sdm._gram_precision(m, xd.shape[1]))
            finally:
                cells.close()
            return dist[:, k - 1:k].contiguous()

        dist = kth_nn(x_dev)
        m = x_dev.shape[0]
        d_max = dev.masked_max(dist, None, ldv=1, n=m)
        curr_s_inds = np.arange(m)
        progress_list = [curr_s_inds.tolist()]
        for i in range(1, n_iter + 1):
            i_d_cutoff = (d_cutoff +
                          (n_iter - i) / n_iter * max(0, d_max - d_cutoff))
            alive = dev.ones(m, torch.uint8)
            scratch = dev.zeros(m, torch.int32)
            kept = dev.rare_update(dist, i_d_cutoff, i, alive, scratch,
                                   ldk=1)
            sel = dev.compact_alive(alive, kept)
            curr_s_inds = curr_s_inds[sel.cpu().numpy()]
            progress_list.append(curr_s_inds.tolist())
            if kept == 0:
                return curr_s_inds.tolist(), progress_list
            if k + 1 > kept:
                # the reference rebuilds the kNN of the kept cells here and
                # sklearn refuses n_neighbors >= n_samples_fit
                raise ValueError(
                    "Expected n_neighbors < n_samples_fit, but n_neighbors = "
                    "{}, n_samples_fit = {}".format(k, kept))
            if kept < m and i < n_iter:
                x_dev = dev.gather_rows(x_dev, sel)
                m = kept
                dist = kth_nn(x_dev)
        return curr_s_inds.tolist(), progress_list

    # detection.py:81-123
    def _pdist_rare_s_detect(self, k, d_cutoff, n_iter, metric=None,
                             use_pca=False, use_hnsw=False,
                             index_params=None, query_params=None):
        param_key = (k, d_cutoff, n_iter)
        if param_key in self._res_lut:
            return self._res_lut[param_key]
        sdm = self._sdm
        sdm._validate_pdist()
        n = sdm._x.shape[0]
        alive = dev.ones(n, torch.uint8)
        last_kept = dev.zeros(n, torch.int32)
        kth = dev.empty(n, torch.float64)

        def kth_of_alive(rank):
            for r0, r1, stripe in sdm._d_stripes():
                dev.masked_kth(stripe, alive, rank, row_begin=r0,
                               out=kth[r0:r1])

        kth_of_alive(k)          # np.sort(d, axis=0)[k, :]
        d_max = dev.masked_max(kth, None)
        m, fresh_rank = n, k
        for i in range(1, n_iter + 1):
            i_d_cutoff = (d_cutoff +
                          (n_iter - i) / n_iter * max(0, d_max - d_cutoff))
            if m == 0:
                continue
            i_k = min(m - 1, k)
            if fresh_rank != i_k:
                kth_of_alive(i_k)
                fresh_rank = i_k
            kept = dev.rare_update(kth, i_d_cutoff, i, alive, last_kept)
            if kept != m:
                fresh_rank = None        # the live set changed
            m = kept
        lk = last_kept.cpu().numpy()
        progress_list = [np.nonzero(lk >= i)[0].tolist()
                         for i in range(n_iter + 1)]
        return progress_list[-1], progress_list

    # detection.py:125-278
    def detect_rare_samples(self, k, d_cutoff, n_iter, nprocs=1, metric=None,
                            use_pca=False, use_hnsw=False,
                            index_params=None, query_params=None):
        """Indices of the non-rare samples of every ``(k, d_cutoff, n_iter)``
        tuple; see the reference docstring for the procedure."""
        k_list = _as_list(k)
        d_cutoff_list = _as_list(d_cutoff)
        n_iter_list = _as_list(n_iter)
        if not (len(k_list) == len(d_cutoff_list) == len(n_iter_list)):
            raise ValueError("Parameter should have the same length."
                             "k: {}, d_cutoff: {}, n_iter: {}.".format(
                                 k, d_cutoff, n_iter))
        n_param_tups = len(k_list)
        for i in range(n_param_tups):
            if k_list[i] < 1 or k_list[i] > self._sdm._x.shape[0] - 1:
                raise ValueError("k should be >= 1 and <= n_samples-1. "
                                 "k: {}".format(k))
            k_list[i] = int(k_list[i])
            if d_cutoff_list[i] <= 0:
                raise ValueError("d_cutoff should be > 0. "
                                 "d_cutoff: {}".format(d_cutoff))
            d_cutoff_list[i] = float(d_cutoff_list[i])
            if n_iter_list[i] < 1:
                raise ValueError("n_iter should be >= 1. "
                                 "n_iter: {}".format(n_iter))
            n_iter_list[i] = int(n_iter_list[i])
        if int(nprocs) < 1:
            raise ValueError("nprocs should be >= 1. nprocs: {}".format(
                nprocs))
        param_tups = [(k_list[i], d_cutoff_list[i], n_iter_list[i], metric,
                       use_pca, use_hnsw, index_params, query_params)
                      for i in range(n_param_tups)]
        run = (self._pdist_rare_s_detect if self._sdm._use_pdist
               else self._no_pdist_rare_s_detect)
        res_list = [run(*ptup) for ptup in param_tups]
        for i in range(n_param_tups):
            param_key = param_tups[i][:3]
            if param_key not in self._res_lut:
                self._res_lut[param_key] = res_list[i]
        return [res[0] for res in res_list]


class FeatureImputation(object):
    """Impute dropped-out features from the k nearest neighbours
    (scedar/knn/imputation.py:13-32): if a feature is below
    ``min_present_val`` in a sample and at least ``n_do`` of its kNN have it
    at or above, it is replaced by the median of the k neighbour values."""

    def __init__(self, sdm):
        super(FeatureImputation, self).__init__()
        self._sdm = sdm
        self._res_lut = {}

    # imputation.py:35-135
    @staticmethod
    def _impute_features_runner(pb_x, knn_ordered_ind_dict, k, n_do,
                                min_present_val, n_iter,
                                statistic_fun=np.median, verbose=False):
        """One parameter set.  ``knn_ordered_ind_dict`` is the LUT of
        ``s_knn_ind_lut`` (or an ``(n, k)`` integer array).  x stays in HBM
        across the iterations (two buffers, Jacobi update)."""
        if statistic_fun is not np.median:
            raise NotImplementedError(
                "the device kernel implements statistic_fun=np.median (the "
                "reference default); other statistics are outside this path")
        start_time = time.time()
        sparse_in = spsp.issparse(pb_x)
        x_host = pb_x.toarray() if sparse_in else np.asarray(pb_x)
        n_samples, n_features = x_host.shape
        if isinstance(knn_ordered_ind_dict, np.ndarray):
            knn = np.ascontiguousarray(knn_ordered_ind_dict, dtype=np.int64)
        else:
            knn = np.array([knn_ordered_ind_dict[s] for s in
                            range(n_samples)], dtype=np.int64)
        knn = knn.reshape(n_samples, -1)
        if knn.shape[1] != k:
            raise ValueError("every sample needs exactly k = {} neighbours, "
                             "got {}".format(k, knn.shape[1]))
        curr = dev.to_device(np.ascontiguousarray(x_host, dtype=np.float64))
        nxt = torch.empty_like(curr)
        knn_dev = dev.to_device(knn)
        idc = dev.zeros((n_samples, n_features), torch.int32)
        counters = dev.zeros(2, torch.int64)
        stats = "Preparation: {:.2f}s\n".format(time.time() - start_time)
        if verbose:
            print(stats)
        for i in range(1, n_iter + 1):
            iter_start_time = time.time()
            # iteratively decreases n_do_i from k to n_do (imputation.py:66)
            n_do_i = n_do + int(np.ceil((n_iter - i) / n_iter * (k - n_do)))
            dev.impute_iter(curr, knn_dev, n_do_i, min_present_val, i, nxt,
                            idc, counters)
            curr, nxt = nxt, curr
            n_pu_entries, n_samples_wfpu = (int(v) for v in
                                            counters.cpu().tolist())
            n_pu_entries_ratio = n_pu_entries / (n_samples * n_features)
            iter_time = time.time() - iter_start_time
            stats_update = str.format(
                "Iteration {} ({:.2f}s): picked up {} total "
                "features in {} ({:.1%}) cells. {:.1%} among "
                "samples * features.\n", i, iter_time,
                n_pu_entries, n_samples_wfpu,
                n_samples_wfpu / n_samples,
                n_pu_entries_ratio)
            stats += stats_update
            if verbose:
                print(stats_update)
        curr_x_pb = curr.cpu().numpy()
        if sparse_in:
            curr_x_pb = spsp.csr_matrix(curr_x_pb)
        impute_idc_pb = spsp.lil_matrix(
            spsp.csr_matrix(idc.cpu().numpy().astype(int)))
        stats_update = "Complete in {:.2f}s\n".format(
            time.time() - start_time)
        if verbose:
            print(stats_update)
        stats += stats_update
        return curr_x_pb, impute_idc_pb, stats

    # imputation.py:137-319
    def impute_features(self, k, n_do, min_present_val, n_iter, nprocs=1,
                        statistic_fun=np.median, metric=None, use_pca=False,
                        use_hnsw=False, index_params=None, query_params=None,
                        verbose=False):
        """kNN imputation for every ``(k, n_do, min_present_val, n_iter)``
        tuple; returns the list of imputed ``SampleDistanceMatrix``."""
        start_time = time.time()
        try:
            if not np.isscalar(np.isreal(statistic_fun([0, 1, 2]))):
                raise ValueError("statistic_fun should be a function of a"
                                 "list of numbers that returns a scalar.")
        except Exception:
            raise ValueError("statistic_fun should be a function of a"
                             "list of numbers that returns a scalar.")
        k_list = _as_list(k)
        n_do_list = _as_list(n_do)
        min_present_val_list = _as_list(min_present_val)
        n_iter_list = _as_list(n_iter)
        if not (len(k_list) == len(n_do_list) ==
                len(min_present_val_list) == len(n_iter_list)):
            raise ValueError("Parameter should have the same length."
                             "k: {}, n_do: {}, min_present_val: {}, "
                             "n_iter: {}.".format(k, n_do, min_present_val,
                                                  n_iter))
        n_param_tups = len(k_list)
        for i in range(n_param_tups):
            if k_list[i] < 1 or k_list[i] >= self._sdm._x.shape[0]:
                raise ValueError("k should be >= 1 and < n_samples. "
                                 "k: {}".format(k))
            k_list[i] = int(k_list[i])
            if n_do_list[i] > k_list[i] or n_do_list[i] < 1:
                raise ValueError("n_do should be  <= k and >= 1. "
                                 "n_do: {}".format(n_do))
            n_do_list[i] = int(n_do_list[i])
            min_present_val_list[i] = float(min_present_val_list[i])
            if n_iter_list[i] < 1:
                raise ValueError("n_iter should be >= 1. "
                                 "n_iter: {}".format(n_iter))
            n_iter_list[i] = int(n_iter_list[i])
        if int(nprocs) < 1:
            raise ValueError("nprocs should be >= 1. nprocs: {}".format(
                nprocs))
        param_tups = [(k_list[i], n_do_list[i], min_present_val_list[i],
                       n_iter_list[i], statistic_fun)
                      for i in range(n_param_tups)]
        res_list = []
        run_param_tups = []
        res_list_run_inds = []
        for i, ptup in enumerate(param_tups):
            if ptup in self._res_lut:
                res_list.append(self._res_lut[ptup])
            else:
                run_param_tups.append(ptup)
                res_list.append(None)
                res_list_run_inds.append(i)
        run_res_list = []
        for ptup in run_param_tups:
            s_knn_ind_lut = self._sdm.s_knn_ind_lut(
                ptup[0], metric=metric, use_pca=use_pca, use_hnsw=use_hnsw,
                index_params=index_params, query_params=query_params,
                verbose=verbose)
            if verbose:
                print("Compute KNNs: {:.2f}s\n".format(
                    time.time() - start_time))
            run_res_list.append(self._impute_features_runner(
                self._sdm._x, s_knn_ind_lut, *ptup, verbose=verbose))
        for i, param_tup in enumerate(run_param_tups):
            if param_tup in self._res_lut:
                # the same tuple given twice in one call: reuse the first run
                res_tup = self._res_lut[param_tup]
            else:
                res_tup = (run_res_list[i][0], run_res_list[i][1],
                           run_res_list[i][2])
                self._res_lut[param_tup] = res_tup
            res_list[res_list_run_inds[i]] = res_tup
        kpu_sdm_list = []
        # the class of the wrapped object: a patched scedar installation gets
        # scedar.eda.SampleDistanceMatrix back (with its tsne / plots /
        # to_classified), the stand-alone package its own class
        sdm_cls = getattr(self, "_result_cls", None) or type(self._sdm)
        if not (isinstance(sdm_cls, type) and hasattr(sdm_cls, "s_knns")):
            sdm_cls = SampleDistanceMatrix
        for res in res_list:
            kpu_sdm = sdm_cls(
                res[0], metric=self._sdm._metric, sids=self._sdm.sids,
                fids=self._sdm.fids, nprocs=self._sdm._nprocs)
            kpu_sdm_list.append(kpu_sdm)
        return kpu_sdm_list
