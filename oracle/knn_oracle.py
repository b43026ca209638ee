"""NumPy restatement of the consumers directly behind the distance-matrix /
kNN path: affinity edge weights,
rare-sample detection and kNN feature imputation.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity: pinned
(golden vectors generated from the live reference by
tests/golden/make_golden.py::consumer_cases + the reference's known-answer
tests, tests/test_knn/test_detection_dense.py:14-24 and
tests/test_knn/test_imputation.py:29-43).

Paths are relative to the upstream checkout.
"""
import numpy as np

from . import sdm_oracle as orc


def affinity_weights(conn_csr, aff_scale=1):
    """scedar/eda/sdm.py:1085-1093 -- edge (source, target, weight) arrays of
    ``knn_conn_mat_to_aff_graph``: CSR order, ``-inf -> 0``,
    ``(max - w) * aff_scale``."""
    sources, targets = conn_csr.nonzero()
    weights = np.asarray(conn_csr[sources, targets]).ravel().copy()
    weights[weights == -np.inf] = 0
    weights = (weights.max() - weights) * aff_scale
    return sources, targets, weights


def rare_pdist(d, k, d_cutoff, n_iter):
    """scedar/knn/detection.py:81-123 -- iterative detection on the distance
    matrix: keep a sample while the ``min(m-1, k)``-th entry of its sorted
    column (self at rank 0) is <= the iteration's cutoff."""
    curr = d
    srt = np.sort(curr, axis=0)
    d_max = srt[k, :].max()
    curr_s_inds = np.arange(curr.shape[0])
    progress = [curr_s_inds.tolist()]
    for i in range(1, n_iter + 1):
        i_d_cutoff = d_cutoff + (n_iter - i) / n_iter * max(0, d_max - d_cutoff)
        i_k = min(curr.shape[0] - 1, k)
        if curr.shape[0]:
            kept = np.nonzero(srt[i_k, :] <= i_d_cutoff)[0]
        else:
            kept = np.zeros(0, dtype=int)
        curr = curr[np.ix_(kept, kept)]
        srt = np.sort(curr, axis=0)
        curr_s_inds = curr_s_inds[kept]
        progress.append(curr_s_inds.tolist())
    return curr_s_inds.tolist(), progress


def rare_no_pdist(x, metric, k, d_cutoff, n_iter):
    """scedar/knn/detection.py:33-79 -- the same procedure from kNN lists
    recomputed on the kept rows of x every iteration (self excluded, so the
    criterion is the k-th neighbour's distance)."""
    x = np.asarray(x, dtype="float64")
    _, dist = orc.knns_raw(x, metric, k)
    d_max = dist[:, -1].max()
    curr_s_inds = np.arange(x.shape[0])
    progress = [curr_s_inds.tolist()]
    for i in range(1, n_iter + 1):
        i_d_cutoff = d_cutoff + (n_iter - i) / n_iter * max(0, d_max - d_cutoff)
        kept = np.nonzero(dist[:, -1] <= i_d_cutoff)[0]
        curr_s_inds = curr_s_inds[kept]
        progress.append(curr_s_inds.tolist())
        if len(curr_s_inds) == 0:
            return curr_s_inds.tolist(), progress
        x = x[kept, :]
        _, dist = orc.knns_raw(x, metric, k)
    return curr_s_inds.tolist(), progress


def impute(x, knn, k, n_do, min_present_val, n_iter):
    """scedar/knn/imputation.py:35-135 with ``statistic_fun=np.median`` on a
    dense matrix.  ``knn``: (n, k) neighbour indices (``s_knn_ind_lut``).
    Returns (imputed x, indicator of the last iteration that imputed an
    entry, (entries, cells) picked up)."""
    curr = np.array(x, dtype="float64")
    knn = np.asarray(knn).reshape(curr.shape[0], k)
    idc = np.zeros(curr.shape, dtype=int)
    for i in range(1, n_iter + 1):
        nxt = curr.copy()
        present = curr >= min_present_val
        n_do_i = n_do + int(np.ceil((n_iter - i) / n_iter * (k - n_do)))
        for s in range(curr.shape[0]):
            knn_at = present[knn[s], :].sum(axis=0) >= n_do_i
            mask = np.logical_and(knn_at, present[s] == 0)
            if mask.any():
                nxt[s, mask] = np.median(curr[knn[s], :][:, mask], axis=0)
                idc[s, mask] = i
        curr = nxt
    per_s = (idc > 0).sum(axis=1)
    return curr, idc, (int(per_s.sum()), int((per_s > 0).sum()))
