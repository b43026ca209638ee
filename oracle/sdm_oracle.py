"""NumPy restatement of scedar's distance-matrix + kNN hot path (the oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity: pinned
(golden vectors from the live reference + the reference's known-answer tests).

Every function cites the reference lines it restates (paths relative to the
upstream checkout, ``scedar/eda/sdm.py`` unless stated).  Third-party
arithmetic that sits under the path (un-vendored, no lock file; versions
installed in the build image: numpy 2.3.5, scipy 1.18.1, scikit-learn 1.9.0)
is restated from its published algorithm:

* ``sklearn.metrics.pairwise.euclidean_distances`` (float64 branch):
  ``D2 = -2 X.Xt ; D2 += |x|^2[:,None] ; D2 += |x|^2[None,:] ; max(D2,0) ;
  diag=0 ; sqrt``  (sklearn/metrics/pairwise.py ``_euclidean_distances``).
* ``sklearn.neighbors.NearestNeighbors.kneighbors(None, k)``: the k+1 nearest
  by distance, then the sample itself removed by index; when the sample is
  not among its own k+1 nearest (more than k duplicates) the first column is
  dropped instead (sklearn/neighbors/_base.py ``kneighbors``).  sklearn
  leaves the order of exactly tied distances unspecified; the oracle (and the
  CUDA path) define it as "lower index first", the rule the reference's own
  stable argsort accessor uses.

Operation order that matters numerically is kept: ``sqrt(1/ss)`` (not
``1/sqrt(ss)``), ``(P*inv).T*inv``, diagonal forced to 1 before ``1-S``,
clamp<0 then zero diagonal then lower triangle mirrored upward.
"""
import warnings

import numpy as np
import scipy.sparse as sps

SPARSE_OK_METRICS = ("cityblock", "cosine", "euclidean", "l1", "l2",
                     "manhattan")
GPU_METRICS = ("cosine", "correlation", "euclidean")


# --------------------------------------------------------------------------
# input contract
# --------------------------------------------------------------------------
def coerce_x(x):
    """scedar/eda/sfm.py:43-59 -- float64 CSR for sparse input, float64
    ndarray otherwise, must be 2-D."""
    if x is None:
        raise ValueError("x cannot be None")
    if sps.issparse(x):
        x = sps.csr_matrix(x, dtype="float64")
    else:
        try:
            x = np.asarray(x, dtype="float64")
        except ValueError as e:
            raise ValueError("Features must be float. {}".format(e))
    if x.ndim != 2:
        raise ValueError("x has shape (n_samples, n_features)")
    return x


# --------------------------------------------------------------------------
# distance matrices
# --------------------------------------------------------------------------
def gram(x):
    """sdm.py:1244-1247 -- the pairwise dot-product matrix X.Xt, dense output
    for dense or CSR input."""
    if sps.issparse(x):
        return (x @ x.T).toarray()
    return np.dot(x, x.T)


def cosine_pdist(x):
    """sdm.py:1223-1266.  Zero vectors get inverse norm 0, hence distance 1 to
    everything (sdm.py:1249-1259); the diagonal of the similarity is forced to
    1 (sdm.py:1265) so self-distance is exactly 0."""
    p = gram(x)
    ss = np.diag(p)
    inv_ss = np.array([1 / s if s != 0 else 0 for s in ss], dtype="float64")
    inv_norm = np.sqrt(inv_ss)
    sim = p * inv_norm            # sim[i, j] = P[i, j] * inv[j]
    sim = sim.T * inv_norm        # sim[i, j] = P[j, i] * inv[i] * inv[j]
    n = sim.shape[0]
    sim[np.arange(n), np.arange(n)] = 1
    return 1 - sim


def correlation_pdist(x):
    """sdm.py:1268-1287 -- centre every row on its mean, then cosine."""
    mu = x.mean(axis=1).reshape(x.shape[0], 1)
    return cosine_pdist(x - mu)


def euclidean_pdist(x):
    """sdm.py:1215-1217 for metric='euclidean' -> sklearn
    ``pairwise_distances`` -> ``euclidean_distances`` float64 branch
    (published algorithm restated, see module docstring)."""
    if sps.issparse(x):
        sq = np.asarray(x.multiply(x).sum(axis=1)).ravel()
    else:
        sq = np.einsum("ij,ij->i", x, x)
    d2 = gram(x)
    d2 *= -2
    d2 += sq[:, None]
    d2 += sq[None, :]
    np.maximum(d2, 0, out=d2)
    n = d2.shape[0]
    d2[np.arange(n), np.arange(n)] = 0
    return np.sqrt(d2, out=d2)


def num_correct_dist_mat(dmat, upper_bound=None):
    """sdm.py:187-222.  In place; returns the same object.  Two sanity checks
    that only warn (diagonal within 1e-5 of 0; triangles agree to rtol 1e-3),
    then: negatives -> 0, diagonal -> 0, optional upper clamp, and the upper
    triangle overwritten with the lower one."""
    if (not isinstance(dmat, np.ndarray) or dmat.ndim != 2
            or dmat.shape[0] != dmat.shape[1]):
        raise ValueError("dmat must be a 2D (n_samples, n_samples) np array")
    n = dmat.shape[0]
    diag = dmat[np.arange(n), np.arange(n)]
    if not np.all(np.abs(diag) <= 1e-5):
        warnings.warn("distance matrix might not be numerically correct. "
                      "diag vals should be close to 0.")
    iu = np.triu_indices(n)
    up, lo = dmat[iu], dmat.T[iu]
    # assert_allclose(actual=up, desired=lo, rtol=1e-3, atol=0) semantics
    with np.errstate(invalid="ignore"):
        bad = ~(np.abs(up - lo) <= 1e-3 * np.abs(lo))
    bad &= ~((up == lo) | (np.isnan(up) & np.isnan(lo)))
    if bad.any():
        warnings.warn("distance matrix might not be numerically correct. "
                      "should be approximately symmetric.")
    dmat[dmat < 0] = 0
    dmat[np.arange(n), np.arange(n)] = 0
    if upper_bound is not None:
        upper_bound = float(upper_bound)
        dmat[dmat > upper_bound] = upper_bound
    dmat[iu] = dmat.T[iu]
    return dmat


def pdist(x, metric):
    """The ``_d`` property, sdm.py:1193-1221 (use_pdist=True branch)."""
    x = coerce_x(x)
    if sps.issparse(x) and metric not in SPARSE_OK_METRICS:
        raise ValueError("Only the following metrics are supportted for "
                         "sparse matrices, {}".format(list(SPARSE_OK_METRICS)))
    n = x.shape[0]
    if x.shape[0] * x.shape[1] == 0:
        return np.zeros((n, n))
    if metric == "cosine":
        d = cosine_pdist(x)
    elif metric == "correlation":
        d = correlation_pdist(x)
    elif metric == "euclidean":
        d = euclidean_pdist(x)
    else:
        raise ValueError("oracle covers only {}".format(GPU_METRICS))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return num_correct_dist_mat(d)


# --------------------------------------------------------------------------
# sorted columns, kNN
# --------------------------------------------------------------------------
def col_sorted_d(d):
    """sdm.py:1289-1296 -- stable sort of every column (row r = r-th nearest)."""
    return np.sort(d, kind="mergesort", axis=0)


def col_argsorted_d(d):
    """sdm.py:1298-1305 -- stable argsort of every column; ties keep the lower
    row index first."""
    return np.argsort(d, kind="mergesort", axis=0)


def knn_ind_lut(d, k):
    """sdm.py:751-767 (use_pdist=True branch): ranks 1..k of the stable column
    argsort -- rank 0 is dropped whatever it is."""
    n = d.shape[0]
    if k < 0 or k > n - 1:
        raise ValueError("k ({}) should >= 0 and <= n_samples-1".format(k))
    k = int(k)
    arg = col_argsorted_d(d)[1:k + 1, :]
    return dict(zip(range(n), arg.T.tolist()))


def knn_drop_first(d, k):
    """Array form of :func:`knn_ind_lut`: (idx[n,k] int64, dist[n,k])."""
    order = np.argsort(d, kind="mergesort", axis=1)[:, 1:k + 1]
    return order.astype(np.int64), np.take_along_axis(d, order, axis=1)


def knns_precomputed(d, k):
    """sdm.py:792-870 + 1053-1083 with use_pdist=True:
    ``NearestNeighbors(k, 'precomputed').fit(d).kneighbors(None, k)``.
    k+1 nearest per row (ties: lower index first), self removed by index, or
    the first column when self is absent.  Returns (idx[n,k], dist[n,k])."""
    if k < 1:
        raise ValueError("k should >= 1")
    n = d.shape[0]
    if k + 1 > n:
        raise ValueError("Expected n_neighbors < n_samples_fit")
    order = np.argsort(d, kind="mergesort", axis=1)[:, :k + 1]
    dist = np.take_along_axis(d, order, axis=1)
    keep = order != np.arange(n)[:, None]
    absent = keep.all(axis=1)
    keep[absent, 0] = False
    idx = order[keep].reshape(n, k).astype(np.int64)
    return idx, dist[keep].reshape(n, k)


def knns_raw(x, metric, k):
    """sdm.py:1053-1083 with use_pdist=False: brute-force kNN on the raw
    matrix.  Distances are the same metric formulas as :func:`pdist`; self
    handling as in :func:`knns_precomputed`."""
    return knns_precomputed(pdist(x, metric), k)


def knn_connectivity(idx, dist, n):
    """sdm.py:872-947 -- CSR matrix with entry (i, idx[i, r]) = dist[i, r],
    exact zeros replaced by -inf."""
    k = idx.shape[1]
    src = np.repeat(np.arange(n), k)
    val = np.array(dist, dtype="float64").ravel()
    val[val == 0] = -np.inf
    return sps.coo_matrix((val, (src, np.asarray(idx).ravel())),
                          shape=(n, n)).tocsr()
