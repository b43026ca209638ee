"""PCA / truncated-SVD producer of the kNN path.

Replaces ``SampleDistanceMatrix._pca_x`` (scedar/eda/sdm.py:1313-1334): for a
dense x, ``sklearn.decomposition.PCA(n_components=min(100, n, p),
svd_solver="auto", random_state=17).fit(x).transform(x)``; for a CSR x,
``TruncatedSVD(n_components=min(100, p - 1), random_state=17)`` (no centring).
It makes the 50-PC input of BASELINE.json configs[4] from the cells x genes
matrix without leaving the device.

Method.  The scores of the first k components are the leading part of
``U S`` of the (column-centred) matrix ``Xc = U S Vt``.  The n p^2- /n^2 p-sized
work runs in the FP64 tensor-core Gram kernel (gram_f64.cu, raw-dot epilogue):

* p <= n (cells >= genes): scatter matrix ``C = Xc^T Xc`` [p, p] as the Gram of
  the TRANSPOSED working copy; ``eigh(C)`` -> V_k; scores ``X V_k - mean V_k``
  as the cross product of the cells with V_k^T;
* n < p: ``G = Xc Xc^T`` [n, n] as the Gram of the column-centred cells;
  ``eigh(G)`` -> U_k, S_k; scores ``U_k S_k``; the signs need V_k =
  ``Xc^T U_k / S_k``, again a cross product.

The symmetric eigenproblem of the small matrix (min(n, p) <= a few thousand;
O(min(n,p)^3), off the n^2 path) is cuSOLVER's ``syevd`` through
``torch.linalg.eigh`` -- a library call, stated as such.

sklearn's solvers are deterministic LAPACK (``full`` for small inputs,
``covariance_eigh`` for n >= 10 p, p <= 1000) or randomized SVD seeded with 17
(everything else, and TruncatedSVD): the exact decomposition computed here is
what the randomized one approximates.  Sign convention: sklearn's ``svd_flip(u,
v, u_based_decision=False)`` -- the entry of largest magnitude of every
component (row of Vt) is positive.
"""
import numpy as np
import scipy.sparse as spsparse
import torch

from . import device as dev

MAX_COMPONENTS = 100


def n_components_of(x):
    """sdm.py:134-138."""
    n, p = x.shape
    if spsparse.issparse(x):
        return max(min(MAX_COMPONENTS, p - 1), 0)
    return min(MAX_COMPONENTS, n, p)


def _flip_signs(v_cols):
    """+1 / -1 per column of V [p, k]: sign of its largest-magnitude entry."""
    idx = v_cols.abs().argmax(dim=0)
    s = torch.sign(v_cols[idx, torch.arange(v_cols.shape[1],
                                            device=v_cols.device)])
    s[s == 0] = 1.0
    return s


def pca_scores_device(cells, k, center):
    """Scores [n, k] (float64 CUDA tensor) of the prepared cells (metric
    "cosine": the working copy is x itself).  ``cells`` is modified when
    n < p and ``center`` (columns shifted in place)."""
    n, p = cells.n, cells.p
    if k <= 0 or n == 0:
        return dev.empty((n, max(k, 0)), torch.float64)
    mean = cells.col_means() if center else None
    if p <= n:
        ct = cells.transposed(mean)                # [p, n], centred
        try:
            scatter = ct.gram_raw()                # Xc^T Xc  [p, p]
        finally:
            ct.close()
        w, v = torch.linalg.eigh(scatter)          # ascending
        vk = v[:, -k:].flip(1).contiguous()        # [p, k], descending
        vk = vk * _flip_signs(vk)
        cv = dev.Cells.from_dense(vk.T.contiguous(), "cosine")
        try:
            y = cells.gram_raw_cross(cv)           # X V_k  [n, k]
        finally:
            cv.close()
        if mean is not None:
            y = y - (mean @ vk)
        return y
    if mean is not None:
        cells.shift_cols(mean)
    gram = cells.gram_raw()                        # Xc Xc^T  [n, n]
    w, u = torch.linalg.eigh(gram)
    uk = u[:, -k:].flip(1).contiguous()            # [n, k]
    sk = torch.sqrt(torch.clamp(w[-k:].flip(0), min=0.0))
    # V_k S_k = Xc^T U_k: only its signs are needed
    ct = cells.transposed()                        # [p, n]
    cu = dev.Cells.from_dense(uk.T.contiguous(), "cosine")
    try:
        vs = ct.gram_raw_cross(cu)                 # [p, k]
    finally:
        ct.close()
        cu.close()
    return uk * (sk * _flip_signs(vs))


def pca_x(x):
    """``_pca_x`` of a dense ndarray or CSR matrix -> ndarray [n, k]."""
    k = n_components_of(x)
    n, p = x.shape
    if k <= 0 or n == 0:
        return np.zeros((n, max(k, 0)))
    if spsparse.issparse(x):
        m = spsparse.csr_matrix(x, dtype="float64")
        if not m.has_canonical_format:
            m = m.copy()
            m.sum_duplicates()
        cells = dev.Cells.from_csr(
            dev.to_device(m.indptr.astype(np.int64, copy=False)),
            dev.to_device(m.indices.astype(np.int32, copy=False)),
            dev.to_device(m.data.astype(np.float64, copy=False)), m.shape,
            "cosine")
        center = False
        k = min(k, n)
    else:
        cells = dev.Cells.from_dense(
            dev.to_device(np.ascontiguousarray(x, dtype=np.float64)),
            "cosine")
        center = True
    try:
        y = pca_scores_device(cells, k, center)
    finally:
        cells.close()
    return y.cpu().numpy()
