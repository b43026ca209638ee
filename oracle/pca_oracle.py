"""NumPy restatement of scedar's PCA producer (the oracle of scedar_b200.pca).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity: pinned against
the live reference's ``_pca_x`` (tests/test_oracle_pca.py and the golden
``tests/golden/pca.npz``) where sklearn's solver is deterministic, and to a
stated tolerance where it is the seeded randomized SVD.

Reference: scedar/eda/sdm.py:134-138 (component count) and 1313-1334
(``_skd_pca`` / ``_pca_x``): dense x -> ``PCA(n_components=min(100, n, p),
svd_solver="auto", random_state=17)``, CSR x -> ``TruncatedSVD(
n_components=min(100, p - 1), random_state=17)``; ``transform(x)``.
Third-party algorithm restated (scikit-learn 1.9.0 installed here, un-vendored):
centre the columns (PCA only), thin SVD ``Xc = U S Vt``, signs by
``svd_flip(u, v, u_based_decision=False)`` -- the largest-magnitude entry of
every row of Vt positive --, scores ``U_k S_k`` (== ``Xc Vt_k^T``).  sklearn's
``auto`` picks LAPACK (``full``: max(n, p) <= 500 or k >= 0.8 min(n, p);
``covariance_eigh``: p <= 1000 and n >= 10 p) or randomized SVD (otherwise, and
always for TruncatedSVD); the exact decomposition below is what the randomized
one approximates.
"""
import numpy as np
import scipy.sparse as sps


def n_components_of(x):
    n, p = x.shape
    if sps.issparse(x):
        return max(min(100, p - 1), 0)
    return min(100, n, p)


def pca_x(x):
    k = n_components_of(x)
    if sps.issparse(x):
        xc = np.asarray(sps.csr_matrix(x, dtype="float64").todense())
        k = min(k, x.shape[0])
    else:
        xc = np.asarray(x, dtype="float64")
        xc = xc - xc.mean(axis=0)
    u, s, vt = np.linalg.svd(xc, full_matrices=False)
    idx = np.abs(vt).argmax(axis=1)
    signs = np.sign(vt[np.arange(vt.shape[0]), idx])
    signs[signs == 0] = 1.0
    return (u * s * signs)[:, :k]
