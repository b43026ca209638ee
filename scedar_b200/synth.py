"""Seeded synthetic cells x genes matrices for the BASELINE.json configs.

Host-side NumPy generators.  All return float64.
"""
import numpy as np
import scipy.sparse as sps


def log_counts(n, p, seed):
    """C1 / C4 shape: Poisson(lambda_g) counts, lambda_g ~ Gamma(0.3, 2.0) per
    gene, then log1p -- continuous, ~70-80 % zeros, no exact ties."""
    rng = np.random.default_rng(seed)
    lam = rng.gamma(0.3, 2.0, size=p)
    x = np.empty((n, p), dtype=np.float64)
    step = max(1, (1 << 24) // max(p, 1))
    for i in range(0, n, step):
        j = min(n, i + step)
        x[i:j] = np.log1p(rng.poisson(lam, size=(j - i, p)))
    return x


def raw_counts(n, p, seed, lam=0.5):
    """C1-ties: raw integer Poisson counts.  Dot products are exact in
    float64, so equal distances are real ties -- the tie-break stress."""
    rng = np.random.default_rng(seed)
    return rng.poisson(lam, size=(n, p)).astype(np.float64)


def sparse_lognormal(n, p, seed, density=0.10):
    """C2: CSR, ~density nonzeros, lognormal(0, 1) values."""
    rng = np.random.default_rng(seed)
    m = sps.random(n, p, density=density, format="csr", random_state=rng,
                   data_rvs=lambda s: rng.lognormal(0.0, 1.0, size=s))
    m.sort_indices()
    return m.astype(np.float64)


def blobs(n, p, seed, n_blobs=10, outlier_frac=0.01):
    """C3: Gaussian blobs (sigma 1, centres N(0, 3^2)) shifted >= 0 plus a
    fraction of uniform outliers (the rare cells)."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(0.0, 3.0, size=(n_blobs, p))
    lab = rng.integers(0, n_blobs, size=n)
    x = np.empty((n, p), dtype=np.float64)
    step = max(1, (1 << 24) // max(p, 1))
    for i in range(0, n, step):
        j = min(n, i + step)
        x[i:j] = centres[lab[i:j]] + rng.standard_normal((j - i, p))
    n_out = int(round(n * outlier_frac))
    if n_out:
        rows = rng.choice(n, size=n_out, replace=False)
        x[rows] = rng.uniform(-12.0, 12.0, size=(n_out, p))
    x -= x.min()
    return x


def pcs(n, p, seed):
    """C5: PC-like coordinates, standard normal times 1/sqrt(1+j)."""
    rng = np.random.default_rng(seed)
    scale = 1.0 / np.sqrt(1.0 + np.arange(p))
    x = np.empty((n, p), dtype=np.float64)
    step = max(1, (1 << 24) // max(p, 1))
    for i in range(0, n, step):
        j = min(n, i + step)
        x[i:j] = rng.standard_normal((j - i, p)) * scale
    return x
