"""GPU: the PCA producer (scedar_b200.pca) against
the live reference's golden vectors and the oracle, and the PCA-space kNN it
feeds (``s_knns(use_pca=True)``, sdm.py:1060-1065)."""
import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    import scedar_b200
    from scedar_b200 import _capi
    _capi.load()
    return scedar_b200


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_pca_golden(b200, golden):
    """The reference's own ``_pca_x`` (deterministic sklearn solvers)."""
    from scedar_b200 import pca
    g = golden("pca")
    for name in ("full_60x9", "full_40x120", "eigh_1200x40"):
        got = pca.pca_x(g[name + "_x"])
        want = g[name + "_pca"]
        assert got.shape == want.shape
        # 40 centred samples span 39 dimensions: the 40th "component" is the
        # null direction, 1e-15 in the reference's SVD and sqrt(1e-15) through
        # a Gram matrix -- noise either way, compared on the absolute scale
        k = got.shape[1] - (1 if name == "full_40x120" else 0)
        assert rel(got[:, :k], want[:, :k]) <= 1e-9, name
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    xs = sps.csr_matrix((g["csr_300x30_data"], g["csr_300x30_indices"],
                         g["csr_300x30_indptr"]), shape=(300, 30))
    got = pca.pca_x(xs)
    assert got.shape == (300, 29)
    assert rel(got, g["csr_300x30_pca"]) <= 1e-9


def test_pca_building_blocks(b200):
    """Column means, transposed copy, raw Gram and cross products against
    NumPy (ragged shapes: padding rows / columns must stay out)."""
    import torch
    from scedar_b200 import device as dev
    rng = np.random.default_rng(3)
    for n, p in ((300, 45), (77, 200), (1030, 129)):
        x = rng.standard_normal((n, p)) + 0.5
        c = dev.Cells.from_dense(dev.to_device(x), "cosine")
        mu = c.col_means().cpu().numpy()
        np.testing.assert_allclose(mu, x.mean(axis=0), rtol=1e-12, atol=1e-14)
        ct = c.transposed(dev.to_device(mu))
        sc = ct.gram_raw().cpu().numpy()
        xc = x - mu
        np.testing.assert_allclose(sc, xc.T @ xc, rtol=1e-11, atol=1e-9)
        assert np.array_equal(sc, sc.T)
        v = rng.standard_normal((7, p))
        cv = dev.Cells.from_dense(dev.to_device(v), "cosine")
        y = c.gram_raw_cross(cv).cpu().numpy()
        np.testing.assert_allclose(y, x @ v.T, rtol=1e-11, atol=1e-10)
        g = c.gram_raw().cpu().numpy()
        np.testing.assert_allclose(g, x @ x.T, rtol=1e-11, atol=1e-9)
        for h in (c, ct, cv):
            h.close()
    torch.cuda.synchronize()


@pytest.mark.parametrize("n,p", [(5000, 300), (400, 1500), (20000, 500)])
def test_pca_against_oracle(b200, n, p):
    """Low-rank signal + noise (a clear spectral gap after 12 components, as
    single-cell data has): every well-separated component must match the exact
    SVD of the oracle; the scores as a whole in the Frobenius sense."""
    from oracle import pca_oracle
    from scedar_b200 import pca
    rng = np.random.default_rng(n + p)
    basis = np.linalg.qr(rng.standard_normal((p, 12)))[0].T
    amp = 30.0 * 0.8 ** np.arange(12)
    x = (rng.standard_normal((n, 12)) * amp) @ basis
    x += 0.1 * rng.standard_normal((n, p)) + rng.standard_normal((1, p))
    got, want = pca.pca_x(x), pca_oracle.pca_x(x)
    assert got.shape == want.shape == (n, min(100, n, p))
    assert rel(got[:, :12], want[:, :12]) <= 1e-8
    # the noise components are near-degenerate: compare the variance captured
    np.testing.assert_allclose((got ** 2).sum(axis=0), (want ** 2).sum(axis=0),
                               rtol=1e-8)


def test_knn_in_pca_space(b200):
    """``s_knns(k, use_pca=True)`` / ``s_knn_ind_lut(k, use_pca=True)``: the
    oracle's kNN on the oracle's PCA coordinates (use_pdist=False route of
    sdm.py:1060-1065)."""
    from oracle import pca_oracle, sdm_oracle as orc
    rng = np.random.default_rng(8)
    basis = np.linalg.qr(rng.standard_normal((200, 10)))[0].T
    x = (rng.standard_normal((1500, 10)) * (20.0 * 0.85 ** np.arange(10))) @ basis
    x += 0.05 * rng.standard_normal((1500, 200))
    pcs = pca_oracle.pca_x(x)
    for use_pdist in (True, False):
        sdm = b200.SampleDistanceMatrix(x, metric="euclidean",
                                        use_pdist=use_pdist)
        idx, dist = sdm.s_knns(15, use_pca=True)
        ridx, rdist = orc.knns_precomputed(orc.pdist(pcs, "euclidean"), 15)
        assert np.array_equal(np.array(idx), ridx)
        np.testing.assert_allclose(np.array(dist), rdist, rtol=1e-8)
        np.testing.assert_allclose(sdm._pca_x[:, :10], pcs[:, :10], rtol=1e-7,
                                   atol=1e-9)
    # config 5's producer at 1/10 scale: 100,000 x 2,000 -> PCs -> kNN
    from scedar_b200 import device as dev, pca
    import torch
    n, p = 100000, 2000
    g = torch.Generator(device="cuda").manual_seed(5)
    bas = torch.linalg.qr(torch.randn((p, 50), generator=g, device="cuda",
                                      dtype=torch.float64))[0].T
    amp = 10.0 / torch.sqrt(1.0 + torch.arange(50, device="cuda",
                                               dtype=torch.float64))
    xd = (torch.randn((n, 50), generator=g, device="cuda",
                      dtype=torch.float64) * amp) @ bas
    xd += 0.01 * torch.randn((n, p), generator=g, device="cuda",
                             dtype=torch.float64)
    cells = dev.Cells.from_dense(xd, "cosine")
    y = pca.pca_scores_device(cells, 50, center=True)
    cells.close()
    assert tuple(y.shape) == (n, 50)
    xc = xd - xd.mean(dim=0)
    # the scores reproduce the centred data up to the noise floor, and their
    # columns are orthogonal with descending norms
    gram = (y.T @ y).cpu().numpy()
    off = gram - np.diag(np.diag(gram))
    assert np.abs(off).max() <= 1e-7 * np.diag(gram).max()
    assert (np.diff(np.diag(gram)) <= 0).all()
    captured = float((y ** 2).sum() / (xc ** 2).sum())
    assert captured > 0.99
    pcs_cells = dev.Cells.from_dense(y.contiguous(), "euclidean")
    from scedar_b200._capi import EXCLUDE_SELF
    ii, dd = pcs_cells.knn_stripe(30, EXCLUDE_SELF, precision="fast")
    assert tuple(ii.shape) == (n, 30) and bool((dd[:, 1:] >= dd[:, :-1]).all())
    pcs_cells.close()
