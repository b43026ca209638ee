"""CPU: the PCA oracle against the live reference's golden vectors
(tests/golden/make_golden_pca.py) and, in the build container, against the
live reference on seeded random cases."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import pca_oracle


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_golden(golden):
    g = golden("pca")
    for name in ("full_60x9", "full_40x120", "eigh_1200x40"):
        want = g[name + "_pca"]
        got = pca_oracle.pca_x(g[name + "_x"])
        assert got.shape == want.shape
        assert rel(got, want) <= 1e-9, name
    xs = sps.csr_matrix((g["csr_300x30_data"], g["csr_300x30_indices"],
                         g["csr_300x30_indptr"]), shape=(300, 30))
    got = pca_oracle.pca_x(xs)
    assert got.shape == g["csr_300x30_pca"].shape == (300, 29)
    assert rel(got, g["csr_300x30_pca"]) <= 1e-9


def test_component_counts():
    """sdm.py:134-138."""
    assert pca_oracle.n_components_of(np.zeros((500, 300))) == 100
    assert pca_oracle.n_components_of(np.zeros((40, 300))) == 40
    assert pca_oracle.n_components_of(np.zeros((400, 30))) == 30
    assert pca_oracle.n_components_of(sps.csr_matrix((400, 30))) == 29
    assert pca_oracle.n_components_of(sps.csr_matrix((400, 300))) == 100


def test_against_live_reference():
    from oracle.ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("live reference only exists in the build container")
    import warnings
    warnings.simplefilter("ignore")
    ref = load_reference()
    rng = np.random.default_rng(5)
    for n, p in ((80, 12), (30, 90), (2500, 60)):
        x = rng.standard_normal((n, p)) * np.linspace(2.0, 0.3, p) + 1.0
        assert rel(pca_oracle.pca_x(x), ref(x, metric="cosine")._pca_x) <= 1e-9
    # randomized solver (n < 10 p, large): the LEADING components of data with
    # a clear spectral gap agree; the trailing ones are the approximation's
    basis = np.linalg.qr(rng.standard_normal((700, 700)))[0][:8]
    x = (rng.standard_normal((1500, 8)) * np.array([40, 30, 22, 16, 12, 9, 7, 5.])
         ) @ basis + 0.05 * rng.standard_normal((1500, 700))
    a, b = pca_oracle.pca_x(x), ref(x, metric="cosine")._pca_x
    assert a.shape == b.shape == (1500, 100)
    assert rel(a[:, :8], b[:, :8]) <= 1e-6
