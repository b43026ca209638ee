"""Seeded randomised stress of the CUDA path: odd shapes, integer data with
exact ties, duplicated / all-zero / constant rows, random k -- FP64 D against
the oracle, every kNN route against selection from that D, the tcgen05 route
against the FP64 route (bit-identical)."""
import numpy as np
import pytest

from oracle import sdm_oracle as orc

pytestmark = pytest.mark.gpu


def make_case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 700))
    p = int(rng.integers(1, 300))
    kind = seed % 4
    if kind == 0:
        x = rng.normal(0.0, 1.0, size=(n, p))
    elif kind == 1:
        x = rng.poisson(0.8, size=(n, p)).astype(np.float64)       # exact ties
    elif kind == 2:
        x = np.log1p(rng.poisson(rng.gamma(0.3, 2.0, size=p), size=(n, p)))
    else:
        x = rng.uniform(-5.0, 5.0, size=(n, p)) * rng.lognormal(0, 2, size=(n, 1))
    if n > 6 and seed % 3 == 0:                 # duplicates, a zero and a constant row
        x[1] = x[0]
        x[n - 1] = x[n // 2]
        x[2] = 0.0
        x[3] = 1.5
    metric = ("euclidean", "cosine", "correlation")[int(rng.integers(0, 3))]
    if metric == "correlation" and p < 2:
        metric = "cosine"
    k = int(rng.integers(1, min(n - 1, 55) + 1))
    return x, metric, k


@pytest.mark.parametrize("seed", range(36))
def test_random_case(seed):
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
    x, metric, k = make_case(seed)
    n, p = x.shape
    cells = dev.Cells.from_dense(torch.from_numpy(x).cuda(), metric)
    d = cells.pdist_symmetric()
    dn = d.cpu().numpy()
    want = orc.pdist(x, metric)
    scale = max(want.max(), 1e-300)
    integer = np.array_equal(x, np.round(x))
    # exactly duplicated non-integer rows: the reference's own value is
    # cancellation noise; everything else to 1e-10
    mask = np.ones_like(want, dtype=bool)
    if metric == "euclidean" and not integer and n > 6 and seed % 3 == 0:
        for a, b in ((0, 1), (n - 1, n // 2)):
            mask[a, b] = mask[b, a] = False
    assert np.abs(dn - want)[mask].max() / scale <= 1e-10, (seed, metric, n, p)
    assert np.array_equal(dn, dn.T) and not dn.diagonal().any() and (dn >= 0).all()
    for ex in (EXCLUDE_FIRST, EXCLUDE_SELF):
        i0, v0 = dev.knn_from_d(d, k, ex)
        i1, v1 = cells.knn_stripe(k, ex)                      # FP64, auto route
        assert torch.equal(i0, i1) and torch.equal(v0, v1), (seed, metric, n, p, k, ex)
        i2, v2 = cells.knn_stripe(k, ex, precision="fast")    # tcgen05 + re-rank
        assert torch.equal(i0, i2) and torch.equal(v0, v2), (seed, metric, n, p, k, ex)
    ridx, _ = orc.knn_drop_first(dn, k)
    assert np.array_equal(dev.knn_from_d(d, k, EXCLUDE_FIRST)[0].cpu().numpy(), ridx)
    # stated tolerance of the tcgen05 matrix: cosine / correlation 1e-5
    # absolute; euclidean 2e-6 * max |x|^2 on the SQUARED distance (a square
    # root near zero turns a 1e-6 error of d^2 into 1e-3 of d: exact duplicates)
    dfast = cells.pdist_symmetric(precision="fast")
    if metric == "euclidean":
        n2max = float((x * x).sum(axis=1).max())
        err2 = (dfast * dfast - d * d).abs().max().item()
        assert err2 <= 2e-6 * max(n2max, 1e-300), (seed, metric, n, p, err2, n2max)
    else:
        assert (dfast - d).abs().max().item() <= 1e-5, (seed, metric, n, p)
    assert torch.equal(dfast, dfast.T)
    cells.close()


@pytest.mark.parametrize("seed", range(12))
def test_random_stripes_and_csr(seed):
    """Random row ranges (the multi-GPU stripes) of D and of the kNN lists equal
    the rows of the whole-matrix result; CSR input equals its densified twin;
    the consumers agree with the oracle on the same random case."""
    import scipy.sparse as sps
    import torch
    from oracle import knn_oracle as korc
    from scedar_b200 import SampleDistanceMatrix, device as dev, knn
    from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
    rng = np.random.default_rng(1000 + seed)
    n, p = int(rng.integers(130, 900)), int(rng.integers(2, 200))
    dens = rng.random((n, p))
    x = np.where(dens < 0.15, rng.lognormal(0, 1, size=(n, p)), 0.0)
    x[5] = 0.0
    x[7] = x[6]
    metric = ("cosine", "euclidean")[seed % 2]
    cells = dev.Cells.from_dense(torch.from_numpy(x).cuda(), metric)
    full = cells.pdist_symmetric()
    k = int(rng.integers(1, 32))
    fi, fv = dev.knn_from_d(full, k, EXCLUDE_SELF)
    for _ in range(3):
        r0 = int(rng.integers(0, n - 1))
        r1 = int(rng.integers(r0 + 1, n + 1))
        assert torch.equal(cells.pdist_stripe(r0, r1), full[r0:r1]), (seed, r0, r1)
        for prec in ("fp64", "fast"):
            si, sv = cells.knn_stripe(k, EXCLUDE_SELF, r0, r1, precision=prec)
            assert torch.equal(si, fi[r0:r1]) and torch.equal(sv, fv[r0:r1]), (seed, prec)
    cells.close()
    a = SampleDistanceMatrix(sps.csr_matrix(x), metric=metric)._d
    assert np.array_equal(a, full.cpu().numpy())
    # consumers on the same case
    sdm = SampleDistanceMatrix(x, metric=metric)
    d = sdm._d
    kk = min(k, 20)
    cut = float(np.median(np.sort(d, axis=0)[kk]) * 1.05) or 1e-3
    res = knn.RareSampleDetection(sdm).detect_rare_samples(kk, cut, 4)[0]
    assert res == korc.rare_pdist(d, kk, cut, 4)[0]
    lut = sdm.s_knn_ind_lut(kk)
    arr = np.array([lut[i] for i in range(n)])
    n_do = int(rng.integers(1, kk + 1))
    got = knn.FeatureImputation._impute_features_runner(x, lut, kk, n_do, 0.5, 2)
    want, idc, _ = korc.impute(x, arr, kk, n_do, 0.5, 2)
    np.testing.assert_array_equal(got[0], want)
    np.testing.assert_array_equal(got[1].toarray(), idc)


def test_fast_knn_adversarial_margin():
    """The exactness of the fast kNN must not rest on the errors that happen to
    be observed among the re-ranked candidates: a wide dynamic range (a tight
    cluster far from the centroid, where the FP32 accumulation of the
    tensor-core pass cancels catastrophically, next to ordinary cells) with
    near-ties at the candidate threshold.  The a-priori bound of margin_kernel
    must send the affected cells to the FP64 kernel: identical lists."""
    import torch
    from scedar_b200 import device as dev
    from scedar_b200._capi import EXCLUDE_SELF
    rng = np.random.default_rng(77)
    p = 96
    base = rng.standard_normal((3000, p))
    far = 4.0e3 + 1e-2 * rng.standard_normal((400, p))      # tight, far away
    # near-ties: groups of cells on a sphere shell around shared centres
    centres = rng.standard_normal((50, p)) * 3
    shell = rng.standard_normal((50 * 40, p))
    shell /= np.linalg.norm(shell, axis=1, keepdims=True)
    ties = np.repeat(centres, 40, axis=0) + shell * (1.0 + 1e-7 * rng.standard_normal((2000, 1)))
    x = np.vstack([base, far, ties])
    rng.shuffle(x)
    for metric in ("euclidean", "cosine", "correlation"):
        cells = dev.Cells.from_dense(dev.to_device(x), metric)
        for k in (5, 15, 40):
            fi, fd = cells.knn_stripe(k, EXCLUDE_SELF, precision="fast")
            ei, ed = cells.knn_stripe(k, EXCLUDE_SELF, precision="fp64")
            assert torch.equal(fi, ei), (metric, k)
            assert torch.equal(fd, ed), (metric, k)
        cells.close()
