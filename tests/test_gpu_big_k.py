"""No cap on k (the reference has none: sdm.py:1064-1079, detection.py:81-123):
beyond the 128 entries of the warp-list kernels the path switches to sorted D
stripes, a radix k-th selection and a block-per-cell CSR kernel.  Every route
against the oracle."""
import numpy as np
import pytest

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


@pytest.fixture(scope="module")
def data():
    from oracle import sdm_oracle as orc
    from scedar_b200 import synth
    x = synth.blobs(700, 30, seed=41, n_blobs=4)
    return x, orc.pdist(x, "euclidean")


@pytest.mark.parametrize("k", [128, 200, 650])
def test_knn_accessors_any_k(b200, data, k):
    from oracle import sdm_oracle as orc
    x, d = data
    ridx, rdist = orc.knns_precomputed(d, k)
    for use_pdist in (True, False):
        sdm = b200.SampleDistanceMatrix(x, metric="euclidean",
                                        use_pdist=use_pdist)
        idx, dist = sdm.s_knns(k)
        assert np.array_equal(np.array(idx), ridx)
        np.testing.assert_allclose(np.array(dist), rdist, rtol=1e-10,
                                   atol=1e-12)
        lut = sdm.s_knn_ind_lut(k)
        ref = orc.knn_ind_lut(d, k)
        assert all(lut[i] == ref[i] for i in range(len(x)))
        conn = sdm.s_knn_connectivity_matrix(k)
        want = orc.knn_connectivity(ridx, rdist, len(x))
        assert np.array_equal(conn.indptr, want.indptr)
        assert np.array_equal(conn.indices, want.indices)
        np.testing.assert_allclose(conn.data, want.data, rtol=1e-10)


@pytest.mark.parametrize("k", [130, 300])
def test_rare_samples_any_k(b200, data, k):
    from oracle import knn_oracle
    x, d = data
    cut = float(np.median(np.sort(d, axis=0)[k]))
    want = knn_oracle.rare_pdist(d, k, cut, 4)
    got = b200.knn.RareSampleDetection(
        b200.SampleDistanceMatrix(x, metric="euclidean"))
    res = got.detect_rare_samples(k, cut, 4)
    assert res[0] == want[0]
    assert got._res_lut[(k, cut, 4)][1] == want[1]
    want_np = knn_oracle.rare_no_pdist(x, "euclidean", k, cut, 4)
    got = b200.knn.RareSampleDetection(
        b200.SampleDistanceMatrix(x, metric="euclidean", use_pdist=False))
    res = got.detect_rare_samples(k, cut, 4)
    assert res[0] == want_np[0]


def test_masked_kth_radix_against_list_kernel(b200, data):
    """The radix selection and the warp-list kernel agree where both apply
    (rank < 128), with dead rows / columns and exact ties."""
    import torch
    from scedar_b200 import device as dev, _capi
    import ctypes
    x, d = data
    dd = np.round(d, 1)                      # many exact ties
    dt = dev.to_device(dd)
    rng = np.random.default_rng(2)
    alive = (rng.random(len(x)) > 0.3).astype(np.uint8)
    at = dev.to_device(alive)
    for rank in (0, 5, 127):
        a = dev.masked_kth(dt, at, rank).cpu().numpy()
        live = np.nonzero(alive)[0]
        sub = np.sort(dd[np.ix_(live, live)], axis=1)
        want = np.full(len(x), np.inf)
        want[live] = sub[:, rank]
        assert np.array_equal(a, want)
    for rank in (128, 300, len(np.nonzero(alive)[0]) - 1):
        a = dev.masked_kth(dt, at, rank).cpu().numpy()
        live = np.nonzero(alive)[0]
        sub = np.sort(dd[np.ix_(live, live)], axis=1)
        want = np.full(len(x), np.inf)
        want[live] = sub[:, rank]
        assert np.array_equal(a, want)
    # more than the number of live columns: +inf like the list kernel
    a = dev.masked_kth(dt, at, int(alive.sum()) + 3).cpu().numpy()
    assert np.isinf(a).all()
