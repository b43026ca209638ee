"""The one-shot device-pointer entry points SURVEY.md 8b lists
(``scedar_pdist_dense``, ``scedar_pdist_csr``, ``scedar_knn_dense``: prepare +
run + free in one call) against the handle-based path and the oracle.

Needs a B200: ``pytest -m gpu``."""
import ctypes

import numpy as np
import pytest

from oracle import sdm_oracle as orc

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


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def test_oneshot_dense_and_knn(b200):
    import torch
    from scedar_b200 import _capi, device as dev, synth
    lib = _capi.load()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n, p, k = 900, 77, 12
    xh = synth.log_counts(n, p, seed=17)
    x = dev.to_device(xh)
    for metric, mid in (("cosine", 0), ("correlation", 1), ("euclidean", 2)):
        cells = dev.Cells.from_dense(x, metric)
        want = cells.pdist_stripe(130, 700)
        got = torch.full((570, n), -1.0, dtype=torch.float64, device="cuda")
        _capi.check(lib.scedar_pdist_dense(_p(x), n, p, mid, 0, 130, 700,
                                           _p(got), stream))
        assert torch.equal(got, want), metric
        np.testing.assert_allclose(got.cpu().numpy(),
                                   orc.pdist(xh, metric)[130:700],
                                   rtol=1e-9, atol=1e-12)
        for ex in (0, 1):
            wi, wd = cells.knn_stripe(k, ex, 64, 900)
            gi = torch.empty((836, k), dtype=torch.int64, device="cuda")
            gd = torch.empty((836, k), dtype=torch.float64, device="cuda")
            _capi.check(lib.scedar_knn_dense(_p(x), n, p, mid, 0, k, ex, 64,
                                             900, _p(gi), _p(gd), stream))
            assert torch.equal(gi, wi) and torch.equal(gd, wd), (metric, ex)
        cells.close()
    # the tensor-core precision through the same entry point: exact lists
    gi2 = torch.empty((836, k), dtype=torch.int64, device="cuda")
    gd2 = torch.empty((836, k), dtype=torch.float64, device="cuda")
    _capi.check(lib.scedar_knn_dense(_p(x), n, p, 2, 1, k, 1, 64, 900,
                                     _p(gi2), _p(gd2), stream))
    assert torch.equal(gi2, gi) and torch.equal(gd2, gd)
    # empty row range: nothing written, no error
    _capi.check(lib.scedar_pdist_dense(_p(x), n, p, 0, 0, 5, 5, None, stream))
    with pytest.raises(ValueError):
        _capi.check(lib.scedar_knn_dense(_p(x), n, p, 0, 0, n, 1, 0, n,
                                         _p(gi), _p(gd), stream))


def test_oneshot_csr(b200):
    import torch
    from scedar_b200 import _capi, device as dev, synth
    lib = _capi.load()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n, p = 400, 300
    xs = synth.sparse_lognormal(n, p, seed=5, density=0.1)
    ip64 = dev.to_device(xs.indptr.astype(np.int64))
    ip32 = dev.to_device(xs.indptr.astype(np.int32))
    ix = dev.to_device(xs.indices.astype(np.int32))
    da = dev.to_device(xs.data.astype(np.float64))
    for metric, mid in (("cosine", 0), ("euclidean", 2)):
        cells = dev.Cells.from_csr(ip64, ix, da, (n, p), metric)
        want = cells.pdist_stripe(0, n)
        for ip, is64 in ((ip64, 1), (ip32, 0)):
            got = torch.empty((n, n), dtype=torch.float64, device="cuda")
            _capi.check(lib.scedar_pdist_csr(_p(ip), is64, _p(ix), _p(da), n,
                                             p, mid, 0, 0, n, _p(got),
                                             stream))
            assert torch.equal(got, want), (metric, is64)
        np.testing.assert_allclose(want.cpu().numpy(), orc.pdist(xs, metric),
                                   rtol=1e-9, atol=1e-12)
        cells.close()
    got = torch.empty((n, n), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):          # sdm.py:1196-1204
        _capi.check(lib.scedar_pdist_csr(_p(ip64), 1, _p(ix), _p(da), n, p, 1,
                                         0, 0, n, _p(got), stream))
