"""Parity of the clustering consumers of D (condensed D
for ``HClustTree``, sub-matrix slicing of ``ind_x`` / MIRAC) against the
oracle, scipy's own ``squareform``, the reference's known-answer tests and
golden linkage matrices from the live reference.  All through the C ABI;
needs a B200: ``pytest -m gpu``."""
import warnings

import numpy as np
import pytest
import scipy.spatial.distance as spdist

from oracle import hclust_oracle as horc
from oracle import sdm_oracle as orc

pytestmark = pytest.mark.gpu

METRICS = ("euclidean", "cosine", "correlation")
LINKAGES = ("single", "complete", "average", "weighted", "ward")


@pytest.fixture(scope="module")
def b200():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    import scedar_b200
    from scedar_b200 import _capi
    _capi.load()
    return scedar_b200


def assert_same_linkage(z, want):
    assert z.shape == want.shape
    np.testing.assert_array_equal(z[:, [0, 1, 3]], want[:, [0, 1, 3]])
    np.testing.assert_allclose(z[:, 2], want[:, 2], rtol=1e-9, atol=1e-12)


# ------------------------------------------------------------------ squareform
@pytest.mark.parametrize("n", [1, 2, 3, 127, 128, 129, 1000, 2311])
def test_squareform_bit_exact(b200, n):
    import torch
    dev = b200.device
    rng = np.random.default_rng(n)
    d = spdist.squareform(spdist.pdist(rng.normal(size=(n, 3)))) if n > 1 \
        else np.zeros((n, n))
    want = spdist.squareform(d)
    t = dev.to_device(d)
    seg, flags = dev.squareform_rows(t, n)
    assert int(flags.item()) == 0
    np.testing.assert_array_equal(seg.cpu().numpy(), want)
    # stripe by stripe into one vector, shared flags, ragged stripes and a
    # padded row stride
    wide = torch.zeros((n, n + 5), dtype=torch.float64, device="cuda")
    wide[:, :n] = t
    out = torch.full((dev.condensed_len(n),), -1.0, dtype=torch.float64,
                     device="cuda")
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    cuts = sorted(set([0, n] + [c for c in (1, 77, 128, 900) if c < n]))
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        b0, b1 = dev.condensed_base(r0, n), dev.condensed_base(r1, n)
        dev.squareform_rows(wide[r0:r1, :n], n, row_begin=r0, out=out[b0:b1],
                            flags=flags)
    assert int(flags.item()) == 0
    np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_squareform_flags(b200):
    from scedar_b200 import _capi, hclust
    dev = b200.device
    d = spdist.squareform(spdist.pdist(np.arange(40.0).reshape(20, 2)))
    bad = d.copy()
    bad[4, 17] = bad[17, 4] = np.nan
    _, flags = dev.squareform_rows(dev.to_device(bad), 20)
    assert int(flags.item()) == _capi.SQF_NAN
    with pytest.raises(ValueError, match="must be symmetric"):
        hclust.condensed_from_device(dev.to_device(bad))
    bad = d.copy()
    bad[19, 19] = 1e-300          # the last row has a diagonal and no run
    _, flags = dev.squareform_rows(dev.to_device(bad), 20)
    assert int(flags.item()) == _capi.SQF_BAD_DIAG
    with pytest.raises(ValueError, match="diagonal must be zero"):
        hclust.condensed_from_device(dev.to_device(bad))
    # a NaN below the diagonal is not read: symmetry is the caller's contract
    low = d.copy()
    low[17, 4] = np.nan
    _, flags = dev.squareform_rows(dev.to_device(low), 20)
    assert int(flags.item()) == 0


# ---------------------------------------------------------------- gather_block
def test_gather_block(b200):
    import torch
    dev = b200.device
    rng = np.random.default_rng(3)
    d = rng.normal(size=(700, 1300))
    t = dev.to_device(d)
    for m, c in ((1, 1), (5, 1300), (700, 3), (333, 1111), (0, 4), (4, 0)):
        rows = rng.integers(0, 700, size=m)
        cols = rng.integers(0, 1300, size=c)
        got = dev.gather_block(t, dev.to_device(rows), dev.to_device(cols))
        assert tuple(got.shape) == (m, c)
        np.testing.assert_array_equal(got.cpu().numpy(), d[rows][:, cols])
    # a strided view of a wider matrix
    got = dev.gather_block(t[:, 100:900], dev.to_device(np.array([699, 0])),
                           dev.to_device(np.array([799, 0, 5])))
    np.testing.assert_array_equal(got.cpu().numpy(),
                                  d[:, 100:900][[699, 0]][:, [799, 0, 5]])
    for rows, cols in (([700], [0]), ([0], [1300]), ([-1], [0]), ([0], [-1])):
        with pytest.raises(IndexError):
            dev.gather_block(t, dev.to_device(np.array(rows)),
                             dev.to_device(np.array(cols)))
    assert torch.cuda.is_available()


def test_d_block_and_ind_x(b200, golden):
    g = golden("hclust")
    x, sel = g["x"], g["indx_sel"].tolist()
    want = orc.pdist(x, "cosine")
    # D only in HBM: the slice is gathered there, the host copy never forms
    sdm = b200.SampleDistanceMatrix(x, metric="cosine")
    blk = sdm.d_block(sel, [0, 149, 7])
    assert sdm._lazy_load_d is None and sdm._d_dev is not None
    np.testing.assert_allclose(blk, want[sel][:, [0, 149, 7]], rtol=1e-10,
                               atol=1e-14)
    sub = sdm.ind_x(sel, [0, 2, 4])
    assert sdm._lazy_load_d is None
    assert sub.sids == sel and sub.fids == [0, 2, 4]
    np.testing.assert_allclose(sub._d, g["indx_d"], rtol=1e-10, atol=1e-14)
    # NumPy's index rules: slices, negative indices, boolean masks
    mask = np.zeros(150, dtype=bool)
    mask[[3, 50, 51]] = True
    np.testing.assert_array_equal(sdm.d_block(slice(10, 40, 7), [-1, -150]),
                                  sdm._d_dev.cpu().numpy()[10:40:7][:, [-1, 0]])
    np.testing.assert_array_equal(sdm.d_block(mask, mask),
                                  sdm._d_dev.cpu().numpy()[mask][:, mask])
    with pytest.raises(IndexError):
        sdm.d_block([150], [0])
    # once the host copy exists it is sliced where it is, same values
    host = sdm._d
    np.testing.assert_array_equal(sdm.d_block(sel, sel), host[sel][:, sel])
    np.testing.assert_array_equal(blk, host[sel][:, [0, 149, 7]])
    # a user-supplied D and the no-pdist form (sdm.py:693-699)
    sdm2 = b200.SampleDistanceMatrix(x, d=host, metric="cosine")
    np.testing.assert_array_equal(sdm2.ind_x(sel)._d, host[sel][:, sel])
    sdm3 = b200.SampleDistanceMatrix(x, metric="cosine", use_pdist=False)
    assert sdm3.ind_x(sel)._lazy_load_d is None
    for s in (sdm, sdm2, sdm3):
        s.release_device()


# ------------------------------------------------------------------- linkage
# reference: tests/test_eda/test_dense_sdm.py:1255-1304
def test_known_tree_5x2(b200, capsys):
    H = b200.HClustTree
    sdm = b200.SampleDistanceMatrix(
        [[0, 0], [100, 100], [1, 1], [101, 101], [80, 80]],
        metric="euclidean")
    hct = H.hclust_tree(sdm.d, linkage="auto")
    assert hct.prev is None
    assert hct.left_count() == 2 and hct.right_count() == 3
    assert hct.count() == 5
    assert hct.leaf_ids() == [0, 2, 4, 1, 3]
    assert hct.left_leaf_ids() == [0, 2]
    assert hct.right_leaf_ids() == [4, 1, 3]
    assert hct.left().left().left().count() == 0
    assert hct.left().left().left().leaf_ids() == []
    assert hct.left().left().left_leaf_ids() == []
    assert hct.left().left().right().count() == 0
    # test_hclust_tree_args / test_hct_from_lkg
    lkg = H.hclust_linkage(sdm.d, linkage="auto", n_eval_rounds=-1,
                           is_euc_dist=True, verbose=True)
    assert "single" in capsys.readouterr().out
    t1, t2 = H.hct_from_lkg(lkg), H.hct_from_lkg(lkg)
    assert t1 is not t2 and t1._left is not t2._left
    # the same tree from the SDM itself (D never leaves HBM as a matrix)
    sdm2 = b200.SampleDistanceMatrix(sdm._x, metric="euclidean")
    assert H.hclust_tree(sdm2, linkage="auto").leaf_ids() == [0, 2, 4, 1, 3]
    assert sdm2._lazy_load_d is None
    # test_hclust_tree_invalid_dmat
    with pytest.raises(ValueError):
        H.hclust_tree(np.arange(5))
    with pytest.raises(ValueError):
        H.hclust_tree(np.arange(10).reshape(2, 5))
    with pytest.raises(ValueError):
        H.hclust_linkage(np.zeros((1, 1)))      # scipy: empty condensed form


# reference: tests/test_eda/test_dense_sdm.py:170-193, test_sparse_sdm.py:163-190
@pytest.mark.parametrize("x", [
    [[0, 5, 30, 10], [1, 5, 30, 10], [0, 5, 33, 10], [2, 5, 30, 7],
     [2, 5, 30, 9]],
    [[0, 0, 30, 10], [1, 2, 30, 10], [0, 3, 33, 10], [2, 4, 30, 7],
     [2, 5, 30, 9]]])
def test_known_sort_x_by_d(b200, x):
    import scipy.sparse as sps
    x1 = np.array(x, dtype=np.float64)
    x2 = x1.copy()
    order = b200.HClustTree.sort_x_by_d(x=x2.T, metric="euclidean",
                                        optimal_ordering=True)
    assert order == [2, 3, 1, 0]
    np.testing.assert_equal(x1, x2)
    order = b200.HClustTree.sort_x_by_d(x=sps.csr_matrix(x2.T),
                                        metric="euclidean",
                                        optimal_ordering=True)
    assert order == [2, 3, 1, 0]


@pytest.mark.parametrize("metric", METRICS)
def test_golden_linkages(b200, golden, metric):
    g = golden("hclust")
    H = b200.HClustTree
    x = g["x"]
    euc = metric == "euclidean"
    sdm = b200.SampleDistanceMatrix(x, metric=metric)
    # (a) from the SDM (condensed on the device), (b) from a host matrix
    # (clean-up + condensed on the device), (c) from a CUDA tensor
    sources = [sdm, orc.pdist(x, metric)]
    for src in sources:
        for lt in LINKAGES:
            assert_same_linkage(H.hclust_linkage(src, linkage=lt),
                                g["z_{}_{}".format(metric, lt)])
        assert_same_linkage(
            H.hclust_linkage(src, linkage="auto", is_euc_dist=euc),
            g["z_{}_auto".format(metric)])
        assert_same_linkage(
            H.hclust_linkage(src, linkage="auto", n_eval_rounds=11,
                             is_euc_dist=euc, optimal_ordering=True),
            g["z_{}_auto_oo".format(metric)])
    assert sdm._lazy_load_d is None
    assert_same_linkage(H.hclust_linkage(sdm._d_dev, linkage="average"),
                        g["z_{}_average".format(metric)])
    assert H.sort_x_by_d(x, metric=metric) == \
        g["leaves_{}".format(metric)].tolist()
    # the condensed vector itself against the oracle's
    np.testing.assert_allclose(sdm._condensed_d(),
                               horc.squareform(orc.pdist(x, metric)),
                               rtol=1e-10, atol=1e-14)
    sdm.release_device()


def test_golden_dirty_matrix(b200, golden):
    g = golden("hclust")
    dm = g["dirty_d"].copy()
    keep = dm.copy()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        z = b200.HClustTree.hclust_linkage(dm, linkage="complete")
    assert any("approximately symmetric" in str(i.message) for i in w)
    assert_same_linkage(z, g["z_dirty_complete"])
    np.testing.assert_array_equal(dm, keep)     # the caller's matrix is a copy
    zo, _ = horc.hclust_linkage(keep.copy(), linkage="complete")
    np.testing.assert_array_equal(z, zo)        # same clean-up, bit for bit


def test_golden_sort_features(b200, golden):
    g = golden("hclust")
    sdm = b200.SampleDistanceMatrix(g["sf_x"], metric="correlation")
    sdm.sort_features(fdist_metric="euclidean", optimal_ordering=True)
    assert sdm.fids == g["sf_fids"].tolist()
    np.testing.assert_array_equal(sdm._x, g["sf_x"][:, g["sf_fids"]])


def test_condensed_striped_equals_resident(b200, monkeypatch):
    """D too large to keep (forced): stripes of D are condensed one by one."""
    from scedar_b200 import sdm as sdm_mod, synth
    x = synth.log_counts(1500, 64, seed=8)
    a = b200.SampleDistanceMatrix(x, metric="correlation")
    whole = a._condensed_d()
    assert a._d_dev is not None
    monkeypatch.setattr(sdm_mod, "DEVICE_D_CACHE_BYTES", 0)
    monkeypatch.setattr(sdm_mod, "STRIPE_BYTES", 8 * 1500 * 300)
    b = b200.SampleDistanceMatrix(x, metric="correlation")
    monkeypatch.setattr(b, "_device_d", lambda: None)
    striped = b._condensed_d()
    assert b._d_dev is None and b._lazy_load_d is None
    np.testing.assert_array_equal(striped, whole)
    np.testing.assert_allclose(whole,
                               horc.squareform(orc.pdist(x, "correlation")),
                               rtol=1e-10, atol=1e-14)
    # a user-supplied host D goes up stripe by stripe as well
    c = b200.SampleDistanceMatrix(x, d=a._d, metric="correlation")
    monkeypatch.setattr(c, "_device_d", lambda: None)
    np.testing.assert_array_equal(c._condensed_d(), spdist.squareform(a._d))
    for s in (a, b, c):
        s.release_device()
