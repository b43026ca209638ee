"""CPU: the host logic above the C ABI (``sdm.py``, ``hclust.py``, ``knn.py``:
caching, routing by k / size / precision, NumPy index rules, list and dict
building, warnings and errors) with the device operators replaced by the
NumPy / oracle stand-ins of ``tests/fake_device.py``.

The test bodies are the ``-m gpu`` tests that go through the public API only
-- the very same functions, re-collected here with a CPU-backed ``b200``
fixture -- so a change to the host code is checked here, without a GPU,
against the same reference known answers and live-reference goldens.  What
this cannot see is the kernels; that is what ``-m gpu`` is for.
"""
import numpy as np
import pytest

import fake_device
import test_gpu_consumers as gpu_consumers
import test_gpu_hclust as gpu_hclust
import test_gpu_parity as gpu_parity


@pytest.fixture
def b200(monkeypatch):
    fake_device.install(monkeypatch)
    import torch
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (0, 0))
    import scedar_b200
    return scedar_b200


def _adopt(module, names):
    """Re-export test functions (and their parametrisation) under this
    module: the module-level ``gpu`` mark of their home does not follow."""
    for name in names:
        globals()[name] = getattr(module, name)


_adopt(gpu_parity, [
    "test_pdist_golden", "test_knn_lut_golden", "test_s_knns_golden",
    "test_precomputed_d_paths", "test_col_sort_golden", "test_csr_golden",
    "test_num_correct_dist_mat_golden", "test_reference_known_answers",
    "test_duplicates_and_zero_rows", "test_empty_and_lazy",
    "test_ragged_shapes"])
_adopt(gpu_hclust, [
    "test_d_block_and_ind_x", "test_known_tree_5x2", "test_known_sort_x_by_d",
    "test_golden_linkages", "test_golden_dirty_matrix",
    "test_golden_sort_features", "test_condensed_striped_equals_resident"])
_adopt(gpu_consumers, [
    "test_connectivity_csr", "test_connectivity_reference_vector",
    "test_affinity_edges", "test_rare_reference_vectors", "test_rare_golden",
    "test_rare_larger_vs_oracle", "test_impute_reference_vectors",
    "test_impute_golden", "test_impute_larger_vs_oracle"])


# ---- routing decisions that only the host code makes -------------------------
def test_routing_by_k_and_precision(b200, monkeypatch):
    from scedar_b200 import sdm as sdm_mod, synth
    calls = fake_device.FakeCells.calls
    x = synth.log_counts(300, 40, seed=2)
    s = b200.SampleDistanceMatrix(x, metric="correlation")
    lut = s.s_knn_ind_lut(5)
    # D kept in HBM, computed once with the symmetric kernel, top-k from it
    assert [c[0] for c in calls] == ["pdist_symmetric"]
    assert s._lazy_load_d is None and s._lazy_load_col_argsorted_d is None
    assert s._d is s._d and s._lazy_load_d is not None      # memoised
    assert [c[0] for c in calls] == ["pdist_symmetric"]
    # k beyond the streaming top-k: the full column sort serves it
    big = s.s_knn_ind_lut(sdm_mod.MAX_TOPK + 3)
    assert s._lazy_load_col_argsorted_d is not None
    assert all(big[i][:5] == lut[i] for i in range(300))
    # once the argsort exists it is re-used for small k as well
    assert s.s_knn_ind_lut(5) == lut
    # precision = fast: exact neighbours come from the candidate kernel, never
    # from the approximate matrix
    del calls[:]
    f = b200.SampleDistanceMatrix(x, metric="correlation")
    f.precision = "fast"
    assert f.s_knn_ind_lut(5) == lut
    assert calls == [("knn_stripe", 5, 0, "fast")]
    # ... unless the matrix is the user's own
    del calls[:]
    u = b200.SampleDistanceMatrix(x, d=s._d, metric="correlation")
    u.precision = "fast"
    assert u.s_knn_ind_lut(5) == lut and calls == []
    # D too large to keep: fused kNN without D for small k, stripes otherwise
    monkeypatch.setattr(sdm_mod, "DEVICE_D_CACHE_BYTES", 0)
    monkeypatch.setattr(sdm_mod, "STRIPE_BYTES", 8 * 300 * 128)
    del calls[:]
    g = b200.SampleDistanceMatrix(x, metric="correlation")
    assert g.s_knn_ind_lut(5) == lut
    assert calls == [("knn_stripe", 5, 0, "fp64")] and g._d_dev is None
    del calls[:]
    assert g.s_knn_ind_lut(40) == {i: big[i][:40] for i in range(300)}
    assert calls == [("pdist_stripe", 0, 128), ("pdist_stripe", 128, 256),
                     ("pdist_stripe", 256, 300)]
    np.testing.assert_array_equal(g._d, s._d)                # streamed D


def test_no_pdist_lut_reuses_last_knns(b200):
    from scedar_b200 import synth
    calls = fake_device.FakeCells.calls
    x = synth.blobs(60, 8, seed=1)
    s = b200.SampleDistanceMatrix(x, metric="euclidean", use_pdist=False)
    l3 = s.s_knn_ind_lut(3)              # computes max(k, 10) neighbours
    assert s._last_k == 10 and calls == [("knn_stripe", 10, 1, "fp64")]
    l7 = s.s_knn_ind_lut(7)              # served from _last_knns
    assert len(calls) == 1 and all(l7[i][:3] == l3[i] for i in range(60))
    s.s_knn_ind_lut(12)
    assert s._last_k == 12 and len(calls) == 2
    t, d = s.s_knns(4)
    assert s._last_k == 4 and len(t) == 60 and t[0].dtype == np.int64
    assert d[0].dtype == np.float64 and np.all(np.diff(d[5]) >= 0)
