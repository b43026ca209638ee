"""CPU: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scedar_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scedar_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from scedar_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _capi.load()


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name


def test_binding_table_matches_header(lib):
    from scedar_b200 import _capi
    assert sorted(_capi.SIGNATURES) == header_functions()


def test_version_and_error_string(lib):
    assert lib.scedar_abi_version() == 2
    assert lib.scedar_max_k() == 127
    assert isinstance(lib.scedar_last_error(), bytes)


def test_argument_errors_need_no_gpu(lib):
    """Argument validation happens before any CUDA call."""
    from scedar_b200 import _capi
    h = ctypes.c_void_p()
    rc = lib.scedar_cells_from_dense(None, 4, 4, 4, 7, None, ctypes.byref(h))
    assert rc == _capi.ERR_ARG
    with pytest.raises(ValueError):
        _capi.check(rc)
    rc = lib.scedar_cells_from_csr(None, 1, None, None, 4, 4, 1, None,
                                   ctypes.byref(h))
    assert rc == _capi.ERR_ARG and b"sparse" in lib.scedar_last_error()
    rc = lib.scedar_knn_from_d(None, 10, 10, 0, 10, 10, 0, None, None, None)
    assert rc == _capi.ERR_ARG      # k > n-1
    rc = lib.scedar_num_correct_dist_mat(None, 2, 3, 0, 0.0, None, None)
    assert rc == _capi.ERR_ARG
    assert lib.scedar_pdist_stripe(None, 0, 0, 0, None, 0, None) == \
        _capi.ERR_ARG


def test_argument_errors_of_every_family(lib):
    """Each family of entry points rejects bad arguments with ERR_ARG and a
    message, before any CUDA call (so this runs without a device)."""
    from scedar_b200 import _capi
    E = _capi.ERR_ARG
    h = ctypes.c_void_p()
    one = ctypes.c_void_p(16)            # any non-NULL pointer: never read
    calls = {
        "cells_from_dense ldx < p": lambda: lib.scedar_cells_from_dense(
            one, 4, 4, 3, 0, None, ctypes.byref(h)),
        "cells_from_dense n < 0": lambda: lib.scedar_cells_from_dense(
            one, -1, 4, 4, 0, None, ctypes.byref(h)),
        "cells_alloc out NULL": lambda: lib.scedar_cells_alloc(
            4, 4, 0, None, None),
        "cells_fill_rows NULL handle": lambda: lib.scedar_cells_fill_rows(
            None, one, 4, 0, 4, None),
        "pdist_lower_rows NULL handle": lambda: lib.scedar_pdist_lower_rows(
            None, 0, 4, one, 4, None),
        "pdist_symmetric NULL handle": lambda: lib.scedar_pdist_symmetric(
            None, 0, one, 4, None),
        "pdist_stripe_p2p NULL handle": lambda: lib.scedar_pdist_stripe_p2p(
            None, 0, 0, 1, None, None, 4, None),
        "knn_stripe NULL handle": lambda: lib.scedar_knn_stripe(
            None, 0, 1, 0, 0, 1, one, one, None),
        "knn_from_d bad exclude": lambda: lib.scedar_knn_from_d(
            one, 10, 10, 0, 10, 3, 5, one, one, None),
        "knn_from_d bad rows": lambda: lib.scedar_knn_from_d(
            one, 10, 10, 4, 2, 3, 0, one, one, None),
        "col_sort ldd < n": lambda: lib.scedar_col_sort(
            one, 3, 4, 4, one, one, None),
        "ncdm ldd < n": lambda: lib.scedar_num_correct_dist_mat(
            one, 2, 3, 0, 0.0, one, None),
        "conn_csr k < 0": lambda: lib.scedar_knn_connectivity_csr(
            one, one, 4, -1, one, one, one, None),
        "affinity nnz < 0": lambda: lib.scedar_knn_affinity(
            one, -1, 1.0, one, None),
        "masked_kth bad rows": lambda: lib.scedar_masked_kth(
            one, 8, 8, 0, 9, one, 1, one, None),
        "masked_kth rank < 0": lambda: lib.scedar_masked_kth(
            one, 8, 8, 0, 8, one, -1, one, None),
        "gather_rows ldx < p": lambda: lib.scedar_gather_rows(
            one, 3, 4, one, 2, one, 4, None),
        "impute_iter k >= n": lambda: lib.scedar_impute_iter(
            one, 4, 4, 4, one, 4, 1, 1.0, 1, 0, 4, one, 4, one, 4, None,
            None),
        "impute_iter n_do > k": lambda: lib.scedar_impute_iter(
            one, 4, 8, 4, one, 3, 4, 1.0, 1, 0, 8, one, 4, one, 4, None,
            None),
        "impute_iter aliasing": lambda: lib.scedar_impute_iter(
            one, 4, 8, 4, one, 3, 2, 1.0, 1, 0, 8, one, 4, one, 4, None,
            None),
        "squareform bad rows": lambda: lib.scedar_squareform_rows(
            one, 8, 8, 3, 9, one, one, None),
        "squareform ldd < n": lambda: lib.scedar_squareform_rows(
            one, 7, 8, 0, 8, one, one, None),
        "squareform NULL flags": lambda: lib.scedar_squareform_rows(
            one, 8, 8, 0, 8, one, None, None),
        "squareform NULL out": lambda: lib.scedar_squareform_rows(
            one, 8, 8, 0, 8, None, one, None),
        "gather_block ldo < c": lambda: lib.scedar_d_gather_block(
            one, 8, 8, 8, one, 2, one, 3, one, 2, one, None),
        "gather_block NULL flags": lambda: lib.scedar_d_gather_block(
            one, 8, 8, 8, one, 2, one, 3, one, 3, None, None),
        "gather_block NULL sel": lambda: lib.scedar_d_gather_block(
            one, 8, 8, 8, None, 2, one, 3, one, 3, one, None),
        "pdist_dense bad metric": lambda: lib.scedar_pdist_dense(
            one, 4, 4, 9, 0, 0, 4, one, None),
        "knn_dense bad precision": lambda: lib.scedar_knn_dense(
            one, 4, 4, 0, 7, 1, 0, 0, 4, one, one, None),
        "pdist_dense bad rows": lambda: lib.scedar_pdist_dense(
            one, 4, 4, 0, 0, 2, 5, one, None),
        "pdist_csr bad precision": lambda: lib.scedar_pdist_csr(
            one, 1, one, one, 4, 4, 0, 9, 0, 4, one, None),
        "knn_dense k > n-1": lambda: lib.scedar_knn_dense(
            one, 4, 4, 0, 0, 4, 0, 0, 4, one, one, None),
        "knn_dense bad exclude": lambda: lib.scedar_knn_dense(
            one, 4, 4, 0, 0, 1, 2, 0, 4, one, one, None),
        "pdist_dense_host NULL out": lambda: lib.scedar_pdist_dense_host(
            one, 4, 4, 0, 0, None, 0),
        "knn_dense_host bad exclude": lambda: lib.scedar_knn_dense_host(
            one, 4, 4, 0, 0, 1, 9, one, one, 0),
        "knn_dense_host k > n-1": lambda: lib.scedar_knn_dense_host(
            one, 4, 4, 0, 0, 4, 0, one, one, 0),
        "stripe_export NULL": lambda: lib.scedar_stripe_export(None, None),
        "stripe_open NULL": lambda: lib.scedar_stripe_open(None, None),
    }
    for what, call in calls.items():
        rc = call()
        assert rc == E, (what, rc, lib.scedar_last_error())
        assert lib.scedar_last_error(), what
    # empty work is accepted without touching the device
    assert lib.scedar_squareform_rows(None, 8, 8, 3, 3, None, one, None) == 0
    assert lib.scedar_d_gather_block(None, 8, 8, 8, None, 0, None, 3, None, 3,
                                     one, None) == 0
    assert lib.scedar_knn_from_d(None, 10, 10, 0, 10, 0, 0, None, None,
                                 None) == 0
    assert lib.scedar_cells_free(None) == 0 and lib.scedar_cells_n(None) == -1


def test_no_oracle_in_product():
    """The product package never imports the oracle or a CPU fallback."""
    pkg = os.path.join(ROOT, "scedar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and \
                    "from oracle" not in text, f
                assert "sklearn" not in text.replace(
                    "scikit-learn", "").replace("sklearn's", "").replace(
                    "sklearn ", "").replace("sklearn:", "") or \
                    "import sklearn" not in text, f
