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
    assert lib.scedar_abi_version() == 1
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
