"""ctypes binding of the C ABI in ``include/scedar_b200.h``.

The shared library is built in-tree (``scedar_b200/csrc/libscedar_b200.so``,
``__graft_entry__.build()``).  There is no fallback: if the library is missing
the import of any compute entry point raises.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int32, c_int64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libscedar_b200.so")

METRIC_IDS = {"cosine": 0, "correlation": 1, "euclidean": 2}
PRECISION_IDS = {"fp64": 0, "fast": 1, "fp64_i8": 2, "fp64_i8s6": 3}
EXCLUDE_FIRST, EXCLUDE_SELF = 0, 1
OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = 0, 1, 2, 3, 4
NCDM_BAD_DIAG, NCDM_ASYMMETRIC = 1, 2
SQF_NAN, SQF_BAD_DIAG = 1, 2
GATHER_BAD_INDEX = 1

# name -> (restype, argtypes); every symbol the header declares
SIGNATURES = {
    "scedar_abi_version": (c_int, []),
    "scedar_last_error": (c_char_p, []),
    "scedar_max_k": (c_int, []),
    "scedar_launch_count": (ctypes.c_longlong, []),
    "scedar_cells_from_dense": (c_int, [c_void_p, c_int64, c_int64, c_int64,
                                        c_int, c_void_p, POINTER(c_void_p)]),
    "scedar_cells_from_csr": (c_int, [c_void_p, c_int, c_void_p, c_void_p,
                                      c_int64, c_int64, c_int, c_void_p,
                                      POINTER(c_void_p)]),
    "scedar_cells_alloc": (c_int, [c_int64, c_int64, c_int, c_void_p,
                                   POINTER(c_void_p)]),
    "scedar_cells_fill_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64,
                                       c_int64, c_void_p]),
    "scedar_cells_fill_rows_ex": (c_int, [c_void_p, c_void_p, c_int64, c_int64,
                                          c_int64, c_int, c_void_p]),
    "scedar_cells_stat_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
    "scedar_pdist_lower_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p,
                                        c_int64, c_void_p]),
    "scedar_pdist_lower_rows_prec": (c_int, [c_void_p, c_int, c_int64, c_int64,
                                             c_void_p, c_int64, c_void_p]),
    "scedar_cells_col_means": (c_int, [c_void_p, c_void_p, c_void_p]),
    "scedar_cells_shift_cols": (c_int, [c_void_p, c_void_p, c_void_p]),
    "scedar_cells_transposed": (c_int, [c_void_p, c_void_p, c_void_p,
                                        POINTER(c_void_p)]),
    "scedar_gram_raw": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "scedar_gram_raw_cross": (c_int, [c_void_p, c_void_p, c_void_p, c_int64,
                                      c_void_p]),
    "scedar_cells_free": (c_int, [c_void_p]),
    "scedar_cells_n": (c_int64, [c_void_p]),
    "scedar_cells_p": (c_int64, [c_void_p]),
    "scedar_cells_metric": (c_int, [c_void_p]),
    "scedar_pdist_stripe": (c_int, [c_void_p, c_int, c_int64, c_int64,
                                    c_void_p, c_int64, c_void_p]),
    "scedar_pdist_stripe_p2p": (c_int, [c_void_p, c_int, c_int, c_int,
                                        POINTER(c_int64), POINTER(c_void_p),
                                        c_int64, c_void_p]),
    "scedar_pdist_p2p_rows": (c_int, [c_void_p, c_int, c_int, c_int,
                                      POINTER(c_int64), POINTER(c_void_p),
                                      c_int64, c_int64, c_int64, c_int,
                                      c_void_p]),
    "scedar_pdist_p2p_pieces": (c_int, [c_void_p, c_int, c_int, c_int,
                                        POINTER(c_int64), POINTER(c_void_p),
                                        c_int64, POINTER(ctypes.c_int32), c_int,
                                        c_int, c_int, c_void_p]),
    "scedar_stripe_alloc": (c_int, [ctypes.c_size_t, POINTER(c_void_p)]),
    "scedar_stripe_free": (c_int, [c_void_p]),
    "scedar_stripe_export": (c_int, [c_void_p, c_void_p]),
    "scedar_stripe_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "scedar_stripe_close": (c_int, [c_void_p]),
    "scedar_pdist_symmetric": (c_int, [c_void_p, c_int, c_void_p, c_int64,
                                       c_void_p]),
    "scedar_knn_stripe": (c_int, [c_void_p, c_int, c_int, c_int, c_int64,
                                  c_int64, c_void_p, c_void_p, c_void_p]),
    "scedar_knn_from_d": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64,
                                  c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "scedar_cells_set_blockmin": (c_int, [c_void_p, c_void_p, c_int, c_int64]),
    "scedar_knn_from_d_blockmin": (c_int, [c_void_p, c_void_p, c_int64, c_void_p,
                                           c_int64, c_int64, c_int64, c_int, c_int,
                                           c_void_p, c_void_p, c_void_p]),
    "scedar_col_sort": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                c_void_p, c_void_p]),
    "scedar_num_correct_dist_mat": (c_int, [c_void_p, c_int64, c_int64, c_int,
                                            c_double, c_void_p, c_void_p]),
    "scedar_knn_connectivity_csr": (c_int, [c_void_p, c_void_p, c_int64, c_int,
                                            c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    "scedar_knn_affinity": (c_int, [c_void_p, c_int64, c_double, c_void_p,
                                    c_void_p]),
    "scedar_masked_kth": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64,
                                  c_void_p, c_int, c_void_p, c_void_p]),
    "scedar_rare_update": (c_int, [c_void_p, c_int64, c_int64, c_double, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "scedar_masked_max": (c_int, [c_void_p, c_int64, c_int64, c_void_p,
                                  c_void_p, c_void_p]),
    "scedar_compact_alive": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "scedar_gather_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p,
                                   c_int64, c_void_p, c_int64, c_void_p]),
    "scedar_impute_iter": (c_int, [c_void_p, c_int64, c_int64, c_int64,
                                   c_void_p, c_int, c_int, c_double, c_int,
                                   c_int64, c_int64, c_void_p, c_int64,
                                   c_void_p, c_int64, c_void_p, c_void_p]),
    "scedar_squareform_rows": (c_int, [c_void_p, c_int64, c_int64, c_int64,
                                       c_int64, c_void_p, c_void_p, c_void_p]),
    "scedar_d_gather_block": (c_int, [c_void_p, c_int64, c_int64, c_int64,
                                      c_void_p, c_int64, c_void_p, c_int64,
                                      c_void_p, c_int64, c_void_p, c_void_p]),
    "scedar_pdist_dense": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int,
                                   c_int64, c_int64, c_void_p, c_void_p]),
    "scedar_pdist_csr": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64,
                                 c_int64, c_int, c_int, c_int64, c_int64,
                                 c_void_p, c_void_p]),
    "scedar_knn_dense": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int,
                                 c_int, c_int, c_int64, c_int64, c_void_p,
                                 c_void_p, c_void_p]),
    "scedar_pdist_dense_host": (c_int, [c_void_p, c_int64, c_int64, c_int,
                                        c_int, c_void_p, c_int]),
    "scedar_knn_dense_host": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int,
                                      c_int, c_int, c_void_p, c_void_p,
                                      c_int]),
}

_lib = None


class ScedarError(RuntimeError):
    """A non-argument failure inside the CUDA library."""


def load():
    """Load the shared library once; raise loudly if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "scedar_b200: {} is missing -- build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)".format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.scedar_abi_version() != 2:
            raise ImportError("scedar_b200: ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    """Turn a return code into the reference's error convention: argument
    errors are ``ValueError``, the rest RuntimeError."""
    if rc == OK:
        return
    msg = load().scedar_last_error().decode("utf-8", "replace")
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise ScedarError(msg)
