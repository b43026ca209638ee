"""Generate the golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

Every array written here is an output of the unmodified reference's own
``SampleDistanceMatrix`` (imported through oracle/ref_loader.py) on a seeded
input; the inputs are stored too so the fixtures are self-contained.  The
fixtures pin the oracle (tests/test_oracle_golden.py) and are compared with
the CUDA path directly (tests/test_gpu_parity.py).
"""
import json
import os
import sys
import warnings

import numpy as np
import scipy
import scipy.sparse as sps
import sklearn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from scedar_b200 import synth  # noqa: E402

SDM = load_reference()
warnings.simplefilter("ignore")


def knn_arrays(res):
    targets, distances = res
    return (np.asarray(targets).astype(np.int32),
            np.asarray(distances, dtype=np.float64))


def dump(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def dense_case(name, x, metrics, k, full):
    out = {"x": x, "k": np.int64(k)}
    for m in metrics:
        sdm = SDM(x, metric=m)
        d = sdm._d
        arg = sdm._col_argsorted_d
        srt = sdm._col_sorted_d
        if full:
            out[m + "_d"] = np.ascontiguousarray(d)
            out[m + "_arg"] = arg.astype(np.int32)
        else:
            out[m + "_d_rows"] = np.ascontiguousarray(d[:8])
            out[m + "_d_grid"] = np.ascontiguousarray(d[::25, ::25])
            out[m + "_arg_top"] = arg[:k + 2].astype(np.int32)
            out[m + "_sorted_top"] = np.ascontiguousarray(srt[:k + 2])
        lut = sdm.s_knn_ind_lut(k)
        out[m + "_lut"] = np.array([lut[i] for i in range(x.shape[0])],
                                   dtype=np.int32)
        idx, dist = knn_arrays(SDM(x, metric=m).s_knns(k))
        out[m + "_knn_idx"], out[m + "_knn_dist"] = idx, dist
        idx, dist = knn_arrays(SDM(x, metric=m, use_pdist=False).s_knns(k))
        out[m + "_raw_knn_idx"], out[m + "_raw_knn_dist"] = idx, dist
    dump(name, **out)


def main():
    # 1. the reference's own sklearn cross-check input
    #    (tests/test_eda/test_dense_sdm.py:1231-1250)
    np.random.seed(222)
    x = np.random.ranf(10000).reshape(500, -1)
    dense_case("ranf222", x, ("cosine", "correlation", "euclidean"), 15,
               full=False)

    # 2. raw integer counts: thousands of exact ties (tie-break stress)
    dense_case("counts_ties", synth.raw_counts(150, 40, seed=7),
               ("cosine", "correlation", "euclidean"), 10, full=True)

    # 3. duplicates, an all-zero row, a constant row
    x = np.array([[0, 0, 0, 0, 0],
                  [1, 1, 1, 1, 1],
                  [3, 1, 0, 2, 5],
                  [3, 1, 0, 2, 5],
                  [3, 1, 0, 2, 5],
                  [6, 2, 0, 4, 10],
                  [0, 4, 1, 1, 0],
                  [2, 0, 0, 7, 1],
                  [2, 0, 0, 7, 1],
                  [9, 9, 1, 0, 3],
                  [0, 0, 0, 0, 0],
                  [5, 5, 5, 5, 5]], dtype=np.float64)
    dense_case("edge_rows", x, ("cosine", "correlation", "euclidean"), 3,
               full=True)

    # 4. C1-like continuous log counts, correlation
    dense_case("logcounts", synth.log_counts(300, 400, seed=3005),
               ("correlation", "cosine"), 15, full=False)

    # 5. CSR input (cosine / euclidean only; correlation must raise)
    xs = synth.sparse_lognormal(150, 60, seed=11, density=0.10).tolil()
    xs[17, :] = 0          # an all-zero cell
    xs = sps.csr_matrix(xs)
    xs.eliminate_zeros()
    out = {"indptr": xs.indptr.astype(np.int64),
           "indices": xs.indices.astype(np.int32), "data": xs.data,
           "shape": np.array(xs.shape, dtype=np.int64), "k": np.int64(15)}
    for m in ("cosine", "euclidean"):
        sdm = SDM(xs, metric=m)
        out[m + "_d"] = np.ascontiguousarray(sdm._d)
        out[m + "_arg_top"] = sdm._col_argsorted_d[:17].astype(np.int32)
        idx, dist = knn_arrays(SDM(xs, metric=m).s_knns(15))
        out[m + "_knn_idx"], out[m + "_knn_dist"] = idx, dist
        idx, dist = knn_arrays(SDM(xs, metric=m, use_pdist=False).s_knns(15))
        out[m + "_raw_knn_idx"], out[m + "_raw_knn_dist"] = idx, dist
        conn = SDM(xs, metric=m).s_knn_connectivity_matrix(5)
        conn.sort_indices()
        out[m + "_conn_indptr"] = conn.indptr.astype(np.int64)
        out[m + "_conn_indices"] = conn.indices.astype(np.int32)
        out[m + "_conn_data"] = conn.data
    dump("csr_lognormal", **out)

    # 6. num_correct_dist_mat on a deliberately broken matrix
    rng = np.random.default_rng(40)
    raw = rng.normal(0.4, 0.5, size=(40, 40))
    raw[np.arange(40), np.arange(40)] = rng.normal(0, 1e-3, size=40)
    with warnings.catch_warnings(record=True) as w0:
        warnings.simplefilter("always")
        fixed = SDM.num_correct_dist_mat(raw.copy())
    with warnings.catch_warnings(record=True) as w1:
        warnings.simplefilter("always")
        fixed_ub = SDM.num_correct_dist_mat(raw.copy(), 0.7)
    sym = np.abs(raw) + np.abs(raw).T
    sym[np.arange(40), np.arange(40)] = 0
    with warnings.catch_warnings(record=True) as w2:
        warnings.simplefilter("always")
        fixed_sym = SDM.num_correct_dist_mat(sym.copy())
    dump("ncdm", raw=raw, fixed=fixed, fixed_ub=fixed_ub, sym=sym,
         fixed_sym=fixed_sym,
         n_warn=np.array([len(w0), len(w1), len(w2)], dtype=np.int64))

    with open(os.path.join(HERE, "PROVENANCE.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py",
                   "reference": "TaylorResearchLab/scedar v0.2.0 "
                                "(/root/reference, imported unmodified)",
                   "numpy": np.__version__, "scipy": scipy.__version__,
                   "sklearn": sklearn.__version__}, f, indent=1)


if __name__ == "__main__":
    main()
