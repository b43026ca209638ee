"""Timing of the BASELINE.json configs other than the bench workload, through
the public drop-in API on one GPU (developer evidence for profiles/, not the
contract bench).  One JSON line per config; times are host wall clock around
a synchronised region (host <-> device copies included where the API does
them), kernel-only figures come from CUDA events.

    python tools/configs_bench.py [c1] [c2] [c3] [c5]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from scedar_b200 import SampleDistanceMatrix, device as dev, knn, synth
from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF

HBM_PEAK = 6558.4  # GB/s, MEASURED_PEAKS.json


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3


def ev(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def c1():
    n, p, k = 3005, 19972, 15
    x = synth.log_counts(n, p, seed=3005)
    SampleDistanceMatrix(x[:256], metric="correlation")._d      # warm-up
    sdm = SampleDistanceMatrix(x, metric="correlation")
    _, t_d = wall(lambda: sdm._d)
    _, t_lut = wall(lambda: sdm.s_knn_ind_lut(k))
    cells = dev.Cells.from_dense(dev.to_device(x), "correlation")
    d = torch.empty((n, n), dtype=torch.float64, device="cuda")
    k_dmma = ev(lambda: cells.pdist_symmetric(out=d, precision="fp64"))
    k_i8 = ev(lambda: cells.pdist_symmetric(out=d, precision="fp64_i8"))
    cells.close()
    return {"config": "C1 3005x19972 correlation + s_knn_ind_lut(15)",
            "d_ms_incl_h2d_d2h": round(t_d, 2), "lut_ms": round(t_lut, 2),
            "d_kernel_ms_dmma": round(k_dmma, 3), "d_kernel_ms_int8_split": round(k_i8, 3),
            "gram_kernel_of_the_api_call": sdm._gram_precision()}


def c2():
    n, p, k = 20000, 20000, 15
    xs = synth.sparse_lognormal(n, p, seed=20000, density=0.10)
    sdm = SampleDistanceMatrix(xs, metric="cosine")
    _, t_knn = wall(lambda: sdm.s_knns(k))
    conn, t_conn = wall(lambda: sdm.s_knn_connectivity_matrix(k))
    _, t_aff = wall(lambda: SampleDistanceMatrix.knn_conn_mat_to_aff_edges(conn))
    cells = sdm._cells("cosine")
    d = sdm._d_dev
    k_dmma = ev(lambda: cells.pdist_symmetric(out=d, precision="fp64"), 2)
    k_i8 = ev(lambda: cells.pdist_symmetric(out=d, precision="fp64_i8"), 2)
    return {"config": "C2 CSR 20000x20000 (10% nnz) cosine + kNN graph k=15",
            "nnz": int(xs.nnz), "first_s_knns_ms_incl_prep_and_D": round(t_knn, 2),
            "connectivity_ms": round(t_conn, 2), "affinity_edges_ms": round(t_aff, 2),
            "d_kernel_ms_dmma": round(k_dmma, 2), "d_kernel_ms_int8_split": round(k_i8, 2),
            "gram_kernel_of_the_api_call": sdm._gram_precision()}


def c3():
    n, p, k = 68579, 2000, 15
    x = synth.blobs(n, p, seed=68579)
    xt = dev.to_device(x)
    out = {"config": "C3 68579x2000 euclidean: kNN + RareSampleDetection + "
                     "FeatureImputation"}
    cells = dev.Cells.from_dense(xt, "euclidean")
    d = cells.pdist_symmetric()
    out["d_symmetric_ms"] = round(ev(lambda: cells.pdist_symmetric(out=d), 1), 2)
    out["d_symmetric_int8_split_ms"] = round(
        ev(lambda: cells.pdist_symmetric(out=d, precision="fp64_i8"), 1), 2)
    cells.pdist_symmetric(out=d)
    idx, dist = dev.knn_from_d(d, k, EXCLUDE_FIRST)
    alive = torch.ones(n, dtype=torch.uint8, device="cuda")
    alive[::7] = 0
    t = ev(lambda: dev.masked_kth(d, alive, k))
    live = int(alive.sum().item())
    out["masked_kth_ms"] = round(t, 3)
    out["masked_kth_GBs"] = round(8.0 * live * n / t * 1e-6, 1)
    out["masked_kth_frac_of_hbm_peak"] = round(out["masked_kth_GBs"] / HBM_PEAK, 3)
    # the whole detection through the drop-in (D resident)
    sdm = SampleDistanceMatrix(x, metric="euclidean")
    sdm._cells_lut["euclidean"] = cells
    sdm._d_dev = d
    cut = float(dist[:, -1].median().item())
    res, t_rare = wall(lambda: knn.RareSampleDetection(sdm).detect_rare_samples(
        k, cut, 5))
    out["rare_detect_5iter_ms"] = round(t_rare, 2)
    out["rare_kept"] = len(res[0])
    # imputation kernel on a thresholded copy
    xi = torch.where(xt > 6.0, xt, torch.zeros_like(xt))
    nxt = torch.empty_like(xi)
    idc = torch.zeros((n, p), dtype=torch.int32, device="cuda")
    st = torch.zeros(2, dtype=torch.int64, device="cuda")
    t = ev(lambda: dev.impute_iter(xi, idx, 5, 6.0, 1, nxt, idc, st))
    out["impute_iter_ms"] = round(t, 3)
    alg = (k + 1) * 8.0 * n * p + 8.0 * n * p
    out["impute_iter_alg_GBs"] = round(alg / t * 1e-6, 1)
    out["impute_stats"] = st.tolist()
    t = ev(lambda: dev.knn_connectivity_csr(idx, dist))
    out["conn_csr_ms"] = round(t, 3)
    # no-pdist detection: gather + fused kNN per iteration (fast precision)
    nd = SampleDistanceMatrix(x, metric="euclidean", use_pdist=False)
    nd.precision = "fast"
    res2, t_rare2 = wall(lambda: knn.RareSampleDetection(nd).detect_rare_samples(
        k, cut, 5))
    out["rare_detect_nopdist_fast_ms"] = round(t_rare2, 2)
    out["rare_nopdist_equals_pdist"] = res2[0] == res[0]
    cells.close()
    return out


def c5():
    n, p, k = 1000000, 50, 30
    x = synth.pcs(n, p, seed=1000000)
    xt = dev.to_device(x)
    cells = dev.Cells.from_dense(xt, "euclidean")
    t_fast = ev(lambda: cells.knn_stripe(k, EXCLUDE_SELF, precision="fast"), 2)
    out = {"config": "C5 1,000,000 x 50 euclidean kNN k=30, no D, one GPU",
           "fast_knn_ms": round(t_fast, 2),
           "pairs_per_s_fast": round(float(n) * n / t_fast * 1e3, 0)}
    cells.close()
    nd = SampleDistanceMatrix(x, metric="euclidean", use_pdist=False)
    nd.precision = "fast"
    _, t_api = wall(lambda: nd.s_knns(k))
    out["s_knns_api_ms_incl_h2d_d2h_lists"] = round(t_api, 2)
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3", "c5"]
    for w in which:
        print(json.dumps(globals()[w]()), flush=True)
        torch.cuda.empty_cache()
