"""Golden vectors for the consumers behind the path: kNN
connectivity / affinity edges, RareSampleDetection, FeatureImputation --
outputs of the LIVE reference (scedar.knn, scedar.eda) on seeded inputs.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_consumers.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from scedar_b200 import synth  # noqa: E402

SDM = load_reference()
import scedar.knn as ref_knn  # noqa: E402  (importable once the stubs are in)

warnings.simplefilter("ignore")


def last_kept(progress, n):
    """progress_list -> per-sample index of the last list that holds it."""
    lk = np.full(n, -1, dtype=np.int64)
    for i, kept in enumerate(progress):
        lk[np.asarray(kept, dtype=np.int64)] = i
    return lk


def main():
    out = {}
    # ---- rare samples: blobs + uniform outliers, euclidean ------------------
    x = synth.blobs(240, 30, seed=5, n_blobs=4, outlier_frac=0.05)
    out["rare_x"] = x
    sdm = SDM(x, metric="euclidean")
    srt = np.sort(sdm._d, axis=0)
    cut = float(np.median(srt[5, :]) * 1.15)
    params = [(5, cut, 4), (9, cut * 1.1, 3), (3, cut * 0.9, 6)]
    out["rare_params"] = np.array(params, dtype=np.float64)
    rsd = ref_knn.RareSampleDetection(sdm)
    rsd.detect_rare_samples([p[0] for p in params], [p[1] for p in params],
                            [p[2] for p in params])
    for j, (k, c, it) in enumerate(params):
        res, prog = rsd._res_lut[(k, float(c), it)]
        out["rare_pdist_lk{}".format(j)] = last_kept(prog, x.shape[0])
    rsd = ref_knn.RareSampleDetection(SDM(x, metric="euclidean",
                                          use_pdist=False))
    rsd.detect_rare_samples([p[0] for p in params], [p[1] for p in params],
                            [p[2] for p in params])
    for j, (k, c, it) in enumerate(params):
        res, prog = rsd._res_lut[(k, float(c), it)]
        out["rare_nopdist_lk{}".format(j)] = last_kept(prog, x.shape[0])
    # cosine on the same cells (distances in [0, 2])
    sdm = SDM(x, metric="cosine")
    srt = np.sort(sdm._d, axis=0)
    ccut = float(np.median(srt[6, :]) * 1.2)
    out["rare_cos_params"] = np.array([6, ccut, 5], dtype=np.float64)
    rsd = ref_knn.RareSampleDetection(sdm)
    rsd.detect_rare_samples(6, ccut, 5)
    out["rare_cos_lk"] = last_kept(rsd._res_lut[(6, ccut, 5)][1], x.shape[0])

    # ---- imputation: log counts with drop-outs, euclidean -------------------
    xi = synth.log_counts(160, 70, seed=41)
    out["imp_x"] = xi
    iparams = [(8, 3, 0.5, 3), (7, 2, 1.0, 2), (5, 5, 0.5, 1), (6, 1, 0.7, 4)]
    out["imp_params"] = np.array(iparams, dtype=np.float64)
    isdm = SDM(xi, metric="euclidean")
    fi = ref_knn.FeatureImputation(isdm)
    res = fi.impute_features([p[0] for p in iparams], [p[1] for p in iparams],
                             [p[2] for p in iparams], [p[3] for p in iparams])
    for j, p in enumerate(iparams):
        key = (p[0], p[1], float(p[2]), p[3], np.median)
        rx, ridc, _ = fi._res_lut[key]
        out["imp_x{}".format(j)] = np.asarray(rx)
        out["imp_idc{}".format(j)] = ridc.toarray().astype(np.int32)
        lut = isdm.s_knn_ind_lut(p[0])
        out["imp_knn{}".format(j)] = np.array(
            [lut[s] for s in range(xi.shape[0])], dtype=np.int32)

    # ---- kNN connectivity + affinity edge weights ---------------------------
    xc = synth.raw_counts(90, 12, seed=3)
    xc[10] = xc[3]                              # exact zero distances occur
    xc[50] = xc[3]
    xc[77] = xc[20]
    out["conn_x"] = xc
    csdm = SDM(xc, metric="euclidean")
    conn = csdm.s_knn_connectivity_matrix(6)
    out["conn_indptr"] = conn.indptr.astype(np.int64)
    out["conn_indices"] = conn.indices.astype(np.int32)
    out["conn_data"] = conn.data
    out["conn_sorted"] = np.array([conn.has_sorted_indices], dtype=np.int64)
    # knn_conn_mat_to_aff_graph's arithmetic (sdm.py:1085-1093); the igraph
    # constructor is stubbed, so capture the arguments it is called with
    import scedar.eda.sdm as ref_sdm_mod
    seen = {}

    class _G(object):
        def __init__(self, edges, directed, edge_attrs):
            seen["edges"] = np.array(edges, dtype=np.int64)
            seen["w"] = np.asarray(edge_attrs["weight"], dtype=np.float64)
    # np.matrix.A1 is what the reference calls on the fancy-indexed result
    ref_sdm_mod.ig.Graph = _G
    try:
        SDM.knn_conn_mat_to_aff_graph(conn, aff_scale=2.5)
        out["aff_edges"] = seen["edges"]
        out["aff_weights"] = seen["w"]
    except Exception as e:                      # scipy without .A1
        print("affinity via the live reference failed:", repr(e))
    path = os.path.join(HERE, "consumers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    for k_ in sorted(out):
        print(" ", k_, out[k_].shape, out[k_].dtype)


if __name__ == "__main__":
    main()
