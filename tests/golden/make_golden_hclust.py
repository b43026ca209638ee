"""Golden vectors for the clustering consumers of D:
``HClustTree.hclust_linkage`` / ``sort_x_by_d`` outputs of the LIVE reference
(scedar.eda) on seeded inputs, plus the slicing of ``ind_x``.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_hclust.py
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
import scedar.eda as ref_eda  # noqa: E402

warnings.simplefilter("ignore")

LINKAGES = ["single", "complete", "average", "weighted", "ward"]


def main():
    out = {}
    # ---- linkage matrices from a reference-computed D, every metric ---------
    x = synth.blobs(150, 12, seed=11, n_blobs=3, outlier_frac=0.04)
    out["x"] = x
    for metric in ("euclidean", "cosine", "correlation"):
        d = SDM(x, metric=metric)._d
        for lt in LINKAGES:
            z = ref_eda.HClustTree.hclust_linkage(d, linkage=lt)
            out["z_{}_{}".format(metric, lt)] = z
        z = ref_eda.HClustTree.hclust_linkage(
            d, linkage="auto", is_euc_dist=(metric == "euclidean"))
        out["z_{}_auto".format(metric)] = z
        z = ref_eda.HClustTree.hclust_linkage(
            d, linkage="auto", n_eval_rounds=11,
            is_euc_dist=(metric == "euclidean"), optimal_ordering=True)
        out["z_{}_auto_oo".format(metric)] = z
        out["leaves_{}".format(metric)] = np.array(
            ref_eda.HClustTree.sort_x_by_d(x, metric=metric), dtype=np.int64)
        t = ref_eda.HClustTree.hclust_tree(d, linkage="average")
        cnt = t.n_round_bipar_cnt(5)
        out["bipar_{}".format(metric)] = np.array(
            [c + [-1] * (32 - len(c)) for c in cnt], dtype=np.int64)

    # ---- a user-supplied, slightly asymmetric matrix with negatives ---------
    rng = np.random.default_rng(7)
    dm = SDM(x[:60], metric="euclidean")._d.copy()
    dm += rng.normal(scale=1e-6, size=dm.shape)
    dm[3, 9] = -0.25
    out["dirty_d"] = dm
    out["z_dirty_complete"] = ref_eda.HClustTree.hclust_linkage(
        dm.copy(), linkage="complete")

    # ---- sort_features (sdm.py:178-184) --------------------------------------
    # continuous, tie-free features: with exactly tied distances (e.g. the
    # all-zero genes of log counts) the merge order follows the rounding
    # noise of the reference's BLAS and no other implementation reproduces it
    xs = synth.blobs(40, 25, seed=19, n_blobs=4, outlier_frac=0.0)
    s = SDM(xs, metric="correlation")
    s.sort_features(fdist_metric="euclidean", optimal_ordering=True)
    out["sf_x"] = xs
    out["sf_fids"] = np.array(s.fids, dtype=np.int64)

    # ---- ind_x slicing of D (sdm.py:684-692) ---------------------------------
    s = SDM(x, metric="cosine")
    sel = [5, 140, 3, 77, 78, 21]
    sub = s.ind_x(sel, [0, 2, 4])
    out["indx_sel"] = np.array(sel, dtype=np.int64)
    out["indx_d"] = sub._d

    path = os.path.join(HERE, "hclust.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    for k_ in sorted(out):
        print(" ", k_, out[k_].shape, out[k_].dtype)


if __name__ == "__main__":
    main()
