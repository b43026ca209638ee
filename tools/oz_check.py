"""Development check of the INT8 split Gram (gram_oz.cu) against the DMMA path
and float64 NumPy: accuracy, bit-symmetry, exactness on counts, the stripe /
row-range / kNN forms, and CUDA-event timings.  Run on the GPU box:
    python tools/oz_check.py [quick|full]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scedar_b200 import device as dev, synth          # noqa: E402
from scedar_b200._capi import EXCLUDE_FIRST           # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
out = []


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def check(n, p, metric, kind="log", precs=("fp64_i8", "fp64_i8s6")):
    x = (synth.raw_counts(n, p, seed=n + p) if kind == "counts"
         else synth.log_counts(n, p, seed=n + p) if kind == "log"
         else synth.blobs(n, p, seed=n + p))
    xt = dev.to_device(x)
    cells = dev.Cells.from_dense(xt, metric)
    ref = cells.pdist_symmetric(precision="fp64")
    rec = {"n": n, "p": p, "metric": metric, "kind": kind}
    for pr in precs:
        d = cells.pdist_symmetric(precision=pr)
        torch.cuda.synchronize()
        err = float((d - ref).abs().max() / ref.max())
        sym = bool(torch.equal(d, d.T))
        diag0 = float(d.diagonal().abs().max()) == 0.0
        rec[pr] = {"rel_err_vs_dmma": err, "bit_symmetric": sym, "zero_diag": diag0,
                   "min": float(d.min()), "equal_to_dmma": bool(torch.equal(d, ref))}
        # stripe form and row-range form against the symmetric one, bit for bit
        if n >= 256:
            r0, r1 = 128, min(n, 128 * max(2, (n // 128) // 2 + 1))
            st = cells.pdist_stripe(r0, r1, precision=pr)
            rec[pr]["stripe_equal"] = bool(torch.equal(st, d[r0:r1]))
            c2 = dev.Cells.alloc(n, p, metric)
            full = torch.full((n, n), -7.0, dtype=torch.float64, device="cuda")
            for b, e in dev.upload_ranges(n, first=256):
                c2.fill_rows(xt[b:e], b)
                c2.pdist_lower_rows(b, e, full, precision=pr)
            rec[pr]["lower_rows_equal"] = bool(torch.equal(full, d))
            c2.close()
        if n > 40:
            i1, d1 = cells.knn_stripe(15, EXCLUDE_FIRST, precision=pr)
            i0, d0 = cells.knn_stripe(15, EXCLUDE_FIRST, precision="fp64")
            rec[pr]["knn_idx_equal_dmma"] = bool(torch.equal(i1, i0))
            rec[pr]["knn_rows_differ"] = int((i1 != i0).any(dim=1).sum())
        del d
    # float64 NumPy on a row sample
    rows = np.unique(np.linspace(0, n - 1, min(n, 64)).astype(int))
    if metric == "euclidean":
        sq = np.einsum("ij,ij->i", x, x)
        d2 = -2.0 * (x[rows] @ x.T) + sq[rows, None] + sq[None, :]
        dref = np.sqrt(np.maximum(d2, 0))
    else:
        xc = x - x.mean(axis=1, keepdims=True) if metric == "correlation" else x
        nr = np.sqrt(np.einsum("ij,ij->i", xc, xc))
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.where(nr > 0, 1.0 / nr, 0.0)
        dref = 1.0 - (xc[rows] @ xc.T) * inv[rows, None] * inv[None, :]
    dref[np.arange(len(rows)), rows] = 0.0
    rt = torch.from_numpy(rows).cuda()
    rec["dmma_vs_numpy"] = float(np.abs(ref[rt].cpu().numpy() - dref).max() / dref.max())
    for pr in precs:
        d = cells.pdist_symmetric(precision=pr)
        rec[pr]["rel_err_vs_numpy"] = float(np.abs(d[rt].cpu().numpy() - dref).max() / dref.max())
        del d
    cells.close()
    print(json.dumps(rec), flush=True)
    out.append(rec)


def timing(n, p, metric="correlation"):
    x = torch.from_numpy(synth.log_counts(n, p, seed=n)).cuda()
    cells = dev.Cells.from_dense(x, metric)
    d = torch.empty((n, n), dtype=torch.float64, device="cuda")
    rec = {"timing": True, "n": n, "p": p}
    for pr in ("fp64", "fp64_i8", "fp64_i8s6", "fast"):
        ms = ev_time(lambda: cells.pdist_symmetric(out=d, precision=pr))
        rec[pr + "_ms"] = ms
        prods = {"fp64": 1, "fp64_i8": 28, "fp64_i8s6": 21, "fast": 3}[pr]
        rec[pr + "_alg_tflops"] = n * (n + 1.0) * p / (ms * 1e-3) * 1e-12
        rec[pr + "_exec_tops"] = rec[pr + "_alg_tflops"] * prods
    cells.close()
    print(json.dumps(rec), flush=True)
    out.append(rec)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    t0 = time.time()
    check(300, 40, "cosine", "counts")
    check(300, 40, "euclidean", "counts")
    check(700, 300, "correlation", "log")
    check(1000, 500, "euclidean", "blobs")
    check(3000, 2000, "correlation", "log")
    check(3000, 2000, "cosine", "counts")
    check(2500, 4000, "euclidean", "log")
    check(600, 4200, "cosine", "counts")        # two K chunks
    check(1500, 9000, "correlation", "log")     # three K chunks
    check(3005, 19972, "correlation", "log")    # BASELINE config 1
    check(900, 19972, "euclidean", "counts")
    timing(20000, 2000)
    timing(3005, 19972)
    if mode == "full":
        timing(60000, 2000)
        timing(100000, 2000)
    print("oz_check done in {:.1f}s".format(time.time() - t0))
    with open(os.path.join(ROOT, "gpurun_out", "oz_check.jsonl"), "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")
