"""Timing of the clustering consumers of D on one GPU:
condensed form (``scedar_squareform_rows``) and sub-matrix gather
(``scedar_d_gather_block``) of a resident D, CUDA events, against the measured
HBM copy peak; the CPU figure beside it is scipy's own ``squareform`` on a
bounded sample.  Developer evidence for profiles/, not the contract bench.

    python tools/hclust_bench.py [n_cells]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from scedar_b200 import device as dev, synth

HBM_PEAK = 6558.4  # GB/s, MEASURED_PEAKS.json


def ev(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.median(times))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    x = synth.log_counts(n, 256, seed=n)
    cells = dev.Cells.from_dense(dev.to_device(x), "correlation")
    d = cells.pdist_symmetric()
    out = torch.empty(dev.condensed_len(n), dtype=torch.float64, device="cuda")
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    ms = ev(lambda: dev.squareform_rows(d, n, out=out, flags=flags))
    gbs = 16.0 * dev.condensed_len(n) / ms / 1e6
    rng = np.random.default_rng(0)
    sel = dev.to_device(np.sort(rng.choice(n, size=n // 4, replace=False)))
    ms_g = ev(lambda: dev.gather_block(d, sel, sel))
    m = sel.numel()
    # CPU: scipy's squareform (checks included, as hclust_linkage calls it)
    from scipy.spatial.distance import squareform
    ns = min(n, 8000)
    dh = d[:ns, :ns].cpu().numpy()
    t0 = time.perf_counter()
    v = squareform(dh)
    cpu_s = time.perf_counter() - t0
    assert np.array_equal(
        v, dev.squareform_rows(d[:ns, :ns].contiguous(), ns)[0].cpu().numpy())
    print(json.dumps({
        "n_cells": n,
        "squareform": {"ms": round(ms, 3), "algorithmic_bytes": 16 * dev.condensed_len(n),
                       "GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / HBM_PEAK, 3),
                       "flags": int(flags.item())},
        "gather_block": {"rows": m, "cols": m, "ms": round(ms_g, 3),
                         "GBps_out": round(8.0 * m * m / ms_g / 1e6, 1),
                         "note": "includes the 4-byte flag read-back (one sync)"},
        "cpu_scipy_squareform": {"n_cells": ns, "s": round(cpu_s, 3),
                                 "pairs_per_s": round(ns * (ns - 1) / 2 / cpu_s)},
        "gpu_pairs_per_s": round(dev.condensed_len(n) / (ms / 1e3)),
    }))


if __name__ == "__main__":
    main()
