"""Timing of the tcgen05 kNN / D kernels only (developer tool).
usage: fast_knn_bench.py NxPxK ..."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scedar_b200 import device as dev
from scedar_b200._capi import EXCLUDE_SELF


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for arg in sys.argv[1:]:
    f = arg.split("x")
    n, p, k = (int(v) for v in f[:3])
    if len(f) > 3 and f[3] == "log":     # the bench workload's data and metric
        from scedar_b200 import synth
        x = dev.to_device(synth.log_counts(n, p, seed=n))
        cells = dev.Cells.from_dense(x, "correlation")
    else:
        g = torch.Generator(device="cuda").manual_seed(n)
        x = torch.randn((n, p), dtype=torch.float64, device="cuda", generator=g)
        cells = dev.Cells.from_dense(x, "euclidean")
    out = {"n": n, "p": p, "k": k, "cg": os.environ.get("SCEDAR_B200_TC_CG", "default")}
    out["fast_knn_ms"] = round(timed(lambda: cells.knn_stripe(k, EXCLUDE_SELF, precision="fast")), 2)
    if 8.0 * n * n < 90e9:
        d = torch.empty((n, n), dtype=torch.float64, device="cuda")
        out["fast_sym_d_ms"] = round(timed(lambda: cells.pdist_symmetric(out=d, precision="fast")), 2)
        del d
    cells.close()
    print(json.dumps(out), flush=True)
    del x
    torch.cuda.empty_cache()
