"""Developer comparison of the three routes of the FP64 kNN (fused Gram ->
top-k, chunked D + top-k, symmetric D + top-k) and the tcgen05 path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scedar_b200 import device as dev
from scedar_b200._capi import EXCLUDE_SELF


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for arg in sys.argv[1:]:
    n, p, k = (int(v) for v in arg.split("x"))
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn((n, p), dtype=torch.float64, device="cuda", generator=g)
    cells = dev.Cells.from_dense(x, "euclidean")
    out = {"n": n, "p": p, "k": k}
    ref = None
    for mode in ("fused", "chunked", "symmetric"):
        if mode == "symmetric" and 8.0 * n * n > 60e9:
            continue
        if mode == "fused" and k > 31:
            continue
        os.environ["SCEDAR_B200_KNN_FP64"] = mode
        out[mode + "_ms"] = round(timed(lambda: cells.knn_stripe(k, EXCLUDE_SELF)), 2)
        i, d = cells.knn_stripe(k, EXCLUDE_SELF)
        if ref is None:
            ref = (i, d)
        else:
            assert torch.equal(ref[0], i) and torch.equal(ref[1], d), mode
    os.environ.pop("SCEDAR_B200_KNN_FP64", None)
    out["auto_ms"] = round(timed(lambda: cells.knn_stripe(k, EXCLUDE_SELF)), 2)
    if k <= 55:
        out["fast_ms"] = round(timed(lambda: cells.knn_stripe(k, EXCLUDE_SELF, precision="fast")), 2)
    cells.close()
    print(json.dumps(out), flush=True)
    del x
    torch.cuda.empty_cache()
