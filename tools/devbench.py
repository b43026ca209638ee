"""Developer timing sweep (not the contract bench): per-kernel CUDA-event
times of the device-level operators at a few shapes."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scedar_b200 import device as dev
from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    shapes = [(3005, 19972), (10000, 2000), (20000, 2000), (40000, 2000),
              (100000, 50)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for n, p in shapes:
        g = torch.Generator(device="cuda").manual_seed(n)
        x = torch.rand((n, p), dtype=torch.float64, device="cuda", generator=g)
        out = {"n": n, "p": p}
        t = timed(lambda: dev.Cells.from_dense(x, "correlation").close())
        out["prep+diag_ms"] = round(t, 3)
        cells = dev.Cells.from_dense(x, "correlation")
        if n * n * 8 < 60e9:
            d = torch.empty((n, n), dtype=torch.float64, device="cuda")
            t = timed(lambda: cells.pdist_stripe(0, n, out=d))
            out["stripe_ms"] = round(t, 3); out["stripe_tflops"] = round(2.0 * n * n * p / t * 1e-9, 2)
            t = timed(lambda: cells.pdist_symmetric(out=d))
            out["sym_ms"] = round(t, 3); out["sym_tflops_alg"] = round(1.0 * n * (n + 1) * p / t * 1e-9, 2)
            t = timed(lambda: dev.knn_from_d(d, 15, EXCLUDE_SELF))
            out["topk_from_d_ms"] = round(t, 3); out["topk_GBs"] = round(8.0 * n * n / t * 1e-6, 1)
            del d
        t = timed(lambda: cells.knn_stripe(15, EXCLUDE_SELF))
        out["fused_knn_ms"] = round(t, 3); out["fused_tflops"] = round(2.0 * n * n * p / t * 1e-9, 2)
        t = timed(lambda: cells.knn_stripe(15, EXCLUDE_SELF, precision="fast"))
        out["fast_knn_ms"] = round(t, 3); out["fast_knn_tflops_alg"] = round(2.0 * n * n * p / t * 1e-9, 2)
        if n * n * 8 < 60e9:
            d = torch.empty((n, n), dtype=torch.float64, device="cuda")
            t = timed(lambda: cells.pdist_stripe(0, n, out=d, precision="fast"))
            out["fast_d_ms"] = round(t, 3); out["fast_d_tflops_alg"] = round(2.0 * n * n * p / t * 1e-9, 2)
            t = timed(lambda: cells.pdist_symmetric(out=d, precision="fast"))
            out["fast_sym_ms"] = round(t, 3); out["fast_sym_tflops_alg"] = round(1.0 * n * (n + 1) * p / t * 1e-9, 2)
            del d
        cells.close(); del x
        torch.cuda.empty_cache()
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
