"""Time the plugin call a scedar user makes at BASELINE configs[3]:
SampleDistanceMatrix(x_numpy, "correlation").s_knn_ind_lut(15) -- python tools/api_time.py [n] [p]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scedar_b200                      # noqa: E402
from scedar_b200 import synth           # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
x = synth.log_counts(n, p, seed=n)
for rep in range(3):
    t0 = time.perf_counter()
    s = scedar_b200.SampleDistanceMatrix(x, metric="correlation")
    t1 = time.perf_counter()
    d = s._device_d()
    import torch
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    lut = s.s_knn_ind_lut(15)
    t3 = time.perf_counter()
    print("call {}: total {:.1f} ms = ctor {:.1f} + D on device {:.1f} + lut {:.1f}".format(
        rep, (t3 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
    s.release_device()
    del s, lut, d
