"""One shape of the INT8 split Gram for ncu: python tools/oz_prof.py [n] [p] [precision] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scedar_b200 import device as dev, synth          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
prec = sys.argv[3] if len(sys.argv) > 3 else "fp64_i8"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.cuda.set_device(0)
x = torch.from_numpy(synth.log_counts(n, p, seed=n)).cuda()
cells = dev.Cells.from_dense(x, "correlation")
d = torch.empty((n, n), dtype=torch.float64, device="cuda")
if os.environ.get("OZ_PROF_BLOCKMIN"):          # block minima ride along
    bmin = dev.new_blockmin(n, n)
    cells.set_blockmin([bmin])
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    cells.pdist_symmetric(out=d, precision=prec)
    b.record()
    if os.environ.get("OZ_PROF_BLOCKMIN"):
        cells.knn_from_d_blockmin(d, bmin, 15)
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("n={} p={} {}: ms {}".format(n, p, prec, ["%.3f" % t for t in ts]))

# OZ_PROF_P2P="world,rank": that rank's share of the peer-store form, all stripes
# local (one GPU) -- tile order and ownership without the NVLink stores
if os.environ.get("OZ_PROF_P2P"):
    from scedar_b200.stripes import partition
    world, rank = (int(v) for v in os.environ["OZ_PROF_P2P"].split(","))
    del d
    torch.cuda.empty_cache()
    parts = partition(n, world)
    peers = dev.PeerStripes(parts, n, blockmin=bool(os.environ.get("OZ_PROF_BLOCKMIN")))
    if os.environ.get("OZ_PROF_BLOCKMIN"):
        cells.set_blockmin(peers.bmin_ptrs(), peers.ldb)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dev.pdist_stripe_p2p(cells, peers, rank, precision=prec)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("p2p world={} rank={}: ms {}".format(world, rank, ["%.3f" % t for t in ts]))
