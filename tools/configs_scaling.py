"""Every BASELINE.json config at N GPUs (run under torchrun, one rank per GPU;
also works as a plain single process):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port 29561 tools/configs_scaling.py [c1 c2 c3 c4 c5]

Inputs are generated on rank 0's GPU with torch's RNG (same shapes and
distributions as scedar_b200.synth; this tool measures throughput, parity is
the test-suite's job) and broadcast.  Times are CUDA-event times, best of 2
after one warm-up, max over ranks.  One JSON line per config on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from scedar_b200 import device as dev
from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
from scedar_b200.stripes import StripedDistanceMatrix

FP64_PEAK = 37.1      # TFLOP/s, profiles/r01_fp64_peak.json
HBM_PEAK = 6558.4     # GB/s, MEASURED_PEAKS.json
RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def timed(fn, reps=2):
    def once():
        torch.cuda.synchronize()
        if WORLD > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if WORLD > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t.item())
    once()
    best, res = 1e30, None
    for _ in range(reps):
        res, t = once()
        best = min(best, t)
    return res, best


def gen(shape, fn):
    """fn() -> tensor on rank 0; other ranks pass None + shape."""
    return fn() if RANK == 0 else None


def fp64_d(out, sdm, n, p, t_d):
    alg = 1.0 * n * (n + 1) * p / WORLD            # per-GPU flops of the symmetric Gram
    out["d_ms"] = round(t_d, 2)
    out["d_pairs_per_s"] = float("%.4g" % (float(n) * n / t_d * 1e3))
    out["k2_tflops_per_gpu_alg"] = round(alg / t_d * 1e-9, 2)
    out["k2_frac_of_dmma_peak"] = round(alg / t_d * 1e-9 / FP64_PEAK, 3)


def striped_d(sdm):
    peers = sdm.make_peer_stripes() if WORLD > 1 else None
    _, t = timed(lambda: sdm.compute_d(peers=peers))
    return peers, t


def c1(device):
    n, p, k = 3005, 19972, 15
    def make():
        g = torch.Generator(device="cuda").manual_seed(3005)
        lam = torch.distributions.Gamma(0.3, 0.5).sample((p,)).cuda().double()
        return torch.log1p(torch.poisson(lam.expand(n, p), generator=g))
    sdm = StripedDistanceMatrix(gen((n, p), make), "correlation", shape=(n, p), device=device)
    out = {"config": "C1 3005x19972 correlation D + kNN k=15", "gpus": WORLD}
    peers, t_d = striped_d(sdm)
    fp64_d(out, sdm, n, p, t_d)
    _, t_k = timed(lambda: sdm.knn(k, EXCLUDE_FIRST))
    out["knn_from_d_ms"] = round(t_k, 3)
    sdm.close()
    if peers:
        peers.close()
    return out


def c2(device):
    n, p, k = 20000, 20000, 15
    csr, nnz = None, 0
    if RANK == 0:
        g = torch.Generator(device="cuda").manual_seed(20000)
        x = torch.zeros((n, p), dtype=torch.float64, device="cuda")
        for r0 in range(0, n, 2000):
            m = torch.rand((2000, p), device="cuda", generator=g) < 0.10
            v = torch.exp(torch.randn((2000, p), device="cuda", generator=g, dtype=torch.float64))
            x[r0:r0 + 2000] = torch.where(m, v, torch.zeros_like(v))
        sp = x.to_sparse_csr()
        csr = (sp.crow_indices().to(torch.int64), sp.col_indices().to(torch.int32),
               sp.values().contiguous())
        nnz = int(csr[2].numel())
        del x, sp
    if WORLD > 1:
        t = torch.tensor([nnz], dtype=torch.int64, device="cuda")
        dist.broadcast(t, src=0)
        nnz = int(t.item())
    out = {"config": "C2 CSR 20000x20000 (10% nnz) cosine + kNN graph k=15", "gpus": WORLD,
           "nnz": nnz}
    (sdm), t_prep = timed(lambda: StripedDistanceMatrix.from_csr(
        csr, "cosine", shape=(n, p), nnz=nnz, device=device), reps=1)
    out["bcast_scatter_diag_ms"] = round(t_prep, 2)
    peers, t_d = striped_d(sdm)
    fp64_d(out, sdm, n, p, t_d)
    (idx, dd), t_k = timed(lambda: sdm.knn(k, EXCLUDE_SELF))
    out["knn_from_d_ms"] = round(t_k, 3)
    _, t_c = timed(lambda: dev.knn_connectivity_csr(idx, dd))
    out["connectivity_csr_ms"] = round(t_c, 3)
    sdm.close()
    if peers:
        peers.close()
    return out


def c3(device):
    n, p, k = 68579, 2000, 15
    def make():
        g = torch.Generator(device="cuda").manual_seed(68579)
        cen = torch.randn((10, p), device="cuda", generator=g, dtype=torch.float64) * 3.0
        lab = torch.randint(0, 10, (n,), device="cuda", generator=g)
        x = cen[lab] + torch.randn((n, p), device="cuda", generator=g, dtype=torch.float64)
        rows = torch.randperm(n, device="cuda", generator=g)[: n // 100]
        x[rows] = (torch.rand((n // 100, p), device="cuda", generator=g, dtype=torch.float64)
                   - 0.5) * 24.0
        return x - x.min()
    x = gen((n, p), make)
    sdm = StripedDistanceMatrix(x, "euclidean", shape=(n, p), device=device)
    out = {"config": "C3 68579x2000 euclidean: D + kNN + rare samples + imputation",
           "gpus": WORLD}
    peers, t_d = striped_d(sdm)
    fp64_d(out, sdm, n, p, t_d)
    (idx, dd), t_k = timed(lambda: sdm.knn(k, EXCLUDE_FIRST))
    out["knn_from_d_ms"] = round(t_k, 3)
    out["knn_GBs_per_gpu"] = round(8.0 * n * n / WORLD / t_k * 1e-6, 1)
    cut = float(dd[:, -1].median().item())
    lk, t_r = timed(lambda: sdm.rare_samples(k, cut, 5))
    out["rare_5iter_ms"] = round(t_r, 2)
    out["rare_kept"] = int((lk == 5).sum().item())
    xi = torch.empty((n, p), dtype=torch.float64, device=device)
    if RANK == 0:
        xi.copy_(torch.where(x > 6.0, x, torch.zeros_like(x)))
    if WORLD > 1:
        dist.broadcast(xi, src=0)
    (_, _, st), t_i = timed(lambda: sdm.impute(xi, idx, k, 5, 6.0, 1))
    out["impute_1iter_ms"] = round(t_i, 2)
    out["impute_entries"] = st[0]
    (_, _), t_f = timed(lambda: sdm.knn(k, EXCLUDE_SELF, from_d=False, precision="fast"))
    out["knn_no_d_fast_exact_ms"] = round(t_f, 2)
    sdm.close()
    if peers:
        peers.close()
    return out


def c4(device):
    n, p, k = 100000, 2000, 15
    def make():
        g = torch.Generator(device="cuda").manual_seed(100000)
        lam = torch.distributions.Gamma(0.3, 0.5).sample((p,)).cuda().double()
        return torch.log1p(torch.poisson(lam.expand(n, p), generator=g))
    sdm = StripedDistanceMatrix(gen((n, p), make), "correlation", shape=(n, p), device=device)
    out = {"config": "C4 100000x2000 correlation D (row-striped) + kNN k=15", "gpus": WORLD}
    peers, t_d = striped_d(sdm)
    if peers is None:
        d = torch.empty((n, n), dtype=torch.float64, device=device)
        _, t_d = timed(lambda: sdm.cells.pdist_symmetric(out=d))
        sdm.d_stripe = d
    fp64_d(out, sdm, n, p, t_d)
    _, t_k = timed(lambda: sdm.knn(k, EXCLUDE_FIRST))
    out["knn_from_d_ms"] = round(t_k, 3)
    out["knn_GBs_per_gpu"] = round(8.0 * n * n / WORLD / t_k * 1e-6, 1)
    sdm.d_stripe = None
    if peers is None:
        del d
    _, t_f = timed(lambda: sdm.knn(k, EXCLUDE_SELF, from_d=False, precision="fast"))
    out["knn_no_d_fast_exact_ms"] = round(t_f, 2)
    sdm.close()
    if peers:
        peers.close()
    return out


def c5(device):
    n, p, k = 1000000, 50, 30
    def make():
        g = torch.Generator(device="cuda").manual_seed(1000000)
        sc = 1.0 / torch.sqrt(1.0 + torch.arange(p, device="cuda", dtype=torch.float64))
        return torch.randn((n, p), device="cuda", generator=g, dtype=torch.float64) * sc
    sdm = StripedDistanceMatrix(gen((n, p), make), "euclidean", shape=(n, p), device=device)
    out = {"config": "C5 1000000x50 euclidean kNN k=30, no D", "gpus": WORLD}
    _, t_f = timed(lambda: sdm.knn(k, EXCLUDE_SELF, from_d=False, precision="fast"))
    out["knn_fast_exact_ms"] = round(t_f, 2)
    out["candidate_pairs_per_s"] = float("%.4g" % (float(n) * n / t_f * 1e3))
    if WORLD >= 4:
        _, t_64 = timed(lambda: sdm.knn(k, EXCLUDE_SELF, from_d=False), reps=1)
        out["knn_fp64_ms"] = round(t_64, 1)
    sdm.close()
    return out


def main():
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(LOCAL)
    device = torch.device("cuda", LOCAL)
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=device)
    which = [a for a in sys.argv[1:]] or ["c1", "c2", "c3", "c4", "c5"]
    for w in which:
        res = globals()[w](device)
        torch.cuda.empty_cache()
        if RANK == 0:
            os.write(real_stdout, (json.dumps(res) + "\n").encode())
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
