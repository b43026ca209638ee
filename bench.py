#!/usr/bin/env python
"""Benchmark of the distance-matrix + kNN hot path (BASELINE.json metric:
cell-pair distances/s and kNN-graph build time).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[3], the one the target is quoted on): the full
correlation distance matrix of 100,000 cells x 2,000 genes (float64, 80 GB,
row-striped over the N GPUs and kept in HBM) plus the k=15 nearest neighbours
of every cell, on seeded synthetic log-count data.  One step = one pass of the
whole path: prepare cells (centre rows, Gram diagonal) -> Gram + metric
epilogue into this rank's D stripe -> row-wise top-k from the stripe ->
all-gather of the kNN lists.

`value` times the step with x already resident in HBM; `e2e` times it through
the public API from pinned HOST memory (at N=1 x is uploaded range by range on
a copy stream while the contraction of the rows already there runs; at N>1
every rank uploads its 1/N row shard over its own PCIe link and the shards are
all-gathered over NVLink -- all inside the timed region) to the kNN lists back
on the host.  The reference arm
(--impl reference) times the CPU oracle port of the reference's own path
(oracle/sdm_oracle.py: pdist + full stable argsort, what s_knn_ind_lut costs)
on a bounded row sample with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRACE = bool(os.environ.get("SCEDAR_BENCH_TRACE"))
METRIC = "cell_pair_distances_per_s"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--p", type=int, default=2000)
    ap.add_argument("--k", type=int, default=15)
    ap.add_argument("--metric", default="correlation")
    ap.add_argument("--cpu-sample", type=int, default=6000,
                    help="rows of the workload the CPU arm is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return ("full {} distance matrix {}x{} (f64, row-striped, HBM-resident) "
            "+ k={} kNN; BASELINE.json configs[3]".format(
                a.metric, a.n, a.p, a.k))


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path on a bounded sample
# ---------------------------------------------------------------------------
def cpu_pass(xs, metric, k):
    from oracle import sdm_oracle as orc
    t0 = time.perf_counter()
    d = orc.pdist(xs, metric)          # SampleDistanceMatrix._d
    lut = orc.knn_ind_lut(d, k)        # s_knn_ind_lut: full stable argsort
    t1 = time.perf_counter()
    assert len(lut) == xs.shape[0]
    return t1 - t0


def cpu_sample(a):
    from scedar_b200 import synth
    ns = min(a.cpu_sample, a.n)
    return synth.log_counts(ns, a.p, seed=a.n)


def cpu_baseline(a, reps=1):
    xs = cpu_sample(a)
    ns = xs.shape[0]
    ts = [cpu_pass(xs, a.metric, a.k) for _ in range(reps)]
    t = min(ts)
    return {"value": ns * ns / t, "unit": UNIT, "cores": os.cpu_count(),
            "kind": "port",
            "sample": "first {} of {} cells x {} genes, same generator; "
                      "pdist + s_knn_ind_lut(k={}) (full stable argsort), "
                      "{:.2f} s per pass; BLAS threads = all host cores, "
                      "NumPy sort single-threaded".format(
                          ns, a.n, a.p, a.k, t)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    xs = cpu_sample(a)
    ns = xs.shape[0]
    for _ in range(min(a.warmup, 1)):       # one warm-up is enough on CPU
        cpu_pass(xs, a.metric, a.k)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_pass(xs, a.metric, a.k)
    dt = (time.perf_counter() - t0) / a.steps
    value = ns * ns / dt
    base = {"value": value, "unit": UNIT, "cores": os.cpu_count(),
            "kind": "port",
            "sample": "first {} of {} cells x {} genes per step".format(
                ns, a.n, a.p)}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "cpu_sample_rows": ns},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0}}), flush=True)


# ---------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------
class ClockSampler(object):
    """Samples SM clock and throttle reasons through NVML from a thread while
    the timed region runs (cheaper and less intrusive than polling the
    nvidia-smi binary)."""
    PERIOD_S = 0.2

    def __init__(self, index):
        self.samples, self.smax, self.ok = [], 0, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = pynvml.nvmlDeviceGetMaxClockInfo(
                self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), clk, rs))
            except Exception:
                pass
            self._stop.wait(self.PERIOD_S)

    def stop(self, t0, t1):
        if not self.ok:
            return None
        self._stop.set()
        self.t.join(timeout=2)
        nv = self.nv
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                 ("hw_thermal_slowdown",
                  "nvmlClocksEventReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown",
                  "nvmlClocksEventReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))
        sm, reasons = [], set()
        for t, clk, rs in self.samples:
            if not (t0 <= t <= t1):
                continue
            sm.append(clk)
            for nm, attr in names:
                bit = getattr(nv, attr, None)
                if bit is None:
                    bit = getattr(nv, attr.replace("ClocksEvent",
                                                   "ClocksThrottle"), 0)
                if rs & bit:
                    reasons.add(nm)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": float(sm[len(sm) // 2]),
                "sm_max_mhz": float(self.smax), "reasons": sorted(reasons),
                "samples": len(sm), "how": "NVML, every 0.2 s during the "
                                           "timed region"}


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from scedar_b200 import _capi, synth
    from scedar_b200._capi import EXCLUDE_FIRST
    from scedar_b200.stripes import StripedDistanceMatrix, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit("--gpus {} but WORLD_SIZE={}".format(a.gpus, world))
    # NCCL prints its version banner on stdout: keep fd 1 clean for the one
    # JSON line the driver reads
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.load()
    n, p, k = a.n, a.p, a.k

    # synthetic input on the host of rank 0 (pinned), broadcast once
    if rank == 0:
        x_host = torch.from_numpy(synth.log_counts(n, p, seed=n)).pin_memory()
    x_dev = torch.empty((n, p), dtype=torch.float64, device=dev)
    if rank == 0:
        x_dev.copy_(x_host, non_blocking=True)
    if world > 1:
        dist.broadcast(x_dev, src=0)
    r0, r1 = partition(n, world)[rank]
    # D stripe: with several GPUs it is a peer-mapped buffer so the symmetric
    # contraction can mirror tiles into the owner's HBM over NVLink
    peers = None
    if world > 1:
        from scedar_b200.device import PeerStripes
        peers = PeerStripes(partition(n, world), n, rank=rank)
        d_buf = peers.tensor(rank)
    else:
        d_buf = torch.empty((r1 - r0, n), dtype=torch.float64, device=dev)
    idx_host = torch.empty((n, k), dtype=torch.int64).pin_memory()
    dist_host = torch.empty((n, k), dtype=torch.float64).pin_memory()
    phase_events = []

    def step(x, broadcast, timed_kernel=False):
        """One pass of the hot path.  With timed_kernel, CUDA events bracket
        the phases on the launching stream (read after the region ends)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] \
            if timed_kernel else None
        trace = [time.perf_counter()] if TRACE else None
        if ev:
            ev[0].record()
        sdm = StripedDistanceMatrix(x, a.metric, shape=(n, p), device=dev,
                                    broadcast=broadcast)   # K1 + Gram diagonal
        if ev:
            ev[1].record()
        if trace:
            trace.append(time.perf_counter())
        sdm.compute_d(out=d_buf, peers=peers)              # K2
        if ev:
            ev[2].record()
        if trace:
            trace.append(time.perf_counter())
        idx, dd = sdm.knn(k, EXCLUDE_FIRST)   # K3, s_knn_ind_lut semantics
        if ev:
            ev[3].record()
            phase_events.append(ev)
        if trace:
            trace.append(time.perf_counter())
        sdm.close()
        if trace:
            trace.append(time.perf_counter())
            print("rank{} host ms: {}".format(rank, [
                round((b - a_) * 1e3, 2) for a_, b in zip(trace, trace[1:])]),
                file=sys.stderr, flush=True)
        return idx, dd

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: x resident in HBM -----------------------------------------
    for _ in range(a.warmup):
        step(x_dev, False)
    sync_all()
    clocks = ClockSampler(local) if (
        rank == 0 and not os.environ.get("SCEDAR_BENCH_NO_CLOCKS")) else None
    launches0 = lib.scedar_launch_count()
    t_host0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        idx, dd = step(x_dev, False, timed_kernel=True)
    e1.record()
    sync_all()
    t_host1 = time.perf_counter()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.scedar_launch_count() - launches0
    clock_info = clocks.stop(t_host0, t_host1) if clocks else None
    ms_per_step = total_ms / a.steps
    value = float(n) * n / (ms_per_step * 1e-3)
    phases = [[ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
              for ev in phase_events]
    prep_ms, gram_avg_ms, topk_ms = [sum(c) / len(c) for c in zip(*phases)]

    # ---- e2e: host x -> kNN lists on the host, through the public API ------
    # With several ranks the rows of x sit sharded over the ranks' pinned host
    # buffers (1/N each, as a per-process loader leaves them): every rank
    # uploads its shard over its own PCIe link and the shards are all-gathered
    # over NVLink (stripes.gather_x) -- 1.6 GB cross the host links once in
    # total, in parallel, instead of through rank 0's link alone.
    from scedar_b200.stripes import gather_x, shard_rows
    sb, se = shard_rows(n, world)[rank]
    if world > 1:
        shard_host = x_dev[sb:se].cpu().pin_memory()      # set-up, not timed
        shard_dev = torch.empty((se - sb, p), dtype=torch.float64, device=dev)

    def e2e_step():
        if world == 1:
            # upload hidden behind the contraction: x goes up range by range
            # on a copy stream, each range is prepared and the rows of the
            # lower triangle it completes are contracted at once
            sdm = StripedDistanceMatrix.from_host(x_host, a.metric, out=d_buf,
                                                  x_dev=x_dev)
            idx, dd = sdm.knn(k, EXCLUDE_FIRST)
            sdm.close()
        else:
            shard_dev.copy_(shard_host, non_blocking=True)
            idx, dd = step(gather_x(shard_dev, n), False)
        if rank == 0:
            idx_host.copy_(idx, non_blocking=True)
            dist_host.copy_(dd, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    sync_all()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / a.steps)
    e2e_value = float(n) * n / e2e_s

    # ---- secondary: the tcgen05 fast path on the same workload --------------
    # (D within 1e-4*max(D) of the FP64 matrix, kNN exact after the FP64
    # re-rank).  Reported beside the headline, never instead of it.
    from scedar_b200.device import Cells
    fast_events = []

    def fast_step(timed=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] \
            if timed else None
        cells = Cells.from_dense(x_dev, a.metric)
        if ev:
            ev[0].record()
        if world == 1:
            cells.pdist_symmetric(out=d_buf, precision="fast")
        else:
            cells.pdist_stripe(r0, r1, out=d_buf, precision="fast")
        if ev:
            ev[1].record()
        fi, fd = cells.knn_stripe(k, EXCLUDE_FIRST, r0, r1, precision="fast")
        if ev:
            ev[2].record()
            fast_events.append(ev)
        cells.close()
        return fi, fd

    fast = None
    if k <= 55:
        for _ in range(a.warmup):
            fast_step()
        sync_all()
        f0 = torch.cuda.Event(enable_timing=True)
        f1 = torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(a.steps):
            fi, fd = fast_step(timed=True)
        f1.record()
        sync_all()
        fast_ms = max_over_ranks(f0.elapsed_time(f1)) / a.steps
        fd_ms = sum(e[0].elapsed_time(e[1]) for e in fast_events) / a.steps
        fk_ms = sum(e[1].elapsed_time(e[2]) for e in fast_events) / a.steps
        same = bool(torch.equal(fi, idx[r0:r1]) and torch.equal(fd, dd[r0:r1]))
        rows_f = r1 - r0
        fast = {
            "what": "same workload with precision=fast: tcgen05/TMA 3xFP16 "
                    "Gram for the D stripe ({}) + fused candidate top-k, FP64 "
                    "re-rank, margin check, FP64 fallback".format(
                        "lower tiles mirrored" if world == 1
                        else "no symmetry used across stripes"),
            "ms_per_step": fast_ms, "value": float(n) * n / (fast_ms * 1e-3),
            "unit": UNIT, "d_stripe_ms": fd_ms, "knn_ms": fk_ms,
            "d_tolerance": "correlation/cosine 1e-5 absolute; euclidean 2e-6 * "
                           "max|x|^2 on d^2 (tests/test_gpu_stress.py)",
            "knn_identical_to_fp64_path": same,
            "roofline": {
                "kernel": "gram_tc_kernel<STORE> (tcgen05.mma kind::f16, TMA, "
                          "TMEM accumulators)",
                "bound": "tensor",
                "achieved": (1.0 * n * (n + 1) * p if world == 1
                             else 2.0 * rows_f * n * p) / (fd_ms * 1e-3) * 1e-12,
                "peak": measured_peaks().get("bf16_tflops", 1590.0),
                "unit": "TFLOP/s",
                "note": "algorithmic flops n(n+1)p at N=1 (2*rows*n*p per "
                        "stripe otherwise); the kernel executes 3x that on "
                        "the tensor pipe (hi.hi + hi.lo + lo.hi)"}}
        fast["roofline"]["frac"] = (fast["roofline"]["achieved"]
                                    / fast["roofline"]["peak"])
    if peers is not None:        # unmap before anyone frees its stripe
        sync_all()
        del d_buf
        peers.close()
        peers = None

    # ---- self-check of the last step's result on 64 rows (not timed) -------
    if rank == 0 and a.metric == "correlation":
        rows = np.arange(0, n, max(1, n // 64))[:64]
        xs = x_host.numpy()
        xc = xs - xs.mean(axis=1, keepdims=True)
        nrm = np.sqrt(np.einsum("ij,ij->i", xc, xc))
        dref = 1.0 - (xc[rows] @ xc.T) / nrm[rows, None] / nrm[None, :]
        dref[np.arange(len(rows)), rows] = -1.0        # self sorts first
        ref = np.argsort(dref, kind="mergesort", axis=1)[:, 1:k + 1]
        got = idx_host[rows].numpy()
        mism = int((ref != got).any(axis=1).sum())
        if mism:
            raise SystemExit("bench self-check failed: {} of {} sampled rows "
                             "disagree with NumPy".format(mism, len(rows)))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    rows_here = r1 - r0
    # entries of D this rank computes: the symmetric contraction is split
    # evenly over the ranks (every tile pair once), n(n+1)/2 in total
    entries = (n * (n + 1.0) / 2.0) * rows_here / n
    alg_flops = 2.0 * p * entries          # n(n+1)p/N
    fp64_peak = fp64_peak_tflops()
    achieved = alg_flops / (gram_avg_ms * 1e-3) * 1e-12
    peaks = measured_peaks()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n_cells": n, "n_genes": p,
                   "k": k, "metric": a.metric,
                   "sharding": "row stripes x{}{}".format(
                       world, ", symmetric tile pairs mirrored into the "
                       "owner's stripe by peer stores over NVLink"
                       if world > 1 else ", lower tiles mirrored"),
                   "l2": "inputs larger than L2 (x {:.1f} GB, D stripe "
                         "{:.1f} GB per GPU)".format(
                             n * p * 8e-9, rows_here * n * 8e-9),
                   "d_resident": "HBM (only the kNN lists leave the GPU)"},
        "knn_graph_build_s": ms_per_step * 1e-3,
        "phase_ms": {"prepare_cells": prep_ms, "gram_epilogue_d": gram_avg_ms,
                     "topk_and_gather": topk_ms},
        "e2e": {"value": e2e_value, "unit": UNIT, "s_per_step": e2e_s,
                "h2d_bytes_per_step": n * p * 8,
                "d2h_bytes_per_step": n * k * 16,
                "input": ("x in pinned host memory, uploaded range by range "
                          "on a copy stream behind the contraction"
                          if world == 1 else
                          "x row-sharded over the {} ranks' pinned host "
                          "buffers, uploaded in parallel, all-gathered over "
                          "NVLink".format(world))},
        "gpu_launches": int(launches),
        "roofline": {
            "kernel": "gram_f64_kernel<STORE> (DMMA m8n8k4 + fused metric "
                      "epilogue, lower tiles mirrored)",
            "bound": "tensor", "achieved": achieved, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": achieved / fp64_peak,
            "peak_source": "FP64 DMMA probe tools/fp64_peak.cu measured on "
                           "this pool (profiles/r01_fp64_peak.json); "
                           "MEASURED_PEAKS.json has no FP64 figure "
                           "(bf16 {} TF/s is not the pipe this kernel "
                           "uses)".format(peaks.get("bf16_tflops")),
            "algorithmic_flops_per_launch": alg_flops,
            "ms_per_launch": gram_avg_ms,
            "share_of_step": gram_avg_ms / ms_per_step,
            "traffic": recorded_traffic()},
        "clocks": clock_info,
        "fast_path": fast,
    }
    if world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def fp64_peak_tflops():
    """FP64 tensor-pipe peak measured by tools/fp64_peak.cu on this pool."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")) as f:
            return max(json.loads(l)["tflops"] for l in f if l.strip())
    except (OSError, ValueError, KeyError):
        return 37.1


def recorded_traffic():
    """DRAM bytes per launch of the Gram kernel from the committed ncu
    capture, when one exists for this workload (else null)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_gram_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
