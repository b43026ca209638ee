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

The Gram kernel is the FP64-grade INT8 split on the tcgen05 tensor cores
(`--precision fp64_i8`, gram_oz.cu: 7 byte slices, exact INT32 accumulation in
TMEM, float64 epilogue; same 1e-10 bar as the DMMA kernel, which is timed
beside it as `dmma_path`).  Every run ends with a parity gate: every rank
compares rows of ITS OWN D stripe (including the tiles a peer computed and
stored over NVLink) and of the gathered kNN lists with a float64 NumPy
evaluation and exits non-zero on any mismatch.

`value` times the step with x already resident in HBM; `e2e` times it through
the public API from pinned HOST memory (at N=1 x is uploaded range by range on
a copy stream while the contraction of the rows already there runs; at N>1
every rank uploads its 1/N row shard over its own PCIe link and the shards are
all-gathered over NVLink -- all inside the timed region) to the kNN lists back
on the host; `e2e_api` (N=1) is the plugin call a scedar user makes,
`SampleDistanceMatrix(x_numpy, metric).s_knn_ind_lut(k)`: pageable upload,
allocation of D, the Python dict of lists.  The reference arm
(--impl reference) times the reference's own CPU path -- the live class where
/root/reference is importable (kind "live"), else the oracle port pinned to it
(oracle/sdm_oracle.py, kind "port") -- on a row sample with all host threads:
20,000 cells when the requested steps fit a few
minutes, fewer otherwise, plus one 20,000-cell pass that times `_d`,
`_col_argsorted_d`, `s_knns(k)` and `s_knn_ind_lut(k)` separately.
"""
import argparse
import json
import os
import sys
import threading
import time

# BLAS / OpenMP threads of the CPU arm and of the parity gate: all host cores,
# also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(os.cpu_count() or 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRACE = bool(os.environ.get("SCEDAR_BENCH_TRACE"))
# tools/umma_probe sustained (profiles/r02_umma_sustained.log): the S = 7 MMA
# schedule alone, random operand bytes, power-capped: 3,819-3,905 TOP/s executed
INT8_MMA_ONLY_SUSTAINED_TOPS = 3860.0
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
    ap.add_argument("--precision", default=os.environ.get(
        "SCEDAR_BENCH_PRECISION", "fp64_i8"),
        choices=["fp64", "fp64_i8", "fp64_i8s6"],
        help="Gram kernel of the headline path (all FP64-grade, 1e-10)")
    ap.add_argument("--cpu-sample", type=int, default=20000,
                    help="rows of the workload the CPU arm is timed on "
                         "(SURVEY.md section 8d: 20,000)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0,
                    help="--impl reference: wall time the warm-up + timed "
                         "steps may take; the sample shrinks to fit")
    ap.add_argument("--check-rows", type=int, default=64)
    ap.add_argument("--e2e-mode", default=os.environ.get(
        "SCEDAR_BENCH_E2E_MODE", "pieces"),
        choices=["simple", "pipelined", "pieces"],
        help="N > 1: 'pieces' (default) holds every rank's own stripe on its "
             "host and overlaps upload, copy-engine exchange and contraction "
             "over growing pieces of every stripe "
             "(StripedDistanceMatrix.from_host_stripes); 'simple' uploads the "
             "row shards, all-gathers x and runs the step; 'pipelined' is the "
             "row-ordered form (from_host_shards)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the dmma_path / fast_path / e2e_api extras")
    return ap.parse_args()


def workload_name(a):
    return ("full {} distance matrix {}x{} (f64, row-striped, HBM-resident) "
            "+ k={} kNN; BASELINE.json configs[3]".format(
                a.metric, a.n, a.p, a.k))


# ---------------------------------------------------------------------------
# CPU arm: the reference's own path on a bounded row sample
# ---------------------------------------------------------------------------
def host_threads():
    """Pin the BLAS pools to all host cores and describe them."""
    info = []
    try:
        import numpy  # noqa: F401  (loads the BLAS the pools belong to)
        import scipy.linalg  # noqa: F401
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
        info = [{k: d.get(k) for k in ("user_api", "internal_api",
                                       "num_threads", "version")}
                for d in threadpoolctl.threadpool_info()]
    except Exception as e:                       # pragma: no cover
        info = [{"error": repr(e)}]
    return info


class CpuArm(object):
    """The four accessors the reference's users pay for, as callables on a
    sample: the live class when the checkout is importable, else the port."""

    def __init__(self, metric, k):
        self.metric, self.k = metric, k
        self.kind, self.cls = "port", None
        try:
            from oracle import ref_loader
            if ref_loader.reference_available():
                import warnings
                warnings.simplefilter("ignore")
                self.cls = ref_loader.load_reference()
                self.kind = "live"
        except Exception:
            self.cls, self.kind = None, "port"
        from oracle import sdm_oracle
        self.orc = sdm_oracle

    def e2e(self, xs):
        """SampleDistanceMatrix(x, metric).s_knn_ind_lut(k): _d + the full
        stable column argsort + the dict of lists."""
        t0 = time.perf_counter()
        if self.cls is not None:
            lut = self.cls(xs, metric=self.metric,
                           nprocs=os.cpu_count()).s_knn_ind_lut(self.k)
        else:
            lut = self.orc.knn_ind_lut(self.orc.pdist(xs, self.metric), self.k)
        dt = time.perf_counter() - t0
        assert len(lut) == xs.shape[0]
        return dt

    def breakdown(self, xs):
        """_d, _col_argsorted_d, s_knns(k) (on the cached D) timed one by
        one, then the end-to-end call on a fresh object."""
        out = {}
        if self.cls is not None:
            sdm = self.cls(xs, metric=self.metric, nprocs=os.cpu_count())
            t0 = time.perf_counter()
            sdm._d
            out["d_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            sdm._col_argsorted_d
            out["col_argsorted_d_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            sdm.s_knns(self.k)
            out["s_knns_s"] = time.perf_counter() - t0
            del sdm
        else:
            t0 = time.perf_counter()
            d = self.orc.pdist(xs, self.metric)
            out["d_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            self.orc.col_argsorted_d(d)
            out["col_argsorted_d_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            # what sklearn's kneighbors does on a precomputed matrix:
            # argpartition of k + 1, then a sort of those (sklearn/neighbors/
            # _base.py:715-755) -- not the full argsort of the oracle's
            # tie-defined restatement
            import numpy as np
            part = np.argpartition(d, self.k + 1, axis=1)[:, :self.k + 1]
            np.take_along_axis(part, np.argsort(
                np.take_along_axis(d, part, axis=1), axis=1), axis=1)
            out["s_knns_s"] = time.perf_counter() - t0
            del d
        out["s_knn_ind_lut_s"] = self.e2e(xs)
        n = xs.shape[0]
        out["rows"] = n
        out["d_pairs_per_s"] = n * n / out["d_s"]
        out["e2e_pairs_per_s"] = n * n / out["s_knn_ind_lut_s"]
        return out


def cpu_sample(a, rows=None):
    from scedar_b200 import synth
    ns = min(a.cpu_sample if rows is None else rows, a.n)
    return synth.log_counts(ns, a.p, seed=a.n)


def cpu_baseline(a):
    """cpu_baseline leg of the default run: one 20,000-cell pass, every
    accessor timed (about a minute or two of CPU work)."""
    threads = host_threads()
    arm = CpuArm(a.metric, a.k)
    xs = cpu_sample(a)
    ns = xs.shape[0]
    br = arm.breakdown(xs)
    return {"value": br["e2e_pairs_per_s"], "unit": UNIT,
            "cores": os.cpu_count(), "kind": arm.kind,
            "sample": "first {} of {} cells x {} genes, same generator; "
                      "SampleDistanceMatrix(x, {!r}).s_knn_ind_lut({}) "
                      "end to end (pdist + num_correct_dist_mat + full stable "
                      "column argsort), one pass; BLAS threads = all host "
                      "cores, NumPy sort single-threaded".format(
                          ns, a.n, a.p, a.metric, a.k),
            "breakdown": br, "threadpools": threads}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    arm = CpuArm(a.metric, a.k)
    x_full = cpu_sample(a)
    # sample size: 20,000 cells if (warm-up + steps) passes fit the budget,
    # else as many as do; calibrated on a small pass (the cost grows like
    # n^2 log n: pairs/s is NOT size-invariant, so the sample is reported)
    n_cal = min(3000, x_full.shape[0])
    t_cal = arm.e2e(x_full[:n_cal])
    warm = min(a.warmup, 1)                 # one warm-up is enough on CPU
    per_step = a.cpu_budget_s / max(a.steps + warm, 1)
    ns = int(n_cal * (per_step / max(t_cal, 1e-3)) ** 0.5 * 0.9)
    ns = max(1000, min(x_full.shape[0], ns // 100 * 100))
    xs = x_full[:ns]
    for _ in range(warm):
        arm.e2e(xs)
    ts = [arm.e2e(xs) for _ in range(a.steps)]
    dt = sum(ts) / len(ts)
    value = ns * ns / dt
    base = {"value": value, "unit": UNIT, "cores": os.cpu_count(),
            "kind": arm.kind,
            "sample": "first {} of {} cells x {} genes per step "
                      "(SampleDistanceMatrix(x, {!r}).s_knn_ind_lut({}) end "
                      "to end); best step {:.2f} s".format(
                          ns, a.n, a.p, a.metric, a.k, min(ts)),
            "best_value": ns * ns / min(ts), "threadpools": threads}
    if ns < x_full.shape[0]:
        # the 20,000-cell figures of SURVEY.md section 8d, one pass, outside
        # the timed steps
        base["breakdown_20k"] = arm.breakdown(x_full)
    else:
        base["breakdown_20k"] = arm.breakdown(xs)
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
    from scedar_b200.stripes import (I8_PRECISIONS, StripedDistanceMatrix,
                                     partition)

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
        peers = PeerStripes(partition(n, world), n, rank=rank,
                            blockmin=a.precision in I8_PRECISIONS)
        d_buf = peers.tensor(rank)
    else:
        d_buf = torch.empty((r1 - r0, n), dtype=torch.float64, device=dev)
    # block minima of D (a by-product of the INT8 split kernel; the selection
    # reads them instead of the whole matrix)
    bmin_buf = None
    if world == 1 and a.precision in I8_PRECISIONS:
        from scedar_b200.device import new_blockmin
        bmin_buf = new_blockmin(r1 - r0, n)
    idx_host = torch.empty((n, k), dtype=torch.int64).pin_memory()
    dist_host = torch.empty((n, k), dtype=torch.float64).pin_memory()
    phase_events = []

    prec = a.precision

    def step(x, broadcast, timed_kernel=False):
        """One pass of the hot path.  With timed_kernel, CUDA events bracket
        the phases on the launching stream (read after the region ends)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] \
            if timed_kernel else None
        trace = [time.perf_counter()] if TRACE else None
        if ev:
            ev[0].record()
        sdm = StripedDistanceMatrix(x, a.metric, shape=(n, p), device=dev,
                                    broadcast=broadcast,
                                    precision=prec)        # K1 + Gram diagonal
        if ev:
            ev[1].record()
        if trace:
            trace.append(time.perf_counter())
        sdm.compute_d(out=d_buf, peers=peers, bmin=bmin_buf)   # K2
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
        if ev:
            ev[4].record()
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
    # (rank 0 alone sets up NVML: before the barrier, or every other rank's timed
    # region starts with the wait for it -- 14 ms at N = 8)
    clocks = ClockSampler(local) if (
        rank == 0 and not os.environ.get("SCEDAR_BENCH_NO_CLOCKS")) else None
    sync_all()
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
    if TRACE:
        print("rank{} device ms: close {} | between steps {}".format(
            rank, [round(ev[3].elapsed_time(ev[4]), 3) for ev in phase_events],
            [round(a_[4].elapsed_time(b_[0]), 3)
             for a_, b_ in zip(phase_events, phase_events[1:])]),
            file=sys.stderr, flush=True)

    # ---- e2e: host x -> kNN lists on the host, through the public API ------
    # With several ranks the rows of x sit sharded over the ranks' pinned host
    # buffers (1/N each, as a per-process loader leaves them): every rank
    # uploads its shard over its own PCIe link and the shards are all-gathered
    # over NVLink (stripes.gather_x) -- 1.6 GB cross the host links once in
    # total, in parallel, instead of through rank 0's link alone.
    from scedar_b200.stripes import (gather_x_from_host, piece_layout,
                                     shard_rows, take_shard)
    sb, se = shard_rows(n, world)[rank]
    pipelined = world > 1 and prec != "fp64" and a.e2e_mode == "pipelined"
    pieces = world > 1 and prec != "fp64" and a.e2e_mode == "pieces"
    if pieces:
        # by-stripe shards: rank r holds the rows of its own D stripe
        from scedar_b200.device import PeerStripes as _PS
        shard_host = x_dev[r0:r1].cpu().pin_memory()      # set-up, not timed
        xpeers = _PS([(0, n)] * world, p, rank=rank)
    elif world > 1 and pipelined:
        # piece-interleaved row shards: rank r holds its 1/N of every row piece
        layout = piece_layout(n, world, 8)
        shard_host = take_shard(x_dev[:n], layout, rank, world).cpu() \
            .pin_memory()                                  # set-up, not timed
        shard_dev = torch.empty(tuple(shard_host.shape), dtype=torch.float64,
                                device=dev)
        cap = max(b + world * m for b, _, m in layout)
        x_e2e = torch.empty((cap, p), dtype=torch.float64, device=dev)
    elif world > 1:
        shard_host = x_dev[sb:se].cpu().pin_memory()      # set-up, not timed
        rows_max = (n + world - 1) // world
        shard_dev = torch.empty((rows_max, p), dtype=torch.float64, device=dev)
        x_e2e = torch.empty((world * rows_max, p), dtype=torch.float64,
                            device=dev)

    def e2e_step():
        if world == 1:
            # upload hidden behind the contraction: x goes up range by range
            # on a copy stream, each range is prepared and the rows of the
            # lower triangle it completes are contracted at once
            sdm = StripedDistanceMatrix.from_host(x_host, a.metric, out=d_buf,
                                                  x_dev=x_dev, precision=prec,
                                                  bmin=bmin_buf)
            idx, dd = sdm.knn(k, EXCLUDE_FIRST)
            sdm.close()
        elif pieces:
            sdm = StripedDistanceMatrix.from_host_stripes(
                shard_host, (n, p), peers, xpeers, metric=a.metric,
                precision=prec,
                n_pieces=int(os.environ["SCEDAR_BENCH_PIECES"])
                if os.environ.get("SCEDAR_BENCH_PIECES") else None)
            idx, dd = sdm.knn(k, EXCLUDE_FIRST)
            sdm.close()
        elif pipelined:
            # upload, all-gather and contraction pipelined over 8 row pieces
            sdm = StripedDistanceMatrix.from_host_shards(
                shard_host, layout, (n, p), peers, metric=a.metric,
                precision=prec, x_dev=x_e2e, shard_dev=shard_dev)
            idx, dd = sdm.knn(k, EXCLUDE_FIRST)
            sdm.close()
        else:
            # the shard goes up in four row ranges, each all-gathered over
            # NVLink as it lands (under the rest of the upload)
            idx, dd = step(gather_x_from_host(shard_host, n, shard_dev=shard_dev,
                                              out=x_e2e), False)
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

    # ---- parity gate on the result of the last end-to-end step (not timed) ---
    # Every rank checks rows of ITS OWN stripe of D -- all columns, so at N > 1
    # also the tiles a peer computed and stored over NVLink -- and the same
    # rows of the gathered kNN lists against float64 NumPy; any mismatch ends
    # the run with a non-zero exit code on every rank.
    gate = parity_gate(a, rank, world, r0, r1, x_dev, d_buf, idx, dd)
    ok = torch.tensor([1.0 if gate["ok"] else 0.0], dtype=torch.float64,
                      device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) != 1.0:
        print("bench parity gate FAILED on rank {}: {}".format(rank, gate),
              file=sys.stderr, flush=True)
        if world > 1:
            dist.destroy_process_group()
        raise SystemExit(3)

    # ---- secondary: the DMMA kernel and the tcgen05 fast path, same workload --
    def other_kernel(precision, what):
        """The same step with another FP64-grade Gram kernel, same buffers."""
        ev_d = []

        def alt_step():
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            sdm = StripedDistanceMatrix(x_dev, a.metric, shape=(n, p),
                                        device=dev, broadcast=False,
                                        precision=precision)
            e[0].record()
            sdm.compute_d(out=d_buf, peers=peers, bmin=bmin_buf)
            e[1].record()
            ev_d.append(e)
            i2, d2 = sdm.knn(k, EXCLUDE_FIRST)
            sdm.close()
            return i2, d2
        alt_step()
        sync_all()
        ev_d.clear()
        g0 = torch.cuda.Event(enable_timing=True)
        g1 = torch.cuda.Event(enable_timing=True)
        reps = min(a.steps, 3)
        g0.record()
        for _ in range(reps):
            i2, d2 = alt_step()
        g1.record()
        sync_all()
        ms = max_over_ranks(g0.elapsed_time(g1)) / reps
        gram = sum(e[0].elapsed_time(e[1]) for e in ev_d) / len(ev_d)
        differ = int((i2 != idx).any(dim=1).sum())
        return {"what": what, "ms_per_step": ms,
                "value": float(n) * n / (ms * 1e-3), "unit": UNIT,
                "gram_ms": gram,
                "knn_indices_identical_to_headline": differ == 0,
                "knn_rows_differing": differ,
                "max_abs_knn_distance_difference": float(
                    (d2 - dd).abs().max())}

    dmma = i8s6 = None
    if not a.no_secondary and prec != "fp64":
        dmma = other_kernel("fp64", "same step with the DMMA Gram kernel "
                            "(mma.sync.m8n8k4.f64, round 1's headline)")
    if not a.no_secondary and prec == "fp64_i8":
        i8s6 = other_kernel(
            "fp64_i8s6", "same step with 6 byte slices instead of 7 (21 INT8 "
            "products instead of 28): D within ~3e-12 of float64 instead of "
            "~4e-14, still inside the 1e-10 bar; not the default because only "
            "7 slices keep the lists identical to the DMMA path down to "
            "4-ulp near-ties")

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
    if k <= 55 and not a.no_secondary:
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
        same = bool(torch.equal(fi, idx[r0:r1]))
        fdiff = float((fd - dd[r0:r1]).abs().max())
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
            "knn_indices_identical_to_headline": same,
            "max_abs_knn_distance_difference": fdiff,
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
        if pieces:
            xpeers.close()

    # ---- e2e_api (N = 1): the plugin call a scedar user makes ----------------
    api = None
    if world == 1 and not a.no_secondary:
        import scedar_b200
        del d_buf
        torch.cuda.empty_cache()
        x_np = x_host.numpy()                     # pageable from torch's view
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            s_api = scedar_b200.SampleDistanceMatrix(x_np, metric=a.metric)
            lut = s_api.s_knn_ind_lut(k)
            ts.append(time.perf_counter() - t0)
            api_ok = all(lut[int(i)] == idx_host[int(i)].tolist()
                         for i in gate["rows"])
            s_api.release_device()
            del s_api, lut
        api = {"what": "SampleDistanceMatrix(x_numpy, metric={!r})."
                       "s_knn_ind_lut({}) through the drop-in class: upload "
                       "of x from pageable host memory, allocation of the "
                       "{:.0f} GB D, Gram, top-k, download, Python dict of "
                       "lists".format(a.metric, k, 8e-9 * n * n),
               "s_per_call": min(ts), "first_call_s": ts[0],
               "value": float(n) * n / min(ts), "unit": UNIT,
               "lists_identical_to_headline_on_gate_rows": bool(api_ok)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    rows_here = r1 - r0
    # entries of D this rank computes: the symmetric contraction is split
    # evenly over the ranks (every tile pair once), n(n+1)/2 in total
    entries = (n * (n + 1.0) / 2.0) * rows_here / n
    alg_flops = 2.0 * p * entries          # n(n+1)p/N
    peaks = measured_peaks()
    achieved = alg_flops / (gram_avg_ms * 1e-3) * 1e-12
    if prec == "fp64":
        peak = fp64_peak_tflops()
        roof = {
            "kernel": "gram_f64_kernel<STORE> (DMMA m8n8k4 + fused metric "
                      "epilogue, lower tiles mirrored)",
            "bound": "tensor", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "FP64 DMMA probe tools/fp64_peak.cu measured on "
                           "this pool (profiles/r01_fp64_peak.json); "
                           "MEASURED_PEAKS.json has no FP64 figure"}
    else:
        slices = 7 if prec == "fp64_i8" else 6
        products = slices * (slices + 1) // 2
        p_oz = (p + 63) // 64 * 64
        executed = achieved * products * p_oz / p     # INT8 TOP/s on the pipe
        bf16 = peaks.get("bf16_tflops_sustained") or peaks.get(
            "bf16_tflops") or 1400.0
        peak = 2.0 * bf16
        roof = {
            "kernel": "gram_oz_kernel<{}> (tcgen05.mma kind::i8, {} byte "
                      "slices, {} products, TMA-fed, INT32 TMEM accumulators, "
                      "float64 metric epilogue, lower tiles mirrored)".format(
                          slices, slices, products),
            "bound": "tensor", "achieved": executed, "peak": peak,
            "unit": "TOP/s (INT8, executed)", "frac": executed / peak,
            "algorithmic_tflops": achieved,
            "executed_over_algorithmic": products * p_oz / p,
            "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops_sustained "
                           "({}; the kernel is timed inside a seconds-long "
                           "step): kind::i8 issues at twice the rate of "
                           "kind::f16 on this part (tools/umma_probe.cu, "
                           "profiles/r02_umma_probe.log: 8,173 vs 4,087 "
                           "MAC/cycle/SM); of measured".format(bf16),
            "frac_of_burst_peak": executed / (2.0 * peaks.get(
                "bf16_tflops", 1640.9)),
            # what THIS instruction schedule sustains on this board from
            # operands resident in shared memory (no TMA, no epilogue), random
            # operand bytes, 170 ms under the power cap:
            # profiles/r02_umma_sustained.log
            "int8_mma_only_sustained_tops": INT8_MMA_ONLY_SUSTAINED_TOPS,
            "frac_of_int8_mma_only_sustained": (
                executed / INT8_MMA_ONLY_SUSTAINED_TOPS),
            "vs_fp64_dmma_peak": achieved / fp64_peak_tflops()}
    roof.update({"algorithmic_flops_per_launch": alg_flops,
                 "ms_per_launch": gram_avg_ms,
                 "share_of_step": gram_avg_ms / ms_per_step,
                 "traffic": recorded_traffic(prec) if world == 1 else None})
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n_cells": n, "n_genes": p,
                   "k": k, "metric": a.metric, "gram_kernel": prec,
                   "sharding": "row stripes x{}{}".format(
                       world, ", symmetric tile pairs stored into both "
                       "owners' stripes (peer stores over NVLink)"
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
                          "NVLink{}".format(
                              world, " -- piece by piece (8 row pieces), the "
                              "contraction of a piece's rows running while "
                              "the next piece is uploaded and gathered"
                              if pipelined else
                              " range by range behind the upload (4 ranges)"
                              if not pieces else
                              " -- by stripe: growing pieces (1/16 ... 1/2) of every stripe, pushed "
                              "into the peers' x buffers by copy engines while "
                              "the pairs of the pieces already there are "
                              "contracted"))},
        "e2e_api": api,
        "gpu_launches": int(launches),
        "roofline": roof,
        "parity_gate": {kk: vv for kk, vv in gate.items() if kk != "rows"},
        "clocks": clock_info,
        "dmma_path": dmma, "i8s6_path": i8s6,
        "fast_path": fast,
    }
    if world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def parity_gate(a, rank, world, r0, r1, x_dev, d_stripe, idx, dd):
    """Rows of this rank's stripe of D and of the gathered kNN lists against a
    float64 NumPy evaluation of the metric over all cells (plain NumPy: the
    oracle package stays test infrastructure).  Tolerance 1e-10 of max(D);
    indices must equal the stable (distance, index) order, a mismatch being
    tolerated only between distances within 4 ulp (none is expected)."""
    import numpy as np
    import torch
    n, k = a.n, a.k
    m = min(a.check_rows, r1 - r0)
    if m <= 0:
        return {"ok": True, "rows_checked": 0, "rows": []}
    rows = np.unique(np.linspace(r0, r1 - 1, m).astype(np.int64))
    x = x_dev.cpu().numpy()
    if a.metric == "euclidean":
        sq = np.einsum("ij,ij->i", x, x)
        d2 = -2.0 * (x[rows] @ x.T)
        d2 += sq[rows, None]
        d2 += sq[None, :]
        ref = np.sqrt(np.maximum(d2, 0.0))
    else:
        xc = x - x.mean(axis=1, keepdims=True) if a.metric == "correlation" \
            else x
        nrm = np.sqrt(np.einsum("ij,ij->i", xc, xc))
        inv = np.where(nrm > 0, 1.0 / np.where(nrm > 0, nrm, 1.0), 0.0)
        ref = 1.0 - (xc[rows] @ xc.T) * inv[rows, None] * inv[None, :]
        np.maximum(ref, 0.0, out=ref)
    ref[np.arange(len(rows)), rows] = 0.0
    rt = torch.from_numpy(rows - r0).to(d_stripe.device)
    got = d_stripe[rt].cpu().numpy()
    err = float(np.abs(got - ref).max() / ref.max())
    diag_ok = not got[np.arange(len(rows)), rows].any()
    # columns == rows inside the stripe's own square block (bit-symmetry)
    blk = d_stripe[:, r0:r1]
    sym_ok = bool(torch.equal(blk[:2048, :2048], blk[:2048, :2048].T))
    # kNN lists of the same rows: rank 0 of the stable order is dropped
    # (s_knn_ind_lut, sdm.py:764)
    order = np.argsort(ref, kind="mergesort", axis=1)[:, 1:k + 1]
    gi = idx[torch.from_numpy(rows).to(idx.device)].cpu().numpy()
    gd = dd[torch.from_numpy(rows).to(dd.device)].cpu().numpy()
    bad = np.nonzero((order != gi).any(axis=1))[0]
    near_ties = 0
    hard = 0
    for r in bad:
        dw, dg = ref[r, order[r]], ref[r, gi[r]]
        if np.abs(dw - dg).max() <= 4 * np.spacing(max(dw.max(), 1e-300)):
            near_ties += 1
        else:
            hard += 1
    derr = float(np.abs(gd - np.take_along_axis(ref, gi, axis=1)).max()
                 / ref.max())
    ok = (err <= 1e-10 and diag_ok and sym_ok and hard == 0 and near_ties == 0
          and derr <= 1e-10)
    return {"ok": bool(ok), "rows_checked": int(len(rows)),
            "stripe": [int(r0), int(r1)], "max_rel_err_d": err,
            "zero_diagonal": bool(diag_ok), "bit_symmetric_block": sym_ok,
            "knn_rows_differing": int(len(bad)), "knn_near_ties": near_ties,
            "knn_max_rel_err_dist": derr,
            "reference": "float64 NumPy over all {} cells, rank-local".format(n),
            "rows": rows.tolist()}


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


def recorded_traffic(prec="fp64"):
    """DRAM bytes per launch of the Gram kernel from the committed ncu
    capture of this workload on ONE GPU, when one exists (else null)."""
    name = {"fp64": "r01_gram_traffic.json"}.get(
        prec, "r02_gram_oz_traffic_{}.json".format(prec))
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
