"""Multi-GPU parity check (run under torchrun on a box with N >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

Row-striped kNN without D (FP64 and tcgen05 fast path), the peer-store
symmetric D, rare-sample detection, imputation and the condensed D / linkage
hand-off over the ranks are compared
with the same computation on one GPU (rank 0 recomputes everything alone).
Prints one JSON line with timings; exits non-zero on any mismatch."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    from scedar_b200 import device as dev, synth
    from scedar_b200._capi import EXCLUDE_FIRST, EXCLUDE_SELF
    from scedar_b200.stripes import StripedDistanceMatrix
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    out = {"world": world}

    def timed(fn):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        dist.barrier()
        return r, (time.perf_counter() - t0) * 1e3

    # ---- C5-like: kNN only, no D ------------------------------------------
    n, p, k = int(os.environ.get("CHECK_N5", 400000)), 50, 30
    x = dev.to_device(synth.pcs(n, p, seed=n)) if rank == 0 else None
    sdm = StripedDistanceMatrix(x, "euclidean", shape=(n, p), device=device)
    for _ in range(2):
        (fi, fd), t_fast = timed(lambda: sdm.knn(k, EXCLUDE_SELF, from_d=False,
                                                 precision="fast"))
    out["knn_fast_ms"] = round(t_fast, 2)
    out["knn_n"] = n
    if n <= 400000:
        (i64, d64), t64 = timed(lambda: sdm.knn(k, EXCLUDE_SELF, from_d=False))
        out["knn_fp64_ms"] = round(t64, 2)
        assert torch.equal(fi, i64) and torch.equal(fd, d64), "fast != fp64"
    if rank == 0:
        cells = dev.Cells.from_dense(x, "euclidean")
        si, sd = cells.knn_stripe(k, EXCLUDE_SELF, precision="fast")
        cells.close()
        assert torch.equal(si, fi) and torch.equal(sd, fd), "striped != single"
    sdm.close()
    del sdm, fi, fd

    # ---- C3-like: D striped with peer stores, rare samples, imputation -----
    n, p, k = int(os.environ.get("CHECK_N3", 20000)), 500, 15
    xh = synth.blobs(n, p, seed=3) if rank == 0 else None
    x = dev.to_device(xh) if rank == 0 else None
    sdm = StripedDistanceMatrix(x, "euclidean", shape=(n, p), device=device)
    peers = dev.PeerStripes(sdm.stripes, n, rank=rank, blockmin=True)
    _, t_d = timed(lambda: sdm.compute_d(peers=peers))
    out["d_p2p_ms"] = round(t_d, 2)
    idx, dd = sdm.knn(k, EXCLUDE_FIRST)
    cut = float(dd[:, -1].median().item())
    lk, t_rare = timed(lambda: sdm.rare_samples(k, cut, 5))
    out["rare_ms"] = round(t_rare, 2)
    xi = torch.empty((n, p), dtype=torch.float64, device=device)
    if rank == 0:
        xi.copy_(torch.where(x > 6.0, x, torch.zeros_like(x)))
    dist.broadcast(xi, src=0)
    (gx, gidc, gstats), t_imp = timed(
        lambda: sdm.impute(xi, idx, k, 5, 6.0, 2))
    out["impute_ms"] = round(t_imp, 2)
    out["impute_stats"] = list(gstats)
    # condensed D (HClustTree hand-off): ragged segments sent to rank 0
    sf, t_sf = timed(lambda: sdm.condensed(dst=0))
    out["condensed_ms"] = round(t_sf, 2)
    z = sdm.hclust_linkage(linkage="single", dst=world - 1)
    assert z.shape == (n - 1, 4)
    if rank == 0:
        # the same operators on one GPU, no process group
        cells = dev.Cells.from_dense(x, "euclidean")
        d = cells.pdist_symmetric()
        assert torch.equal(d[sdm.row_begin:sdm.row_end], sdm.d_stripe), "D"
        sf1, fl = dev.squareform_rows(d, n)
        assert int(fl.item()) == 0 and torch.equal(sf, sf1), "condensed D"
        import scipy.cluster.hierarchy as sch
        assert np.array_equal(z, sch.linkage(sf1.cpu().numpy(), "single"))
        del sf1
        alive = torch.ones(n, dtype=torch.uint8, device=device)
        lk1 = torch.zeros(n, dtype=torch.int32, device=device)
        kth = dev.masked_kth(d, alive, k)
        d_max = dev.masked_max(kth)
        m, fresh = n, k
        for i in range(1, 6):
            c = cut + (5 - i) / 5 * max(0, d_max - cut)
            if m == 0:
                continue
            i_k = min(m - 1, k)
            if fresh != i_k:
                kth = dev.masked_kth(d, alive, i_k)
                fresh = i_k
            kept = dev.rare_update(kth, c, i, alive, lk1)
            if kept != m:
                fresh = None
            m = kept
        assert torch.equal(lk, lk1), "rare samples"
        out["rare_kept"] = int((lk1 == 5).sum().item())
        i1, _ = dev.knn_from_d(d, k, EXCLUDE_FIRST)
        assert torch.equal(i1, idx), "kNN lists"
        cur, nxt = xi.clone(), torch.empty_like(xi)
        idc = torch.zeros((n, p), dtype=torch.int32, device=device)
        st = torch.zeros(2, dtype=torch.int64, device=device)
        for i in range(1, 3):
            n_do_i = 5 + int(np.ceil((2 - i) / 2 * (k - 5)))
            dev.impute_iter(cur, idx, n_do_i, 6.0, i, nxt, idc, st)
            cur, nxt = nxt, cur
        assert torch.equal(cur, gx) and torch.equal(idc, gidc), "imputation"
        assert tuple(st.tolist()) == gstats
        cells.close()
    dist.barrier()
    del gx, gidc
    sdm.close()
    # ---- the same D with the INT8 split kernel: every tile pair computed by one
    # rank and stored into both owners' stripes (own HBM or the peer's over
    # NVLink), against the single-GPU matrix of the same kernel, bit for bit
    xb = torch.empty((n, p), dtype=torch.float64, device=device)
    if rank == 0:
        xb.copy_(x)
    dist.broadcast(xb, src=0)
    for prec in ("fp64_i8", "fp64_i8s6"):
        s8 = StripedDistanceMatrix(xb, "euclidean", shape=(n, p), device=device,
                                   broadcast=False, precision=prec)
        peers.tensor(rank).fill_(-3.0)
        for _ in range(2):
            _, t_d = timed(lambda: s8.compute_d(peers=peers))
        out["d_p2p_{}_ms".format(prec)] = round(t_d, 2)
        cells = dev.Cells.from_dense(xb, "euclidean")
        d1 = cells.pdist_symmetric(precision=prec)
        mine = peers.tensor(rank)
        same = bool(torch.equal(d1[s8.row_begin:s8.row_end], mine))
        cells.close()
        flag = torch.tensor([1.0 if same else 0.0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert float(flag.item()) == 1.0, "peer-store D ({}) differs".format(prec)
        # the pipelined form: x sharded over the ranks' host buffers, uploaded,
        # gathered and contracted piece by piece -- the same matrix again
        from scedar_b200.stripes import piece_layout, take_shard
        lay = piece_layout(n, world, 4)
        shard = take_shard(xb, lay, rank, world).cpu().pin_memory()
        peers.tensor(rank).fill_(-5.0)
        for _ in range(2):
            s9, t_pipe = timed(lambda: StripedDistanceMatrix.from_host_shards(
                shard, lay, (n, p), peers, metric="euclidean", precision=prec))
            s9.close()
        out["d_pipelined_{}_ms".format(prec)] = round(t_pipe, 2)
        same = bool(torch.equal(d1[s8.row_begin:s8.row_end], peers.tensor(rank)))
        flag = torch.tensor([1.0 if same else 0.0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert float(flag.item()) == 1.0, "pipelined D ({}) differs".format(prec)
        # the by-stripe pipeline: every rank holds its own stripe's rows on the
        # host; pieces of every stripe at once, copy-engine exchange over NVLink
        xpeers = dev.PeerStripes([(0, n)] * world, p, rank=rank)
        shard2 = xb[s8.row_begin:s8.row_end].cpu().pin_memory()
        peers.tensor(rank).fill_(-6.0)
        for _ in range(2):
            s10, t_pc = timed(lambda: StripedDistanceMatrix.from_host_stripes(
                shard2, (n, p), peers, xpeers, metric="euclidean", precision=prec))
            k_i, k_d = s10.knn(k, EXCLUDE_SELF)
            s10.close()
        out["d_pieces_{}_ms".format(prec)] = round(t_pc, 2)
        same = bool(torch.equal(d1[s8.row_begin:s8.row_end], peers.tensor(rank)))
        r_i, r_d = dev.knn_from_d(d1, k, EXCLUDE_SELF)
        same = same and bool(torch.equal(k_i, r_i) and torch.equal(k_d, r_d))
        flag = torch.tensor([1.0 if same else 0.0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert float(flag.item()) == 1.0, "by-stripe pipelined D ({}) differs".format(prec)
        torch.cuda.synchronize()
        dist.barrier()
        xpeers.close()
        del d1
        s8.close()
    peers.close()
    if rank == 0:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
