"""CPU: host-side logic of the drop-in that needs no device -- constructor
contract, laziness, error conventions (reference tests cited inline)."""
import numpy as np
import pytest
import scipy.sparse as sps

from scedar_b200 import SampleDistanceMatrix as SDM
from scedar_b200.stripes import partition

x_3x2 = [[0, 0], [1, 1], [2, 2]]


def test_ctor_is_lazy_and_validates():
    # tests/test_eda/test_dense_sdm.py:72-96: unknown metric accepted lazily
    for m in ("unknown", 1, 1.0, ("euclidean",), ["euclidean"]):
        sdm = SDM(x_3x2, metric=m)
        with pytest.raises(Exception):
            sdm.d
    with pytest.raises(ValueError):
        SDM(x_3x2, metric="precomputed")
    with pytest.raises(ValueError):
        SDM(np.empty(0), metric="euclidean")
    with pytest.raises(ValueError):
        SDM(None)
    with pytest.raises(ValueError):
        SDM([["a", "b"]])
    sdm = SDM(x_3x2, metric="euclidean", nprocs=5)
    assert sdm._nprocs == 5 and sdm._lazy_load_d is None
    assert SDM(x_3x2, nprocs=0)._nprocs == 1
    assert sdm.sids == [0, 1, 2] and sdm.fids == [0, 1]
    assert sdm.x == [[0.0, 0.0], [1.0, 1.0], [2.0, 2.0]]
    assert sdm._x.dtype == np.float64


def test_ids():
    with pytest.raises(ValueError):
        SDM(x_3x2, sids=[0, 0, 1])
    with pytest.raises(ValueError):
        SDM(x_3x2, sids=[0, 1])
    with pytest.raises(ValueError):
        SDM(x_3x2, fids=["a", 1])
    sdm = SDM(x_3x2, sids=["a", "b", "c"], fids=[10, 20], use_pdist=False)
    assert sdm.s_id_to_ind(["c", "a"]) == [2, 0]
    assert sdm.f_id_to_ind([20]) == [1]
    sub = sdm.id_x(["c", "a"], [20])
    assert sub.sids == ["c", "a"] and sub._x.tolist() == [[2.0], [0.0]]


def test_no_pdist_errors():
    # tests/test_eda/test_no_pdist_sdm.py:25-61, 398-411
    sdm = SDM(x_3x2, metric="euclidean", use_pdist=False)
    for attr in ("d", "_d", "_col_sorted_d", "_col_argsorted_d"):
        with pytest.raises(ValueError):
            getattr(sdm, attr)
    with pytest.raises(ValueError):
        sdm.s_ith_nn_d(1)
    with pytest.raises(ValueError):
        sdm.s_ith_nn_ind(1)
    with pytest.raises(ValueError):
        SDM(x_3x2, d=np.zeros((3, 3)), use_pdist=False)


def test_bad_d_and_k():
    with pytest.raises(ValueError):
        SDM(x_3x2, d=[[0, 1], [1, 0]])
    with pytest.raises(ValueError):
        SDM(x_3x2, d=[[0, "1a1", 2], [1, 0, 1], [2, 1, 0]])
    sdm = SDM(x_3x2, metric="euclidean")
    for bad in (-1, -0.5, 3, 3.5, 7):
        with pytest.raises(ValueError):
            sdm.s_knn_ind_lut(bad)
    with pytest.raises(ValueError):
        sdm.s_knns(0)
    with pytest.raises(ValueError):
        sdm.s_knns(1, index_params={})
    with pytest.raises(ValueError):
        sdm.s_knns(1, query_params={})
    with pytest.raises(NotImplementedError):
        sdm.s_knns(1, use_hnsw=True)
    with pytest.raises(ValueError):
        SDM.num_correct_dist_mat(np.zeros((3, 2)))
    with pytest.raises(ValueError):
        SDM.num_correct_dist_mat([[0, 1], [1, 0]])


def test_sparse_contract():
    xs = sps.csr_matrix(np.eye(3))
    sdm = SDM(xs, metric="correlation")       # lazy: ctor fine
    assert sdm._is_sparse and sps.isspmatrix_csr(sdm._x)
    with pytest.raises(ValueError):           # test_sparse_sdm.py:39-41
        sdm._d
    assert sps.issparse(SDM(sps.coo_matrix(np.eye(3))).x)


def test_empty_shapes_need_no_device():
    # tests/test_eda/test_dense_sdm.py:60-70
    sdm = SDM(np.empty((0, 0)), metric="euclidean")
    assert sdm._d.shape == (0, 0)
    assert sdm._col_sorted_d.shape == (0, 0)
    assert sdm._col_argsorted_d.shape == (0, 0)
    assert len(sdm.sids) == 0 and len(sdm.fids) == 0


def test_partition_is_tile_aligned_and_covers():
    for n in (0, 1, 127, 128, 129, 3005, 100000, 1000000):
        for w in (1, 2, 3, 4, 8):
            parts = partition(n, w)
            assert len(parts) == w and parts[0][0] == 0 and parts[-1][1] == n
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1 and b0 <= e0
            assert all(b % 128 == 0 or b == n for b, _ in parts)
    sizes = [e - b for b, e in partition(100000, 8)]
    assert max(sizes) - min(sizes) <= 128 + 96


def test_piece_layout_and_shards():
    """Row pieces of the pipelined multi-rank upload (stripes.piece_layout /
    take_shard): pieces tile [0, n) on 128 * world boundaries, and gathering
    the ranks' shares piece by piece rebuilds x."""
    import torch
    from scedar_b200.stripes import piece_layout, take_shard
    for n, w, q in ((100000, 8, 8), (100000, 2, 8), (3005, 4, 8), (700, 3, 8),
                    (128, 8, 8), (1000, 1, 4), (5, 2, 3)):
        lay = piece_layout(n, w, q)
        assert lay[0][0] == 0 and lay[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(lay, lay[1:]))
        assert all(b % (128 * w) == 0 for b, _, _ in lay)
        assert all(m * w >= e - b for b, e, m in lay)
        x = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
        shards = [take_shard(x, lay, r, w) for r in range(w)]
        assert all(s.shape[0] == sum(m for _, _, m in lay) for s in shards)
        cap = max(b + w * m for b, _, m in lay)
        xd = torch.full((cap, 3), -1.0, dtype=torch.float64)
        off = 0
        for b, e, m in lay:
            for r in range(w):               # what all_gather_into_tensor does
                xd[b + r * m:b + (r + 1) * m] = shards[r][off:off + m]
            off += m
        assert torch.equal(xd[:n], x)


def test_stripe_pieces_cover_every_tile_once():
    """Pieces of the by-stripe pipelined upload: every row tile belongs to one
    piece, a piece takes a contiguous part of every stripe, in order."""
    from scedar_b200.stripes import partition, stripe_pieces
    for n, world, nq in ((100000, 8, 8), (1900, 3, 8), (130, 2, 4), (3005, 8, 3),
                         (128, 1, 8), (1, 2, 2)):
        parts = partition(n, world)
        if nq == 3:
            from scedar_b200.stripes import GROWING_PIECES
            table, ranges = stripe_pieces(parts, n, cuts=GROWING_PIECES)
            nq = len(GROWING_PIECES)
        else:
            table, ranges = stripe_pieces(parts, n, nq)
        tiles = (n + 127) // 128
        assert len(table) == tiles and all(0 <= q < nq for q in table)
        for r, (b, e) in enumerate(parts):
            pos = b
            for q in range(nq):
                rb, re_ = ranges[q][r]
                if re_ > rb:
                    assert rb == pos and rb % 128 == 0
                    assert all(table[t] == q for t in range(rb // 128, (re_ + 127) // 128))
                    pos = re_
            assert pos == e or e <= b
        # non-decreasing inside a stripe
        for b, e in parts:
            seq = table[b // 128:(e + 127) // 128]
            assert seq == sorted(seq)
