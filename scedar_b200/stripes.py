"""Row-stripe sharding of the distance matrix / kNN path over several GPUs.

One process per GPU (``torch.distributed``, NCCL over NVLink).  Row i of D and
the kNN list of cell i depend on x[i] and all of x and on nothing else, so the
path shards by contiguous, tile-aligned row stripes:

* x is broadcast once from rank 0 (<= 1.6 GB at 100k x 2000);
* every rank prepares the cells (the K1 pass is cheaper than shipping its
  output) and owns rows [row_begin, row_end) of D and of the kNN lists;
* only the (index, distance) lists are all-gathered -- D stripes stay in the
  HBM of their owner;
* the consumers behind the path shard the same way: rare-sample detection
  (``rare_samples``) exchanges one float64 per cell and iteration (the k-th
  neighbour distances, all-gathered; the keep / drop decision is replicated),
  imputation (``impute``) all-gathers the updated rows of x once per iteration.

The compute operators are injected (``ops``) so the sharding logic can be
tested on CPU with the ``gloo`` backend; the default is the CUDA library.
"""
import os

import torch
import torch.distributed as dist

TILE = 128


I8_PRECISIONS = ("fp64_i8", "fp64_i8s6")


def partition(n, world):
    """Tile-aligned contiguous row stripes: list of (row_begin, row_end)."""
    tiles = (n + TILE - 1) // TILE
    base, extra = divmod(tiles, world)
    out, t0 = [], 0
    for r in range(world):
        t1 = t0 + base + (1 if r < extra else 0)
        out.append((min(t0 * TILE, n), min(t1 * TILE, n)))
        t0 = t1
    return out


def shard_rows(n, world):
    """Even split of the rows of x over the ranks' HOST buffers (independent
    of the tile-aligned D stripes): list of (row_begin, row_end)."""
    chunk = (n + world - 1) // world
    return [(min(r * chunk, n), min((r + 1) * chunk, n)) for r in range(world)]


def gather_x(x_rows, n, group=None):
    """All-gather of the ranks' row shards of x (each uploaded over the rank's
    own PCIe link) into the full [n, p] matrix on every rank -- N links feed
    the node instead of one upload on rank 0 followed by a broadcast."""
    world = dist.get_world_size(group)
    chunk = (n + world - 1) // world
    p = x_rows.shape[1]
    pad = x_rows
    if x_rows.shape[0] != chunk:
        pad = torch.zeros((chunk, p), dtype=x_rows.dtype, device=x_rows.device)
        pad[:x_rows.shape[0]] = x_rows
    out = torch.empty((world * chunk, p), dtype=x_rows.dtype,
                      device=x_rows.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n]


def gather_x_from_host(shard_host, n, group=None, shard_dev=None, out=None,
                       device=None, chunks=4):
    """``gather_x`` of row shards that still sit in the ranks' (pinned) HOST
    buffers: the shard goes up in ``chunks`` row ranges on a copy stream and
    every range is all-gathered on a communication stream as soon as it has
    landed, so the NVLink exchange runs under the rest of the PCIe upload.
    Returns the full [n, p] matrix (a view of ``out``, [world * ceil(n / world),
    p]); work queued later on the current stream waits for it."""
    world = dist.get_world_size(group)
    chunk = (n + world - 1) // world
    rows, p = shard_host.shape
    device = device if device is not None else (
        shard_dev.device if shard_dev is not None else shard_host.device)
    if shard_dev is None:
        shard_dev = torch.empty((chunk, p), dtype=shard_host.dtype, device=device)
    if out is None:
        out = torch.empty((world * chunk, p), dtype=shard_host.dtype,
                          device=device)
    assert shard_dev.shape[0] >= chunk and out.shape[0] >= world * chunk
    on_gpu = out.is_cuda
    if on_gpu:
        from . import device as dev_
        main = torch.cuda.current_stream()
        copy, comm = dev_._copy_stream(), dev_._comm_stream()
        copy.wait_stream(main)      # the buffers may still be read by earlier work
        comm.wait_stream(main)
    chunks = max(1, min(chunks, chunk))
    for c in range(chunks):
        c0, c1 = chunk * c // chunks, chunk * (c + 1) // chunks
        if c1 <= c0:
            continue
        h1 = min(c1, rows)          # the last rank's shard may be short
        views = [out[r * chunk + c0:r * chunk + c1] for r in range(world)]
        if on_gpu:
            with torch.cuda.stream(copy):
                if h1 > c0:
                    shard_dev[c0:h1].copy_(shard_host[c0:h1], non_blocking=True)
                up = torch.cuda.Event()
                up.record(copy)
            with torch.cuda.stream(comm):
                comm.wait_event(up)
                dist.all_gather(views, shard_dev[c0:c1], group=group)
        else:
            if h1 > c0:
                shard_dev[c0:h1].copy_(shard_host[c0:h1])
            dist.all_gather(views, shard_dev[c0:c1].contiguous(), group=group)
    if on_gpu:
        main.wait_stream(comm)
    return out[:n]


def piece_layout(n, world, n_pieces=8):
    """Row pieces for the pipelined multi-rank upload of x (``from_host_shards``):
    a list of ``(row_begin, row_end, rows_per_rank)``.  Piece boundaries are
    multiples of 128 * world (the last piece ends at n), so a piece is a whole
    number of row tiles and splits evenly over the ranks; inside piece q rank r
    holds the rows ``[b + r m, min(b + (r + 1) m, e))``, m = ceil((e - b) /
    world) -- the last piece may leave some ranks short (padding rows)."""
    unit = TILE * world
    units = (n + unit - 1) // unit
    n_pieces = max(1, min(n_pieces, units))
    out, u0 = [], 0
    for q in range(n_pieces):
        u1 = units * (q + 1) // n_pieces
        b, e = min(u0 * unit, n), min(u1 * unit, n)
        if e > b:
            out.append((b, e, (e - b + world - 1) // world))
        u0 = u1
    return out


GROWING_PIECES = (1 / 16, 1 / 8, 1 / 4, 1 / 2, 1.0)
"""Cumulative fractions of a stripe that the pieces of the by-stripe pipeline
bring: the work a piece unlocks grows with the square of what has arrived, so
the first piece is small (the contraction starts almost at once) and every
later upload hides behind a longer contraction."""


def stripe_pieces(stripes, n, n_pieces=8, cuts=None):
    """Pieces for the by-stripe pipelined upload (``from_host_stripes``): piece q
    is the q-th part of EVERY stripe, in whole row tiles -- ``n_pieces`` equal
    parts, or up to the cumulative fractions ``cuts`` (ascending, last = 1).
    Returns ``(piece_of_tile, ranges)``: the piece of each of the ceil(n / 128)
    row tiles, and ``ranges[q][r]`` = the rows ``(begin, end)`` of stripe r that
    piece q brings (possibly empty)."""
    if cuts is None:
        cuts = [(q + 1) / n_pieces for q in range(n_pieces)]
    cuts = list(cuts)
    assert cuts and abs(cuts[-1] - 1.0) < 1e-12 and cuts == sorted(cuts)
    n_pieces = len(cuts)
    tiles = (n + TILE - 1) // TILE
    piece_of_tile = [0] * tiles
    ranges = [[(0, 0)] * len(stripes) for _ in range(n_pieces)]
    for r, (b, e) in enumerate(stripes):
        t0, t1 = b // TILE, (e + TILE - 1) // TILE
        if e <= b:
            t1 = t0
        for q in range(n_pieces):
            a0 = t0 + (int((t1 - t0) * cuts[q - 1] + 1e-9) if q else 0)
            a1 = t0 + (int((t1 - t0) * cuts[q] + 1e-9) if q + 1 < n_pieces
                       else t1 - t0)
            for t in range(a0, a1):
                piece_of_tile[t] = q
            ranges[q][r] = (min(a0 * TILE, n), min(a1 * TILE, n))
    return piece_of_tile, ranges


def take_shard(x, layout, rank, world):
    """Rank ``rank``'s host shard of x for ``layout``: its sub-range of every
    piece, one after the other, each padded with zero rows to the piece's
    rows-per-rank (what a per-process loader would hold)."""
    parts = []
    for b, e, m in layout:
        lo, hi = min(b + rank * m, e), min(b + (rank + 1) * m, e)
        blk = x[lo:hi]
        if hi - lo < m:
            pad = torch.zeros((m - (hi - lo), x.shape[1]), dtype=x.dtype,
                              device=blk.device)
            blk = torch.cat([blk, pad], dim=0)
        parts.append(blk)
    return torch.cat(parts, dim=0).contiguous()


class _DeviceOps(object):
    """The CUDA library (default operators)."""

    def prepare(self, x, metric):
        from . import device
        return device.Cells.from_dense(x, metric)

    def prepare_csr(self, indptr, indices, data, shape, metric):
        from . import device
        return device.Cells.from_csr(indptr, indices, data, shape, metric)

    def pdist_stripe(self, cells, r0, r1, out=None, precision="fp64"):
        return cells.pdist_stripe(r0, r1, out=out, precision=precision)

    def knn_from_d(self, d, k, exclude, row_begin):
        from . import device
        return device.knn_from_d(d, k, exclude, row_begin=row_begin)

    # block minima of the INT8 split launches (device.Cells.set_blockmin);
    # SCEDAR_B200_BLOCKMIN=0 selects from the whole of D instead (A/B timing)
    blockmin = os.environ.get("SCEDAR_B200_BLOCKMIN", "1") != "0"

    def new_blockmin(self, rows, n):
        from . import device
        return device.new_blockmin(rows, n)

    def knn_from_d_blockmin(self, cells, d, bmin, k, exclude, row_begin):
        return cells.knn_from_d_blockmin(d, bmin, k, exclude,
                                         row_begin=row_begin)

    def peer_stripes(self, stripes, n, rank, group, blockmin=False):
        from . import device
        return device.PeerStripes(stripes, n, rank=rank, group=group,
                                  blockmin=blockmin)

    def pdist_stripe_p2p(self, cells, peers, rank, precision="fp64"):
        from . import device
        device.pdist_stripe_p2p(cells, peers, rank, precision=precision)

    def knn_stripe(self, cells, k, exclude, r0, r1, precision="fp64"):
        return cells.knn_stripe(k, exclude, r0, r1, precision=precision)

    # -- consumers ------------------------------------------
    def squareform_rows(self, d, n, row_begin, out=None):
        from . import device
        return device.squareform_rows(d, n, row_begin=row_begin, out=out)

    def new_mask(self, n, device):
        return torch.ones(n, dtype=torch.uint8, device=device)

    def masked_kth(self, d, alive, rank, row_begin):
        from . import device
        return device.masked_kth(d, alive, rank, row_begin=row_begin)

    def rare_update(self, kth, cutoff, it, alive, last_kept):
        from . import device
        return device.rare_update(kth, cutoff, it, alive, last_kept)

    def vec_max(self, v):
        from . import device
        return device.masked_max(v, None)

    def impute_iter(self, x_curr, knn, n_do_i, min_present_val, it, x_next,
                    idc, stats, r0, r1):
        from . import device
        device.impute_iter(x_curr, knn, n_do_i, min_present_val, it, x_next,
                           idc, stats, row_begin=r0, row_end=r1)


class StripedDistanceMatrix(object):
    """Distance matrix + kNN of a cells x genes matrix, row-striped over the
    ranks of a process group (or a single process when torch.distributed is
    not initialised).

    Parameters
    ----------
    x : float64 tensor [n, p] on this rank's device, or None on ranks != src
        (it is then received by broadcast).
    shape : (n, p), needed on ranks that pass x=None.
    broadcast : False when every rank already holds x.
    """

    @classmethod
    def from_host(cls, x_host, metric="correlation", out=None, x_dev=None,
                  ops=None, precision="fp64", bmin=None):
        """Single-GPU form for a pinned HOST matrix: the upload of x is hidden
        behind the contraction (``device.pdist_symmetric_from_host``); on
        return the cells are prepared and D (float64, symmetric) is already
        being computed into ``out``.  With several ranks use the row-sharded
        upload (``gather_x``) and the regular constructor instead."""
        if dist.is_available() and dist.is_initialized() and \
                dist.get_world_size() > 1:
            raise ValueError("from_host is the single-GPU form")
        from . import device
        self = cls.__new__(cls)
        self.ops = ops or _DeviceOps()
        self.group, self.distributed, self.rank, self.world = None, False, 0, 1
        self.n, self.p = x_host.shape
        self.metric = metric
        self.precision = precision
        self.stripes = partition(self.n, 1)
        self.row_begin, self.row_end = self.stripes[0]
        self.bmin = None
        if precision in I8_PRECISIONS and getattr(self.ops, "blockmin", False):
            self.bmin = (bmin if bmin is not None
                         else self.ops.new_blockmin(self.n, self.n))
        self.cells, self.d_stripe = device.pdist_symmetric_from_host(
            x_host, metric, out=out, x_dev=x_dev, precision=precision,
            bmin=self.bmin)
        self.d_device = self.d_stripe.device
        return self

    @classmethod
    def from_host_shards(cls, shard_host, layout, shape, peers, metric="correlation",
                         precision="fp64_i8", x_dev=None, shard_dev=None,
                         group=None, reserve_sms=8):
        """Several ranks, x row-sharded over the ranks' pinned HOST buffers
        (``take_shard``): the upload, the all-gather and the contraction run as
        a pipeline over the row pieces of ``layout``.  Every rank copies its
        share of piece q over its own PCIe link on a copy stream, the shares
        are all-gathered over NVLink on a communication stream, and as soon as
        a piece has landed its cells are prepared and the lower-triangle tiles
        of its rows -- which need nothing beyond this piece -- are contracted
        and stored into the owners' stripes (``device.pdist_p2p_rows``, INT8
        split kernels), while the next piece is on its way.  ``reserve_sms``
        SMs are left to the all-gather kernels.  On return D (this rank's
        stripe of ``peers``) is complete on every rank."""
        from . import device
        self = cls.__new__(cls)
        self.ops = _DeviceOps()
        self.group = group
        self.distributed = True
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n, self.p = shape
        n, p, world, rank = self.n, self.p, self.world, self.rank
        self.metric, self.precision = metric, precision
        self.stripes = partition(n, world)
        self.row_begin, self.row_end = self.stripes[rank]
        dev_ = peers.tensor(rank).device
        self.d_device = dev_
        cap = max(b + world * m for b, _, m in layout)
        if x_dev is None:
            x_dev = torch.empty((cap, p), dtype=torch.float64, device=dev_)
        assert x_dev.shape[0] >= cap and x_dev.shape[1] == p
        rows_total = sum(m for _, _, m in layout)
        if shard_dev is None:
            shard_dev = torch.empty((rows_total, p), dtype=torch.float64,
                                    device=dev_)
        main = torch.cuda.current_stream()
        copy, comm = device._copy_stream(), device._comm_stream()
        copy.wait_stream(main)
        comm.wait_stream(main)
        self._fence()              # nobody still reads a stripe of the last step
        self.cells = device.Cells.alloc(n, p, metric)
        self.bmin = None
        if peers.blockmin and precision in I8_PRECISIONS:
            self.cells.set_blockmin(peers.bmin_ptrs(), peers.ldb)
            self.bmin = peers.bmin_tensor(rank)
        off = 0
        for b, e, m in layout:
            with torch.cuda.stream(copy):
                shard_dev[off:off + m].copy_(shard_host[off:off + m],
                                             non_blocking=True)
                up = torch.cuda.Event()
                up.record(copy)
            with torch.cuda.stream(comm):
                comm.wait_event(up)
                dist.all_gather_into_tensor(x_dev[b:b + world * m],
                                            shard_dev[off:off + m], group=group)
                landed = torch.cuda.Event()
                landed.record(comm)
            main.wait_event(landed)
            self.cells.fill_rows(x_dev[b:e], b)
            device.pdist_p2p_rows(self.cells, peers, rank, b, e,
                                  precision=precision, reserve_sms=reserve_sms)
            off += m
        self._fence()              # every peer store has landed
        self.d_stripe = peers.tensor(rank)
        return self

    @classmethod
    def from_host_stripes(cls, shard_host, shape, peers, xpeers,
                          metric="correlation", precision="fp64_i8", group=None,
                          n_pieces=None, reserve_sms=2):
        """Several ranks, x row-sharded BY STRIPE over the ranks' pinned HOST
        buffers (``shard_host`` = the rows of this rank's own D stripe): upload,
        exchange and contraction as a pipeline over pieces, piece q being the
        q-th part of every stripe (``stripe_pieces``: ``n_pieces`` equal parts,
        default the growing ``GROWING_PIECES``).  A rank copies
        its part of piece q over its own PCIe link into its x buffer, pushes it
        into every peer's x buffer with copy-engine transfers over NVLink
        (``xpeers``: ``device.PeerStripes([(0, n)] * world, p)``, one [n, p]
        buffer per rank) and a one-element all-reduce tells everybody that the
        piece is complete; the cells it brought are then prepared and the pairs
        of row tiles whose later tile just arrived are contracted
        (``device.pdist_p2p_pieces``: the ownership and orientation of the
        whole-matrix form, so every rank gets an equal share of every piece and
        only mirrored tiles cross NVLink) while the next piece is on its way.
        ``reserve_sms`` SMs stay free for the all-reduce kernels.  On return D
        (this rank's stripe of ``peers``, with its block minima) is complete on
        every rank, bit-identical to ``compute_d(peers=...)``."""
        from . import device
        self = cls.__new__(cls)
        self.ops = _DeviceOps()
        self.group = group
        self.distributed = True
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n, self.p = shape
        n, p, world, rank = self.n, self.p, self.world, self.rank
        self.metric, self.precision = metric, precision
        self.stripes = partition(n, world)
        self.row_begin, self.row_end = self.stripes[rank]
        assert tuple(shard_host.shape) == (self.row_end - self.row_begin, p)
        if n_pieces is None:
            piece_of_tile, ranges = stripe_pieces(self.stripes, n,
                                                  cuts=GROWING_PIECES)
        else:
            piece_of_tile, ranges = stripe_pieces(self.stripes, n, n_pieces)
        n_pieces = len(ranges)
        xs = [xpeers.tensor(r) for r in range(world)]       # [n, p] per rank
        mine = xs[rank]
        self.d_device = mine.device
        main = torch.cuda.current_stream()
        copy, comm = device._copy_stream(), device._comm_stream()
        self._fence()          # nobody still reads a D stripe or an x buffer
        copy.wait_stream(main)
        comm.wait_stream(main)
        self.cells = device.Cells.alloc(n, p, metric)
        self.bmin = None
        if getattr(peers, "blockmin", False) and precision in I8_PRECISIONS:
            self.cells.set_blockmin(peers.bmin_ptrs(), peers.ldb)
            self.bmin = peers.bmin_tensor(rank)
        for q in range(n_pieces):
            b, e = ranges[q][rank]
            with torch.cuda.stream(copy):
                if e > b:
                    mine[b:e].copy_(shard_host[b - self.row_begin:e - self.row_begin],
                                    non_blocking=True)
                up = torch.cuda.Event()
                up.record(copy)
            with torch.cuda.stream(comm):
                comm.wait_event(up)
                if e > b:
                    for r in range(world):
                        if r != rank:
                            xs[r][b:e].copy_(mine[b:e], non_blocking=True)
                self._fence()          # on comm: every rank has pushed piece q
                landed = torch.cuda.Event()
                landed.record(comm)
            main.wait_event(landed)
            for r in range(world):
                rb, re_ = ranges[q][r]
                if re_ > rb:
                    self.cells.fill_rows(mine[rb:re_], rb, with_stat=False)
            device.pdist_p2p_pieces(
                self.cells, peers, rank, piece_of_tile, q, n_pieces,
                precision=precision,
                reserve_sms=reserve_sms if q + 1 < n_pieces else 0)
        self.cells.stat_rows()     # one pass: the handle is complete for any kernel
        self._fence()              # every peer store has landed
        self.d_stripe = peers.tensor(rank)
        return self

    @classmethod
    def from_csr(cls, csr, metric="cosine", shape=None, nnz=None, group=None,
                 src=0, device=None, ops=None, broadcast=True):
        """CSR cells (cosine / euclidean): ``csr = (indptr int64 [n+1], indices
        int32 [nnz], data float64 [nnz])`` device tensors on the source rank
        (None elsewhere; ``shape``, ``nnz`` and ``device`` are then required).
        The three arrays are broadcast and every rank scatters them into its
        own dense working copy."""
        self = cls.__new__(cls)
        self.ops = ops or _DeviceOps()
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        n, p = shape
        if csr is None:
            csr = (torch.empty(n + 1, dtype=torch.int64, device=device),
                   torch.empty(nnz, dtype=torch.int32, device=device),
                   torch.empty(nnz, dtype=torch.float64, device=device))
        if self.world > 1 and broadcast:
            for t in csr:
                dist.broadcast(t, src=src, group=group)
        self.n, self.p = n, p
        self.d_device = csr[2].device
        self.metric = metric
        self.precision = "fp64"
        self.stripes = partition(self.n, self.world)
        self.row_begin, self.row_end = self.stripes[self.rank]
        self.cells = self.ops.prepare_csr(csr[0], csr[1], csr[2], (n, p), metric)
        self.d_stripe = None
        self.bmin = None
        return self

    def __init__(self, x, metric="correlation", shape=None, group=None,
                 src=0, device=None, ops=None, broadcast=True,
                 precision="fp64"):
        """``precision``: Gram kernel of the D stripes -- "fp64" (DMMA) or
        "fp64_i8" / "fp64_i8s6" (INT8 split on tcgen05, same 1e-10 bar)."""
        self.ops = ops or _DeviceOps()
        self.precision = precision
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        if x is None:
            if shape is None or device is None:
                raise ValueError("ranks without x need shape and device")
            x = torch.empty(shape, dtype=torch.float64, device=device)
        if x.dim() != 2 or x.dtype != torch.float64:
            raise ValueError("x has shape (n_samples, n_features), float64")
        if self.world > 1 and broadcast:
            dist.broadcast(x, src=src, group=group)
        self.n, self.p = x.shape
        self.d_device = x.device
        self.metric = metric
        self.stripes = partition(self.n, self.world)
        self.row_begin, self.row_end = self.stripes[self.rank]
        self.cells = self.ops.prepare(x, metric)
        self.d_stripe = None
        self.bmin = None

    @property
    def rows(self):
        return self.row_end - self.row_begin

    def compute_d(self, out=None, peers=None, bmin=None):
        """This rank's rows of D, kept in HBM: tensor [rows, n].

        With ``peers`` (a ``device.PeerStripes`` shared by all ranks) the
        contraction is the balanced symmetric one: every tile pair of the
        whole matrix is computed by one rank and mirrored into the other
        owner's stripe over NVLink.  The two barriers fence the peer writes:
        nobody may still be reading a stripe when the stores start, and every
        store must have landed before anyone reads."""
        extra = {} if self.precision == "fp64" else {
            "precision": self.precision}
        # the INT8 split kernels leave the minimum of every run of 32 entries
        # behind; the selection that follows then skips almost all of D
        use_bmin = (self.precision in I8_PRECISIONS
                    and getattr(self.ops, "blockmin", False))
        self.bmin = None
        if peers is not None and self.world > 1:
            if use_bmin and getattr(peers, "blockmin", False):
                self.cells.set_blockmin(peers.bmin_ptrs(), peers.ldb)
                self.bmin = peers.bmin_tensor(self.rank)
            self._fence()
            self.ops.pdist_stripe_p2p(self.cells, peers, self.rank, **extra)
            self._fence()
            self.d_stripe = peers.tensor(self.rank)
            return self.d_stripe
        if use_bmin:
            if bmin is None:
                bmin = self.ops.new_blockmin(self.rows, self.n)
            self.cells.set_blockmin([bmin])
            self.bmin = bmin
        self.d_stripe = self.ops.pdist_stripe(self.cells, self.row_begin,
                                              self.row_end, out=out, **extra)
        return self.d_stripe

    def _fence(self):
        """Stream-ordered cross-rank barrier (a one-element all-reduce): no
        host synchronisation, later kernels on this stream wait for it."""
        if getattr(self, "_flag", None) is None:
            self._flag = torch.zeros(1, dtype=torch.int32,
                                     device=self.d_device)
        dist.all_reduce(self._flag, group=self.group)

    def make_peer_stripes(self):
        """Allocate this rank's peer-mappable D stripe (followed by its block
        minima when the INT8 split kernels run) and map the others."""
        if self.precision in I8_PRECISIONS and getattr(self.ops, "blockmin", False):
            return self.ops.peer_stripes(self.stripes, self.n, self.rank,
                                         self.group, blockmin=True)
        return self.ops.peer_stripes(self.stripes, self.n, self.rank,
                                     self.group)

    def knn_local(self, k, exclude, from_d=True, precision="fp64"):
        """(idx[rows, k] int64, dist[rows, k]) of this rank's rows.
        ``precision="fast"`` (only without D): tensor-core candidates + FP64
        re-rank, the same lists bit for bit."""
        if from_d:
            if self.d_stripe is None:
                self.compute_d()
            if self.bmin is not None:
                return self.ops.knn_from_d_blockmin(
                    self.cells, self.d_stripe, self.bmin, k, exclude,
                    self.row_begin)
            return self.ops.knn_from_d(self.d_stripe, k, exclude,
                                       self.row_begin)
        if precision == "fp64":
            return self.ops.knn_stripe(self.cells, k, exclude, self.row_begin,
                                       self.row_end)
        return self.ops.knn_stripe(self.cells, k, exclude, self.row_begin,
                                   self.row_end, precision=precision)

    def knn(self, k, exclude, from_d=True, precision="fp64"):
        """kNN lists of ALL cells on every rank: the all-gather of the
        per-stripe lists (the only data-path collective)."""
        idx, dd = self.knn_local(k, exclude, from_d=from_d,
                                 precision=precision)
        if self.world == 1:
            return idx, dd
        return self._gather(idx), self._gather(dd)

    # -- consumers behind the path ---------------------------
    def rare_samples(self, k, d_cutoff, n_iter):
        """``RareSampleDetection._pdist_rare_s_detect``
        (scedar/knn/detection.py:81-123) on the row-striped D: every rank
        takes the ``min(m-1, k)``-th smallest live entry of its own rows, the
        n values are all-gathered, and the keep / drop decision is applied
        identically on every rank.  Returns ``last_kept`` (int32 [n]): the
        last iteration that kept each cell (== n_iter for the non-rare)."""
        if self.d_stripe is None:
            self.compute_d()
        n = self.n
        alive = self.ops.new_mask(n, self.d_device)
        last_kept = torch.zeros(n, dtype=torch.int32, device=self.d_device)

        def kth_all(rank):
            mine = self.ops.masked_kth(self.d_stripe, alive, rank,
                                       self.row_begin)
            return mine if self.world == 1 else self._gather(mine)

        kth = kth_all(k)
        d_max = self.ops.vec_max(kth)
        m, fresh_rank = n, k
        for i in range(1, n_iter + 1):
            cutoff = d_cutoff + (n_iter - i) / n_iter * max(0, d_max - d_cutoff)
            if m == 0:
                continue
            i_k = min(m - 1, k)
            if fresh_rank != i_k:
                kth = kth_all(i_k)
                fresh_rank = i_k
            kept = self.ops.rare_update(kth, cutoff, i, alive, last_kept)
            if kept != m:
                fresh_rank = None
            m = kept
        return last_kept

    def impute(self, x, knn, k, n_do, min_present_val, n_iter):
        """``FeatureImputation._impute_features_runner``
        (scedar/knn/imputation.py:35-135, median statistic) with the cells
        row-striped: every rank imputes its own rows from the full current x
        and the updated rows are all-gathered after each iteration.  ``x``
        [n, p] float64 and ``knn`` [n, k] int64 are replicated.  Returns
        (imputed x, indicator int32 [n, p], (entries, cells) picked up)."""
        import math
        n, p = x.shape
        curr = x.clone(memory_format=torch.contiguous_format)
        nxt = torch.empty_like(curr)
        idc = torch.zeros((n, p), dtype=torch.int32, device=x.device)
        stats = torch.zeros(2, dtype=torch.int64, device=x.device)
        r0, r1 = self.row_begin, self.row_end
        for i in range(1, n_iter + 1):
            n_do_i = n_do + int(math.ceil((n_iter - i) / n_iter * (k - n_do)))
            self.ops.impute_iter(curr, knn, n_do_i, min_present_val, i, nxt,
                                 idc, stats, r0, r1)
            if self.world > 1:
                self._exchange_imputed(curr, nxt, idc, i, r0, r1)
            else:
                curr, nxt = nxt, curr
        if self.world > 1:
            idc = self._gather(idc[r0:r1])
            dist.all_reduce(stats, group=self.group)
        return curr, idc, tuple(int(v) for v in stats.tolist())

    def _exchange_imputed(self, curr, nxt, idc, it, r0, r1):
        """After iteration ``it`` every rank holds the new values of its own
        rows in ``nxt``; they differ from ``curr`` exactly where the indicator
        says ``it``.  Only those (position, value) pairs are exchanged and
        written into every rank's ``curr`` -- bytes proportional to what the
        iteration picked up, not to n * p."""
        p = curr.shape[1]
        own = (idc[r0:r1] == it).reshape(-1).nonzero().squeeze(1)
        vals = nxt[r0:r1].reshape(-1)[own]
        own = own + r0 * p                       # positions in the whole matrix
        cnt = torch.tensor([own.numel()], dtype=torch.int64, device=curr.device)
        cnts = torch.empty(self.world, dtype=torch.int64, device=curr.device)
        dist.all_gather_into_tensor(cnts, cnt, group=self.group)
        cnts = cnts.tolist()
        m = max(cnts)
        if m == 0:
            return
        pos_pad = torch.zeros(m, dtype=torch.int64, device=curr.device)
        val_pad = torch.zeros(m, dtype=curr.dtype, device=curr.device)
        pos_pad[:own.numel()] = own
        val_pad[:own.numel()] = vals
        pos_all = torch.empty(self.world * m, dtype=torch.int64, device=curr.device)
        val_all = torch.empty(self.world * m, dtype=curr.dtype, device=curr.device)
        dist.all_gather_into_tensor(pos_all, pos_pad, group=self.group)
        dist.all_gather_into_tensor(val_all, val_pad, group=self.group)
        flat = curr.view(-1)
        for r, c in enumerate(cnts):
            if c:
                flat[pos_all[r * m:r * m + c]] = val_all[r * m:r * m + c]

    def condensed(self, dst=0):
        """``scipy.spatial.distance.squareform(D)`` (the input of
        ``sch.linkage``, sdm.py:1666) assembled on rank ``dst``: every rank
        packs the strict upper triangle of its own rows -- one contiguous
        segment of the vector -- and sends exactly that segment (rank 0's is
        the longest, the last rank's the shortest: a padded gather would move
        up to twice the bytes).  Returns the float64 device vector
        [n(n-1)/2] on ``dst``, None elsewhere; what scipy's validity check
        rejects raises ``ValueError`` on every rank."""
        from . import hclust
        if self.d_stripe is None:
            self.compute_d()
        n = self.n
        base = [b * (2 * n - b - 1) // 2 for b, _ in self.stripes]
        base.append(n * (n - 1) // 2 if n > 1 else 0)
        total = base[-1]
        if self.world == 1:
            seg, flags = self.ops.squareform_rows(self.d_stripe, n, 0)
            hclust._condensed_flags_check(int(flags.item()))
            return seg
        out = None
        if self.rank == dst:
            out = torch.empty(total, dtype=torch.float64,
                              device=self.d_stripe.device)
            mine = out[base[self.rank]:base[self.rank + 1]]
            seg, flags = self.ops.squareform_rows(
                self.d_stripe, n, self.row_begin, out=mine)
            if seg.data_ptr() != mine.data_ptr():
                mine.copy_(seg)
            for r in range(self.world):
                if r != dst and base[r + 1] > base[r]:
                    dist.recv(out[base[r]:base[r + 1]], src=self._global(r),
                              group=self.group)
        else:
            seg, flags = self.ops.squareform_rows(self.d_stripe, n,
                                                  self.row_begin)
            if seg.numel():
                dist.send(seg.contiguous(), dst=self._global(dst),
                          group=self.group)
        # NCCL has no bitwise reduction: one slot per flag bit, MAX
        bits = torch.stack([(flags.reshape(-1)[0] & 1), (flags.reshape(-1)[0] & 2)]
                           ).to(torch.int32)
        dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=self.group)
        bits = bits.tolist()
        hclust._condensed_flags_check((1 if bits[0] else 0) |
                                      (2 if bits[1] else 0))
        return out

    def _global(self, group_rank):
        """Rank in the default group of a rank of ``self.group``."""
        if self.group is None:
            return group_rank
        return dist.get_global_rank(self.group, group_rank)

    def hclust_linkage(self, linkage="complete", n_eval_rounds=None,
                       is_euc_dist=False, optimal_ordering=False,
                       verbose=False, dst=0):
        """``HClustTree.hclust_linkage`` (sdm.py:1629-1669) of the striped D:
        the condensed vector is assembled on rank ``dst``, scipy's
        (sequential) linkage runs there, and the (n-1) x 4 linkage matrix is
        broadcast, so every rank returns it."""
        from . import hclust
        sf = self.condensed(dst=dst)
        z = None
        if self.rank == dst or self.world == 1:
            z = hclust.linkage_from_condensed(
                sf.cpu().numpy(), self.n, linkage=linkage,
                n_eval_rounds=n_eval_rounds, is_euc_dist=is_euc_dist,
                optimal_ordering=optimal_ordering, verbose=verbose)
        if self.world > 1:
            box = [z]
            dist.broadcast_object_list(box, src=self._global(dst),
                                       group=self.group)
            z = box[0]
        return z

    def _gather(self, t):
        """all-gather of ragged stripes: pad to the widest stripe, trim."""
        widest = max(e - b for b, e in self.stripes)
        pad = torch.zeros((widest,) + tuple(t.shape[1:]), dtype=t.dtype,
                          device=t.device)
        pad[:t.shape[0]] = t
        out = torch.empty((self.world * widest,) + tuple(t.shape[1:]),
                          dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        parts = [out[r * widest:r * widest + (e - b)]
                 for r, (b, e) in enumerate(self.stripes)]
        return torch.cat(parts, dim=0)

    def close(self):
        if hasattr(self.cells, "close"):
            self.cells.close()
        self.d_stripe = None
