"""Device-level operators: torch tensors in HBM in, torch tensors out.

torch is only the allocator / stream provider here; every kernel is in
``csrc/`` behind the C ABI.  All functions run on the current CUDA device and
the current torch stream.
"""
import ctypes
import os

import numpy as np
import torch

from . import _capi
from ._capi import EXCLUDE_FIRST, EXCLUDE_SELF, METRIC_IDS, PRECISION_IDS

_INIT_PID = os.getpid()


def _guard():
    """A CUDA context does not survive fork(); the reference's ``parmap``
    (scedar/utils.py:45-52) forks.  Refuse instead of crashing the GPU."""
    if os.getpid() != _INIT_PID and torch.cuda.is_initialized():
        raise RuntimeError(
            "scedar_b200: CUDA was initialised in the parent process; the GPU "
            "path cannot be entered from a forked child (run with nprocs=1)")
    if not torch.cuda.is_available():
        raise RuntimeError("scedar_b200: no CUDA device -- this package has no "
                           "CPU fallback")


# Every tensor the host modules allocate themselves goes through these three,
# so the package has exactly one place that names the device.
DEVICE = "cuda"


def zeros(shape, dtype):
    _guard()
    return torch.zeros(shape, dtype=dtype, device=DEVICE)


def ones(shape, dtype):
    _guard()
    return torch.ones(shape, dtype=dtype, device=DEVICE)


def empty(shape, dtype):
    _guard()
    return torch.empty(shape, dtype=dtype, device=DEVICE)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _f64(t):
    assert t.is_cuda and t.dtype == torch.float64, "need a CUDA float64 tensor"
    return t


class Cells(object):
    """Prepared cells x genes matrix resident in HBM (``scedar_cells_t``).

    The constructor arguments are kept alive until the preparation kernels
    have been enqueued; the handle owns its own padded working copy.
    """

    def __init__(self, handle, n, p, metric):
        self._h = handle
        self.n, self.p, self.metric = int(n), int(p), metric

    @classmethod
    def from_dense(cls, x, metric):
        _guard()
        lib = _capi.load()
        x = _f64(x)
        if x.dim() != 2:
            raise ValueError("x has shape (n_samples, n_features)")
        if x.stride(1) != 1 and x.shape[1] > 1:
            x = x.contiguous()
        n, p = x.shape
        ldx = x.stride(0) if n > 1 and p > 0 else max(p, 1)
        if ldx < p:
            x, ldx = x.contiguous(), max(p, 1)
        h = ctypes.c_void_p()
        _capi.check(lib.scedar_cells_from_dense(
            _ptr(x), n, p, ldx, METRIC_IDS[metric], _stream(),
            ctypes.byref(h)))
        return cls(h, n, p, metric)

    @classmethod
    def alloc(cls, n, p, metric):
        """Empty cells to be filled range by range (``fill_rows``)."""
        _guard()
        h = ctypes.c_void_p()
        _capi.check(_capi.load().scedar_cells_alloc(
            int(n), int(p), METRIC_IDS[metric], _stream(), ctypes.byref(h)))
        return cls(h, n, p, metric)

    def fill_rows(self, x_rows, row_begin, with_stat=True):
        """Prepare the cells [row_begin, row_begin + len(x_rows)) from a
        float64 device tensor (128-aligned ranges, the last one ends at n).
        ``with_stat=False`` leaves the Gram diagonal (the statistic of the DMMA
        and 3xFP16 kernels; the INT8 split kernels do not read it) to a later
        ``stat_rows`` pass."""
        x_rows = _f64(x_rows)
        if x_rows.dim() != 2 or x_rows.shape[1] != self.p:
            raise ValueError("x_rows has shape (rows, n_features)")
        if x_rows.stride(1) != 1 and self.p > 1:
            x_rows = x_rows.contiguous()
        ldx = x_rows.stride(0) if x_rows.shape[0] > 1 else max(self.p, 1)
        _capi.check(_capi.load().scedar_cells_fill_rows_ex(
            self._h, _ptr(x_rows), max(ldx, self.p), int(row_begin),
            int(row_begin) + x_rows.shape[0], int(bool(with_stat)), _stream()))

    def stat_rows(self, row_begin=0, row_end=None):
        """The Gram diagonal of rows filled with ``with_stat=False``."""
        row_end = self.n if row_end is None else row_end
        _capi.check(_capi.load().scedar_cells_stat_rows(
            self._h, int(row_begin), int(row_end), _stream()))

    def pdist_lower_rows(self, row_begin, row_end, out, precision="fp64"):
        """Rows [row_begin, row_end) of the lower triangle of the whole
        matrix ``out`` [n, n], mirrored; needs the cells [0, row_end) only."""
        _capi.check(_capi.load().scedar_pdist_lower_rows_prec(
            self._h, PRECISION_IDS[precision], int(row_begin), int(row_end),
            _ptr(out), out.stride(0) if self.n > 1 else self.n, _stream()))
        return out

    @classmethod
    def from_csr(cls, indptr, indices, data, shape, metric):
        _guard()
        lib = _capi.load()
        n, p = shape
        assert indptr.dtype in (torch.int32, torch.int64)
        assert indices.dtype == torch.int32 and data.dtype == torch.float64
        h = ctypes.c_void_p()
        _capi.check(lib.scedar_cells_from_csr(
            _ptr(indptr), int(indptr.dtype == torch.int64), _ptr(indices),
            _ptr(data), n, p, METRIC_IDS[metric], _stream(), ctypes.byref(h)))
        torch.cuda.current_stream().synchronize()  # inputs may now be freed
        return cls(h, n, p, metric)

    def close(self):
        if self._h is not None and self._h.value:
            _capi.load().scedar_cells_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- block minima (by-product of the INT8 split launches) ----------------
    def set_blockmin(self, bufs, ldb=None):
        """Attach one block-minima buffer per output stripe of the launches
        that follow (tensors [rows, ldb] or raw device pointers); ``None`` /
        empty detaches.  Only the INT8 split kernels fill them."""
        bufs = list(bufs or [])
        ptrs = [b.data_ptr() if torch.is_tensor(b) else int(b) for b in bufs]
        if bufs and ldb is None:
            ldb = bufs[0].stride(0)
        table = (ctypes.c_void_p * max(len(ptrs), 1))(*ptrs)
        _capi.check(_capi.load().scedar_cells_set_blockmin(
            self._h, table, len(ptrs), int(ldb or 0)))
        self._bmin_keep = bufs          # the handle holds raw pointers

    def knn_from_d_blockmin(self, d, bmin, k, exclude=EXCLUDE_SELF, row_begin=0):
        """``knn_from_d`` of rows that an INT8 split launch of this handle wrote
        together with their block minima ``bmin`` [rows, ldb]: reads n / 32
        minima and the few runs of 32 entries that can hold a neighbour."""
        rows, n = d.shape
        assert n == self.n and bmin.shape[0] == rows
        idx = torch.empty((rows, k), dtype=torch.int64, device="cuda")
        dist = torch.empty((rows, k), dtype=torch.float64, device="cuda")
        _capi.check(_capi.load().scedar_knn_from_d_blockmin(
            self._h, _ptr(d), d.stride(0) if rows > 1 else n, _ptr(bmin),
            bmin.stride(0) if rows > 1 else bmin.shape[1], row_begin,
            row_begin + rows, k, exclude, _ptr(idx), _ptr(dist), _stream()))
        return idx, dist

    # -- distance matrix ----------------------------------------------------
    def pdist_stripe(self, row_begin=0, row_end=None, out=None,
                     precision="fp64"):
        """Rows [row_begin, row_end) of D -> float64 tensor [rows, n]."""
        row_end = self.n if row_end is None else row_end
        rows = row_end - row_begin
        if out is None:
            out = torch.empty((max(rows, 0), self.n), dtype=torch.float64,
                              device="cuda")
        _capi.check(_capi.load().scedar_pdist_stripe(
            self._h, PRECISION_IDS[precision], row_begin, row_end, _ptr(out),
            out.stride(0) if out.dim() == 2 and out.shape[0] > 1 else self.n,
            _stream()))
        return out

    def pdist_symmetric(self, out=None, precision="fp64"):
        """All of D, lower-triangle tiles computed once and mirrored."""
        if out is None:
            out = torch.empty((self.n, self.n), dtype=torch.float64,
                              device="cuda")
        _capi.check(_capi.load().scedar_pdist_symmetric(
            self._h, PRECISION_IDS[precision], _ptr(out),
            out.stride(0) if self.n > 1 else self.n, _stream()))
        return out

    # -- PCA producer pieces --------------------------
    def col_means(self):
        """Mean of every column of the working copy over the real rows."""
        out = torch.empty(self.p, dtype=torch.float64, device="cuda")
        _capi.check(_capi.load().scedar_cells_col_means(
            self._h, _ptr(out), _stream()))
        return out

    def shift_cols(self, shift):
        """working copy -= shift[column], in place (column centring)."""
        shift = _f64(shift).contiguous()
        assert shift.numel() == self.p
        _capi.check(_capi.load().scedar_cells_shift_cols(
            self._h, _ptr(shift), _stream()))

    def transposed(self, shift=None):
        """New handle over the transposed working copy (rows = features),
        optionally minus ``shift[feature]``."""
        if shift is not None:
            shift = _f64(shift).contiguous()
            assert shift.numel() == self.p
        h = ctypes.c_void_p()
        _capi.check(_capi.load().scedar_cells_transposed(
            self._h, _ptr(shift), _stream(), ctypes.byref(h)))
        torch.cuda.current_stream().synchronize()   # shift may now be freed
        return Cells(h, self.p, self.n, self.metric)

    def gram_raw(self, out=None):
        """Bare Gram matrix of the prepared rows [n, n] (FP64 tensor cores)."""
        if out is None:
            out = torch.empty((self.n, self.n), dtype=torch.float64,
                              device="cuda")
        _capi.check(_capi.load().scedar_gram_raw(
            self._h, _ptr(out), out.stride(0) if self.n > 1 else self.n,
            _stream()))
        return out

    def gram_raw_cross(self, other, out=None):
        """self.rows . other.rows^T -> [self.n, other.n]."""
        if out is None:
            out = torch.empty((self.n, other.n), dtype=torch.float64,
                              device="cuda")
        _capi.check(_capi.load().scedar_gram_raw_cross(
            self._h, other._h, _ptr(out),
            out.stride(0) if self.n > 1 else other.n, _stream()))
        return out

    # -- kNN without D ------------------------------------------------------
    def knn_stripe(self, k, exclude=EXCLUDE_SELF, row_begin=0, row_end=None,
                   precision="fp64"):
        row_end = self.n if row_end is None else row_end
        rows = max(row_end - row_begin, 0)
        idx = torch.empty((rows, k), dtype=torch.int64, device="cuda")
        dist = torch.empty((rows, k), dtype=torch.float64, device="cuda")
        _capi.check(_capi.load().scedar_knn_stripe(
            self._h, PRECISION_IDS[precision], k, exclude, row_begin, row_end,
            _ptr(idx), _ptr(dist), _stream()))
        return idx, dist


def blockmin_ld(n):
    """Row stride of a block-minima buffer: ceil(n / 32) rounded up to 4."""
    return ((n + 31) // 32 + 3) // 4 * 4


def new_blockmin(rows, n):
    return torch.empty((max(rows, 0), blockmin_ld(n)), dtype=torch.float64,
                       device="cuda")


def knn_from_d(d, k, exclude=EXCLUDE_SELF, row_begin=0):
    """Top-k of every row of a materialised stripe ``d`` [rows, n]; row r is
    cell ``row_begin + r``."""
    _guard()
    d = _f64(d)
    rows, n = d.shape
    idx = torch.empty((rows, k), dtype=torch.int64, device="cuda")
    dist = torch.empty((rows, k), dtype=torch.float64, device="cuda")
    _capi.check(_capi.load().scedar_knn_from_d(
        _ptr(d), d.stride(0) if rows > 1 else n, n, row_begin,
        row_begin + rows, k, exclude, _ptr(idx), _ptr(dist), _stream()))
    return idx, dist


def col_sort(d, want_sorted=True, want_arg=True):
    """Stable sort of every row of a stripe ``d`` [rows, n] of a symmetric D
    (row j of D == column j).  Returns cell-major [rows, n] tensors -- the
    transpose of the reference's (rank, cell) layout; either may be None."""
    _guard()
    d = _f64(d)
    rows, n = d.shape
    srt = (torch.empty((rows, n), dtype=torch.float64, device="cuda")
           if want_sorted else None)
    arg = (torch.empty((rows, n), dtype=torch.int64, device="cuda")
           if want_arg else None)
    _capi.check(_capi.load().scedar_col_sort(
        _ptr(d), d.stride(0) if rows > 1 else max(n, 1), rows, n, _ptr(srt),
        _ptr(arg), _stream()))
    return srt, arg


def num_correct_dist_mat_(d, upper_bound=None):
    """In-place clean-up of a device matrix; returns the warning flag bits."""
    _guard()
    d = _f64(d)
    n = d.shape[0]
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    _capi.check(_capi.load().scedar_num_correct_dist_mat(
        _ptr(d), d.stride(0) if n > 1 else max(n, 1), n,
        int(upper_bound is not None),
        float(upper_bound) if upper_bound is not None else 0.0, _ptr(flags),
        _stream()))
    return int(flags.item())


def to_device(a):
    """Host ndarray -> CUDA tensor.  A one-shot copy from pageable memory goes
    through the driver's own staging buffers; pinning the source first
    (cudaHostAlloc + a host copy) costs several times the transfer itself
    (measured: 548 -> 94 ms for the 480 MB of config 1)."""
    _guard()
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def to_host(t, out):
    """Device tensor -> the caller's ndarray (one D2H copy, no staging array)."""
    torch.from_numpy(out).copy_(t)
    return out


class _RawCuda(object):
    """Exposes a raw device allocation through __cuda_array_interface__ so
    torch can view it without a copy."""

    def __init__(self, ptr, shape, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
            "version": 3, "strides": None}


class PeerStripes(object):
    """The D stripes of all ranks of one node, each allocated by its owner with
    ``scedar_stripe_alloc`` and mapped into every peer process through CUDA
    IPC, so a Gram epilogue can store a mirrored tile straight into the
    owner's HBM over NVLink (``Cells.pdist_stripe_p2p``).

    ``stripes``: list of (row_begin, row_end) per rank (stripes.partition).
    With ``group=None`` and ``rank=None`` all stripes are allocated locally
    (single-process emulation of the N ranks on one GPU, used by the tests).
    """

    def __init__(self, stripes, n, rank=None, group=None, blockmin=False):
        """``blockmin``: every stripe is followed, in the same allocation, by its
        block-minima buffer [rows, ldb] (``Cells.set_blockmin``)."""
        _guard()
        import torch.distributed as dist
        self.lib = _capi.load()
        self.stripes, self.n, self.rank = list(stripes), int(n), rank
        self.world = len(self.stripes)
        self.blockmin = bool(blockmin)
        self.ldb = blockmin_ld(self.n)
        self.ptrs = [None] * self.world
        self._own, self._opened = [], []
        mine = range(self.world) if rank is None else [rank]
        for r in mine:
            b, e = self.stripes[r]
            p = ctypes.c_void_p()
            extra = max(e - b, 0) * self.ldb * 8 if self.blockmin else 0
            _capi.check(self.lib.scedar_stripe_alloc(
                max(e - b, 0) * self.n * 8 + extra, ctypes.byref(p)))
            self.ptrs[r] = p.value
            self._own.append(p.value)
        if rank is not None and self.world > 1:
            blob = ctypes.create_string_buffer(64)
            _capi.check(self.lib.scedar_stripe_export(
                ctypes.c_void_p(self.ptrs[rank]), blob))
            handles = [None] * self.world
            dist.all_gather_object(handles, blob.raw, group=group)
            for r, h in enumerate(handles):
                if r == rank:
                    continue
                p = ctypes.c_void_p()
                buf = ctypes.create_string_buffer(h, 64)
                _capi.check(self.lib.scedar_stripe_open(buf, ctypes.byref(p)))
                self.ptrs[r] = p.value
                self._opened.append(p.value)

    def tensor(self, r=None):
        """Zero-copy torch view [rows, n] of stripe r (default: own)."""
        r = self.rank if r is None else r
        b, e = self.stripes[r]
        if e - b == 0:
            return torch.empty((0, self.n), dtype=torch.float64, device="cuda")
        return torch.as_tensor(_RawCuda(self.ptrs[r], (e - b, self.n), self),
                               device="cuda")

    def bmin_ptrs(self):
        """Device pointers of the stripes' block-minima buffers."""
        assert self.blockmin
        return [p + max(e - b, 0) * self.n * 8
                for p, (b, e) in zip(self.ptrs, self.stripes)]

    def bmin_tensor(self, r=None):
        r = self.rank if r is None else r
        b, e = self.stripes[r]
        if e - b == 0:
            return torch.empty((0, self.ldb), dtype=torch.float64, device="cuda")
        return torch.as_tensor(
            _RawCuda(self.bmin_ptrs()[r], (e - b, self.ldb), self), device="cuda")

    def tables(self):
        begins = [b for b, _ in self.stripes] + [self.stripes[-1][1]]
        return ((ctypes.c_int64 * len(begins))(*begins),
                (ctypes.c_void_p * self.world)(*self.ptrs))

    def close(self):
        torch.cuda.synchronize()
        for p in self._opened:
            self.lib.scedar_stripe_close(ctypes.c_void_p(p))
        for p in self._own:
            self.lib.scedar_stripe_free(ctypes.c_void_p(p))
        self._opened, self._own = [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pdist_stripe_p2p(cells, peers, rank, precision="fp64"):
    # precision: "fp64" (DMMA) or "fp64_i8" / "fp64_i8s6" (INT8 split)
    """Rank ``rank``'s share of the balanced symmetric contraction: its tiles
    go to its own stripe, their transposes to the peers' stripes.  All ranks
    must barrier before reading (stripes.StripedDistanceMatrix does)."""
    begins, ptrs = peers.tables()
    _capi.check(_capi.load().scedar_pdist_stripe_p2p(
        cells._h, PRECISION_IDS[precision], rank, peers.world, begins, ptrs,
        peers.n, _stream()))


def pdist_p2p_rows(cells, peers, rank, row_begin, row_end, precision="fp64_i8",
                   reserve_sms=0):
    """Row-range form of ``pdist_stripe_p2p`` (INT8 split kernels): the
    lower-triangle tiles of rows [row_begin, row_end) this rank takes, stored
    with their transposes into the owners' stripes; needs the cells
    [0, row_end) only."""
    begins, ptrs = peers.tables()
    _capi.check(_capi.load().scedar_pdist_p2p_rows(
        cells._h, PRECISION_IDS[precision], rank, peers.world, begins, ptrs,
        peers.n, int(row_begin), int(row_end), int(reserve_sms), _stream()))


def pdist_p2p_pieces(cells, peers, rank, piece_of_tile, piece, n_pieces,
                     precision="fp64_i8", reserve_sms=0):
    """Piece form of ``pdist_stripe_p2p`` (INT8 split kernels): the pairs of
    row tiles whose later tile arrived with ``piece`` (``piece_of_tile``: int32
    per row tile), owned and oriented as in the whole-matrix form."""
    begins, ptrs = peers.tables()
    table = (ctypes.c_int32 * len(piece_of_tile))(*[int(v) for v in piece_of_tile])
    _capi.check(_capi.load().scedar_pdist_p2p_pieces(
        cells._h, PRECISION_IDS[precision], rank, peers.world, begins, ptrs,
        peers.n, table, int(piece), int(n_pieces), int(reserve_sms), _stream()))


# ---------------------------------------------------------------------------
# consumers behind the path: kNN graph, rare samples, imputation
# ---------------------------------------------------------------------------
def knn_connectivity_csr(idx, dist):
    """kNN lists [n, k] (device) -> CSR (indptr int64, indices int32, data
    float64) of ``s_knn_connectivity_matrix`` (sdm.py:934-947)."""
    _guard()
    dist = _f64(dist)
    assert idx.is_cuda and idx.dtype == torch.int64 and idx.shape == dist.shape
    idx, dist = idx.contiguous(), dist.contiguous()
    n, k = idx.shape
    indptr = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    indices = torch.empty(n * k, dtype=torch.int32, device="cuda")
    data = torch.empty(n * k, dtype=torch.float64, device="cuda")
    _capi.check(_capi.load().scedar_knn_connectivity_csr(
        _ptr(idx), _ptr(dist), n, k, _ptr(indptr), _ptr(indices), _ptr(data),
        _stream()))
    return indptr, indices, data


def knn_affinity(data, aff_scale=1.0):
    """CSR data -> affinity weights (sdm.py:1085-1093)."""
    _guard()
    data = _f64(data).contiguous()
    out = torch.empty_like(data)
    _capi.check(_capi.load().scedar_knn_affinity(
        _ptr(data), data.numel(), float(aff_scale), _ptr(out), _stream()))
    return out


def masked_kth(d, alive, rank, row_begin=0, out=None):
    """Value at 0-based ``rank`` of every alive row of the stripe ``d``
    [rows, n] restricted to the alive columns (+inf for dead rows)."""
    _guard()
    d = _f64(d)
    rows, n = d.shape
    assert alive.is_cuda and alive.dtype == torch.uint8 and alive.numel() == n
    if out is None:
        out = torch.empty(rows, dtype=torch.float64, device="cuda")
    _capi.check(_capi.load().scedar_masked_kth(
        _ptr(d), d.stride(0) if rows > 1 else n, n, row_begin,
        row_begin + rows, _ptr(alive), int(rank), _ptr(out), _stream()))
    return out


def rare_update(kth, cutoff, it, alive, last_kept, ldk=1):
    """One keep / drop decision (detection.py:63-66, 111-115); returns the
    number of cells kept (one 8-byte device -> host read)."""
    _guard()
    n = alive.numel()
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    _capi.check(_capi.load().scedar_rare_update(
        _ptr(kth), int(ldk), n, float(cutoff), int(it), _ptr(alive),
        _ptr(last_kept), _ptr(counter), _stream()))
    return int(counter.item())


def masked_max(v, alive=None, ldv=1, n=None):
    """max over the (alive) entries of a non-negative strided vector."""
    _guard()
    n = v.numel() // ldv if n is None else n
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    _capi.check(_capi.load().scedar_masked_max(
        _ptr(v), int(ldv), n, _ptr(alive), _ptr(out), _stream()))
    return float(out.item())


def compact_alive(alive, count):
    """Ascending indices of the alive cells (int64 [count])."""
    _guard()
    sel = torch.empty(max(count, 1), dtype=torch.int64, device="cuda")
    _capi.check(_capi.load().scedar_compact_alive(
        _ptr(alive), alive.numel(), _ptr(sel), _stream()))
    return sel[:count]


def gather_rows(x, sel):
    """x[sel, :] on the device."""
    _guard()
    x = _f64(x)
    m, p = sel.numel(), x.shape[1]
    out = torch.empty((m, p), dtype=torch.float64, device="cuda")
    _capi.check(_capi.load().scedar_gather_rows(
        _ptr(x), x.stride(0) if x.shape[0] > 1 else max(p, 1), p, _ptr(sel),
        m, _ptr(out), max(p, 1), _stream()))
    return out


def impute_iter(x_curr, knn, n_do_i, min_present_val, it, x_next, idc,
                stats=None, row_begin=0, row_end=None):
    """One imputation iteration (imputation.py:56-101), median statistic, on
    the cells [row_begin, row_end) (default: all)."""
    _guard()
    x_curr, x_next = _f64(x_curr), _f64(x_next)
    n, p = x_curr.shape
    assert knn.is_cuda and knn.dtype == torch.int64 and knn.is_contiguous()
    assert idc.is_cuda and idc.dtype == torch.int32 and idc.is_contiguous()
    assert x_curr.is_contiguous() and x_next.is_contiguous()
    _capi.check(_capi.load().scedar_impute_iter(
        _ptr(x_curr), max(p, 1), n, p, _ptr(knn), knn.shape[1], int(n_do_i),
        float(min_present_val), int(it), int(row_begin),
        int(n if row_end is None else row_end), _ptr(x_next), max(p, 1),
        _ptr(idc), max(p, 1), _ptr(stats), _stream()))


# ---------------------------------------------------------------------------
# clustering consumers of D: condensed form, sub-matrix
# ---------------------------------------------------------------------------
def condensed_len(n):
    """Length of the condensed vector of an n x n distance matrix."""
    return n * (n - 1) // 2 if n > 1 else 0


def condensed_base(i, n):
    """Offset of row i's run (i, i+1) .. (i, n-1) in the condensed vector."""
    return i * (2 * n - i - 1) // 2


def squareform_rows(d, n, row_begin=0, out=None, flags=None):
    """Strict upper triangle of the rows of a stripe ``d`` [rows, n] of a
    symmetric D, row after row: the segment ``[condensed_base(row_begin),
    condensed_base(row_begin + rows))`` of ``scipy.spatial.distance.squareform
    (D)`` (sdm.py:1666).  ``flags`` (int32[1], zeroed by the caller, shared by
    all stripes) collects the SQF_* bits.  Returns (segment, flags)."""
    _guard()
    d = _f64(d)
    rows = d.shape[0]
    assert d.dim() == 2 and d.shape[1] == n and row_begin + rows <= n
    if d.stride(1) != 1 and n > 1:
        d = d.contiguous()
    seg = condensed_base(row_begin + rows, n) - condensed_base(row_begin, n)
    if out is None:
        out = torch.empty(seg, dtype=torch.float64, device="cuda")
    assert out.is_cuda and out.dtype == torch.float64 and out.numel() == seg
    assert out.is_contiguous()
    if flags is None:
        flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    _capi.check(_capi.load().scedar_squareform_rows(
        _ptr(d), d.stride(0) if rows > 1 else max(n, 1), n, int(row_begin),
        int(row_begin) + rows, _ptr(out), _ptr(flags), _stream()))
    return out, flags


def gather_block(d, row_sel, col_sel):
    """``d[row_sel][:, col_sel]`` on the device for int64 index tensors
    (sdm.py:684-692, mirac.py:316-331); raises IndexError like NumPy when a
    selected index is out of range."""
    _guard()
    d = _f64(d)
    assert d.dim() == 2
    if d.stride(1) != 1 and d.shape[1] > 1:
        d = d.contiguous()
    for sel in (row_sel, col_sel):
        assert sel.is_cuda and sel.dtype == torch.int64 and sel.dim() == 1
    row_sel, col_sel = row_sel.contiguous(), col_sel.contiguous()
    m, c = row_sel.numel(), col_sel.numel()
    out = torch.empty((m, c), dtype=torch.float64, device="cuda")
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    _capi.check(_capi.load().scedar_d_gather_block(
        _ptr(d), d.stride(0) if d.shape[0] > 1 else max(d.shape[1], 1),
        d.shape[0], d.shape[1], _ptr(row_sel), m, _ptr(col_sel), c, _ptr(out),
        max(c, 1), _ptr(flags), _stream()))
    if int(flags.item()) & _capi.GATHER_BAD_INDEX:
        raise IndexError("index out of bounds for a matrix of shape "
                         "{}".format(tuple(d.shape)))
    return out


def upload_ranges(n, first=2048):
    """128-aligned row ranges that double in size: the first is small so the
    contraction starts almost at once, every later range is uploaded while the
    (longer) contraction of the previous ones runs."""
    out, b = [], 0
    e = min(n, max(128, first // 128 * 128))
    while b < n:
        out.append((b, e))
        b, e = e, min(n, 2 * e)
    return out


class _PageableStager(object):
    """Upload of PAGEABLE host rows at pinned-memory speed: a few host threads
    copy row ranges into a small ring of pinned buffers (NumPy releases the GIL
    for the copy), each buffer goes up with an asynchronous copy as soon as it
    is full.  The driver's own path for pageable sources is a single-threaded
    staging copy (about 16 GB/s here against 50 GB/s from pinned memory), and
    pinning the caller's array costs more than the transfer."""
    CHUNK_BYTES = 32 << 20
    SLOTS = 4

    def __init__(self):
        from concurrent.futures import ThreadPoolExecutor
        self.threads = max(1, min(8, (os.cpu_count() or 2) // 2))
        self.pool = ThreadPoolExecutor(max_workers=self.threads)
        self.bufs = [torch.empty(self.CHUNK_BYTES // 8, dtype=torch.float64)
                     .pin_memory() for _ in range(self.SLOTS)]
        self.events = [None] * self.SLOTS
        self.turn = 0

    def copy_rows(self, dst, src, stream):
        """dst: device tensor [rows, p] (contiguous rows); src: C-contiguous
        float64 ndarray of the same shape.  The copies are queued on
        ``stream``; returns when the last buffer has been handed to it."""
        rows, p = src.shape
        per = max(1, self.CHUNK_BYTES // (8 * max(p, 1)))
        for b in range(0, rows, per):
            e = min(rows, b + per)
            slot = self.turn % self.SLOTS
            self.turn += 1
            if self.events[slot] is not None:
                self.events[slot].synchronize()       # its last upload is done
            host = self.bufs[slot][:(e - b) * p].view(e - b, p)
            hn = host.numpy()
            t = min(self.threads, e - b)
            cuts = [b + (e - b) * i // t for i in range(t + 1)]
            jobs = [self.pool.submit(np.copyto, hn[c0 - b:c1 - b], src[c0:c1])
                    for c0, c1 in zip(cuts, cuts[1:]) if c1 > c0]
            for j in jobs:
                j.result()
            with torch.cuda.stream(stream):
                dst[b:e].copy_(host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            self.events[slot] = ev


_STAGER = None


def _stager():
    global _STAGER
    if _STAGER is None:
        _STAGER = _PageableStager()
    return _STAGER


def pdist_symmetric_from_host(x_host, metric, out=None, x_dev=None,
                              precision="fp64", bmin=None):
    """FP64 distance matrix of a pinned HOST matrix with the upload hidden
    behind the contraction: x is copied range by range on a second stream,
    each range is prepared (K1 + diagonal) as it lands and the rows of the
    lower triangle it completes are contracted and mirrored at once
    (``scedar_pdist_lower_rows``) -- the result is bit-identical to
    ``Cells.from_dense(x).pdist_symmetric()``.  Returns (cells, d).

    ``x_host`` may also be a plain (pageable) C-contiguous float64 ndarray: its
    ranges then go through the pinned staging ring of ``_PageableStager``."""
    _guard()
    pageable = isinstance(x_host, np.ndarray)
    if pageable:
        assert x_host.dtype == np.float64 and x_host.ndim == 2 \
            and x_host.flags.c_contiguous
    else:
        assert x_host.dtype == torch.float64 and x_host.dim() == 2
    n, p = x_host.shape
    if out is None:
        out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    if x_dev is None:
        x_dev = torch.empty((n, p), dtype=torch.float64, device="cuda")
    main = torch.cuda.current_stream()
    copy = _copy_stream()
    copy.wait_stream(main)          # x_dev may still be read by earlier work
    cells = Cells.alloc(n, p, metric)
    if bmin is not None:            # block minima ride along (INT8 split only)
        cells.set_blockmin([bmin])
    for b, e in upload_ranges(n):
        if pageable:
            _stager().copy_rows(x_dev[b:e], x_host[b:e], copy)
        with torch.cuda.stream(copy):
            if not pageable:
                x_dev[b:e].copy_(x_host[b:e], non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(copy)
        main.wait_event(landed)
        cells.fill_rows(x_dev[b:e], b)
        cells.pdist_lower_rows(b, e, out, precision=precision)
    return cells, out


_COPY_STREAMS = {}


def _copy_stream():
    d = torch.cuda.current_device()
    if d not in _COPY_STREAMS:
        _COPY_STREAMS[d] = torch.cuda.Stream(device=d)
    return _COPY_STREAMS[d]


_COMM_STREAMS = {}


def _comm_stream():
    """Stream the collectives of a pipelined upload are issued on."""
    d = torch.cuda.current_device()
    if d not in _COMM_STREAMS:
        _COMM_STREAMS[d] = torch.cuda.Stream(device=d)
    return _COMM_STREAMS[d]
