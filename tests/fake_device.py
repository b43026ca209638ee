"""CPU stand-in for ``scedar_b200.device`` -- TEST INFRASTRUCTURE ONLY.

``install(monkeypatch)`` replaces the device operators the host modules
(``sdm.py``, ``hclust.py``, ``knn.py``) call with NumPy / oracle versions that
honour the same contract (argument meaning, shapes, dtypes, orientation) on
CPU tensors, so the host logic above the C ABI -- caching, routing by k and
size, index rules, list / dict building, error conventions -- runs under
``-m "not gpu"``.  Nothing here is a product path: the real operators fail
loudly without a CUDA device, and the parity of the kernels themselves is what
the ``-m gpu`` tests establish.
"""
import numpy as np
import scipy.sparse as sps
import torch

from oracle import sdm_oracle as orc


def _stable_topk(d, k, exclude, row_begin):
    order = np.argsort(d, kind="mergesort", axis=1)[:, :k + 1]
    rows = np.arange(d.shape[0])[:, None] + row_begin
    if exclude == 0:
        idx = order[:, 1:]
    else:
        keep = order != rows
        keep[keep.all(axis=1), 0] = False
        idx = order[keep].reshape(d.shape[0], k)
    return idx.astype(np.int64), np.take_along_axis(d, idx, axis=1)


class FakeCells(object):
    calls = []          # (what, args) log the tests can inspect

    def __init__(self, x, metric):
        self.x, self.metric = x, metric
        self.n, self.p = x.shape
        self._d = None

    @classmethod
    def from_dense(cls, x, metric):
        return cls(x.numpy(), metric)

    @classmethod
    def from_csr(cls, indptr, indices, data, shape, metric):
        # densified, as the scatter kernel does (duplicates summed)
        return cls(sps.csr_matrix((data.numpy(), indices.numpy(),
                                   indptr.numpy()), shape=shape).toarray(),
                   metric)

    def _full(self):
        if self._d is None:
            self._d = orc.pdist(self.x, self.metric)
        return self._d

    def pdist_symmetric(self, out=None, precision="fp64"):
        FakeCells.calls.append(("pdist_symmetric", precision))
        return torch.from_numpy(self._full().copy())

    def pdist_stripe(self, row_begin=0, row_end=None, out=None,
                     precision="fp64"):
        FakeCells.calls.append(("pdist_stripe", row_begin, row_end))
        row_end = self.n if row_end is None else row_end
        return torch.from_numpy(self._full()[row_begin:row_end].copy())

    def knn_stripe(self, k, exclude=1, row_begin=0, row_end=None,
                   precision="fp64"):
        FakeCells.calls.append(("knn_stripe", k, exclude, precision))
        row_end = self.n if row_end is None else row_end
        i, d = _stable_topk(self._full()[row_begin:row_end], k, exclude,
                            row_begin)
        return torch.from_numpy(i), torch.from_numpy(d)

    def close(self):
        pass


def to_device(a):
    return torch.from_numpy(np.array(a, copy=True))


def to_host(t, out):
    out[...] = t.numpy()
    return out


def knn_from_d(d, k, exclude=1, row_begin=0):
    i, dd = _stable_topk(d.numpy(), k, exclude, row_begin)
    return torch.from_numpy(i), torch.from_numpy(dd)


def col_sort(d, want_sorted=True, want_arg=True):
    a = d.numpy()
    arg = np.argsort(a, kind="mergesort", axis=1)
    srt = np.take_along_axis(a, arg, axis=1)
    return (torch.from_numpy(srt) if want_sorted else None,
            torch.from_numpy(arg.astype(np.int64)) if want_arg else None)


def num_correct_dist_mat_(d, upper_bound=None):
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        orc.num_correct_dist_mat(d.numpy(), upper_bound)
    flags = 0
    for i in w:
        flags |= 1 if "diag" in str(i.message) else 2
    return flags


def squareform_rows(d, n, row_begin=0, out=None, flags=None):
    a = d.numpy()
    assert a.shape[1] == n and row_begin + a.shape[0] <= n
    runs = [a[r, row_begin + r + 1:] for r in range(a.shape[0])]
    v = np.concatenate(runs) if runs else np.zeros(0)
    if flags is None:
        flags = torch.zeros(1, dtype=torch.int32)
    if np.isnan(v).any():
        flags |= 1
    if any(not (a[r, row_begin + r] == 0) for r in range(a.shape[0])):
        flags |= 2
    seg = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
    if out is not None:
        out.copy_(seg)
        seg = out
    return seg, flags


def gather_block(d, row_sel, col_sel):
    a, r, c = d.numpy(), row_sel.numpy(), col_sel.numpy()
    if ((r < 0) | (r >= a.shape[0])).any() or \
            ((c < 0) | (c >= a.shape[1])).any():
        raise IndexError("index out of bounds")
    return torch.from_numpy(a[r][:, c].copy())


def knn_connectivity_csr(idx, dist):
    n, k = idx.shape
    m = orc.knn_connectivity(idx.numpy(), dist.numpy(), n)
    return (torch.from_numpy(m.indptr.astype(np.int64)),
            torch.from_numpy(m.indices.astype(np.int32)),
            torch.from_numpy(m.data.copy()))


def knn_affinity(data, aff_scale=1.0):
    w = data.numpy().copy()
    w[np.isneginf(w)] = 0
    return torch.from_numpy((w.max() - w) * aff_scale if w.size else w)


def masked_kth(d, alive, rank, row_begin=0, out=None):
    a, live = d.numpy(), alive.numpy().astype(bool)
    res = np.full(a.shape[0], np.inf)
    for r in range(a.shape[0]):
        if live[row_begin + r]:
            res[r] = np.sort(a[r][live])[rank]
    res = torch.from_numpy(res)
    if out is not None:
        out.copy_(res)
        return out
    return res


def rare_update(kth, cutoff, it, alive, last_kept, ldk=1):
    n = alive.numel()
    v = kth.numpy().reshape(-1)[::ldk][:n] if kth.dim() == 1 else \
        kth.numpy()[:, 0]
    live = alive.numpy().astype(bool)
    keep = live & (v <= cutoff)
    last_kept.numpy()[keep] = it
    alive.numpy()[live & ~keep] = 0
    return int(keep.sum())


def masked_max(v, alive=None, ldv=1, n=None):
    a = v.numpy()
    a = a[:, 0] if a.ndim == 2 else a.reshape(-1)[::ldv]
    if n is not None:
        a = a[:n]
    if alive is not None:
        a = a[alive.numpy().astype(bool)]
    return float(a.max()) if a.size else 0.0


def compact_alive(alive, count):
    return torch.from_numpy(np.nonzero(alive.numpy())[0].astype(np.int64))


def gather_rows(x, sel):
    return torch.from_numpy(x.numpy()[sel.numpy()].copy())


def impute_iter(x_curr, knn, n_do_i, min_present_val, it, x_next, idc,
                stats=None, row_begin=0, row_end=None):
    x, kn = x_curr.numpy(), knn.numpy()
    row_end = x.shape[0] if row_end is None else row_end
    present = x >= min_present_val
    for s in range(row_begin, row_end):
        mask = (present[kn[s]].sum(axis=0) >= n_do_i) & ~present[s]
        x_next.numpy()[s] = x[s]
        if mask.any():
            x_next.numpy()[s, mask] = np.median(x[kn[s]][:, mask], axis=0)
            idc.numpy()[s, mask] = it
    if stats is not None:
        sub = idc.numpy()[row_begin:row_end] > 0
        stats[0] = int(sub.sum())
        stats[1] = int(sub.any(axis=1).sum())


_OPS = ("to_device", "to_host", "knn_from_d", "col_sort",
        "num_correct_dist_mat_", "squareform_rows", "gather_block",
        "knn_connectivity_csr", "knn_affinity", "masked_kth", "rare_update",
        "masked_max", "compact_alive", "gather_rows", "impute_iter")


def install(monkeypatch):
    """Swap the operators of ``scedar_b200.device`` for the stand-ins."""
    from scedar_b200 import device as dev
    monkeypatch.setattr(dev, "DEVICE", "cpu")
    monkeypatch.setattr(dev, "_guard", lambda: None)
    monkeypatch.setattr(dev, "Cells", FakeCells)
    for name in _OPS:
        monkeypatch.setattr(dev, name, globals()[name])
    FakeCells.calls = []
    return dev
