"""Hierarchical clustering on top of the B200 distance matrix.

Mirror of the part of ``scedar.eda.HClustTree`` that touches the distance
matrix (scedar/eda/sdm.py:1369-1436, 1564-1582, 1629-1693): same static
entry points (``hclust_linkage``, ``hclust_tree``, ``hct_from_lkg``,
``sort_x_by_d``), same arguments, same ``ValueError``s and warnings.  What
moves to the device is the data side of ``hclust_linkage`` -- the clean-up
``num_correct_dist_mat`` and ``scipy.spatial.distance.squareform`` -- so that
a D that already lives in HBM reaches ``scipy.cluster.hierarchy.linkage`` as
one condensed vector (n(n-1)/2 doubles device -> host) instead of as an n x n
host matrix that is then copied, checked and condensed on one core.  The
agglomeration itself (``sch.linkage``, third-party, sequential) stays where
the reference has it, on the CPU.

Extension over the reference: ``dmat`` may also be a
``scedar_b200.SampleDistanceMatrix`` (its D is taken from HBM, or computed
there stripe by stripe, and never becomes a host matrix) or a CUDA float64
tensor.

``bi_partition`` / ``cluster_id_to_lab_list`` (sdm.py:1438-1627) belong to
MIRAC's tree walk, not to this path, and are not provided.
"""
import numpy as np
import scipy.cluster.hierarchy as sch
import torch

from . import _capi
from . import device as dev


def _condensed_flags_check(flags):
    """The two rejections of scipy's ``squareform(checks=True)`` ->
    ``is_valid_dm(X, throw=True, name='X')``, in its order."""
    if flags & _capi.SQF_NAN:
        raise ValueError("Distance matrix 'X' must be symmetric.")
    if flags & _capi.SQF_BAD_DIAG:
        raise ValueError("Distance matrix 'X' diagonal must be zero.")


def condensed_from_device(d):
    """Condensed form (host float64 vector) of a whole symmetric D in HBM."""
    n = d.shape[0]
    seg, flags = dev.squareform_rows(d, n)
    out = np.empty(dev.condensed_len(n), dtype=np.float64)
    if out.size:
        dev.to_host(seg, out)
    _condensed_flags_check(int(flags.item()))
    return out


def _condensed_of(dmat):
    """``squareform(num_correct_dist_mat(np.array(dmat)))`` (sdm.py:1633-1634,
    1666) -> (condensed host vector, n)."""
    from .sdm import SampleDistanceMatrix
    if hasattr(dmat, "_condensed_d"):
        # a (drop-in or patched) SDM: its D is clean by construction (fused
        # epilogue, or the constructor's num_correct_dist_mat), so the
        # clean-up is the identity
        return dmat._condensed_d(), dmat._x.shape[0]
    if isinstance(dmat, torch.Tensor):
        if dmat.dim() != 2 or dmat.shape[0] != dmat.shape[1]:
            raise ValueError("dmat must be a 2D (n_samples, n_samples)"
                             " np array")
        t = dmat.detach().to(device=dev.DEVICE, dtype=torch.float64).clone()
    else:
        a = np.array(dmat, dtype="float")
        if a.ndim != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("dmat must be a 2D (n_samples, n_samples)"
                             " np array")
        t = dev.to_device(a)
    n = t.shape[0]
    if n:
        SampleDistanceMatrix._warn_ncdm(dev.num_correct_dist_mat_(t))
    return condensed_from_device(t), n


def _multinomial_mdl(x):
    """``MultinomialMdl(x).mdl`` (scedar/eda/mdl.py:86-106): code length of
    the values of ``x`` under their own empirical distribution."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    _, cnts = np.unique(x, return_counts=True)
    if len(cnts) > 1:
        return (-np.log(cnts / n) * cnts).sum()
    if len(cnts) == 1:
        return np.log(n)
    return 0


def bipartition_counts(hac_z, n_rounds):
    """``HClustTree.hct_from_lkg(hac_z).n_round_bipar_cnt(n_rounds)``
    (sdm.py:1564-1582) read straight off the linkage matrix: round r lists
    the leaf counts of the (left, right) children of every node of the
    previous round, empty subtrees (children of a leaf) counting 0."""
    assert n_rounds > 0
    z = np.asarray(hac_z)
    n = z.shape[0] + 1
    child = z[:, :2].astype(np.int64)
    size = z[:, 3].astype(np.int64)
    frontier = np.array([2 * n - 2], dtype=np.int64)   # root; -1 = empty
    rounds = []
    for _ in range(n_rounds):
        inner = frontier >= n
        row = np.where(inner, frontier - n, 0)
        nxt = np.empty(2 * frontier.shape[0], dtype=np.int64)
        nxt[0::2] = np.where(inner, child[row, 0], -1)
        nxt[1::2] = np.where(inner, child[row, 1], -1)
        cnt = np.where(nxt < 0, 0,
                       np.where(nxt < n, 1, size[np.maximum(nxt - n, 0)]))
        rounds.append(cnt.tolist())
        frontier = nxt
    return rounds


def linkage_from_condensed(dmat_sf, n, linkage="complete", n_eval_rounds=None,
                           is_euc_dist=False, optimal_ordering=False,
                           verbose=False):
    """The part of ``hclust_linkage`` behind ``squareform`` (sdm.py:1636-1669):
    the ``linkage="auto"`` selection and ``sch.linkage`` on the condensed
    vector of an n x n matrix.  Host only."""
    if linkage == "auto":
        try_linkages = ("single", "complete", "average", "weighted")
        if is_euc_dist:
            try_linkages += ("centroid", "median", "ward")
        if n_eval_rounds is None:
            n_eval_rounds = int(np.ceil(np.log2(n)))
        else:
            n_eval_rounds = int(np.ceil(max(np.log2(n), n_eval_rounds)))
        scores = []
        for lt in try_linkages:
            # the clean-up and squareform are idempotent, so the condensed
            # vector is shared by every candidate (the reference redoes both
            # per candidate, sdm.py:1651)
            z = sch.linkage(dmat_sf, method=lt, optimal_ordering=False)
            mdls = np.array([_multinomial_mdl(c) for c in
                             bipartition_counts(z, n_eval_rounds)])
            scores.append(np.sum(mdls / np.arange(1, n_eval_rounds + 1)))
        linkage = try_linkages[scores.index(max(scores))]
        if verbose:
            print(linkage, tuple(zip(try_linkages, scores)), sep="\n")
    return sch.linkage(dmat_sf, method=linkage,
                       optimal_ordering=optimal_ordering)


class HClustTree(object):
    """Hierarchical clustering tree over ``scipy.cluster.hierarchy.ClusterNode``
    (sdm.py:1369-1436): a binary tree whose missing children are empty trees
    of count 0."""

    def __init__(self, node, prev=None):
        super(HClustTree, self).__init__()
        self._node = node
        self._prev = prev
        self._left = None if node is None else node.get_left()
        self._right = None if node is None else node.get_right()

    @property
    def prev(self):
        return self._prev

    def count(self):
        return 0 if self._node is None else self._node.get_count()

    def left(self):
        return HClustTree(self._left, self)

    def left_count(self):
        return self.left().count()

    def right(self):
        return HClustTree(self._right, self)

    def right_count(self):
        return self.right().count()

    def leaf_ids(self):
        """Leaf ids from left to right."""
        if self._node is None:
            return []
        return self._node.pre_order(lambda xn: xn.get_id())

    def left_leaf_ids(self):
        return self.left().leaf_ids()

    def right_leaf_ids(self):
        return self.right().leaf_ids()

    def n_round_bipar_cnt(self, n):
        """sdm.py:1564-1582."""
        assert n > 0
        out, level = [], [self]
        for _ in range(n):
            kids = []
            for t in level:
                kids += [t.left(), t.right()]
            out.append([t.count() for t in kids])
            level = kids
        return out

    @staticmethod
    def hclust_linkage(dmat, linkage="complete", n_eval_rounds=None,
                       is_euc_dist=False, optimal_ordering=False,
                       verbose=False):
        """sdm.py:1629-1669 -> the scipy linkage matrix."""
        dmat_sf, n = _condensed_of(dmat)
        return linkage_from_condensed(
            dmat_sf, n, linkage=linkage, n_eval_rounds=n_eval_rounds,
            is_euc_dist=is_euc_dist, optimal_ordering=optimal_ordering,
            verbose=verbose)

    @staticmethod
    def hclust_tree(dmat, linkage="complete", n_eval_rounds=None,
                    is_euc_dist=False, optimal_ordering=False, verbose=False):
        """sdm.py:1671-1678."""
        return HClustTree.hct_from_lkg(HClustTree.hclust_linkage(
            dmat=dmat, linkage=linkage, n_eval_rounds=n_eval_rounds,
            is_euc_dist=is_euc_dist, optimal_ordering=optimal_ordering,
            verbose=verbose))

    @staticmethod
    def hct_from_lkg(hac_z):
        """sdm.py:1680-1682."""
        return HClustTree(sch.to_tree(hac_z))

    @staticmethod
    def sort_x_by_d(x, dmat=None, metric="cosine", linkage="auto",
                    n_eval_rounds=None, optimal_ordering=False,
                    nprocs=None, verbose=False):
        """sdm.py:1684-1693: leaf order of the auto-linkage tree of the rows
        of ``x``.  As in the reference, ``linkage``, ``n_eval_rounds`` and
        ``verbose`` are accepted and not used.  D stays in HBM."""
        from .sdm import SampleDistanceMatrix
        sdm = SampleDistanceMatrix(x, d=dmat, metric=metric, nprocs=nprocs)
        sdm._validate_pdist()
        tree = HClustTree.hclust_tree(
            sdm, linkage="auto", is_euc_dist=(metric == "euclidean"),
            optimal_ordering=optimal_ordering)
        return tree.leaf_ids()
