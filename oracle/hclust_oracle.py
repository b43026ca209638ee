"""CPU restatement of the distance-matrix side of ``scedar.eda.HClustTree``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker of
``scedar_b200.hclust`` and of the ``scedar_squareform_rows`` /
``scedar_d_gather_block`` kernels, never a product path.

Follows scedar/eda/sdm.py:1629-1693 (``hclust_linkage``, ``hclust_tree``,
``sort_x_by_d``), 1564-1582 (``n_round_bipar_cnt``) and
scedar/eda/mdl.py:86-106 (``MultinomialMdl.mdl``).  Third-party under it, not
vendored by the reference and used here as the reference uses it:
``scipy.spatial.distance.squareform`` and ``scipy.cluster.hierarchy.linkage``
/ ``to_tree`` (installed: scipy 1.18.1; the reference does not pin scipy).

Parity status: PINNED -- tests/test_oracle_hclust.py holds it to the
reference's known answers (tests/test_eda/test_dense_sdm.py:170-193,
1255-1304) and to linkage matrices produced by the live reference
(tests/golden/make_golden_hclust.py -> tests/golden/hclust.npz).
"""
import numpy as np
import scipy.cluster.hierarchy as sch
import scipy.spatial.distance as spdist

from . import sdm_oracle as orc


def squareform(dmat):
    """``spspatial.distance.squareform(dmat)`` (sdm.py:1666) spelled out: the
    strict upper triangle, row after row; symmetric and zero-diagonal input
    required (scipy's ``is_valid_dm``)."""
    dmat = np.asarray(dmat, dtype=np.float64)
    n = dmat.shape[0]
    if not (dmat == dmat.T).all():
        raise ValueError("Distance matrix 'X' must be symmetric.")
    if not (np.diagonal(dmat) == 0).all():
        raise ValueError("Distance matrix 'X' diagonal must be zero.")
    if n <= 1:
        return np.array([], dtype=np.float64)
    return np.concatenate([dmat[i, i + 1:] for i in range(n - 1)])


def multinomial_mdl(x):
    """mdl.py:86-106."""
    x = np.array(x, dtype=np.float64)
    n = x.shape[0]
    vals, cnts = np.unique(x, return_counts=True)
    if len(vals) > 1:
        return (-np.log(cnts / n) * cnts).sum()
    if len(vals) == 1:
        return np.log(n)
    return 0


def n_round_bipar_cnt(root, n_rounds):
    """sdm.py:1564-1582 on a ``ClusterNode`` tree: per round the leaf counts of
    the left / right child of every subtree of the previous round; a missing
    child is an empty subtree of count 0 whose children are empty too."""
    assert n_rounds > 0
    rounds, level = [], [root]
    for _ in range(n_rounds):
        nxt = []
        for node in level:
            if node is None:
                nxt += [None, None]
            else:
                nxt += [node.get_left(), node.get_right()]
        rounds.append([0 if c is None else c.get_count() for c in nxt])
        level = nxt
    return rounds


def hclust_linkage(dmat, linkage="complete", n_eval_rounds=None,
                   is_euc_dist=False, optimal_ordering=False):
    """sdm.py:1629-1669."""
    dmat = np.array(dmat, dtype="float")
    dmat = orc.num_correct_dist_mat(dmat)
    n = dmat.shape[0]
    if linkage == "auto":
        tries = ("single", "complete", "average", "weighted")
        if is_euc_dist:
            tries += ("centroid", "median", "ward")
        if n_eval_rounds is None:
            n_eval_rounds = int(np.ceil(np.log2(n)))
        else:
            n_eval_rounds = int(np.ceil(max(np.log2(n), n_eval_rounds)))
        scores = []
        for lt in tries:
            z = sch.linkage(squareform(dmat), method=lt)
            cnts = n_round_bipar_cnt(sch.to_tree(z), n_eval_rounds)
            mdls = np.array([multinomial_mdl(np.array(c)) for c in cnts])
            scores.append(np.sum(mdls / np.arange(1, n_eval_rounds + 1)))
        linkage = tries[scores.index(max(scores))]
    return sch.linkage(squareform(dmat), method=linkage,
                       optimal_ordering=optimal_ordering), linkage


def leaf_ids(hac_z):
    """``HClustTree.hct_from_lkg(hac_z).leaf_ids()`` (sdm.py:1421-1430)."""
    return sch.to_tree(hac_z).pre_order(lambda nd: nd.get_id())


def sort_x_by_d(x, metric="cosine", optimal_ordering=False):
    """sdm.py:1684-1693 (always linkage="auto", as the reference)."""
    d = orc.pdist(orc.coerce_x(x), metric)
    z, _ = hclust_linkage(d, linkage="auto",
                          is_euc_dist=(metric == "euclidean"),
                          optimal_ordering=optimal_ordering)
    return leaf_ids(z)


def _self_check():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(9, 3))
    d = spdist.squareform(spdist.pdist(x))
    assert np.array_equal(squareform(d), spdist.squareform(d))
