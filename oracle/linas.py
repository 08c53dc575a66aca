"""NumPy restatement of the LINAS-engine scoring / ranking / metric path (oracle; test-only).

Every function cites the reference lines it follows (paths relative to ``/root/reference``).
The restatement is vectorised where that cannot change a result bit (integer ranks, the
order of floating-point additions is kept where the reference's order is observable).
Pinned by ``tests/test_oracle_golden.py`` against outputs of the imported reference.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "l2norm", "cal_error", "cal_simi", "norm_score", "get_gt", "gt_ranks", "eval_q2m",
    "ap_score", "t2v_map", "v2t_map", "cal_perf", "topk_ids", "t2v", "v2t", "t2v_inv_rank",
    "v2t_inv_rank", "v2t_inv_rank_multi", "fused_errors",
]


# ---------------------------------------------------------------------------------------------
# scoring
# ---------------------------------------------------------------------------------------------
def l2norm(X):
    """Row-wise ``X / ||X||_2`` with no epsilon; dtype follows the input.

    LINAS-engine/evaluation.py:10-14 (a zero row gives NaN, as there).
    """
    X = np.asarray(X)
    return 1.0 * X / np.linalg.norm(X, axis=1, keepdims=True)


def cal_error(videos, captions, measure="cosine"):
    """``errors[q, v] = -<q_hat, v_hat>``; LINAS-engine/evaluation.py:17-21 (cosine branch).

    ``cal_error_batch`` (evaluation.py:41-45) is the same expression for cosine.
    """
    if measure != "cosine":
        raise NotImplementedError("oracle restates the cosine branch only (SURVEY.md §8a A2)")
    return -1 * np.dot(l2norm(captions), l2norm(videos).T)


def cal_simi(captions, videos, measure="cosine"):
    """``+<q_hat, v_hat>`` with the argument order swapped; LINAS-engine/evaluation.py:75-79."""
    if measure != "cosine":
        raise NotImplementedError
    return np.dot(l2norm(captions), l2norm(videos).T)


def norm_score(t2v_all_errors):
    """Global min-max normalisation of the score matrix; LINAS-engine/validate.py:7-11."""
    s = -t2v_all_errors
    s = s - np.min(s)
    s = s / np.max(s)
    return -s


def fused_errors(video_spaces, caption_spaces, weights, mode="weighted-cosine"):
    """Multi-space fusion as SURVEY.md §8a row F defines it (the reference has vestiges only).

    ``weighted-cosine``: ``errors = sum_s w_s * cal_error(V_s, Q_s)``.
    ``norm_score``     : ``errors = sum_s w_s * norm_score(cal_error(V_s, Q_s))``.
    """
    acc = None
    for V, Q, w in zip(video_spaces, caption_spaces, weights):
        e = cal_error(V, Q)
        if mode == "norm_score":
            e = norm_score(e)
        elif mode != "weighted-cosine":
            raise ValueError(mode)
        acc = w * e if acc is None else acc + w * e
    return acc


def topk_ids(errors_row, k):
    """``np.argsort(errors[0])[:topK]``; LINAS-engine/inference.py:79."""
    return np.argsort(errors_row)[:k]


# ---------------------------------------------------------------------------------------------
# ground truth
# ---------------------------------------------------------------------------------------------
def get_gt(video_ids, caption_ids):
    """Ground-truth containers; LINAS-engine/util/metrics.py:106-120.

    ``v2t_gt[i]`` lists (ascending) the caption rows whose id up to the first ``'#'`` equals
    ``video_ids[i]``; ``t2v_gt`` is the inverse dict, built by walking ``v2t_gt`` in video order
    so each value list is in ascending video index.  The reference does this with an
    O(Nv*Nq) double loop; grouping captions by key first gives the same containers.
    """
    by_key = {}
    for i, cap_id in enumerate(caption_ids):
        by_key.setdefault(cap_id.split("#", 1)[0], []).append(i)
    v2t_gt = [list(by_key.get(vid, [])) for vid in video_ids]
    t2v_gt = {}
    for i, caps in enumerate(v2t_gt):
        for c in caps:
            t2v_gt.setdefault(c, []).append(i)
    return v2t_gt, t2v_gt


# ---------------------------------------------------------------------------------------------
# ranks and metrics
# ---------------------------------------------------------------------------------------------
def _inverse_argsort(row):
    order = np.argsort(row)
    pos = np.empty_like(order)
    pos[order] = np.arange(order.shape[0])
    return pos


def gt_ranks(scores, q2m_gts):
    """Best 1-based rank of any ground-truth item per query row (``n_m + 1`` if the row has none).

    The loop body of LINAS-engine/util/metrics.py:138-147: position of each GT id in
    ``np.argsort(row)`` plus one, minimum over the row's GT ids.
    """
    n_q, n_m = scores.shape
    out = np.zeros((n_q,), np.int32)
    for i in range(n_q):
        pos = _inverse_argsort(scores[i])
        rank = n_m + 1
        for k in q2m_gts[i]:
            rank = min(rank, int(pos[k]) + 1)
        out[i] = rank
    return out


def metrics_from_ranks(ranks):
    """R@1/5/10, MedR, MeanR from int32 1-based ranks; LINAS-engine/util/metrics.py:149-157."""
    n_q = ranks.shape[0]
    r1 = 100.0 * len(np.where(ranks <= 1)[0]) / n_q
    r5 = 100.0 * len(np.where(ranks <= 5)[0]) / n_q
    r10 = 100.0 * len(np.where(ranks <= 10)[0]) / n_q
    return (r1, r5, r10, np.median(ranks), ranks.mean())


def eval_q2m(scores, q2m_gts):
    """LINAS-engine/util/metrics.py:124-157."""
    return metrics_from_ranks(gt_ranks(scores, q2m_gts))


def ap_score(sorted_labels, k=0):
    """``APScorer(k).score``; LINAS-engine/basic/metric.py:13-17,31-46.

    ``nr_relevant`` counts the whole list, the sum runs over the first ``k`` positions
    (all of them if ``k <= 0`` or ``k > len``), additions in rank order in double precision.
    """
    nr_relevant = sum(1 for x in sorted_labels if x > 0)
    if nr_relevant == 0:
        return 0.0
    length = k if 0 < k <= len(sorted_labels) else len(sorted_labels)
    ap, rel = 0.0, 0
    for i in range(length):
        if sorted_labels[i] >= 1:
            rel += 1
            ap += float(rel) / (i + 1.0)
    return ap / nr_relevant


def ap_from_ranks(ranks_1based, nr_relevant=None, k=0, list_len=None):
    """AP of a label list whose relevant items sit at the given 1-based ranks.

    Same additions in the same order as :func:`ap_score` (which only adds at relevant
    positions), so the result is bit-identical to scoring the full label list.
    """
    ranks = sorted(int(r) for r in ranks_1based)
    nr = len(ranks) if nr_relevant is None else nr_relevant
    if nr == 0:
        return 0.0
    ap = 0.0
    for j, r in enumerate(ranks, 1):
        if k > 0 and (list_len is None or k <= list_len) and r > k:
            break
        ap += float(j) / (r - 1 + 1.0)
    return ap / nr


def t2v_map(c2i, t2v_gts):
    """Text->video mAP; LINAS-engine/util/metrics.py:61-79 (only the FIRST GT video is marked)."""
    perf = []
    for i in range(c2i.shape[0]):
        pos = _inverse_argsort(c2i[i, :])
        perf.append(ap_from_ranks([pos[t2v_gts[i][0]] + 1]))
    return np.mean(perf)


def v2t_map(c2i, v2t_gts):
    """Video->text mAP; LINAS-engine/util/metrics.py:83-102 (all GT captions are marked)."""
    perf = []
    for i in range(c2i.shape[1]):
        pos = _inverse_argsort(c2i[:, i])
        # the reference builds a label list, so duplicate GT ids collapse into one label
        gts = sorted(set(int(x) for x in v2t_gts[i]))
        perf.append(ap_from_ranks([pos[x] + 1 for x in gts]))
    return np.mean(perf)


def cal_perf(t2v_all_errors, v2t_gt, t2v_gt):
    """Return tuple of LINAS-engine/validate.py:15-54 (logging / tensorboard side effects omitted)."""
    t2v_r = eval_q2m(t2v_all_errors, t2v_gt)
    t2v_m = t2v_map(t2v_all_errors, t2v_gt)
    v2t_r = eval_q2m(t2v_all_errors.T, v2t_gt)
    v2t_m = v2t_map(t2v_all_errors, v2t_gt)
    return (*v2t_r, v2t_m), (*t2v_r, t2v_m)


# ---------------------------------------------------------------------------------------------
# legacy fixed-n_caption metrics (never reached from cal_perf)
# ---------------------------------------------------------------------------------------------
def _legacy_summary(ranks0):
    n = len(ranks0)
    r1 = 100.0 * len(np.where(ranks0 < 1)[0]) / n
    r5 = 100.0 * len(np.where(ranks0 < 5)[0]) / n
    r10 = 100.0 * len(np.where(ranks0 < 10)[0]) / n
    return [float(x) for x in (r1, r5, r10, np.floor(np.median(ranks0)) + 1, ranks0.mean() + 1)]


def t2v(c2i, n_caption=5):
    """LINAS-engine/util/metrics.py:5-29 (0-based ranks, ``medr = floor(median) + 1``)."""
    assert c2i.shape[0] // c2i.shape[1] == n_caption, c2i.shape
    ranks = np.zeros(c2i.shape[0])
    for i in range(len(ranks)):
        ranks[i] = _inverse_argsort(c2i[i])[i // n_caption]
    return _legacy_summary(ranks)


def v2t(c2i, n_caption=5):
    """LINAS-engine/util/metrics.py:34-57 (first position holding any caption of video i)."""
    assert c2i.shape[0] // c2i.shape[1] == n_caption, c2i.shape
    ranks = np.zeros(c2i.shape[1])
    for i in range(len(ranks)):
        inds = np.argsort(c2i[:, i])
        ranks[i] = np.where(inds // n_caption == i)[0][0]
    return _legacy_summary(ranks)


def t2v_inv_rank(c2i, n_caption=1):
    """LINAS-engine/util/metrics.py:161-177."""
    assert c2i.shape[0] // c2i.shape[1] == n_caption, c2i.shape
    inv = np.zeros(c2i.shape[0])
    for i in range(len(inv)):
        rank = np.where(np.argsort(c2i[i, :]) == i // n_caption)[0]
        inv[i] = sum(1.0 / (rank + 1))
    return np.mean(inv)


def v2t_inv_rank(c2i, n_caption=1):
    """LINAS-engine/util/metrics.py:181-197."""
    assert c2i.shape[0] // c2i.shape[1] == n_caption, c2i.shape
    inv = np.zeros(c2i.shape[1])
    for i in range(len(inv)):
        rank = np.where(np.argsort(c2i[:, i]) // n_caption == i)[0]
        inv[i] = sum(1.0 / (rank + 1))
    return np.mean(inv)


def v2t_inv_rank_multi(c2i, n_caption=2):
    """LINAS-engine/util/metrics.py:202-218."""
    assert c2i.shape[0] // c2i.shape[1] == n_caption, c2i.shape
    return [v2t_inv_rank(c2i[list(range(i, c2i.shape[0], n_caption)), :], n_caption=1)
            for i in range(n_caption)]
