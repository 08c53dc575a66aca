"""Ad-hoc video search (TRECVID AVS, V3C1 shape): top-1000 shots per query + AP@1000 / mAP.

The reference names this use (``util/TEMPLATE_do_test_avs.sh:10`` calls a ``tester_avs.py`` that is not in the
tree) and ships the scorer it would use: ``getScorer('AP@k')`` -> ``APScorer(k)``
(``LINAS-engine/basic/metric.py:13-17,25-46,118-125``): over the first ``k`` positions of the ranked list,
``ap = sum_j j / rank_j`` over the relevant items met in rank order, divided by the number of relevant items in
the WHOLE list.  The ranking itself is ``np.argsort(errors)[:topK]``
(``LINAS-engine/inference.py:79-80``) with ``topK = 1000``; here it is one ``search`` of the resident corpus.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N
from .engine import search_shards


def search_avs(stores, queries, k=1000, weights=None, comm=None, n_total=None):
    """Top-``k`` shots per query: ``(scores fp64 [nq, k], idx int64 [nq, k])`` with global shot rows."""
    stores = list(stores) if isinstance(stores, (list, tuple)) else [stores]
    return search_shards(stores, queries, k, weights=weights, comm=comm, n_total=n_total)


def _list_ranks(idx, relevant, n_mem):
    """1-based position of every relevant row in its query's ranked list (``n_mem + 1`` if absent) as a CSR:
    returns ``(off int64 [nq+1], rank int32 [n_entries], device)``."""
    idx = idx if torch.is_tensor(idx) else torch.as_tensor(np.asarray(idx))
    dev = idx.device if idx.is_cuda else torch.device("cuda", torch.cuda.current_device())
    idx = idx.to(dev, torch.int64)
    nq, kk = idx.shape
    assert len(relevant) == nq, "one relevant set per query"
    sizes = np.fromiter((len(r) for r in relevant), dtype=np.int64, count=nq)
    off = np.zeros(nq + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    off_d = torch.from_numpy(off).to(dev)
    if int(off[-1]) == 0:
        return off_d, torch.zeros(1, dtype=torch.int32, device=dev), dev
    rel = torch.from_numpy(np.concatenate([np.asarray(r, dtype=np.int64).reshape(-1) for r in relevant])).to(dev)
    idx = idx.contiguous() if idx.stride(1) != 1 else idx
    rank = torch.empty(int(off[-1]), dtype=torch.int32, device=dev)
    N.call("xmve_list_ranks", N.ptr(idx), nq, kk, idx.stride(0), N.ptr(off_d), N.ptr(rel), int(off[-1]), int(n_mem) + 1,
           N.ptr(rank), N.stream_ptr())
    return off_d, rank, dev


def ap_at_k(idx, relevant, n_shots, k=None):
    """``APScorer(k).score`` of every query's ranked list against its relevant set, on the device.

    ``idx`` int64 ``[nq, kk]`` ranked shot rows (``-1`` padded), ``relevant`` one sequence of shot rows per query,
    ``n_shots`` the corpus size (the length of the full list the reference scorer would be given).  Relevant
    shots that are not in the returned list lie beyond position ``kk`` and contribute nothing, exactly as in
    ``basic/metric.py:36-44``; the denominator is the size of the relevant set.  Returns ``(ap float64 [nq], mAP)``
    with ``mAP = np.mean(ap)`` (``util/metrics.py:75-79`` style).
    """
    N.require_device()
    nq, kk = idx.shape
    k = kk if k is None else min(int(k), kk)
    off_d, rank, dev = _list_ranks(idx, relevant, n_shots)
    ap = torch.zeros(nq, dtype=torch.float64, device=dev)
    if nq == 0 or int(off_d[-1]) == 0:
        out = ap.cpu().numpy()
        return out, (np.mean(out) if nq else np.float64("nan"))
    from .metrics import rank_metrics
    rank_metrics(rank, off_d, nq, n_shots, False, k, max(len(r) for r in relevant), None, ap, None, None)
    out = ap.cpu().numpy()
    return out, np.mean(out)


def write_run_file(path, query_ids, idx, scores, shot_ids, run_tag="xmve"):
    """TREC run file ``qid Q0 shot rank score tag`` (one line per returned shot), the format AVS submissions and
    ``trec_eval`` read; replaces the ``pred_errors_matrix.pth.tar`` dump of ``tester.py:140`` for corpora whose
    score matrix cannot exist."""
    idx = idx.cpu().numpy() if torch.is_tensor(idx) else np.asarray(idx)
    scores = scores.cpu().numpy() if torch.is_tensor(scores) else np.asarray(scores)
    with open(path, "w") as f:
        for q, qid in enumerate(query_ids):
            for r in range(idx.shape[1]):
                if idx[q, r] < 0:
                    break
                f.write("%s Q0 %s %d %.17g %s\n" % (qid, shot_ids[int(idx[q, r])], r + 1, scores[q, r], run_tag))
