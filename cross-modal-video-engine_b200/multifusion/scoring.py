"""Composed (text + reference video) retrieval scoring: the arithmetic of
``MultiFusion/src/validate.py:compute_cirr_val_metrics`` (:44-55 index preparation, :65-113 score /
rank / drop-the-reference / labels, :119 top-100 names, :135-141 recalls) and of
``MultiFusion/src/inference.py:51,63-65`` (single-query top-1), on the corpus-resident engine.

The reference builds ``predicted_features`` with CLIP + Combiner inside the same function
(validate.py:41-42 -> generate_cirr_val_predictions); those encoders are upstream of the scoring path, so the
functions here take the already-built query features.  :mod:`.validate` / :mod:`.inference` wrap them with the
reference's own signatures.
"""
from __future__ import annotations

import numpy as np
import torch

from ..engine import CorpusStore


def build_index(index_features, device="cuda", index_offset=0):
    """``[N, T, D]`` frame features (or pre-pooled ``[N, D]``) -> resident store.

    Fuses ``Combiner.time_process`` (mean over the T frames, combiner.py:140-143, run by the reference in
    128-row chunks, validate.py:44-53) and ``F.normalize(dim=-1).float()`` (validate.py:55) into K1.
    """
    n, d = index_features.shape[0], index_features.shape[-1]
    store = CorpusStore(n, (d,), device=device, norm_mode="eps", index_offset=index_offset)
    step = 1 << 18
    for lo in range(0, n, step):
        store.add(index_features[lo:lo + step])
    return store


def name_rows(index_names, wanted):
    """Corpus row of each id in ``wanted`` (``-1`` when it is not an index item).  The reference compares NAMES
    (validate.py:76-77), so with duplicated ids every copy would be dropped; ids are unique in its datasets
    (utils.py:57) and the first occurrence is taken here."""
    rows = {}
    for r, n in enumerate(np.asarray(index_names).tolist()):
        rows.setdefault(n, r)
    return np.fromiter((rows.get(n, -1) for n in np.asarray(wanted).tolist()), dtype=np.int64, count=len(wanted))


class CirrEvaluator:
    """The scoring stage of ``compute_cirr_val_metrics`` against ONE resident index, for callers that evaluate many
    query sets (every epoch: combiner_train.py:398): the store, the ``name -> row`` map and the device copy of the
    names are built once.  :meth:`metrics` is then one top-100 search with the reference item excluded, one
    ``xmve_list_ranks`` for the position of every target and four counts."""

    def __init__(self, index_features, index_names, store=None):
        self.names = np.asarray(index_names, dtype=np.int64)
        self.store = store if store is not None else build_index(index_features)
        self.rows = {}
        for r, n in enumerate(self.names.tolist()):
            self.rows.setdefault(n, r)
        self.names_dev = torch.from_numpy(self.names).to(self.store.device)

    def _rows_of(self, wanted):
        return np.fromiter((self.rows.get(n, -1) for n in np.asarray(wanted).tolist()), dtype=np.int64,
                           count=len(wanted))

    def search(self, predicted_features, reference_names, k, defer=False):
        """Top-``k`` rows per query with the query's own reference item dropped (validate.py:76-83)."""
        store = self.store
        ref_row = self._rows_of(reference_names)
        excl = np.where(ref_row >= 0, ref_row + store.index_offset, -1)
        q = predicted_features if torch.is_tensor(predicted_features) else \
            torch.from_numpy(np.asarray(predicted_features))
        return store.search(q.float(), k, exclude=excl, defer=defer)

    def metrics(self, predicted_features, reference_names, target_names, top_names=100):
        """``((-1, -1, -1, r@1, r@5, r@10, r@50), sorted_index_names[:, :top_names])`` (validate.py:135-143, :119)."""
        from .. import _native as N
        store = self.store
        dev = store.device
        n_q = len(target_names)
        k = min(max(top_names, 50), store.n - 1)
        _, idx = self.search(predicted_features, reference_names, k)
        # labels[q, j] = sorted_names[q, j] == target[q] (validate.py:84-87): names are unique ids, so this is the
        # position of the target's ROW in the list
        t_rows = self._rows_of(target_names)
        tgt_row = torch.from_numpy(np.where(t_rows >= 0, t_rows + store.index_offset, -2)).to(dev)   # -2: never listed
        off = torch.arange(n_q + 1, dtype=torch.int64, device=dev)
        pos = torch.empty((max(n_q, 1),), dtype=torch.int32, device=dev)
        if n_q:
            N.call("xmve_list_ranks", N.ptr(idx), n_q, min(50, k), idx.stride(0), N.ptr(off), N.ptr(tgt_row), n_q,
                   1 << 30, N.ptr(pos), N.stream_ptr())
        pos = pos[:n_q].cpu().numpy()
        sorted_names = self.names_dev[(idx[:, :top_names] - store.index_offset).clamp(min=0)].cpu().numpy()
        # torch: int64 sum / python int -> float32 division, then .item() * 100 (validate.py:135-138)
        recalls = [float(np.float32(np.sum(pos <= kk)) / np.float32(n_q)) * 100 for kk in (1, 5, 10, 50)]
        return (-1, -1, -1, *recalls), sorted_names


def cirr_metrics_from_features(predicted_features, index_features, index_names, reference_names, target_names,
                               top_names=100, store=None):
    """Returns ``((group_r1, group_r2, group_r3, r@1, r@5, r@10, r@50), sorted_index_names[:, :top_names])``.

    The three group recalls are the constant -1 of validate.py:139-141.  ``index_names`` are the integer
    ids of utils.py:57; the query's own reference item is removed from its ranked list (validate.py:76-83)
    by excluding that corpus row in the selection kernel.  Unlike the reference (whose empty trailing block
    raises in ``reshape(0, -1)``, validate.py:96-97) a query count that is a multiple of 32 is fine.
    One-shot form of :class:`CirrEvaluator`.
    """
    return CirrEvaluator(index_features, index_names, store=store).metrics(
        predicted_features, reference_names, np.asarray(target_names, dtype=np.int64), top_names)


def top1_name(query_feature, index_features, index_names, store=None):
    """inference.py:51,63-65: name of the nearest index item (no reference removal there)."""
    if store is None:
        store = build_index(index_features)
    q = query_feature if torch.is_tensor(query_feature) else torch.from_numpy(np.asarray(query_feature))
    _, idx = store.search(q.float().reshape(1, -1), 1)
    return index_names[int(idx[0, 0]) - store.index_offset]
