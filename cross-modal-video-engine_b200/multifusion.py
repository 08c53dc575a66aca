"""Composed (text + reference video) retrieval scoring: the arithmetic of
``MultiFusion/src/validate.py:compute_cirr_val_metrics`` (:44-55 index preparation, :65-113 score /
rank / drop-the-reference / labels, :119 top-100 names, :135-141 recalls) and of
``MultiFusion/src/inference.py:51,63-65`` (single-query top-1), on the corpus-resident engine.

The reference builds ``predicted_features`` with CLIP + Combiner inside the same function
(validate.py:48-49 -> generate_cirr_val_predictions); those encoders are upstream of the scoring path,
so the entry points here take the already-built query features.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import CorpusStore


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


def compute_cirr_val_metrics(predicted_features, index_features, index_names, reference_names, target_names,
                             top_names=100, store=None):
    """Returns ``((group_r1, group_r2, group_r3, r@1, r@5, r@10, r@50), sorted_index_names[:, :top_names])``.

    The three group recalls are the constant -1 of validate.py:139-141.  ``index_names`` are the integer
    ids of utils.py:57; the query's own reference item is removed from its ranked list (validate.py:76-83)
    by excluding that corpus row in the selection kernel.
    """
    names = np.asarray(index_names, dtype=np.int64)
    ref = np.asarray(reference_names, dtype=np.int64)
    tgt = np.asarray(target_names, dtype=np.int64)
    if store is None:
        store = build_index(index_features)
    n_v = store.n
    # name -> corpus row of the reference item (names are unique ids)
    order = np.argsort(names, kind="stable")
    pos = np.searchsorted(names[order], ref)
    pos = np.clip(pos, 0, n_v - 1)
    ref_row = np.where(names[order][pos] == ref, order[pos], -1).astype(np.int64)
    k = min(max(top_names, 50), n_v - 1)
    q = predicted_features if torch.is_tensor(predicted_features) else torch.from_numpy(np.asarray(predicted_features))
    excl = np.where(ref_row >= 0, ref_row + store.index_offset, -1)
    _, idx = store.search(q.float(), k, exclude=excl)
    idx = idx.cpu().numpy()
    sorted_names = names[idx]                                              # [Nq, k]
    labels = sorted_names[:, :50] == tgt[:, None]
    n_q = len(labels)
    # torch: int64 sum / python int -> float32 division, then .item() * 100 (validate.py:135-138)
    recalls = [float(np.float32(np.sum(labels[:, :kk])) / np.float32(n_q)) * 100 for kk in (1, 5, 10, 50)]
    return (-1, -1, -1, *recalls), sorted_names[:, :top_names]


def top1_name(query_feature, index_features, index_names, store=None):
    """inference.py:51,63-65: name of the nearest index item (no reference removal there)."""
    if store is None:
        store = build_index(index_features)
    q = query_feature if torch.is_tensor(query_feature) else torch.from_numpy(np.asarray(query_feature))
    _, idx = store.search(q.float().reshape(1, -1), 1)
    return index_names[int(idx[0, 0])]
