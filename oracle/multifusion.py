"""torch-CPU fp32 restatement of MultiFusion's composed-retrieval scoring (oracle; test-only).

Pinned: ``oracle/make_golden_mf.py`` runs the UNMODIFIED ``MultiFusion/src/validate.py`` and ``inference.py``
(stand-in modules for the absent ``clip`` / ``decord`` / ``h5py`` / ``comet_ml`` / ``ftfy`` imports, seeded
predictions in place of the CLIP + Combiner call) and commits their outputs under ``tests/golden/mf_cirr*``;
``tests/test_oracle_golden.py`` checks this file against them bit for bit.  It restates, line range by line range,

* ``MultiFusion/src/validate.py:44-55``  index preparation (8-frame mean in 128-row chunks via
  ``Combiner.time_process`` = ``fea.mean(dim=1)``, combiner.py:140-143; then
  ``F.normalize(dim=-1).float()``),
* ``MultiFusion/src/validate.py:65-113`` 32-query blocks: ``1 - P @ index.T``, ``torch.argsort`` on
  the CPU copy, removal of the query's own reference item, top-50 target labels,
* ``MultiFusion/src/validate.py:119``     the top-100 name dump,
* ``MultiFusion/src/validate.py:135-141`` recall@1/5/10/50 (+ three constant -1 group recalls),
* ``MultiFusion/src/inference.py:51,63-65`` the single-query top-1.

``Combiner.time_process`` is additionally pinned on its own (``tests/golden/mf_time_process.npz``).  One
deliberate difference: a query count that is a multiple of 32 makes the reference raise (empty trailing block,
``reshape(0, -1)``, validate.py:96-97; recorded in the golden); the restatement evaluates the full blocks.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def time_process(fea):
    """``fea.mean(dim=1)``; MultiFusion/src/combiner.py:140-143."""
    return fea.mean(dim=1)


def prepare_index(index_features, chunk=128):
    """validate.py:44-55.  ``[N, T, D]`` -> unit-norm fp32 ``[N, D]`` (``[N, D]`` passes through the pool)."""
    with torch.no_grad():
        if index_features.dim() == 3:
            n = len(index_features)
            parts = []
            for b in range(int(n / chunk) + 1):     # includes the trailing (possibly empty) block
                lo = b * chunk
                hi = (b + 1) * chunk if b < int(n / chunk) else n
                parts.append(time_process(index_features[lo:hi]))
            index_features = torch.cat(parts, dim=0)
        return F.normalize(index_features, dim=-1).float()


def compute_cirr_val_metrics(predicted_features, index_features, index_names, reference_names,
                             target_names, block=32, top_names=100):
    """validate.py:65-141 on already-built query features.

    Returns ``(7-tuple of metrics, sorted_index_names[:, :top_names] as int64 ndarray)``.
    ``index_names`` / ``reference_names`` / ``target_names`` are integer ids (utils.py:57).
    """
    index = prepare_index(index_features)
    names = torch.as_tensor(np.asarray(index_names))
    ref = torch.as_tensor(np.asarray(reference_names))
    tgt = torch.as_tensor(np.asarray(target_names))
    n_q, n_v = len(predicted_features), len(names)
    labels, ranked = [], []
    for b in range(int(n_q / block) + 1):
        lo = b * block
        hi = (b + 1) * block if b < int(n_q / block) else n_q
        tmp = 1 - predicted_features[lo:hi].float() @ index.T
        order = torch.argsort(tmp.cpu(), dim=-1)
        sorted_names = names[order]
        keep = sorted_names != ref[lo:hi].unsqueeze(1)
        sorted_names = sorted_names[keep].reshape(sorted_names.shape[0], n_v - 1)
        labels.append(sorted_names[:, :50] == tgt[lo:hi].unsqueeze(1))
        ranked.append(sorted_names[:, :top_names])
    labels = torch.cat(labels, dim=0)
    ranked = torch.cat(ranked, dim=0)
    recalls = [(torch.sum(labels[:, :k]) / len(labels)).item() * 100 for k in (1, 5, 10, 50)]
    return (-1, -1, -1, *recalls), ranked.numpy()


def top1_name(query_feature, index_features, index_names):
    """inference.py:51,63-65: ``index_names[argsort(1 - q @ normalize(index).T)[0][0]]``."""
    index = F.normalize(index_features, dim=-1).float()
    scores = 1 - query_feature.float() @ index.T
    return index_names[int(torch.argsort(scores.cpu(), dim=-1)[0][0])]
