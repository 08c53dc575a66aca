"""``getScorer('AP')`` / ``getScorer('AP@k')`` of ``LINAS-engine/basic/metric.py:25-46,118-125``.

Only the AP scorer is ever reached from the scoring path (util/metrics.py:66,88); the label list is turned
into the 1-based ranks of its relevant entries and reduced by the same device kernel as the matrix path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N


class APScorer:
    def __init__(self, k=0):
        self.k = k

    def name(self):
        return "AP@%d" % self.k if self.k > 0 else "AP"

    def getLength(self, sorted_labels):
        length = self.k
        if length > len(sorted_labels) or length <= 0:
            length = len(sorted_labels)
        return length

    def score(self, sorted_labels):
        labels = np.asarray(sorted_labels)
        ranks = (np.nonzero(labels > 0)[0] + 1).astype(np.int32)
        if ranks.size == 0:
            return 0.0
        N.require_device()
        dev = torch.device("cuda", torch.cuda.current_device())
        r = torch.from_numpy(ranks).to(dev)
        off = torch.tensor([0, ranks.size], dtype=torch.int64, device=dev)
        ap = torch.empty(1, dtype=torch.float64, device=dev)
        from .metrics import rank_metrics
        rank_metrics(r, off, 1, len(labels), False, self.k, int(ranks.size), None, ap, None, None)
        return float(ap.item())


def getScorer(name):
    elems = name.split("@")
    if elems[0] != "AP":
        raise NotImplementedError("only the AP scorer is on the retrieval scoring path (got %r)" % name)
    return APScorer(int(elems[1]) if len(elems) == 2 else 0)
