"""Forward value of the training-time triplet ranking loss of ``LINAS-engine/loss.py`` (:7-73 similarity functions,
:87-153 ``TripletLoss``) on the scoring kernels -- SURVEY.md section 8f row 4.

``TripletLoss(...)(s, im)`` forms the batch score matrix ``sim(im, s)`` with ``xmve_pairwise_f64`` (the same tiled
CUDA-core kernel as the non-cosine measures of ``cal_error``; the embeddings arrive l2-normalised from the model, so
'cosine' is a plain dot product, loss.py:7-10) and reduces the hinge costs with ``xmve_triplet_cost``.  Same
constructor arguments, same value (fp64 arithmetic, returned as a float32 scalar tensor like the reference's).

This is the VALUE only: there is no backward pass here (training is outside the scoring path; use it for validation
loss curves or to check a training run's loss on the evaluation node).
"""
from __future__ import annotations

import torch

from . import _native as N

#: measure name -> (kernel measure, alpha as a function of the dim K, beta): score = alpha * f(im_i, s_j) + beta
_SIMS = {
    'cosine': (N.MEASURE_DOT, lambda k: 1.0, 0.0),                 # cosine_sim: im.mm(s.t())
    'order': (N.MEASURE_ORDER, lambda k: -1.0, 0.0),               # order_sim: -sqrt(sum clamp(s - im, 0)^2)
    'euclidean': (N.MEASURE_SQL2, lambda k: -1.0, 0.0),            # euclidean_sim: -sum (s - im)^2
    'jaccard': (N.MEASURE_JACCARD, lambda k: 1.0, 0.0),            # jaccard_sim: sum min / sum max
    'l1': (N.MEASURE_L1, lambda k: -1.0, 0.0),                     # L1_sim
    'l2': (N.MEASURE_SQL2, lambda k: -1.0, 0.0),                   # L2_sim (no root, like euclidean_sim)
    'l1_norm': (N.MEASURE_L1, lambda k: 1.0 / k, -1.0),            # L1_sim_norm: sum |d| / K - 1
    'l2_norm': (N.MEASURE_SQL2, lambda k: 1.0 / k, -1.0),          # L2_sim_norm: sum d^2 / K - 1
}


def get_sim(name):
    """``loss.get_sim`` (:76-78): a callable ``sim(im, s) -> [n_im, n_s]`` score matrix (fp32 device tensor)."""
    assert name in _SIMS, '%s not supported.' % name
    return lambda im, s: score_matrix(im, s, name).float()


def score_matrix(im, s, measure='cosine'):
    """``sim(im, s)`` of loss.py:7-73 as an fp64 device tensor ``[n_im, n_s]``."""
    N.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    a = torch.as_tensor(im).to(dev, torch.float64).contiguous()
    b = torch.as_tensor(s).to(dev, torch.float64).contiguous()
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1], "im [n, K] and s [m, K]"
    code, alpha, beta = _SIMS[measure]
    k = a.shape[1]
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device=dev)
    if out.numel():
        N.call("xmve_pairwise_f64", N.ptr(a), a.shape[0], k, N.ptr(b), b.shape[0], k, k, code, float(alpha(k)),
               float(beta), N.ptr(out), out.stride(0), N.stream_ptr())
    return out


class TripletLoss(torch.nn.Module):
    """triplet ranking loss (forward value); constructor as ``loss.TripletLoss`` (:87-110)."""

    def __init__(self, margin=0, measure=False, max_violation=False, cost_style='sum', direction='all'):
        super().__init__()
        self.margin = margin
        self.cost_style = cost_style
        self.direction = direction
        self.measure = measure if measure in _SIMS and measure != 'cosine' else 'cosine'   # :96-109: else cosine_sim
        self.max_violation = max_violation

    def forward(self, s, im):
        # compute video-sentence score matrix (:114); it is square: scores.diag() pairs caption i with video i
        scores = score_matrix(im, s, self.measure)
        n = scores.shape[0]
        assert scores.shape[1] == n, "TripletLoss pairs caption i with video i: equal batch sizes"
        out = torch.empty((2,), dtype=torch.float64, device=scores.device)
        N.call("xmve_triplet_cost", N.ptr(scores), n, scores.stride(0), float(self.margin),
               1 if self.max_violation else 0, N.ptr(out), N.stream_ptr())
        cost_s = out[0] if self.direction in ('v2t', 'all') else None       # caption retrieval (:129-132)
        cost_im = out[1] if self.direction in ('t2v', 'all') else None      # video retrieval (:134-137)
        zero = torch.zeros((), dtype=torch.float64, device=scores.device)   # the torch.zeros(1) of :146-149
        count = float(n if self.max_violation else n * n)                    # elements of cost_s / cost_im
        total = zero
        for c in (cost_s, cost_im):
            if c is None:
                continue
            total = total + (c if self.cost_style == 'sum' else c / count)   # .sum() / .mean() (:151-154)
        return total.float()
