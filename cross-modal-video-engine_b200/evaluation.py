"""Drop-in for the scoring functions of ``LINAS-engine/evaluation.py`` (same names, argument order,
return conventions), running on hand-written sm_100a kernels through libxmve.

* ``l2norm(X)``                       evaluation.py:10-14
* ``cal_error(videos, captions)``     evaluation.py:17-36 (cosine branch) -> errors = -cosine, [Nq, Nv]
* ``cal_error_batch(...)``            evaluation.py:41-72 (cosine branch is identical)
* ``cal_simi(captions, videos)``      evaluation.py:75-84 (+cosine; NOTE the swapped argument order)
* ``encode_vid`` / ``encode_text``    evaluation.py:87-171, without the per-batch ``.cpu().numpy()`` round trip: the
  embeddings stay on the encoder's device as one float64 tensor (the reference keeps float64 arrays too), ready for
  ``cal_error`` / ``CorpusStore.add``

dtype follows the input, like the reference: float64 arrays (what ``encode_vid`` / ``encode_text``
produce, evaluation.py:102,134) are scored by the exact fp64 kernel; float32 arrays by the tcgen05
kernel with split-bf16 (x3) operands, |error| ~ 1e-6.  ``measure='cosine'`` is the hot path (SURVEY.md section 8a
A2); the cdist / jaccard measures (section 8f row 4) run on a tiled CUDA-core kernel in fp64.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N

_BM, _BN = 128, 256


def _dev():
    N.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def _host_in(x):
    """ndarray / tensor -> (CUDA tensor fp32|fp64 contiguous, was_numpy)."""
    was_numpy = not torch.is_tensor(x)
    t = torch.from_numpy(np.ascontiguousarray(x)) if was_numpy else x
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t.to(_dev(), non_blocking=True).contiguous(), was_numpy


def _out(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


def _is_f64(x):
    return (x.dtype == torch.float64) if torch.is_tensor(x) else (getattr(x, "dtype", None) == np.float64)


def _dt(t):
    return N.F64 if t.dtype == torch.float64 else N.F32


def l2norm(X):
    """Row-wise ``X / ||X||`` (no epsilon), dtype preserved; evaluation.py:10-14."""
    x, was_numpy = _host_in(X)
    n, d = x.shape
    out = torch.empty((n, d), dtype=torch.float64, device=x.device)
    N.call("xmve_normalize_f64", N.ptr(x), _dt(x), n, d, x.stride(0), N.ptr(out), d, N.NORM_PLAIN, N.stream_ptr())
    return _out(out.to(x.dtype), was_numpy)


def _round_up(x, m):
    return (x + m - 1) // m * m


def score_matrix(queries, corpus, alpha, norm_mode=N.NORM_PLAIN, fuse_into=None, fuse_w=1.0):
    """``alpha * l2norm(queries) @ l2norm(corpus).T`` on the device, dtype of the inputs.

    ``fuse_into`` (float64 inputs with an even dim only; returns None otherwise and the caller composes): instead of
    returning the matrix E, accumulate it into ``fuse_into`` in the kernel's epilogue -- ``fuse_into = fuse_w * E`` when
    ``fuse_into`` is a fresh ``("first", tensor)``, ``fuse_into += fuse_w * E`` for ``("add", tensor)`` -- with every
    product and sum rounded on its own, like ``xmve_fuse_accumulate`` on the stored matrix."""
    q, _ = _host_in(queries)
    v, _ = _host_in(corpus)
    if q.dtype != v.dtype:
        q, v = q.to(torch.float64), v.to(torch.float64)
    nq, d = q.shape
    nv = v.shape[0]
    assert v.shape[1] == d, "embedding dims differ"
    st = N.stream_ptr()
    if nq == 0 or nv == 0:
        return torch.empty((nq, nv), dtype=q.dtype, device=q.device)
    if fuse_into is not None and not (q.dtype == torch.float64 and d % 2 == 0 and nq and nv):
        return None
    if q.dtype == torch.float64:
        qn = torch.empty((nq, d), dtype=torch.float64, device=q.device)
        vn = torch.empty((nv, d), dtype=torch.float64, device=q.device)
        N.call("xmve_normalize_f64", N.ptr(q), N.F64, nq, d, q.stride(0), N.ptr(qn), d, norm_mode, st)
        N.call("xmve_normalize_f64", N.ptr(v), N.F64, nv, d, v.stride(0), N.ptr(vn), d, norm_mode, st)
        step = 1 << 21
        if fuse_into is not None:
            how, out = fuse_into
            assert out.shape == (nq, nv) and out.dtype == torch.float64 and out.stride(1) == 1
            for q0 in range(0, nq, step):
                q1 = min(nq, q0 + step)
                N.call("xmve_score_f64_fused", N.ptr(qn[q0:]), q1 - q0, d, N.ptr(vn), nv, d, d, float(alpha),
                       float(fuse_w), 1 if how == "first" else 0, N.ptr(out[q0:]), out.stride(0), st)
            return out
        out = torch.empty((nq, nv), dtype=torch.float64, device=q.device)
        for q0 in range(0, nq, step):
            q1 = min(nq, q0 + step)
            N.call("xmve_score_f64", N.ptr(qn[q0:]), q1 - q0, d, N.ptr(vn), nv, d, d, float(alpha),
                   N.ptr(out[q0:]), nv, st)
        return out
    # float32: tensor-core path, split-bf16 operands  A = [hi | hi | lo],  B = [hi | lo | hi]
    dpad = _round_up(d, 64)
    a_op = torch.zeros((_round_up(nq, _BM), 3 * dpad), dtype=torch.bfloat16, device=q.device)
    b_op = torch.zeros((_round_up(nv, _BN), 3 * dpad), dtype=torch.bfloat16, device=q.device)
    N.call("xmve_prepare_rows", N.ptr(q), N.F32, nq, d, 1, q.stride(0), None, 0, 0, None, None, N.ptr(a_op), 3 * dpad, 0,
           N.OP_X3_QUERY, 1.0, norm_mode, st)
    N.call("xmve_prepare_rows", N.ptr(v), N.F32, nv, d, 1, v.stride(0), None, 0, 0, None, None, N.ptr(b_op), 3 * dpad, 0,
           N.OP_X3_CORPUS, 1.0, norm_mode, st)
    out = torch.empty((nq, nv), dtype=torch.float32, device=q.device)
    N.call("xmve_score_store", N.ptr(a_op), nq, 3 * dpad, N.ptr(b_op), nv, 3 * dpad, 1, 3 * dpad, float(alpha),
           N.ptr(out), nv, st)
    return out


#: measure -> (kernel measure, alpha as a function of the dim, beta); evaluation.py:22-35
_PAIRWISE = {
    'euclidean': (N.MEASURE_L2, lambda d: 1.0, 0.0), 'l2': (N.MEASURE_L2, lambda d: 1.0, 0.0),
    'l1': (N.MEASURE_L1, lambda d: 1.0, 0.0),
    'l1_norm': (N.MEASURE_L1, lambda d: -1.0 / d, -1.0), 'l2_norm': (N.MEASURE_L2, lambda d: -1.0 / d, -1.0),
    'jaccard': (N.MEASURE_JACCARD, lambda d: -1.0, 0.0),
}


def pairwise_matrix(queries, corpus, measure):
    """The scipy ``cdist`` / ``jaccard_sim`` branches of ``cal_error`` on the CUDA cores, in fp64.  ``cdist`` always
    returns float64; ``jaccard_sim`` runs on ``torch.Tensor(...)`` = float32 in the reference (returned as float32)."""
    q, _ = _host_in(queries)
    v, _ = _host_in(corpus)
    nq, d = q.shape
    nv = v.shape[0]
    assert v.shape[1] == d, "embedding dims differ"
    if measure == 'jaccard':                                     # torch.Tensor(x): values rounded to float32 first
        q, v = q.float(), v.float()
    q, v = q.double().contiguous(), v.double().contiguous()
    code, alpha, beta = _PAIRWISE[measure]
    out = torch.empty((nq, nv), dtype=torch.float64, device=q.device)
    step = 1 << 21
    for q0 in range(0, nq, step):
        q1 = min(nq, q0 + step)
        N.call("xmve_pairwise_f64", N.ptr(q[q0:]), q1 - q0, d, N.ptr(v), nv, d, d, code, float(alpha(d)), float(beta),
               N.ptr(out[q0:]), nv, N.stream_ptr())
    return out.float() if measure == 'jaccard' else out


def cal_error(videos, captions, measure='cosine'):
    """errors[q, v] = -cos(caption q, video v) (evaluation.py:17-21) or one of the distance measures (:22-35)."""
    was_numpy = not torch.is_tensor(captions)
    if measure == 'cosine':
        return _out(score_matrix(captions, videos, -1.0), was_numpy)
    if measure not in _PAIRWISE:
        raise ValueError("unknown measure %r" % (measure,))
    return _out(pairwise_matrix(captions, videos, measure), was_numpy)


def cal_error_batch(videos, captions, measure='cosine', batch_size=2000):
    """evaluation.py:41-72: same results as ``cal_error`` (the reference batches only to bound the memory of its
    broadcasted jaccard; the kernel here never materialises the [Nq, Nv, D] tensor)."""
    return cal_error(videos, captions, measure)


def cal_simi(captions, videos, measure='cosine'):
    """+cos(caption, video); evaluation.py:75-79 (captions FIRST here, unlike ``cal_error``)."""
    was_numpy = not torch.is_tensor(captions)
    if measure == 'jaccard':                                     # evaluation.py:80-83: +jaccard_sim
        return _out(-pairwise_matrix(captions, videos, 'jaccard'), was_numpy)
    if measure != 'cosine':
        raise ValueError("cal_simi knows 'cosine' and 'jaccard' (evaluation.py:75-84), got %r" % (measure,))
    return _out(score_matrix(captions, videos, 1.0), was_numpy)


def fused_errors(video_spaces, caption_spaces, weights, mode='weighted-cosine'):
    """Multi-space fusion of error matrices (the reference keeps only vestiges: the
    ``--space`` / ``--measure_2`` flags, ``norm_score``, teacher + student embedding pairs):

    * ``'weighted-cosine'``: ``errors = sum_s w_s * cal_error(V_s, Q_s)``
    * ``'norm_score'``     : ``errors = sum_s w_s * norm_score(cal_error(V_s, Q_s))`` (validate.py:7-11 per space)

    One array per embedding space on each side; dtype and container follow the inputs like ``cal_error``.  For
    corpora whose matrix cannot exist use ``CorpusStore(dims=(D_1, D_2, ...)).search(q, k, weights=...)``, which
    folds the weights into ONE tensor-core contraction."""
    from .validate import norm_score
    if mode not in ('weighted-cosine', 'norm_score'):
        raise ValueError(mode)
    assert len(video_spaces) == len(caption_spaces) == len(weights) and len(weights) >= 1
    was_numpy = not torch.is_tensor(caption_spaces[0])
    acc = None
    for s_i, (V, Q, w) in enumerate(zip(video_spaces, caption_spaces, weights)):
        if mode == 'weighted-cosine':
            # float64 inputs: w_s * E_s is accumulated in the epilogue of the score kernel (no [Nq, Nv] round trip)
            nq_, nv_ = len(Q), len(V)
            if acc is None and _is_f64(Q) and _is_f64(V):
                acc = torch.empty((nq_, nv_), dtype=torch.float64, device=_dev())
                if score_matrix(Q, V, -1.0, fuse_into=("first", acc), fuse_w=w) is not None:
                    continue
                acc = None
            elif acc is not None and acc.dtype == torch.float64 and _is_f64(Q) and _is_f64(V) \
                    and score_matrix(Q, V, -1.0, fuse_into=("add", acc), fuse_w=w) is not None:
                continue
        e = score_matrix(Q, V, -1.0)
        if mode == 'norm_score':
            e = norm_score(e)
        if acc is None:
            acc = torch.empty_like(e)
        assert e.dtype == acc.dtype and e.shape == acc.shape, "spaces must agree in dtype and row counts"
        if e.numel():
            N.call("xmve_fuse_accumulate", N.ptr(acc), acc.stride(0), N.ptr(e), e.stride(0), _dt(e), e.shape[0], e.shape[1],
                   float(w), 1 if s_i == 0 else 0, N.stream_ptr())
    return _out(acc, was_numpy)


def _encode(encoder, data_loader, n_inputs, return_ids):
    embeddings = None
    n = len(data_loader.dataset)
    ids = [''] * n
    for batch in data_loader:
        datas, (idxs, data_ids) = batch[:n_inputs], batch[n_inputs:]
        emb = encoder(*datas)
        if embeddings is None:                                   # np.zeros((N, D)) in the reference: float64
            embeddings = torch.zeros((n, emb.size(1)), dtype=torch.float64, device=emb.device)
        rows = torch.as_tensor(list(idxs), dtype=torch.int64, device=emb.device)
        embeddings[rows] = emb.detach().to(torch.float64)        # stays on the device: no PCIe round trip
        for j, idx in enumerate(idxs):
            ids[idx] = data_ids[j]
        del datas
    if return_ids:
        return embeddings, ids
    return embeddings


def encode_vid(encoder, data_loader, return_ids=True):
    """evaluation.py:87-115: embeddings ``[len(dataset), D]`` scattered by dataset index, as ONE device tensor."""
    return _encode(encoder, data_loader, 1, return_ids)


def encode_text(encoder, data_loader, style, return_ids=True):
    """evaluation.py:118-171: ``style='distill_from_best_model'`` batches are ``(datas, idxs, ids)``, ``style='GT'``
    batches carry ``support_datas`` as a second encoder input."""
    if style == 'distill_from_best_model':
        return _encode(encoder, data_loader, 1, return_ids)
    if style == 'GT':
        return _encode(encoder, data_loader, 2, return_ids)
    return None                                                  # the reference falls through for other styles
