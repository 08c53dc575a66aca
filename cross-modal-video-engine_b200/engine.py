"""Corpus-resident store and the search pipeline (host orchestration of the sm_100a kernels).

Replaces, for a corpus that is encoded once and queried many times, the per-call
``cal_error(video_embs, cap_emb)`` + ``np.argsort(errors[0])[:topK]`` of
``LINAS-engine/inference.py:78-80`` (which re-normalises the whole corpus on every call,
``evaluation.py:19-20``) and the 32-query ``1 - P @ index.T`` / ``torch.argsort`` blocks of
``MultiFusion/src/validate.py:65-113``.

The pipeline is :func:`search_shards` (its docstring lists the steps): K1 on the queries, a sampled per-query
threshold, the fused tcgen05 score + threshold filter over the resident operand (the score matrix never reaches
HBM), an exact fp64 rescore of the survivors in two rounds, a bitonic top-k with a certificate that nothing
outside the candidate set can belong to the top-k, and a re-run of the rare rows that miss the certificate.
The first pass never waits for the host: the error bound is a device scalar, the number of uncertified rows comes
back through one asynchronous copy (``defer=True`` hands the caller a :class:`PendingSearch` to resolve later).
:meth:`CorpusStore.search` is the one-shard case; ``distributed.sharded_search`` joins the shards of several GPUs.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _native as N

#: a-priori bound on |bf16-operand score - exact cosine| per unit of fusion weight: bf16 keeps 8 significant bits,
#: so rounding moves a vector by at most 2^-8 of its norm and a product of two unit vectors by at most 2 * 2^-8
#: (Cauchy-Schwarz); plus fp32 accumulation slack.  Used only when no measured residuals are available
#: (search(eps=...) callers, NaN rows); the search itself uses :func:`measured_eps` (typically 3.5e-3 - 4e-3).
EPS_X1 = 8.5e-3
SMALL_NV = 16384
#: multiples of eps subtracted from the sampled threshold kth(sample, j) -- see :func:`thr_margin`.  Round 1 always
#: used 2: "the k-th best exact score is at least kth_approx - eps, a row's approximate score at most eps above its
#: exact one".  But j already carries a 5.5-sigma sampling slack (plan()): the j-th largest of the sample sits near
#: corpus rank j * step (1 500 at C5), an order of magnitude below the k-th best, so the certificate
#: kth_exact - eps >= thr holds without the extra margin (P(miss) ~ 1e-7 per row; a miss only costs a re-run of that
#: row).  Dropping it cuts the candidate lists from 5 800 to 1 500 per query at C5 and from 2 350 to 1 200 at C4.
THR_EPS_MARGIN = 0.0


def thr_margin(lam):
    """Multiples of eps to subtract from ``kth(sample, j)`` for ``lam = k / step`` expected top-k rows in the sample.

    The margin-free threshold lives on the score gap between corpus rank ``k`` and rank ``j * step``.  The rank ratio
    ``j / lam = 1 + 5.5 / sqrt(lam) + 4 / lam`` is >= 3.3 for ``lam <= 8.5`` (a gap of ~0.4 sigma of the score
    distribution against eps ~ 0.2 sigma) but tends to 1 for deep lists over small corpora (k = 1000 of 200 k rows:
    1.9), where the missing margin showed up as re-runs -- a whole extra FILTER pass for the rows concerned
    (profiles/r2_summary.md section 6).  So: none up to lam = 8.5, one eps up to 64, the rigorous two beyond.  A wrong
    guess costs time, never correctness (the certificate decides)."""
    if lam <= 8.5:
        return THR_EPS_MARGIN
    return max(THR_EPS_MARGIN, 1.0 if lam <= 64.0 else 2.0)
ROW_TOPJ_MAX = 4096
#: size of the strided corpus sample that sets the per-query threshold: n / SAMPLE_DIV rows, clamped to
#: [SAMPLE_MIN, SAMPLE_MAX] (plan()).  A larger sample costs a longer sampling pass and buys a tighter threshold,
#: i.e. fewer candidates for the FILTER epilogue, the order statistics and the rescore.
SAMPLE_DIV = int(os.environ.get("XMVE_SAMPLE_DIV", 128))
SAMPLE_MIN = int(os.environ.get("XMVE_SAMPLE_MIN", 8192))
SAMPLE_MAX = int(os.environ.get("XMVE_SAMPLE_MAX", 65536))
_BM, _BN = 128, 256


def _round_up(x, m):
    return (x + m - 1) // m * m


def _as_spaces(x, dims):
    """Split ``x`` (tensor ``[n, sum(dims)]`` / ``[n, T, sum(dims)]`` or a per-space list) into per-space views."""
    if isinstance(x, (list, tuple)):
        assert len(x) == len(dims), "expected one array per embedding space"
        return list(x)
    if len(dims) == 1:
        return [x]
    out, o = [], 0
    for d in dims:
        out.append(x[..., o:o + d])
        o += d
    return out


def _to_device(x, device):
    """numpy / torch, host / device -> CUDA tensor (fp32 or fp64), last dim contiguous."""
    if not torch.is_tensor(x):
        import numpy as np
        x = torch.from_numpy(np.ascontiguousarray(x))
    if x.dtype not in (torch.float32, torch.float64):
        x = x.float()
    x = x.to(device, non_blocking=True)
    return x


def _prepare(src, d, raw, raw_off, norm, resid, op, op_off, layout, weight, norm_mode):
    """One K1 launch for one embedding space of one batch of rows."""
    frames = 1
    if src.dim() == 3:
        frames = src.shape[1]
    n = src.shape[0]
    if n == 0:
        return
    if src.stride(-1) != 1 or (frames > 1 and src.stride(1) != d):
        src = src.contiguous()
    if frames > 1 and not src.is_contiguous():
        src = src.contiguous()
    N.call("xmve_prepare_rows", N.ptr(src), N.F64 if src.dtype == torch.float64 else N.F32, n, d, frames,
           src.stride(0),
           N.ptr(raw), raw.stride(0) if raw is not None else 0, raw_off,
           N.ptr(norm), N.ptr(resid),
           N.ptr(op), op.stride(0) if op is not None else 0, op_off, layout,
           float(weight), norm_mode, N.stream_ptr())


class CorpusStore:
    """Video-embedding corpus resident in HBM: bf16 operand rows + raw fp32 rows + fp64 norms.

    ``dims`` lists the embedding spaces (e.g. ``(1536, 512)`` for latent + concept); rows are appended
    with :meth:`add` (the analogue of ``encode_vid``'s batch loop, evaluation.py:98-105) and
    normalised once, on the device, as they arrive.  ``index_offset`` is this shard's first global
    row (multi-GPU sharding).

    Memory: ``capacity * (2 * sum(dpad) + 4 * sum(dims) + 12 * S)`` bytes -- 12.3 KB per row at 1536 + 512 dims, i.e.
    123 GB for the 10 M-row C5 corpus (41 GB bf16 operand + 82 GB fp32 raw rows kept for the exact rescore) and a
    ceiling of about 14 M such rows per 180 GB B200; larger corpora are sharded (``distributed.shard_range``).  The
    constructor checks the free device memory first and raises a sized ``MemoryError`` instead of a CUDA OOM.

    Inputs are taken as the reference keeps them: fp32 model outputs, possibly stored in float64 arrays
    (evaluation.py:102-105).  float64 inputs are therefore narrowed to fp32 by K1 (exact for such arrays); the
    fp64 arithmetic of the rescore then equals ``cal_error`` on the float64 copies.  Genuinely double-precision
    embeddings would be ranked after rounding to fp32 -- use ``evaluation.cal_error`` for those.
    """

    def __init__(self, capacity, dims, device="cuda", norm_mode="plain", index_offset=0):
        N.require_device()
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dims = tuple(int(d) for d in (dims if isinstance(dims, (list, tuple)) else (dims,)))
        self.dpads = tuple(_round_up(d, 64) for d in self.dims)
        self.k = sum(self.dpads)
        self.dtot = sum(self.dims)
        self.raw_ld = _round_up(self.dtot, 4)
        self.capacity = int(capacity)
        self.norm_mode = {"plain": N.NORM_PLAIN, "eps": N.NORM_EPS}[norm_mode]
        self.index_offset = int(index_offset)
        self.n = 0
        rows = _round_up(max(self.capacity, 1), _BN)
        need = rows * self.k * 2 + max(self.capacity, 1) * (self.raw_ld * 4 + len(self.dims) * 12)
        free = torch.cuda.mem_get_info(self.device)[0] + torch.cuda.memory_reserved(self.device) \
            - torch.cuda.memory_allocated(self.device)
        if need > free:
            raise MemoryError(
                "CorpusStore: %d rows x %s dims need %.1f GB of HBM (bf16 operand + fp32 raw rows + norms) but only "
                "%.1f GB are free on %s; shard the corpus over more GPUs (distributed.shard_range) or lower the "
                "capacity" % (self.capacity, list(self.dims), need / 1e9, free / 1e9, self.device))
        self.op = torch.zeros((rows, self.k), dtype=torch.bfloat16, device=self.device)
        self.raw = torch.empty((max(self.capacity, 1), self.raw_ld), dtype=torch.float32, device=self.device)
        self.norm = torch.empty((len(self.dims), max(self.capacity, 1)), dtype=torch.float64, device=self.device)
        self.resid = torch.zeros((len(self.dims), max(self.capacity, 1)), dtype=torch.float32, device=self.device)
        self._resid_max2 = None
        self.space_off = (C.c_int32 * (len(self.dims) + 1))(*([0] + list(_cumsum(self.dims))))

    # -- ingest ----------------------------------------------------------------------------------
    def add(self, rows):
        """Append rows: tensor / ndarray ``[n, sum(dims)]`` (or ``[n, T, D]``: mean over T first), or a per-space list."""
        spaces = [_to_device(s, self.device) for s in _as_spaces(rows, self.dims)]
        n = spaces[0].shape[0]
        if self.n + n > self.capacity:
            raise ValueError("CorpusStore capacity %d exceeded" % self.capacity)
        op_off = raw_off = 0
        with torch.cuda.device(self.device):                  # launch on the store's device, whatever is current
            for s, (src, d, dp) in enumerate(zip(spaces, self.dims, self.dpads)):
                assert src.shape[-1] == d, "space %d: expected dim %d, got %d" % (s, d, src.shape[-1])
                _prepare(src, d, self.raw[self.n:], raw_off, self.norm[s, self.n:], self.resid[s, self.n:],
                         self.op[self.n:], op_off, N.OP_X1, 1.0, self.norm_mode)
                op_off += dp
                raw_off += d
        self.n += n
        return self

    # -- query side ------------------------------------------------------------------------------
    def prepare_queries(self, queries, weights=None):
        """K1 on the query batch: returns (a_op bf16 [nq_pad, K], q_raw fp32 [nq, raw_ld], q_norm fp64 [S, nq],
        q_resid fp32 [S, nq] squared bf16 residuals, nq)."""
        weights = _weights(weights, len(self.dims))
        spaces = [_to_device(s, self.device) for s in _as_spaces(queries, self.dims)]
        nq = spaces[0].shape[0]
        a_op = torch.zeros((_round_up(max(nq, 1), _BM), self.k), dtype=torch.bfloat16, device=self.device)
        q_raw = torch.empty((max(nq, 1), self.raw_ld), dtype=torch.float32, device=self.device)
        q_norm = torch.empty((len(self.dims), max(nq, 1)), dtype=torch.float64, device=self.device)
        q_res = torch.zeros((len(self.dims), max(nq, 1)), dtype=torch.float32, device=self.device)
        op_off = raw_off = 0
        for s, (src, d, dp) in enumerate(zip(spaces, self.dims, self.dpads)):
            _prepare(src, d, q_raw, raw_off, q_norm[s], q_res[s], a_op, op_off, N.OP_X1, weights[s], self.norm_mode)
            op_off += dp
            raw_off += d
        return a_op, q_raw, q_norm, q_res, nq

    # -- search ----------------------------------------------------------------------------------
    def search(self, queries, k, weights=None, exclude=None, eps=None, small_nv=SMALL_NV, stats=None, defer=False):
        """Top-``k`` corpus rows per query by fused cosine score, exact (fp64) scores, descending.

        Returns ``(scores float64 [nq, k], idx int64 [nq, k])`` on the device; ``idx`` are global row
        numbers (``index_offset`` + local), ``-1`` / ``-inf`` padded if the corpus has fewer than ``k``
        rows.  ``exclude[q]`` (global row or -1) is dropped from row q's list -- MultiFusion's removal
        of the query's own reference item (validate.py:76-83).  ``eps`` overrides the measured bound on
        |tensor-core score - exact score| (see :func:`measured_eps`).  ``defer=True`` returns a
        :class:`PendingSearch` as soon as the work is enqueued (``.result()`` gives the tuple).
        """
        return search_shards([self], queries, k, weights=weights, exclude=exclude, eps=eps, small_nv=small_nv,
                             stats=stats, defer=defer)

    def plan(self, k, n=None):
        return plan(k, self.n if n is None else n)

    def resid_max2(self):
        """max over rows of the squared bf16 quantisation residual of the operand (device scalar, cached)."""
        if self._resid_max2 is None or self._resid_max2[0] != self.n:
            r = self.resid[:, :self.n].sum(0).max() if self.n else torch.zeros((), device=self.device)
            self._resid_max2 = (self.n, r)
        return self._resid_max2[1]

    def _weights_arr(self, wts):
        return (C.c_double * len(wts))(*[float(w) for w in wts])

    def _exact_small(self, q_raw, q_norm, nq, k, wts, excl):
        """fp64 score matrix + bitonic top-k (corpora up to ``small_nv`` rows)."""
        dev, st = self.device, N.stream_ptr()
        acc = None
        o = 0
        for s, d in enumerate(self.dims):
            qn = torch.empty((nq, d), dtype=torch.float64, device=dev)
            vn = torch.empty((self.n, d), dtype=torch.float64, device=dev)
            qs, vs = q_raw[:nq, o:o + d], self.raw[:self.n, o:o + d]
            N.call("xmve_normalize_f64", N.ptr(qs), N.F32, nq, d, q_raw.stride(0), N.ptr(qn), d, self.norm_mode, st)
            N.call("xmve_normalize_f64", N.ptr(vs), N.F32, self.n, d, self.raw.stride(0), N.ptr(vn), d,
                   self.norm_mode, st)
            sc = torch.empty((nq, self.n), dtype=torch.float64, device=dev)
            for q0 in range(0, nq, 1 << 21):
                q1 = min(nq, q0 + (1 << 21))
                N.call("xmve_score_f64", N.ptr(qn[q0:]), q1 - q0, d, N.ptr(vn), self.n, d, d, float(wts[s]),
                       N.ptr(sc[q0:]), self.n, st)
            acc = sc if acc is None else acc.add_(sc)
            o += d
        kk = min(k, self.n)
        out_s = torch.full((nq, k), float("-inf"), dtype=torch.float64, device=dev)
        out_i = torch.full((nq, k), -1, dtype=torch.int64, device=dev)
        s_ = torch.empty((nq, kk), dtype=torch.float64, device=dev)
        i_ = torch.empty((nq, kk), dtype=torch.int64, device=dev)
        N.call("xmve_select_topk_i32", N.ptr(acc), None, nq, self.n, None, self.index_offset, N.ptr(excl), kk,
               None, 0.0, None, None, N.ptr(s_), N.ptr(i_), None, None, None, None, st)
        out_s[:, :kk] = s_
        out_i[:, :kk] = i_
        return out_s, out_i

    # one shard's share of a filtered pass ----------------------------------------------------------
    def _sample(self, a_op, nq, step):
        """K2 STORE over every ``step``-th row of this shard -> fp32 ``[nq, ceil(n / step)]``."""
        n_s = (self.n + step - 1) // step
        sample = torch.empty((nq, n_s), dtype=torch.float32, device=self.device)
        N.call("xmve_score_store", N.ptr(a_op), nq, a_op.stride(0), N.ptr(self.op), n_s, self.op.stride(0), step,
               self.k, 1.0, N.ptr(sample), sample.stride(0), N.stream_ptr())
        return sample

    def _filter(self, a_op, nq, thr, cap, step=1):
        """K2 FILTER over the shard (or over every ``step``-th row): candidates (approximate score, local row --
        sampled-row number when ``step > 1``) above ``thr`` per query."""
        dev = self.device
        rows = (self.n + step - 1) // step
        cand_count = torch.zeros((nq,), dtype=torch.int32, device=dev)
        cand_score = torch.empty((nq, cap), dtype=torch.float32, device=dev)
        cand_idx = torch.empty((nq, cap), dtype=torch.int32, device=dev)
        N.call("xmve_score_filter", N.ptr(a_op), nq, a_op.stride(0), N.ptr(self.op), rows, self.op.stride(0), step,
               self.k, N.ptr(thr), None, None, N.ptr(cand_count), N.ptr(cand_score), N.ptr(cand_idx), cap,
               N.stream_ptr())
        return cand_count, cand_score, cand_idx

    def _filter_band(self, a_op, nq, lo, hi, cap):
        """K2 FILTER with a two-sided window per row: rows scoring above ``hi`` are counted, rows in ``(lo, hi]``
        are listed -> (above int32 [nq], count int32 [nq], scores fp32 [nq, cap], local rows int32 [nq, cap])."""
        dev = self.device
        above = torch.zeros((nq,), dtype=torch.int32, device=dev)
        count = torch.zeros((nq,), dtype=torch.int32, device=dev)
        c_s = torch.empty((nq, cap), dtype=torch.float32, device=dev)
        c_i = torch.empty((nq, cap), dtype=torch.int32, device=dev)
        N.call("xmve_score_filter", N.ptr(a_op), nq, a_op.stride(0), N.ptr(self.op), self.n, self.op.stride(0), 1,
               self.k, N.ptr(lo), N.ptr(hi), N.ptr(above), N.ptr(count), N.ptr(c_s), N.ptr(c_i), cap, N.stream_ptr())
        return above, count, c_s, c_i

    def _sample_top(self, a_op, nq, step, big_j):
        """The largest scores of the ``step``-strided sample of this shard WITHOUT writing the sample matrix:
        a coarse STORE pass (every ``r * step``-th row, a few thousand columns) gives a per-query floor ``thr0``
        that about ``4 * big_j`` of the fine sample's scores exceed; a FILTER pass over the fine sample keeps those.
        Returns ``(scores fp32 [nq, cap_s], counts int32 [nq], thr0 fp32 [nq])``; rows whose count is below
        ``big_j`` (the floor came out too high -- vanishingly rare) are handled by the caller through ``thr0``."""
        n_s = (self.n + step - 1) // step
        r = max(2, min(16, n_s // 2048))
        coarse = self._sample(a_op, nq, step * r)
        j0 = min(coarse.shape[1], int(math.ceil(4.0 * big_j / r)))
        thr0 = _row_kth(coarse, None, j0, 0.0, 0)
        cap_s = 1 << max(10, int(math.ceil(math.log2(16 * big_j))))
        count, score, _ = self._filter(a_op, nq, thr0, cap_s, step=step)
        return score, count, thr0

    def _rescore(self, q_raw, q_norm, nq, wts, cand, bound, bound_hi=None, exact=None):
        """Exact fp64 scores of the candidates whose approximate score reaches ``bound`` (second round: of those in
        ``[bound, bound_hi)``, the entries at or above ``bound_hi`` keep their first-round values) -> fp64 [nq, cap]."""
        cand_count, cand_score, cand_idx = cand
        cap = cand_score.shape[1]
        if exact is None:
            exact = torch.empty((nq, cap), dtype=torch.float64, device=self.device)
        N.call("xmve_rescore", N.ptr(q_raw), nq, q_raw.stride(0), N.ptr(q_norm), N.ptr(self.raw), self.n,
               self.raw.stride(0), N.ptr(self.norm), len(self.dims), self.space_off, self._weights_arr(wts),
               self.norm_mode, N.ptr(cand_score), N.ptr(cand_idx), N.ptr(cand_count), cap, N.ptr(bound),
               N.ptr(bound_hi), N.ptr(exact), N.stream_ptr())
        return exact

    def _pilot_top(self, exact, cand, nq, excl, m):
        """The ``m`` largest first-round exact scores per query, descending, -inf padded -> fp64 [nq, m]."""
        cand_count, cand_score, cand_idx = cand
        out = torch.empty((nq, m), dtype=torch.float64, device=self.device)
        N.call("xmve_pilot_top", N.ptr(exact), N.ptr(cand_idx), N.ptr(cand_count), nq, cand_score.shape[1],
               self.index_offset, N.ptr(excl), m, N.ptr(out), N.stream_ptr())
        return out

    def _select(self, exact, cand, nq, k, excl, thr, eps_t, bound, out_s, out_i, certify, n_bad=None):
        """Local top-k (global row ids) of the rescored candidates into ``out_s`` / ``out_i`` (+ certificate)."""
        cand_count, cand_score, cand_idx = cand
        cap = cand_score.shape[1]
        cert = thr_next = None
        if certify:
            cert = torch.empty((nq,), dtype=torch.int32, device=self.device)
            thr_next = torch.empty((nq,), dtype=torch.float32, device=self.device)
        N.call("xmve_select_topk_i32", N.ptr(exact), N.ptr(cand_idx), nq, cap, N.ptr(cand_count),
               self.index_offset, N.ptr(excl), k, N.ptr(thr) if certify else None, 1.0, N.ptr(eps_t),
               N.ptr(bound) if certify else None, N.ptr(out_s), N.ptr(out_i), None, N.ptr(cert), N.ptr(thr_next),
               N.ptr(n_bad) if certify else None, N.stream_ptr())
        return cert, thr_next


def plan(k, n):
    """Sampling step, order statistics and candidate capacity for a top-``k`` search of ``n`` corpus rows."""
    n = int(n)
    n_s = min(n, max(SAMPLE_MIN, min(SAMPLE_MAX, n // SAMPLE_DIV)))
    step = max(1, n // max(n_s, 1))
    n_s = (n + step - 1) // step
    lam = k / step
    j = int(math.ceil(lam + 5.5 * math.sqrt(lam) + 4))
    cap = 1 << max(11, int(math.ceil(math.log2(8 * step * j))))
    cap = min(cap, 32768)
    j_cap = max(j, int(0.5 * cap / step))
    return {"step": step, "n_sample": n_s, "j": min(j, n_s), "j_cap": min(j_cap, n_s), "cap": cap,
            "margin": thr_margin(lam)}


def measured_eps(dq, dv, wts, n_space, k_len):
    """Rigorous bound on |tensor-core score - exact score| from the MEASURED bf16 residuals.

    With Q = concat_s(w_s q_hat_s), V = concat_s(v_hat_s) and Qb, Vb their bf16 roundings,
    ``<Q,V> - <Qb,Vb> = <Q-Qb, V> + <Qb, V-Vb>``, so by Cauchy-Schwarz the operand rounding costs at most
    ``dq * |V| + |Qb| * dv`` with ``dq = max_q |Q-Qb|``, ``dv = max_v |V-Vb|`` (K1 measures both), ``|V| = sqrt(S)``
    and ``|Qb| <= sqrt(sum w^2) + dq``.  The bf16 x bf16 products are exact in fp32; accumulating ``k_len`` of them
    in fp32 (with truncation at worst) costs at most ``k_len * 2^-22 * |Qb| * |Vb|``.  The a-priori worst case is
    2 * 2^-8 per unit weight (``EPS_X1``); Gaussian-like rows measure ``dq, dv ~ 1.7e-3`` of the norm, i.e. about half.
    """
    qn = math.sqrt(sum(w * w for w in wts)) + dq
    vn = math.sqrt(n_space) * (1.0 + 2.0 ** -8)
    e = dq * vn + qn * dv + k_len * 2.0 ** -22 * qn * vn + 1e-6
    if not math.isfinite(e) or e <= 0.0:                   # NaN rows (zero vectors under 'plain' normalisation)
        return EPS_X1 * max(1.0, sum(abs(w) for w in wts))
    return e


class _Phases:
    """CUDA-event marks at the phase boundaries of a search (only when the caller passes ``stats``)."""

    def __init__(self, on):
        self.on, self.marks = on, []

    def mark(self, name):
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def result(self):
        if len(self.marks) < 2:
            return {}
        self.marks[-1][1].synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


class SoloComm:
    """The collective interface of :func:`search_shards` for a single process (world size 1)."""
    world = 1

    def gather(self, t):
        return t.unsqueeze(0)

    def max_(self, t):
        return t

    def sum_int(self, v):
        return int(v)

    def sum_(self, t):
        return t


def _row_topj(vals, counts, j):
    rows, cols = vals.shape
    out = torch.empty((rows, j), dtype=torch.float32, device=vals.device)
    if rows:
        N.call("xmve_row_topj", N.ptr(vals), rows, cols, vals.stride(0), N.ptr(counts), j, N.ptr(out), N.stream_ptr())
    return out


def _row_kth(vals, counts, j1, sub, j2, sub_dev=None):
    """``max(kth(row, j1) - sub * sub_dev[0], kth(row, j2))`` per row (``sub_dev`` None: ``- sub``)."""
    rows, cols = vals.shape
    out = torch.empty((rows,), dtype=torch.float32, device=vals.device)
    if rows:
        N.call("xmve_row_kth", N.ptr(vals), rows, cols, vals.stride(0), N.ptr(counts), j1, float(sub), N.ptr(sub_dev), j2,
               N.ptr(out), N.stream_ptr())
    return out


def _eps_device(q_res, dv2, wts, n_space, k_len):
    """:func:`measured_eps` on the device: fp32 ``[1]`` tensor (no host round trip)."""
    out = torch.empty((1,), dtype=torch.float32, device=q_res.device)
    N.call("xmve_eps_bound", N.ptr(q_res), n_space, q_res.shape[1], N.ptr(dv2), math.sqrt(sum(w * w for w in wts)),
           int(k_len), EPS_X1 * max(1.0, sum(abs(w) for w in wts)), N.ptr(out), N.stream_ptr())
    return out


def _count_before(exact, cand_idx, counts, idx_offset, s_gt, g, out):
    """``out[e] += #{candidates of entry e that precede its ground truth}`` (exact score above ``s_gt[e]``, or equal
    with a smaller global row than ``g[e]``)."""
    n_ent, cap = exact.shape
    if n_ent:
        N.call("xmve_count_before", N.ptr(exact), N.ptr(cand_idx), N.ptr(counts), n_ent, cap, int(idx_offset),
               N.ptr(s_gt), N.ptr(g), N.ptr(out), N.stream_ptr())


def _pilot_bound(lists, k, eps_t):
    """``lists`` fp64 ``[n_seg, nq, m]`` (descending per segment) -> round_down(k-th largest of the union - eps)."""
    n_seg, nq, m = lists.shape
    out = torch.empty((nq,), dtype=torch.float32, device=lists.device)
    if nq:
        N.call("xmve_pilot_bound", N.ptr(lists), n_seg, nq, m, k, 1.0, N.ptr(eps_t), N.ptr(out), N.stream_ptr())
    return out


def packed_bytes(rows, length):
    """Bytes of one packed top-k block: scores fp64 [rows, len] | rows int64 [rows, len] | overflow int32 [rows]."""
    return int(N.lib.xmve_packed_topk_bytes(int(rows), int(length)))


def packed_views(block, rows, length):
    """Typed views (scores, idx, flags) into one packed block (a uint8 tensor of :func:`packed_bytes` bytes)."""
    n = rows * length
    return (block[: n * 8].view(torch.float64).view(rows, length),
            block[n * 8: n * 16].view(torch.int64).view(rows, length),
            block[n * 16: n * 16 + rows * 4].view(torch.int32))


def _merge_packed(packed, rows, length, k, thr, eps_t, n_bad):
    """K3 on the all-gathered packed blocks ``[n_seg, seg_bytes]`` -> (scores, idx, cert, thr_next)."""
    n_seg, seg_bytes = packed.shape
    out_s = torch.empty((rows, k), dtype=torch.float64, device=packed.device)
    out_i = torch.empty((rows, k), dtype=torch.int64, device=packed.device)
    cert = torch.empty((rows,), dtype=torch.int32, device=packed.device)
    thr_next = torch.empty((rows,), dtype=torch.float32, device=packed.device)
    if rows:
        N.call("xmve_merge_topk_packed", N.ptr(packed), n_seg, seg_bytes, rows, length, None, k, N.ptr(thr), 1.0,
               N.ptr(eps_t), N.ptr(out_s), N.ptr(out_i), N.ptr(cert), N.ptr(thr_next), N.ptr(n_bad), N.stream_ptr())
    return out_s, out_i, cert, thr_next


def _union(parts, comm):
    """Per-row lists from the local shards ``[nq, m]`` -> the union over all shards of all ranks ``[nq, G*L*m]``."""
    loc = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
    if comm.world == 1:
        return loc
    g = comm.gather(loc.contiguous())                                   # [G, nq, L*m]
    return g.permute(1, 0, 2).reshape(loc.shape[0], -1).contiguous()


def _segments(parts, comm):
    """Per-shard tensors of one shape -> ``[G * L, *shape]`` over all shards of all ranks (one collective)."""
    loc = parts[0].unsqueeze(0) if len(parts) == 1 else torch.stack(parts)
    if comm.world == 1:
        return loc
    g = comm.gather(loc)                                                # [G, L, ...]
    return g.view((-1,) + tuple(loc.shape[1:]))


#: the pilot holds at most this many exact scores per query and shard (xmve_pilot_top)
PILOT_MAX = 1000
#: ... and the union of the pilots of all shards at most this many (xmve_pilot_bound sorts it in shared memory)
PILOT_UNION_MAX = 4096


class PendingSearch:
    """A search whose first pass has been enqueued.  :meth:`result` waits for the count of uncertified rows (one
    int32 copied asynchronously to pinned host memory right after the merge), re-runs those rows if there are any,
    and returns ``(scores, idx)``.  Resolving a search one step late lets the host enqueue the next batch while the
    GPU still scores this one (bench.py does that); ``search_shards(defer=False)`` resolves at once."""

    def __init__(self, ctx):
        self._ctx = ctx
        self._ctx_template = None
        self._done = ctx is None
        self.scores = self.idx = None
        self.reran = False                    # True once result() had to re-run rows (scores / idx were rewritten)

    @classmethod
    def ready(cls, scores, idx):
        p = cls(None)
        p.scores, p.idx = scores, idx
        return p

    def result(self):
        if not self._done:
            if self._ctx.dev.type == "cuda":
                with torch.cuda.device(self._ctx.dev):
                    self.reran = self._ctx.resolve()
            else:
                self.reran = self._ctx.resolve()
            self._done = True
            self._ctx = None
        return self.scores, self.idx


class _Search:
    """State of one filtered search (every rank holds the same queries, thresholds and outputs)."""

    def __init__(self, stores, comm, k, k_eff, kk, wts, excl, eps_t, pl, a_op, q_raw, q_norm, nq, thr, out_s, out_i,
                 stats, ph):
        self.stores, self.comm, self.k, self.k_eff, self.kk, self.wts = stores, comm, k, k_eff, kk, wts
        self.excl, self.eps_t, self.pl = excl, eps_t, pl
        self.a_op, self.q_raw, self.q_norm, self.nq, self.thr = a_op, q_raw, q_norm, nq, thr
        self.out_s, self.out_i, self.stats, self.ph = out_s, out_i, stats, ph
        self.ref = stores[0]
        self.dev = self.ref.device
        self.live = [s for s in stores if s.n]
        self.n_shards = len(stores) * comm.world
        self.solo = self.n_shards == 1
        cap = pl["cap"]
        if not self.solo:
            cap = max(2048, min(cap, 1 << int(math.ceil(math.log2(4.0 * cap / self.n_shards)))))
        self.cap = cap
        self.n_shard_max = max(s.n for s in self.live) if self.live else 1
        self.pending = None

    # one pass over all rows (rows None) or over the rows that are re-run --------------------------------
    def run_pass(self, rows):
        ref, dev, comm, ph, k_eff, kk, eps_t = self.ref, self.dev, self.comm, self.ph, self.k_eff, self.kk, self.eps_t
        if rows is None:
            a_sub, q_sub, qn_sub, thr_sub, ex_sub, n_sub = self.a_op, self.q_raw, self.q_norm, self.thr, self.excl, self.nq
        else:
            n_sub = rows.numel()
            a_sub = torch.zeros((_round_up(n_sub, _BM), ref.k), dtype=torch.bfloat16, device=dev)
            a_sub[:n_sub] = self.a_op[rows]
            q_sub = self.q_raw[rows].contiguous()
            qn_sub = self.q_norm[:, rows].contiguous()
            thr_sub = self.thr[rows].contiguous()
            ex_sub = self.excl[rows].contiguous() if self.excl is not None else None
        cap = self.cap
        # 3: fused score + threshold filter on every local shard
        ph.mark("alloc")
        cands = [s._filter(a_sub, n_sub, thr_sub, cap) for s in self.live]
        ph.mark("filter")
        # 4: what needs an exact score
        m = kk if self.solo else min(kk, int(math.ceil(1.5 * kk / self.n_shards)) + 10)
        if kk <= PILOT_MAX and (self.solo or self.n_shards * m <= PILOT_UNION_MAX):
            # two rounds.  One: every shard rescores its m best approximate candidates; the k-th largest exact score
            # of the union of these pilots (kth1) is a lower bound on the true k-th best.  Two: only candidates whose
            # approximate score reaches kth1 - eps can still belong to the top-k.
            firsts = [_row_kth(c[1], c[0], m, 0.0, 0) for c in cands]
            exacts = [s._rescore(q_sub, qn_sub, n_sub, self.wts, c, b1) for s, c, b1 in zip(self.live, cands, firsts)]
            pilots = [s._pilot_top(e, c, n_sub, ex_sub, m) for s, c, e in zip(self.live, cands, exacts)]
            if not pilots:
                pilots = [torch.full((n_sub, m), float("-inf"), dtype=torch.float64, device=dev)]
            bound = _pilot_bound(_segments(pilots, comm), k_eff, eps_t)
            ph.mark("pilot")
            for s, c, b1, e in zip(self.live, cands, firsts, exacts):
                s._rescore(q_sub, qn_sub, n_sub, self.wts, c, bound, b1, e)
        else:
            # very deep lists (k > 1000): one round against (kk-th largest approximate score over all shards) - 2 eps
            if self.solo:
                bound = _row_kth(cands[0][1], cands[0][0], kk, 2.0, 0, eps_t)
            elif kk <= ROW_TOPJ_MAX:
                tops = [_row_topj(c[1], c[0], kk) for c in cands]
                if not tops:
                    tops = [torch.full((n_sub, kk), float("-inf"), dtype=torch.float32, device=dev)]
                bound = _row_kth(_union(tops, comm), None, kk, 2.0, 0, eps_t)
            else:
                # deeper than xmve_row_topj extracts: the best shard's own kk-th largest approximate score is a lower
                # bound on the kk-th largest of the union (looser, so a few more rows are rescored)
                bound = torch.full((n_sub,), float("-inf"), dtype=torch.float32, device=dev)
                for c in cands:
                    bound = torch.maximum(bound, _row_kth(c[1], c[0], kk, 2.0, 0, eps_t))
                bound = comm.max_(bound)
            ph.mark("bound")
            exacts = [s._rescore(q_sub, qn_sub, n_sub, self.wts, c, bound) for s, c in zip(self.live, cands)]
        ph.mark("rescore")
        # 5: selection.  One shard: the local kernel certifies.  Several: packed local lists, ONE gather, K3.
        n_bad = torch.zeros((1,), dtype=torch.int32, device=dev)
        if self.solo:
            s_ = torch.empty((n_sub, k_eff), dtype=torch.float64, device=dev)
            i_ = torch.empty((n_sub, k_eff), dtype=torch.int64, device=dev)
            cert, thr_next = self.live[0]._select(exacts[0], cands[0], n_sub, k_eff, ex_sub, thr_sub, eps_t, bound,
                                                  s_, i_, True, n_bad)
            over = ("lists", [cands[0][0]], cap)
        else:
            seg = packed_bytes(n_sub, k_eff)
            blocks = torch.empty((max(len(self.live), 1), seg), dtype=torch.uint8, device=dev)
            for li, (s, c, e) in enumerate(zip(self.live, cands, exacts)):
                b_s, b_i, b_f = packed_views(blocks[li], n_sub, k_eff)
                s._select(e, c, n_sub, k_eff, ex_sub, None, eps_t, None, b_s, b_i, False)
                b_f.copy_(c[0] > cap)
            if not self.live:
                b_s, b_i, b_f = packed_views(blocks[0], n_sub, k_eff)
                b_s.fill_(float("-inf"))
                b_i.fill_(-1)
                b_f.zero_()
            gathered = blocks if comm.world == 1 else comm.gather(blocks).view(-1, seg)
            s_, i_, cert, thr_next = _merge_packed(gathered, n_sub, k_eff, k_eff, thr_sub, eps_t, n_bad)
            over = ("packed", gathered, n_sub, k_eff)
        ph.mark("select_merge")
        if self.stats is not None and rows is None:
            st = self.stats
            st["eps"] = float(eps_t)
            st["cap"] = cap
            st["pilot_m"] = m
            st["cand_count"] = [c[0] for c in cands]
            ar = torch.arange(cap, device=dev)
            st["rescored_per_query"] = sum(
                float(((ar[None, :] < c[0][:, None]) & (e[:, :cap] > float("-inf"))).sum())
                for c, e in zip(cands, exacts)) / max(n_sub, 1)
            ph.mark("stats_bookkeeping")
        return s_, i_, cert, thr_next, over, n_bad, thr_sub

    def first_pass(self, in_graph=False, flag_host=None):
        self._cap0 = self.cap
        s_, i_, cert, thr_next, over, n_bad, thr_sub = self.run_pass(None)
        self.out_s[:, :self.k_eff] = s_
        self.out_i[:, :self.k_eff] = i_
        self._last = (cert, thr_next, over, thr_sub)
        if self.dev.type == "cuda":
            # (pinned memory is allocated outside a graph capture: GraphSearch passes its own buffer)
            self._flag = torch.empty((1,), dtype=torch.int32).pin_memory() if flag_host is None else flag_host
            self._flag.copy_(n_bad, non_blocking=True)
            self._event = None
            if not in_graph:                                  # a captured event cannot be waited on: see replayed()
                self._event = torch.cuda.Event()
                self._event.record()
        else:
            self._flag, self._event = n_bad, None
        self._thr0 = self.thr

    def replayed(self):
        """The captured first pass has just been replayed (GraphSearch): mark the point the host waits for."""
        self.thr = self._thr0                                 # a re-run of the previous batch may have replaced it
        self.cap = self._cap0                                 # ... or grown the candidate lists
        self._event = torch.cuda.Event()
        self._event.record()

    def resolve(self):
        """Wait for the first pass's certificate count; re-run (all ranks alike) the rows that missed it."""
        if self._event is not None:
            self._event.synchronize()
        self.ph.mark("certify_sync")
        reran = int(self._flag[0]) != 0
        if reran:
            self._rerun()
        if self.stats is not None:
            self.stats["phases_ms"] = self.ph.result()
        return reran

    @staticmethod
    def _overflowed(over):
        """Per-row overflow flags of a pass (formed only when rows are re-run)."""
        if over[0] == "lists":
            return over[1][0] > over[2]
        _, gathered, n_sub, k_eff = over
        flags = torch.stack([packed_views(gathered[g], n_sub, k_eff)[2] for g in range(gathered.shape[0])])
        return flags.max(dim=0).values != 0

    def _rerun(self):
        cert, thr_next, over, thr_sub = self._last
        rows = None
        for attempt in range(12):
            bad = torch.nonzero(cert == 0).flatten()            # device -> host sync (identical on every rank)
            if bad.numel() == 0:
                return
            over = self._overflowed(over)
            # an overflowed row needs a HIGHER threshold; if the kernel cannot propose one, grow the lists
            stuck = over[bad] & (thr_next[bad] <= thr_sub[bad])
            if rows is None:
                self.thr = self.thr.clone()
                self.thr[bad] = thr_next[bad]
                rows = bad
            else:
                self.thr[rows[bad]] = thr_next[bad]
                rows = rows[bad]
            if self.stats is not None:
                self.stats["reruns"] = self.stats.get("reruns", 0) + 1
                self.stats["rerun_rows"] = self.stats.get("rerun_rows", 0) + int(rows.numel())
            if bool(stuck.any()):
                # overflow that a tighter threshold cannot fix (dense neighbourhoods within eps of the k-th best):
                # grow the lists -- for the few rows left they may grow until they hold a whole shard
                room = max(32768, min(1 << int(math.ceil(math.log2(self.n_shard_max))),
                                      (1 << 27) // max(int(rows.numel()), 1)))
                self.cap = min(room, self.cap * 4)
            s_, i_, cert, thr_next, over, _, thr_sub = self.run_pass(rows)
            self.out_s[rows, :self.k_eff] = s_
            self.out_i[rows, :self.k_eff] = i_
        bad = torch.nonzero(cert == 0).flatten()
        if bad.numel():
            raise N.XmveError("search: %d row(s) could not be certified after 12 passes (increase eps headroom "
                              "or candidate capacity; heavy score ties?)" % int(bad.numel()))


class GraphSearch:
    """A search of fixed shape captured ONCE into a CUDA graph and replayed per query batch.

    The first pass of :func:`search_shards` never waits for the host and has static shapes for a given
    ``(number of queries, k, weights, exclude yes/no)``: about 25 kernel launches and as many small allocations, which
    cost the host ~0.4 ms -- more than the GPU needs for a handful of queries (online search, the 60-query AVS batch).
    Replaying the captured graph costs ~10 us of host time.  The rare uncertified rows are re-run eagerly by
    :meth:`PendingSearch.result`, exactly as without a graph.

        gs = GraphSearch(store, n_queries=60, k=1000)
        scores, idx = gs(queries)                 # or  p = gs(queries, defer=True); ...; p.result()

    With ``comm`` (``distributed.GroupComm``) the NCCL gathers of a sharded search are captured inside the graph and
    every rank replays in lockstep (``tools/check_sharded.py`` / ``tools/graph_sharded_check.py``: identical to the
    eager search; 0.94 -> 0.64 ms for 60 queries over two 135 k-row shards, where the ~40 launches of the eager step
    are what bounds it).  The returned tensors are the graph's static output buffers and are overwritten by the
    next call.
    """

    def __init__(self, stores, n_queries, k, weights=None, with_exclude=False, comm=None, n_total=None,
                 small_nv=SMALL_NV):
        self.stores = list(stores) if isinstance(stores, (list, tuple)) else [stores]
        ref = self.stores[0]
        self.dev = ref.device
        self.nq, self.k = int(n_queries), int(k)
        self.comm = comm or SoloComm()
        self.q_in = torch.randn((self.nq, ref.dtot), dtype=torch.float32, device=self.dev)   # warm-up queries
        self.excl_in = torch.full((self.nq,), -1, dtype=torch.int64, device=self.dev) if with_exclude else None
        kw = dict(weights=weights, exclude=self.excl_in, eps=None, small_nv=small_nv, stats=None, comm=self.comm,
                  n_total=n_total)
        with torch.cuda.device(self.dev):
            # warm-up outside the capture: opt-in shared-memory attributes, the cached corpus residual, NCCL
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    _search_shards(self.stores, self.q_in, self.k, kw["weights"], kw["exclude"], None, small_nv, None,
                                   self.comm, n_total).result()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._flag_host = torch.empty((1,), dtype=torch.int32).pin_memory()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._pending = _search_shards(self.stores, self.q_in, self.k, kw["weights"], kw["exclude"], None,
                                               small_nv, None, self.comm, n_total, in_graph=True,
                                               flag_host=self._flag_host)
        self.scores, self.idx = self._pending.scores, self._pending.idx

    def __call__(self, queries, exclude=None, defer=False):
        q = queries if torch.is_tensor(queries) else torch.as_tensor(queries)
        assert tuple(q.shape) == tuple(self.q_in.shape), "GraphSearch was captured for %s queries" % (self.q_in.shape,)
        with torch.cuda.device(self.dev):
            self.q_in.copy_(q, non_blocking=True)
            if self.excl_in is not None:
                e = torch.full((self.nq,), -1, dtype=torch.int64) if exclude is None else torch.as_tensor(exclude)
                self.excl_in.copy_(e.to(torch.int64), non_blocking=True)
            else:
                assert exclude is None, "capture with with_exclude=True to pass exclusions"
            self.graph.replay()
            ctx = self._pending._ctx_template
            ctx.replayed()
            p = PendingSearch(ctx)
            p.scores, p.idx = self.scores, self.idx
        return p if defer else p.result()


def _dv2_global(stores, comm):
    """max over ALL corpus rows (every shard of every rank) of the squared bf16 residual: device fp32 [1], cached per
    corpus state -- the one all-reduce it needs is paid when the corpus changes, not per search."""
    ref = stores[0]
    key = (tuple(s.n for s in stores), comm.world)
    cached = getattr(ref, "_dv2_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    v = torch.stack([s.resid_max2().reshape(()) for s in stores]).max().reshape(1).float()
    v = comm.max_(v)
    ref._dv2_cache = (key, v)
    return v


def search_shards(stores, queries, k, weights=None, exclude=None, eps=None, small_nv=SMALL_NV, stats=None, comm=None,
                  n_total=None, defer=False, head_stream=None):
    """Exact top-``k`` over a corpus cut into shards: ``stores`` are this process's shards (normally one), ``comm``
    joins the processes of a ``torch.distributed`` group (``distributed.GroupComm``; every rank calls this function
    with the same queries).  All shards work against ONE per-query threshold:

    1. K1 on the query batch; the measured error bound ``eps`` as a device scalar (the corpus-side maximum is
       all-reduced once per corpus state, the query side is identical on every rank).
    2. the largest scores of a ``step``-strided sample of each shard (a coarse K2 STORE pass sets a floor, a K2
       FILTER pass over the sample keeps what exceeds it); the top-J of every shard are gathered and the global
       threshold is ``max(kth(union, j) - thr_margin(k / step) * eps, kth(union, j_cap))`` -- what one GPU would compute on the whole
       corpus, so each shard appends only its share of the candidates.
    3. K2 FILTER over each shard (the score matrix never reaches HBM).
    4. exact fp64 rescore in two rounds: every shard rescores its best ``m`` approximate candidates; the pilots are
       gathered and the k-th largest exact score of their union, minus eps, bounds what the second round rescores
       (about half of what the one-round window ``approximate k-th - 2 eps`` needs).
    5. local top-k per shard into a packed block; ONE gather of the blocks; merge (K3) with the certificate
       ``kth_exact - eps >= threshold`` and no overflowed list.
    6. rows without a certificate are re-run (by all ranks alike) with the threshold the merge proposes.  Their
       number is the only thing the host reads back, asynchronously: with ``defer=True`` the function returns a
       :class:`PendingSearch` right after enqueueing steps 1-5 and ``.result()`` does step 6.

    Corpora of at most ``small_nv`` rows skip 2-4: every shard forms its fp64 score matrix directly.
    With one shard and one rank no gather happens and the certificate comes from the local selection kernel.

    ``head_stream`` (a ``torch.cuda.Stream``; use with ``defer=True``): steps 1-2 are enqueued on that stream and the
    current stream only waits for their result before step 3.  Batches are independent, so when the caller keeps one
    search in flight the K1 / sampling / threshold work of batch i+1 runs next to the rescore / selection / merge of
    batch i as soon as FILTER(i) has left the SMs, instead of behind it -- both are short, latency-bound kernel
    chains (1.2 ms + 1.9 ms of a 31 ms step on 8 GPUs).  The queries must be complete when the call is made, or have
    been produced on ``head_stream`` (it does not wait for the current stream).
    """
    comm = comm or SoloComm()
    ref = stores[0]
    if ref.device.type == "cuda":
        with torch.cuda.device(ref.device):                     # kernels, streams and allocations follow the store
            pending = _search_shards(stores, queries, k, weights, exclude, eps, small_nv, stats, comm, n_total,
                                     head_stream=head_stream)
            return pending if defer else pending.result()
    pending = _search_shards(stores, queries, k, weights, exclude, eps, small_nv, stats, comm, n_total)
    return pending if defer else pending.result()


def _head_on(stream):
    import contextlib
    return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()


def _search_shards(stores, queries, k, weights, exclude, eps, small_nv, stats, comm, n_total, in_graph=False,
                   flag_host=None, head_stream=None):
    ref = stores[0]
    dev, n_space = ref.device, len(ref.dims)
    n_shards = len(stores) * comm.world
    solo = n_shards == 1
    n_local = sum(s.n for s in stores)
    if n_total is None:
        n_total = comm.sum_int(n_local)
    if n_total == 0:
        raise ValueError("empty corpus")
    wts = _weights(weights, n_space)
    ph = _Phases(stats is not None and ref.device.type == "cuda")
    ph.mark("start")
    nq_in = (queries[0] if isinstance(queries, (list, tuple)) else queries).shape[0]
    # the head (steps 1-2) on its own stream: filtered searches only, and not while phases are being timed
    if dev.type != "cuda" or stats is not None or in_graph or n_total <= small_nv or nq_in == 0:
        head_stream = None
    if head_stream is not None:
        for t in (queries if isinstance(queries, (list, tuple)) else (queries, exclude)):
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(head_stream)                  # the caller may drop it while the head still reads it
    with _head_on(head_stream):
        a_op, q_raw, q_norm, q_res, nq = ref.prepare_queries(queries, wts)
        excl = None
        if exclude is not None:
            excl = torch.as_tensor(exclude, dtype=torch.int64).to(dev)
    ph.mark("prepare_queries")
    if nq == 0:
        return PendingSearch.ready(torch.empty((0, k), dtype=torch.float64, device=dev),
                                   torch.empty((0, k), dtype=torch.int64, device=dev))
    k_eff = min(int(k), n_total)
    out_s = torch.full((nq, k), float("-inf"), dtype=torch.float64, device=dev)
    out_i = torch.full((nq, k), -1, dtype=torch.int64, device=dev)

    if n_total <= small_nv:
        parts = [s._exact_small(q_raw, q_norm, nq, k_eff, wts, excl) for s in stores if s.n]
        if not parts:
            parts = [(torch.full((nq, k_eff), float("-inf"), dtype=torch.float64, device=dev),
                      torch.full((nq, k_eff), -1, dtype=torch.int64, device=dev))]
        if solo:
            s_, i_ = parts[0]
        else:
            s_, i_ = _merge(_union([p[0] for p in parts], comm), _union([p[1] for p in parts], comm), k_eff)
        out_s[:, :k_eff] = s_
        out_i[:, :k_eff] = i_
        return PendingSearch.ready(out_s, out_i)

    kk = k_eff + (1 if excl is not None else 0)               # one extra in case the excluded row is among them
    pl = plan(kk, n_total)
    with _head_on(head_stream):
        eps_t, thr = _threshold(stores, comm, solo, wts, eps, pl, a_op, q_res, nq, ph)
    if head_stream is not None:
        # step 3 onwards runs on the caller's stream: it waits for the head, and the caching allocator must not hand
        # the head's tensors back to the head stream while this stream still reads them
        main = torch.cuda.current_stream()
        main.wait_stream(head_stream)
        for t in (a_op, q_raw, q_norm, q_res, eps_t, thr, excl):
            if t is not None:
                t.record_stream(main)
    search = _Search(stores, comm, k, k_eff, kk, wts, excl, eps_t, pl, a_op, q_raw, q_norm, nq, thr, out_s, out_i,
                     stats, ph)
    search.first_pass(in_graph, flag_host)
    pending = PendingSearch(search)
    pending._ctx_template = search
    pending.scores, pending.idx = out_s, out_i
    return pending


def _threshold(stores, comm, solo, wts, eps, pl, a_op, q_res, nq, ph):
    """Steps 1b-2 of :func:`search_shards`: the error bound and the per-query threshold -> (eps_t, thr)."""
    ref = stores[0]
    dev, n_space = ref.device, len(ref.dims)
    w_abs = max(1.0, sum(abs(w) for w in wts))
    if eps is None:
        eps_t = _eps_device(q_res, _dv2_global(stores, comm), wts, n_space, ref.k)
    else:
        eps_t = torch.full((1,), float(eps) * w_abs, dtype=torch.float32, device=dev)
    ph.mark("eps")
    live = [s for s in stores if s.n]
    # 2: one global threshold from the shards' samples
    j_cap = pl["j_cap"] if solo else min(pl["j_cap"], ROW_TOPJ_MAX)      # xmve_row_topj holds 4096 values per row
    big_j = max(pl["j"], j_cap)
    lists, floor = [], torch.full((nq,), float("-inf"), dtype=torch.float32, device=dev)
    for s in live:
        if (s.n + pl["step"] - 1) // pl["step"] >= 16384:
            # two-level: the sample matrix (2.1 GB at 10 M rows) is never written or radix-selected
            sc, cnt, thr0 = s._sample_top(a_op, nq, pl["step"], big_j)
            lists.append((sc, cnt))
            floor = thr0 if len(lists) == 1 else torch.minimum(floor, thr0)
        else:                                                 # small shard: the few launches of the plain way win
            lists.append((s._sample(a_op, nq, pl["step"]), None))
            floor = torch.full_like(floor, float("-inf"))     # a complete list needs no floor
    if solo:
        thr = _row_kth(lists[0][0], lists[0][1], pl["j"], pl["margin"], j_cap, eps_t)
    elif big_j > ROW_TOPJ_MAX:
        # more order statistics than xmve_row_topj extracts (k in the thousands over a small corpus): the j-th largest
        # of ONE shard's sample is a lower bound on the j-th largest of the union -- a valid, lower threshold; lists
        # that overflow because of it go through the re-run path
        thr = torch.full((nq,), float("-inf"), dtype=torch.float32, device=dev)
        for sc, cnt in lists:
            thr = torch.maximum(thr, _row_kth(sc, cnt, pl["j"], pl["margin"], 0, eps_t))
        thr = comm.max_(thr)
        floor = -comm.max_(-floor)
    else:
        tops = [_row_topj(sc, cnt, big_j) for sc, cnt in lists]
        if not tops:
            tops = [torch.full((nq, big_j), float("-inf"), dtype=torch.float32, device=dev)]
        tops.append(floor.unsqueeze(1))                       # the floors ride along with the order statistics
        u = _union(tops, comm)                                # [nq, G * (L * big_j + 1)]
        per = u.shape[1] // comm.world
        floor = u.view(nq, comm.world, per)[:, :, -1].min(dim=1).values
        u.view(nq, comm.world, per)[:, :, -1] = float("-inf")
        thr = _row_kth(u, None, pl["j"], pl["margin"], j_cap, eps_t)
    # a two-level list that came out too short gives -inf (or a value below the floor): fall back to the coarse
    # floor, which ~0.4 % of the corpus exceeds -- far more than k rows, so it is below the k-th best score
    thr = torch.maximum(thr, floor - 2.0 * eps_t)
    del lists
    ph.mark("sample_threshold")
    return eps_t, thr


def rank_of_gt(stores, queries, gt_off, gt_rows, weights=None, comm=None, n_total=None, cap=8192, eps=None,
               stats=None):
    """EXACT rank of every ground-truth item when the score matrix cannot exist (C3-C5 scale), sharded.

    ``gt_off`` int64 ``[nq + 1]`` / ``gt_rows`` int64 ``[E]`` are a CSR of global corpus rows per query (host arrays
    or tensors).  Returns ``ranks`` int32 ``[E]`` on the device: ``1 + #{v : s(v) > s(g)} + #{v < g : s(v) == s(g)}``
    with ``s`` the exact fp64 fused score -- the position of ``g`` in the stable ascending argsort of the reference's
    errors row (``util/metrics.py:139-145``).  Feed them to ``metrics.RankResult.from_ranks`` for R@K / MedR / MeanR /
    mAP.

    1. the shard that owns ``g`` computes ``s(g)`` with the rescore kernel; one all-reduce(max) shares it.
    2. the tensor-core FILTER runs with one operand row per ENTRY and the window ``(s(g) - eps, s(g) + eps]``:
       rows above the window are counted (they are certainly better: ``|approx - exact| <= eps``), rows inside it
       are listed, rescored exactly and compared with ``s(g)`` (``xmve_count_before``); rows below are certainly
       worse.  Shards add their counts with ONE all-reduce(sum).
    3. entries whose window holds more than ``cap`` rows (ground truths ranked tens of thousands deep) fall back to
       an exact fp64 pass over the corpus (chunked ``xmve_score_f64`` + ``xmve_count_band_f64``).
    """
    comm = comm or SoloComm()
    stores = list(stores) if isinstance(stores, (list, tuple)) else [stores]
    ref = stores[0]
    dev, n_space = ref.device, len(ref.dims)
    wts = _weights(weights, n_space)
    if n_total is None:
        n_total = comm.sum_int(sum(s.n for s in stores))
    import contextlib
    with (torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()):
        a_op, q_raw, q_norm, q_res, nq = ref.prepare_queries(queries, wts)
        off = torch.as_tensor(gt_off, dtype=torch.int64)
        g = torch.as_tensor(gt_rows, dtype=torch.int64).to(dev)
        n_ent = int(g.numel())
        assert off.numel() == nq + 1 and int(off[-1]) == n_ent, "gt_off must be a CSR over the query rows"
        if n_ent == 0:
            return torch.zeros((0,), dtype=torch.int32, device=dev)
        owner = torch.repeat_interleave(torch.arange(nq), off[1:] - off[:-1]).to(dev)        # entry -> query row
        if eps is None:
            eps_t = _eps_device(q_res, _dv2_global(stores, comm), wts, n_space, ref.k)
        else:
            eps_t = torch.full((1,), float(eps) * max(1.0, sum(abs(w) for w in wts)), dtype=torch.float32, device=dev)
        # one operand / raw row per entry
        a_ent = torch.zeros((_round_up(n_ent, _BM), ref.k), dtype=torch.bfloat16, device=dev)
        a_ent[:n_ent] = a_op[owner]
        q_ent = q_raw[owner].contiguous()
        qn_ent = q_norm[:, owner].contiguous()
        live = [s for s in stores if s.n]
        # 1: exact score of every ground-truth item, from the shard that owns it
        s_gt = torch.full((n_ent,), float("-inf"), dtype=torch.float64, device=dev)
        for s in live:
            mine = (g >= s.index_offset) & (g < s.index_offset + s.n)
            loc = torch.where(mine, g - s.index_offset, torch.zeros_like(g)).to(torch.int32).unsqueeze(1).contiguous()
            cnt = mine.to(torch.int32)
            ex = s._rescore(q_ent, qn_ent, n_ent, wts, (cnt, torch.zeros((n_ent, 1), device=dev), loc), None)
            s_gt = torch.where(mine, ex[:, 0], s_gt)
        s_gt = comm.max_(s_gt)
        if bool((s_gt == float("-inf")).any()):
            raise IndexError("rank_of_gt: a ground-truth row is outside the corpus (or its score is not finite)")
        # 2: window around s(g), widened by the float roundings of its ends
        e64 = eps_t.double()
        lo = (s_gt - 1.001 * e64 - 1e-7).float()
        hi = (s_gt + 1.001 * e64 + 1e-7).float()
        tally = torch.zeros((2, n_ent), dtype=torch.int64, device=dev)                        # [before, overflowed]
        for s in live:
            above, count, c_s, c_i = s._filter_band(a_ent, n_ent, lo, hi, cap)
            ex = s._rescore(q_ent, qn_ent, n_ent, wts, (count, c_s, c_i), None)
            tally[0] += above
            _count_before(ex, c_i, count, s.index_offset, s_gt, g, tally[0])
            tally[1] += (count > cap)
            del ex, c_s, c_i
        tally = comm.sum_(tally)
        deep = torch.nonzero(tally[1] != 0).flatten()                                         # host sync
        if stats is not None:
            stats["eps"] = float(eps_t)
            stats["deep_entries"] = int(deep.numel())
        # 3: exact fp64 pass for the entries whose window overflowed
        if deep.numel():
            fix = _rank_deep(live, q_ent[deep].contiguous(), qn_ent[:, deep].contiguous(), s_gt[deep].contiguous(),
                             g[deep].contiguous(), wts)
            tally[0, deep] = comm.sum_(fix)
        return (tally[0] + 1).to(torch.int32)


def _rank_deep(live, q_ent, qn_ent, s_gt, g, wts, chunk=1 << 17, band_cap=64, delta=1e-13):
    """#{rows before the ground truth} over this rank's shards from an exact fp64 score matrix, chunk by chunk.
    Scores more than ``delta`` above ``s_gt`` are counted; the handful within ``delta`` (the fp64 matrix and the
    rescore kernel sum in different orders, ~1e-15 apart) are settled by the rescore kernel itself."""
    dev = q_ent.device
    n_ent = q_ent.shape[0]
    before = torch.zeros((n_ent,), dtype=torch.int64, device=dev)
    st = N.stream_ptr()
    for s in live:
        for e0 in range(0, n_ent, 4096):
            e1 = min(n_ent, e0 + 4096)
            ne = e1 - e0
            qs_n, o = [], 0
            for d in s.dims:
                qn = torch.empty((ne, d), dtype=torch.float64, device=dev)
                src = q_ent[e0:e1, o:o + d]
                N.call("xmve_normalize_f64", N.ptr(src), N.F32, ne, d, q_ent.stride(0), N.ptr(qn), d, s.norm_mode, st)
                qs_n.append(qn)
                o += d
            for r0 in range(0, s.n, chunk):
                r1 = min(s.n, r0 + chunk)
                acc, o = None, 0
                for si, d in enumerate(s.dims):
                    vn = torch.empty((r1 - r0, d), dtype=torch.float64, device=dev)
                    vs = s.raw[r0:r1, o:o + d]
                    N.call("xmve_normalize_f64", N.ptr(vs), N.F32, r1 - r0, d, s.raw.stride(0), N.ptr(vn), d,
                           s.norm_mode, st)
                    sc = torch.empty((ne, r1 - r0), dtype=torch.float64, device=dev)
                    N.call("xmve_score_f64", N.ptr(qs_n[si]), ne, d, N.ptr(vn), r1 - r0, d, d, float(wts[si]),
                           N.ptr(sc), r1 - r0, st)
                    acc = sc if acc is None else acc.add_(sc)
                    o += d
                bcount = torch.zeros((ne,), dtype=torch.int32, device=dev)
                bidx = torch.zeros((ne, band_cap), dtype=torch.int32, device=dev)
                N.call("xmve_count_band_f64", N.ptr(acc), ne, r1 - r0, r1 - r0, r0, N.ptr(s_gt[e0:]), float(delta),
                       N.ptr(before[e0:]), N.ptr(bcount), N.ptr(bidx), band_cap, st)
                if bool((bcount > band_cap).any()):
                    raise N.XmveError("rank_of_gt: more than %d corpus rows tie with a ground truth to 1e-13" % band_cap)
                bidx += r0                                                                    # shard-local rows
                ex = s._rescore(q_ent[e0:e1], qn_ent[:, e0:e1].contiguous(), ne, wts,
                                (bcount, torch.zeros((ne, band_cap), device=dev), bidx), None)
                _count_before(ex, bidx, bcount, s.index_offset, s_gt[e0:e1], g[e0:e1], before[e0:e1])
                del acc, ex
    return before


def search_norm_score(stores, queries, k, weights=None, comm=None, n_total=None, **kw):
    """Top-``k`` under ``norm_score`` fusion: every space's score matrix is
    min-max normalised over the WHOLE matrix (LINAS-engine/validate.py:7-11) before the weighted sum,
    ``fused[q, v] = sum_s w_s * (cos_s[q, v] - min_s) / (max_s - min_s)``.

    For corpora whose matrices cannot exist this is an affine map of the cosines, so the ranking equals that of a
    weighted-cosine search with ``w_s / (max_s - min_s)``; the global extremes are exact: ``max_s`` is the largest
    top-1 score of a one-hot search of space s, ``min_s`` minus the largest top-1 score of the negated queries.
    Costs ``2 S`` extra top-1 searches.  Returns ``(fused scores fp64 [nq, k], idx)``; the errors the reference
    would rank are ``-fused``."""
    stores = list(stores) if isinstance(stores, (list, tuple)) else [stores]
    ref = stores[0]
    n_space = len(ref.dims)
    wts = _weights(weights, n_space)
    q_dev = [_to_device(x, ref.device) for x in _as_spaces(queries, ref.dims)]
    q_all = q_dev[0] if n_space == 1 else torch.cat(q_dev, dim=-1)
    adj, shift = [], 0.0
    kw_ext = {key: v for key, v in kw.items() if key not in ("exclude", "stats", "defer")}   # extremes: all rows count
    kw.pop("defer", None)
    for s in range(n_space):
        onehot = [1.0 if t == s else 0.0 for t in range(n_space)]
        top, _ = search_shards(stores, q_all, 1, weights=onehot, comm=comm, n_total=n_total, **kw_ext)
        bot, _ = search_shards(stores, -q_all, 1, weights=onehot, comm=comm, n_total=n_total, **kw_ext)
        hi, lo = float(top.max()), -float(bot.max())            # global extremes of cos_s over all (q, v)
        rng = hi - lo                                            # s / np.max(s) after s -= np.min(s)
        if not rng > 0.0:
            raise ValueError("search_norm_score: space %d has a constant score matrix (max == min); norm_score "
                             "divides by zero there, as the reference's would (validate.py:10)" % s)
        adj.append(wts[s] / rng)
        shift += wts[s] * lo / rng
    scores, idx = search_shards(stores, q_all, k, weights=adj, comm=comm, n_total=n_total, **kw)
    return scores - shift, idx


def _merge(scores, idx, k, thr=None, eps=0.0, overflow=None):
    """K3 on ``[nq, m]`` (score, global index) pairs -> top-``k`` (+ certificate when ``thr`` is given)."""
    nq, m = scores.shape
    out_s = torch.empty((nq, k), dtype=torch.float64, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    cert = thr_next = None
    if thr is not None:
        cert = torch.empty((nq,), dtype=torch.int32, device=scores.device)
        thr_next = torch.empty((nq,), dtype=torch.float32, device=scores.device)
    if nq:
        N.call("xmve_select_topk_i64", N.ptr(scores), N.ptr(idx), nq, m, None, k, N.ptr(thr), float(eps), None,
               N.ptr(overflow), N.ptr(out_s), N.ptr(out_i), None, N.ptr(cert), N.ptr(thr_next), None, N.stream_ptr())
    if thr is None:
        return out_s, out_i
    return out_s, out_i, cert, thr_next


def _cumsum(xs):
    t = 0
    for x in xs:
        t += x
        yield t


def _weights(weights, n_space):
    if weights is None:
        return [1.0] * n_space if n_space == 1 else [1.0 / n_space] * n_space
    w = [float(x) for x in weights]
    assert len(w) == n_space, "one fusion weight per embedding space"
    return w


def merge_topk(scores, idx, k):
    """K3: ``scores/idx [G, nq, kk]`` gathered from G shards -> global top-``k`` (same ordering rule)."""
    g, nq, kk = scores.shape
    s = scores.permute(1, 0, 2).reshape(nq, g * kk).contiguous()
    i = idx.permute(1, 0, 2).reshape(nq, g * kk).contiguous()
    return _merge(s, i, k)
