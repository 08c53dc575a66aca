"""Corpus-resident store and the search pipeline (host orchestration of the sm_100a kernels).

Replaces, for a corpus that is encoded once and queried many times, the per-call
``cal_error(video_embs, cap_emb)`` + ``np.argsort(errors[0])[:topK]`` of
``LINAS-engine/inference.py:78-80`` (which re-normalises the whole corpus on every call,
``evaluation.py:19-20``) and the 32-query ``1 - P @ index.T`` / ``torch.argsort`` blocks of
``MultiFusion/src/validate.py:65-113``.

Pipeline of :meth:`CorpusStore.search` (all on the caller's CUDA stream, no host sync until the
certification flags are read):

1. K1  queries -> bf16 operand (per-space weight / norm folded in), fp32 raw copy, fp64 norms
2. K2  STORE pass over every ``step``-th corpus row (strided TMA view)      -> score sample
3.     radix select of the sample                                              -> per-row threshold
4. K2  FILTER pass over the whole shard; the epilogue appends (score, index) above the threshold
5.     radix select of the candidates' approximate scores                     -> rescore bound
6.     exact fp64 rescore of the survivors from the raw fp32 rows
7.     bitonic top-k of the exact scores + certification that nothing outside the candidate set
       can belong to the top-k (given |bf16 score - exact| <= eps)
8.     rows that are not certified are re-run with the threshold the kernel proposes.

Corpora of at most ``small_nv`` rows skip 2-5: the fp64 score matrix is formed directly.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _native as N

#: worst-case |bf16-operand score - exact cosine| for unit vectors: 2 * 2^-9 relative per product
#: (Cauchy-Schwarz) plus fp32 accumulation slack.
EPS_X1 = 4.5e-3
SMALL_NV = 16384
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


def _prepare(src, d, raw, raw_off, norm, op, op_off, layout, weight, norm_mode):
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
           N.ptr(norm),
           N.ptr(op), op.stride(0) if op is not None else 0, op_off, layout,
           float(weight), norm_mode, N.stream_ptr())


class CorpusStore:
    """Video-embedding corpus resident in HBM: bf16 operand rows + raw fp32 rows + fp64 norms.

    ``dims`` lists the embedding spaces (e.g. ``(1536, 512)`` for latent + concept); rows are appended
    with :meth:`add` (the analogue of ``encode_vid``'s batch loop, evaluation.py:98-105) and
    normalised once, on the device, as they arrive.  ``index_offset`` is this shard's first global
    row (multi-GPU sharding).
    """

    def __init__(self, capacity, dims, device="cuda", norm_mode="plain", index_offset=0):
        N.require_device()
        self.device = torch.device(device)
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
        self.op = torch.zeros((rows, self.k), dtype=torch.bfloat16, device=self.device)
        self.raw = torch.empty((max(self.capacity, 1), self.raw_ld), dtype=torch.float32, device=self.device)
        self.norm = torch.empty((len(self.dims), max(self.capacity, 1)), dtype=torch.float64, device=self.device)
        self.space_off = (C.c_int32 * (len(self.dims) + 1))(*([0] + list(_cumsum(self.dims))))

    # -- ingest ----------------------------------------------------------------------------------
    def add(self, rows):
        """Append rows: tensor / ndarray ``[n, sum(dims)]`` (or ``[n, T, D]``: mean over T first), or a per-space list."""
        spaces = [_to_device(s, self.device) for s in _as_spaces(rows, self.dims)]
        n = spaces[0].shape[0]
        if self.n + n > self.capacity:
            raise ValueError("CorpusStore capacity %d exceeded" % self.capacity)
        op_off = raw_off = 0
        for s, (src, d, dp) in enumerate(zip(spaces, self.dims, self.dpads)):
            assert src.shape[-1] == d, "space %d: expected dim %d, got %d" % (s, d, src.shape[-1])
            _prepare(src, d, self.raw[self.n:], raw_off, self.norm[s, self.n:], self.op[self.n:], op_off,
                     N.OP_X1, 1.0, self.norm_mode)
            op_off += dp
            raw_off += d
        self.n += n
        return self

    # -- query side ------------------------------------------------------------------------------
    def prepare_queries(self, queries, weights=None):
        """K1 on the query batch: returns (a_op bf16 [nq_pad, K], q_raw fp32 [nq, raw_ld], q_norm fp64 [S, nq])."""
        weights = _weights(weights, len(self.dims))
        spaces = [_to_device(s, self.device) for s in _as_spaces(queries, self.dims)]
        nq = spaces[0].shape[0]
        a_op = torch.zeros((_round_up(max(nq, 1), _BM), self.k), dtype=torch.bfloat16, device=self.device)
        q_raw = torch.empty((max(nq, 1), self.raw_ld), dtype=torch.float32, device=self.device)
        q_norm = torch.empty((len(self.dims), max(nq, 1)), dtype=torch.float64, device=self.device)
        op_off = raw_off = 0
        for s, (src, d, dp) in enumerate(zip(spaces, self.dims, self.dpads)):
            _prepare(src, d, q_raw, raw_off, q_norm[s], a_op, op_off, N.OP_X1, weights[s], self.norm_mode)
            op_off += dp
            raw_off += d
        return a_op, q_raw, q_norm, nq

    # -- search ----------------------------------------------------------------------------------
    def search(self, queries, k, weights=None, exclude=None, eps=EPS_X1, small_nv=SMALL_NV, stats=None):
        """Top-``k`` corpus rows per query by fused cosine score, exact (fp64) scores, descending.

        Returns ``(scores float64 [nq, k], idx int64 [nq, k])`` on the device; ``idx`` are global row
        numbers (``index_offset`` + local), ``-1`` / ``-inf`` padded if the shard has fewer than ``k``
        rows.  ``exclude[q]`` (global row or -1) is dropped from row q's list -- MultiFusion's removal
        of the query's own reference item (validate.py:76-83).
        """
        if self.n == 0:
            raise ValueError("empty corpus")
        wts = _weights(weights, len(self.dims))
        eps = float(eps) * max(1.0, sum(abs(w) for w in wts))      # the error bound scales with sum |w_s|
        a_op, q_raw, q_norm, nq = self.prepare_queries(queries, wts)
        if nq == 0:
            return (torch.empty((0, k), dtype=torch.float64, device=self.device),
                    torch.empty((0, k), dtype=torch.int64, device=self.device))
        excl = None
        if exclude is not None:
            excl = torch.as_tensor(exclude, dtype=torch.int64).to(self.device)
        k_eff = min(int(k), self.n)
        out_s = torch.full((nq, k), float("-inf"), dtype=torch.float64, device=self.device)
        out_i = torch.full((nq, k), -1, dtype=torch.int64, device=self.device)
        if self.n <= small_nv:
            s_, i_ = self._search_exact_small(q_raw, q_norm, nq, k_eff, wts, excl)
        else:
            s_, i_ = self._search_filtered(a_op, q_raw, q_norm, nq, k_eff, wts, excl, eps, stats)
        out_s[:, :k_eff] = s_
        out_i[:, :k_eff] = i_
        return out_s, out_i

    def _weights_arr(self, wts):
        return (C.c_double * len(wts))(*[float(w) for w in wts])

    def _search_exact_small(self, q_raw, q_norm, nq, k, wts, excl):
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
        out_s = torch.empty((nq, k), dtype=torch.float64, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        N.call("xmve_select_topk_i32", N.ptr(acc), None, nq, self.n, None, self.index_offset, N.ptr(excl), k,
               None, 0.0, None, N.ptr(out_s), N.ptr(out_i), None, None, None, st)
        return out_s, out_i

    def plan(self, k, n=None):
        """Sampling step, order statistics and candidate capacity for a top-``k`` search of ``n`` rows."""
        n = self.n if n is None else int(n)
        n_s = min(n, max(8192, min(65536, n // 128)))
        step = max(1, n // max(n_s, 1))
        n_s = (n + step - 1) // step
        lam = k / step
        j = int(math.ceil(lam + 5.5 * math.sqrt(lam) + 4))
        cap = 1 << max(11, int(math.ceil(math.log2(8 * step * j))))
        cap = min(cap, 32768)
        j_cap = max(j, int(0.5 * cap / step))
        return {"step": step, "n_sample": n_s, "j": min(j, n_s), "j_cap": min(j_cap, n_s), "cap": cap}

    def _search_filtered(self, a_op, q_raw, q_norm, nq, k, wts, excl, eps, stats):
        dev, st = self.device, N.stream_ptr()
        kk = k + (1 if excl is not None else 0)           # one extra in case the excluded row is among them
        pl = self.plan(kk)
        # 2-3: sampled threshold  thr = max(kth(sample, j) - 2 eps, kth(sample, j_cap))
        sample = torch.empty((nq, pl["n_sample"]), dtype=torch.float32, device=dev)
        N.call("xmve_score_store", N.ptr(a_op), nq, a_op.stride(0), N.ptr(self.op), pl["n_sample"],
               self.op.stride(0), pl["step"], self.k, 1.0, N.ptr(sample), sample.stride(0), st)
        thr = torch.empty((nq,), dtype=torch.float32, device=dev)
        N.call("xmve_row_kth", N.ptr(sample), nq, pl["n_sample"], sample.stride(0), None, pl["j"], 2.0 * eps,
               pl["j_cap"], N.ptr(thr), st)
        del sample
        out_s = torch.empty((nq, k), dtype=torch.float64, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        rows = None                                        # None = all rows; else LongTensor of rows to re-run
        cap = pl["cap"]
        for attempt in range(12):
            if rows is None:
                a_sub, q_sub, qn_sub, thr_sub, ex_sub, n_sub = a_op, q_raw, q_norm, thr, excl, nq
            else:
                n_sub = rows.numel()
                a_sub = torch.zeros((_round_up(n_sub, _BM), self.k), dtype=torch.bfloat16, device=dev)
                a_sub[:n_sub] = a_op[rows]
                q_sub = q_raw[rows].contiguous()
                qn_sub = q_norm[:, rows].contiguous()
                thr_sub = thr[rows].contiguous()
                ex_sub = excl[rows].contiguous() if excl is not None else None
            s_, i_, cert, thr_next, cnt = self._filter_pass(a_sub, q_sub, qn_sub, n_sub, k, kk, wts, ex_sub, thr_sub,
                                                            eps, cap)
            bad = torch.nonzero(cert == 0).flatten()       # device -> host sync (the step's only one)
            if rows is None:
                out_s, out_i = s_, i_
            else:
                out_s[rows] = s_
                out_i[rows] = i_
            if bad.numel() == 0:
                break
            # an overflowed row needs a HIGHER threshold; if the kernel cannot propose one, grow the lists
            stuck = (cnt[bad] > cap) & (thr_next[bad] <= thr_sub[bad])
            if rows is None:
                thr = thr.clone()
                thr[bad] = thr_next[bad]
                rows = bad
            else:
                thr[rows[bad]] = thr_next[bad]
                rows = rows[bad]
            if stats is not None:
                stats["reruns"] = stats.get("reruns", 0) + 1
                stats["rerun_rows"] = stats.get("rerun_rows", 0) + int(rows.numel())
            if bool(stuck.any()) and cap < 32768:
                cap = min(32768, cap * 4)                  # overflow that a tighter threshold cannot fix
        else:
            raise N.XmveError("search: %d row(s) could not be certified after 12 passes (increase eps headroom "
                              "or candidate capacity; heavy score ties?)" % int(rows.numel()))
        return out_s, out_i

    def _filter_pass(self, a_op, q_raw, q_norm, nq, k, kk, wts, excl, thr, eps, cap):
        dev, st = self.device, N.stream_ptr()
        cand_count = torch.zeros((nq,), dtype=torch.int32, device=dev)
        cand_score = torch.empty((nq, cap), dtype=torch.float32, device=dev)
        cand_idx = torch.empty((nq, cap), dtype=torch.int32, device=dev)
        # 4: fused score + threshold filter; the score matrix never reaches HBM
        N.call("xmve_score_filter", N.ptr(a_op), nq, a_op.stride(0), N.ptr(self.op), self.n, self.op.stride(0),
               self.k, N.ptr(thr), None, None, N.ptr(cand_count), N.ptr(cand_score), N.ptr(cand_idx), cap, st)
        # 5: bound = (kk-th largest approximate candidate score) - 2 eps: nothing below it can reach the top-k
        bound = torch.empty((nq,), dtype=torch.float32, device=dev)
        N.call("xmve_row_kth", N.ptr(cand_score), nq, cap, cap, N.ptr(cand_count), kk, 2.0 * eps, 0, N.ptr(bound), st)
        # 6: exact fp64 rescore of the survivors
        exact = torch.empty((nq, cap), dtype=torch.float64, device=dev)
        N.call("xmve_rescore", N.ptr(q_raw), nq, q_raw.stride(0), N.ptr(q_norm), N.ptr(self.raw), self.n,
               self.raw.stride(0), N.ptr(self.norm), len(self.dims), self.space_off, self._weights_arr(wts),
               self.norm_mode, N.ptr(cand_score), N.ptr(cand_idx), N.ptr(cand_count), cap, N.ptr(bound),
               N.ptr(exact), st)
        # 7: final top-k + certification
        out_s = torch.empty((nq, k), dtype=torch.float64, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        cert = torch.empty((nq,), dtype=torch.int32, device=dev)
        thr_next = torch.empty((nq,), dtype=torch.float32, device=dev)
        N.call("xmve_select_topk_i32", N.ptr(exact), N.ptr(cand_idx), nq, cap, N.ptr(cand_count),
               self.index_offset, N.ptr(excl), k, N.ptr(thr), float(eps), N.ptr(bound), N.ptr(out_s), N.ptr(out_i),
               None, N.ptr(cert), N.ptr(thr_next), st)
        self.last_cand_count = cand_count
        return out_s, out_i, cert, thr_next, cand_count


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
    out_s = torch.empty((nq, k), dtype=torch.float64, device=s.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=s.device)
    if nq:
        N.call("xmve_select_topk_i64", N.ptr(s), N.ptr(i), nq, g * kk, None, k, N.ptr(out_s), N.ptr(out_i), None,
               N.stream_ptr())
    return out_s, out_i
