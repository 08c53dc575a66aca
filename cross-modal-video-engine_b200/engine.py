"""Corpus-resident store and the search pipeline (host orchestration of the sm_100a kernels).

Replaces, for a corpus that is encoded once and queried many times, the per-call
``cal_error(video_embs, cap_emb)`` + ``np.argsort(errors[0])[:topK]`` of
``LINAS-engine/inference.py:78-80`` (which re-normalises the whole corpus on every call,
``evaluation.py:19-20``) and the 32-query ``1 - P @ index.T`` / ``torch.argsort`` blocks of
``MultiFusion/src/validate.py:65-113``.

The pipeline is :func:`search_shards` (its docstring lists the steps): K1 on the queries, a sampled per-query
threshold, the fused tcgen05 score + threshold filter over the resident operand (the score matrix never reaches
HBM), an exact fp64 rescore of the survivors, a bitonic top-k with a certificate that nothing outside the
candidate set can belong to the top-k, and a re-run of the rare rows that miss the certificate.
:meth:`CorpusStore.search` is the one-shard case; ``distributed.sharded_search`` joins the shards of several GPUs.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _native as N

#: a-priori bound on |bf16-operand score - exact cosine| per unit of fusion weight: bf16 keeps 8 significant bits,
#: so rounding moves a vector by at most 2^-8 of its norm and a product of two unit vectors by at most 2 * 2^-8
#: (Cauchy-Schwarz); plus fp32 accumulation slack.  Used only when no measured residuals are available
#: (search(eps=...) callers, NaN rows); the search itself uses :func:`measured_eps` (typically 3.5e-3 - 4e-3).
EPS_X1 = 8.5e-3
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
    def search(self, queries, k, weights=None, exclude=None, eps=None, small_nv=SMALL_NV, stats=None):
        """Top-``k`` corpus rows per query by fused cosine score, exact (fp64) scores, descending.

        Returns ``(scores float64 [nq, k], idx int64 [nq, k])`` on the device; ``idx`` are global row
        numbers (``index_offset`` + local), ``-1`` / ``-inf`` padded if the corpus has fewer than ``k``
        rows.  ``exclude[q]`` (global row or -1) is dropped from row q's list -- MultiFusion's removal
        of the query's own reference item (validate.py:76-83).  ``eps`` overrides the measured bound on
        |tensor-core score - exact score| (see :func:`measured_eps`).
        """
        return search_shards([self], queries, k, weights=weights, exclude=exclude, eps=eps, small_nv=small_nv,
                             stats=stats)

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
               None, 0.0, None, N.ptr(s_), N.ptr(i_), None, None, None, st)
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

    def _rescore_select(self, q_raw, q_norm, nq, k, wts, excl, cand, bound, thr, eps, certify):
        """Exact fp64 rescore of the candidates above ``bound`` + local top-k (global row ids)."""
        dev, st = self.device, N.stream_ptr()
        cand_count, cand_score, cand_idx = cand
        cap = cand_score.shape[1]
        exact = torch.empty((nq, cap), dtype=torch.float64, device=dev)
        N.call("xmve_rescore", N.ptr(q_raw), nq, q_raw.stride(0), N.ptr(q_norm), N.ptr(self.raw), self.n,
               self.raw.stride(0), N.ptr(self.norm), len(self.dims), self.space_off, self._weights_arr(wts),
               self.norm_mode, N.ptr(cand_score), N.ptr(cand_idx), N.ptr(cand_count), cap, N.ptr(bound),
               N.ptr(exact), st)
        out_s = torch.empty((nq, k), dtype=torch.float64, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        cert = thr_next = None
        if certify:
            cert = torch.empty((nq,), dtype=torch.int32, device=dev)
            thr_next = torch.empty((nq,), dtype=torch.float32, device=dev)
        N.call("xmve_select_topk_i32", N.ptr(exact), N.ptr(cand_idx), nq, cap, N.ptr(cand_count),
               self.index_offset, N.ptr(excl), k, N.ptr(thr) if certify else None, float(eps),
               N.ptr(bound) if certify else None, N.ptr(out_s), N.ptr(out_i), None, N.ptr(cert), N.ptr(thr_next), st)
        return out_s, out_i, cert, thr_next


def plan(k, n):
    """Sampling step, order statistics and candidate capacity for a top-``k`` search of ``n`` corpus rows."""
    n = int(n)
    n_s = min(n, max(8192, min(65536, n // 128)))
    step = max(1, n // max(n_s, 1))
    n_s = (n + step - 1) // step
    lam = k / step
    j = int(math.ceil(lam + 5.5 * math.sqrt(lam) + 4))
    cap = 1 << max(11, int(math.ceil(math.log2(8 * step * j))))
    cap = min(cap, 32768)
    j_cap = max(j, int(0.5 * cap / step))
    return {"step": step, "n_sample": n_s, "j": min(j, n_s), "j_cap": min(j_cap, n_s), "cap": cap}


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


def _row_topj(vals, counts, j):
    rows, cols = vals.shape
    out = torch.empty((rows, j), dtype=torch.float32, device=vals.device)
    if rows:
        N.call("xmve_row_topj", N.ptr(vals), rows, cols, vals.stride(0), N.ptr(counts), j, N.ptr(out), N.stream_ptr())
    return out


def _row_kth(vals, counts, j1, sub, j2):
    rows, cols = vals.shape
    out = torch.empty((rows,), dtype=torch.float32, device=vals.device)
    if rows:
        N.call("xmve_row_kth", N.ptr(vals), rows, cols, vals.stride(0), N.ptr(counts), j1, float(sub), j2, N.ptr(out),
               N.stream_ptr())
    return out


def _union(parts, comm):
    """Per-row lists from the local shards ``[nq, m]`` -> the union over all shards of all ranks ``[nq, G*L*m]``."""
    loc = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
    if comm.world == 1:
        return loc
    g = comm.gather(loc.contiguous())                                   # [G, nq, L*m]
    return g.permute(1, 0, 2).reshape(loc.shape[0], -1).contiguous()


def search_shards(stores, queries, k, weights=None, exclude=None, eps=None, small_nv=SMALL_NV, stats=None, comm=None,
                  n_total=None):
    """Exact top-``k`` over a corpus cut into shards: ``stores`` are this process's shards (normally one), ``comm``
    joins the processes of a ``torch.distributed`` group (``distributed.GroupComm``; every rank calls this function
    with the same queries).  All shards work against ONE per-query threshold:

    1. K1 on the query batch; the measured error bound ``eps`` (max over ranks).
    2. the largest scores of a ``step``-strided sample of each shard (a coarse K2 STORE pass sets a floor, a K2
       FILTER pass over the sample keeps what exceeds it); the top-J of every shard are gathered and the global
       threshold is ``max(kth(union, j) - 2 eps, kth(union, j_cap))`` -- what one GPU would compute on the whole
       corpus, so each shard appends only its share of the candidates.
    3. K2 FILTER over each shard (the score matrix never reaches HBM).
    4. the kk-th largest approximate candidate score over all shards - 2 eps bounds what needs an exact score.
    5. exact fp64 rescore + local top-k per shard; ONE gather of the ``[nq, k]`` lists; merge (K3) with the
       certificate ``kth_exact - eps >= threshold`` and no overflowed list.
    6. rows without a certificate are re-run (by all ranks alike) with the threshold the merge proposes.

    Corpora of at most ``small_nv`` rows skip 2-4: every shard forms its fp64 score matrix directly.
    With one shard and one rank no gather happens and the certificate comes from the local selection kernel.
    """
    comm = comm or SoloComm()
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
    a_op, q_raw, q_norm, q_res, nq = ref.prepare_queries(queries, wts)
    ph.mark("prepare_queries")
    if nq == 0:
        return (torch.empty((0, k), dtype=torch.float64, device=dev), torch.empty((0, k), dtype=torch.int64, device=dev))
    excl = None
    if exclude is not None:
        excl = torch.as_tensor(exclude, dtype=torch.int64).to(dev)
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
        return out_s, out_i

    if eps is None:
        m = torch.stack([q_res.sum(0).max(), torch.stack([s.resid_max2() for s in stores]).max()]).double()
        dq2, dv2 = comm.max_(m).tolist()                       # host sync (tiny); the bound must be a host float
        eps = measured_eps(math.sqrt(dq2), math.sqrt(dv2), wts, n_space, ref.k) if dq2 == dq2 and dv2 == dv2 \
            else EPS_X1 * max(1.0, sum(abs(w) for w in wts))
    else:
        eps = float(eps) * max(1.0, sum(abs(w) for w in wts))
    ph.mark("eps_sync")
    kk = k_eff + (1 if excl is not None else 0)               # one extra in case the excluded row is among them
    pl = plan(kk, n_total)
    live = [s for s in stores if s.n]
    # 2: one global threshold from the shards' samples
    big_j = max(pl["j"], pl["j_cap"])
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
        thr = _row_kth(lists[0][0], lists[0][1], pl["j"], 2.0 * eps, pl["j_cap"])
    else:
        tops = [_row_topj(sc, cnt, big_j) for sc, cnt in lists]
        if not tops:
            tops = [torch.full((nq, big_j), float("-inf"), dtype=torch.float32, device=dev)]
        thr = _row_kth(_union(tops, comm), None, pl["j"], 2.0 * eps, pl["j_cap"])
        floor = -comm.max_(-floor)                            # min over the ranks
    # a two-level list that came out too short gives -inf (or a value below the floor): fall back to the coarse
    # floor, which ~0.4 % of the corpus exceeds -- far more than k rows, so it is below the k-th best score
    thr = torch.maximum(thr, floor - 2.0 * eps)
    del lists
    ph.mark("sample_threshold")
    cap = pl["cap"] if solo else max(2048, min(pl["cap"], 1 << int(math.ceil(math.log2(4.0 * pl["cap"] / n_shards)))))
    n_shard_max = max(s.n for s in live) if live else 1
    rows = None                                               # None = all rows; else LongTensor of rows to re-run
    for attempt in range(12):
        if rows is None:
            a_sub, q_sub, qn_sub, thr_sub, ex_sub, n_sub = a_op, q_raw, q_norm, thr, excl, nq
        else:
            n_sub = rows.numel()
            a_sub = torch.zeros((_round_up(n_sub, _BM), ref.k), dtype=torch.bfloat16, device=dev)
            a_sub[:n_sub] = a_op[rows]
            q_sub = q_raw[rows].contiguous()
            qn_sub = q_norm[:, rows].contiguous()
            thr_sub = thr[rows].contiguous()
            ex_sub = excl[rows].contiguous() if excl is not None else None
        # 3: fused score + threshold filter on every local shard
        ph.mark("alloc")
        cands = [s._filter(a_sub, n_sub, thr_sub, cap) for s in live]
        ph.mark("filter")
        # 4: nothing below (kk-th largest approximate candidate score) - 2 eps can reach the top-k
        if solo:
            bound = _row_kth(cands[0][1], cands[0][0], kk, 2.0 * eps, 0)
        else:
            tops = [_row_topj(c[1], c[0], kk) for c in cands]
            if not tops:
                tops = [torch.full((n_sub, kk), float("-inf"), dtype=torch.float32, device=dev)]
            bound = _row_kth(_union(tops, comm), None, kk, 2.0 * eps, 0)
        ph.mark("bound")
        # 5: exact rescore + selection
        if solo:
            s_, i_, cert, thr_next = live[0]._rescore_select(q_sub, qn_sub, n_sub, k_eff, wts, ex_sub, cands[0], bound,
                                                             thr_sub, eps, True)
            over = cands[0][0] > cap
        else:
            parts = [s._rescore_select(q_sub, qn_sub, n_sub, k_eff, wts, ex_sub, c, bound, thr_sub, eps, False)
                     for s, c in zip(live, cands)]
            over = torch.zeros((n_sub,), dtype=torch.int32, device=dev)
            for c in cands:
                over |= (c[0] > cap).int()
            over = comm.max_(over)
            if not parts:
                parts = [(torch.full((n_sub, k_eff), float("-inf"), dtype=torch.float64, device=dev),
                          torch.full((n_sub, k_eff), -1, dtype=torch.int64, device=dev), None, None)]
            s_, i_, cert, thr_next = _merge(_union([p[0] for p in parts], comm), _union([p[1] for p in parts], comm),
                                            k_eff, thr=thr_sub, eps=eps, overflow=over)
            over = over != 0
        ph.mark("rescore_select_merge")
        if stats is not None and rows is None:
            stats["eps"] = eps
            stats["cap"] = cap
            stats["cand_count"] = [c[0] for c in cands]
            ar = torch.arange(cap, device=dev)
            stats["rescored_per_query"] = sum(
                float(((ar[None, :] < c[0][:, None]) & (c[1] >= bound[:, None])).sum()) for c in cands) / n_sub
            ph.mark("stats_bookkeeping")
        bad = torch.nonzero(cert == 0).flatten()              # device -> host sync (identical on every rank)
        ph.mark("certify_sync")
        if rows is None:
            out_s[:, :k_eff] = s_
            out_i[:, :k_eff] = i_
        else:
            out_s[rows, :k_eff] = s_
            out_i[rows, :k_eff] = i_
        if bad.numel() == 0:
            break
        # an overflowed row needs a HIGHER threshold; if the kernel cannot propose one, grow the lists
        stuck = over[bad] & (thr_next[bad] <= thr_sub[bad])
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
        if bool(stuck.any()):
            # overflow that a tighter threshold cannot fix (dense neighbourhoods within eps of the k-th best): grow
            # the lists -- for the few rows left they may grow until they hold a whole shard (cannot overflow then)
            room = max(32768, min(1 << int(math.ceil(math.log2(n_shard_max))), (1 << 27) // max(int(rows.numel()), 1)))
            cap = min(room, cap * 4)
    else:
        raise N.XmveError("search: %d row(s) could not be certified after 12 passes (increase eps headroom "
                          "or candidate capacity; heavy score ties?)" % int(rows.numel()))
    if stats is not None:
        stats["phases_ms"] = ph.result()
    return out_s, out_i


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
    for s in range(n_space):
        onehot = [1.0 if t == s else 0.0 for t in range(n_space)]
        top, _ = search_shards(stores, q_all, 1, weights=onehot, comm=comm, n_total=n_total, **kw)
        bot, _ = search_shards(stores, -q_all, 1, weights=onehot, comm=comm, n_total=n_total, **kw)
        hi, lo = float(top.max()), -float(bot.max())            # global extremes of cos_s over all (q, v)
        rng = hi - lo                                            # s / np.max(s) after s -= np.min(s)
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
        N.call("xmve_select_topk_i64", N.ptr(scores), N.ptr(idx), nq, m, None, k, N.ptr(thr), float(eps),
               N.ptr(overflow), N.ptr(out_s), N.ptr(out_i), None, N.ptr(cert), N.ptr(thr_next), N.stream_ptr())
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
