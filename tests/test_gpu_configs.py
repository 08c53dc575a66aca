"""The five BASELINE.json configurations at their FULL sizes (B200 only).

C1 is covered by the golden fixture ``c1_1k`` (tests/test_gpu_parity.py).  C2 runs the oracle in full (it finishes
in well under a minute).  C3-C5 are too large for the CPU oracle, so they are checked through size-independent
properties and against a plain torch fp64 statement of the same scores computed on the device in chunks (a
library matmul used ONLY as the checker): identical top-k index lists, scores within 1e-12, planted items found,
AP@1000 identical to the oracle's scorer on the exact ranks.  The checker never reads what the product wrote
(``store.raw`` / ``store.norm``): it REGENERATES every corpus chunk from the seed, re-applies the planted rows,
pools the frames and takes the norms itself -- a row-placement or norm bug of K1 (e.g. beyond 2^31 bytes) would
show up as a mismatch.
"""
import time

import numpy as np
import pytest
import torch

from oracle import linas

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def X():
    from cross_modal_video_engine_b200 import (_native, avs, distributed, engine, evaluation, metrics, multifusion,
                                               synth, validate)
    _native.require_device()

    class NS:
        pass
    ns = NS()
    ns.__dict__.update(dict(engine=engine, evaluation=evaluation, metrics=metrics, validate=validate, avs=avs,
                            multifusion=multifusion, synth=synth, distributed=distributed))
    return ns


def _dedupe_plant(rows, vecs):
    """Keep the FIRST vector planted at each row (an index_put with duplicate rows is not deterministic)."""
    pos = torch.arange(rows.numel(), device=rows.device)
    uniq, inv = torch.unique(rows, return_inverse=True)
    first = torch.full((uniq.numel(),), rows.numel(), dtype=torch.int64, device=rows.device)
    first.scatter_reduce_(0, inv, pos, reduce="amin")
    return uniq, vecs[first]


def _chunks(X, nv, d, seed, chunk=250000, plant=None, frames=1):
    """The synthetic corpus chunk by chunk, generated on the device from the seed: yields ``(lo, rows)`` with
    ``rows`` fp32 ``[n, d]`` or ``[n, frames, d]``; ``plant = (rows, vectors)`` overwrites some rows."""
    for c, lo in enumerate(range(0, nv, chunk)):
        n = min(chunk, nv - lo)
        buf = X.synth.device_gaussian(n * frames, d, seed * 100003 + c, "cuda")
        if frames > 1:
            buf = buf.view(n, frames, d)
        if plant is not None:
            rows, vecs = plant
            m = (rows >= lo) & (rows < lo + n)
            if bool(m.any()):
                if frames > 1:
                    buf[rows[m] - lo] = vecs[m].unsqueeze(1).expand(-1, frames, -1).contiguous()
                else:
                    buf[rows[m] - lo] = vecs[m]
        yield lo, buf


def _fp64_topk(X, q_raw, nv, dims, weights, k, seed, chunk=250000, plant=None, frames=1, eps_norm=False,
               exclude=None):
    """Exact top-k of sum_s w_s * cos_s in torch fp64 on the device, from corpus rows REGENERATED from the seed
    (independent of the store): frames pooled as fp32 ``(f0 + f1 + ...) / T`` (Combiner.time_process on fp32
    features), norms in fp64 (``eps_norm``: ``max(norm, 1e-12)`` as F.normalize), order (score desc, row asc)."""
    nq = q_raw.shape[0]
    dev = q_raw.device
    offs = np.cumsum((0,) + tuple(dims))
    qn = []
    for s, (a, b) in enumerate(zip(offs[:-1], offs[1:])):
        q = q_raw[:, a:b].double()
        nrm = torch.linalg.vector_norm(q, dim=1, keepdim=True)
        qn.append(weights[s] * q / (nrm.clamp(min=1e-12) if eps_norm else nrm))
    best_s = torch.full((nq, 0), 0.0, dtype=torch.float64, device=dev)
    best_i = torch.full((nq, 0), 0, dtype=torch.int64, device=dev)
    for lo, buf in _chunks(X, nv, sum(dims), seed, chunk, plant, frames):
        if frames > 1:
            acc = buf[:, 0].clone()
            for f in range(1, frames):
                acc += buf[:, f]
            buf = acc / float(frames)
        hi = lo + buf.shape[0]
        sc = torch.zeros((nq, hi - lo), dtype=torch.float64, device=dev)
        for s, (a, b) in enumerate(zip(offs[:-1], offs[1:])):
            v = buf[:, a:b].double()
            nrm = torch.linalg.vector_norm(v, dim=1, keepdim=True)
            v /= nrm.clamp(min=1e-12) if eps_norm else nrm
            sc += qn[s] @ v.T
            del v
        if exclude is not None:
            hit = (exclude >= lo) & (exclude < hi)
            sc[torch.nonzero(hit).flatten(), (exclude[hit] - lo)] = float("-inf")
        ids = torch.arange(lo, hi, device=dev).expand(nq, -1)
        cs, ci = torch.cat([best_s, sc], 1), torch.cat([best_i, ids], 1)
        # (score desc, index asc): candidates are kept index-ascending, the sort by score is stable
        order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = torch.gather(cs, 1, order), torch.gather(ci, 1, order)
        order = torch.argsort(best_i, dim=1, stable=True)
        best_s, best_i = torch.gather(best_s, 1, order), torch.gather(best_i, 1, order)
        del sc, cs, ci, buf
    order = torch.argsort(best_s, dim=1, descending=True, stable=True)
    return torch.gather(best_s, 1, order), torch.gather(best_i, 1, order)


def _device_store(X, nv, dims, seed, chunk=250000, plant=None, frames=1, norm_mode="plain"):
    """Corpus generated on the device chunk by chunk (:func:`_chunks`) and appended to a resident store."""
    import gc
    gc.collect()
    torch.cuda.empty_cache()                     # the stores of earlier tests are gone; hand their blocks back
    store = X.engine.CorpusStore(nv, dims, norm_mode=norm_mode)
    for _, buf in _chunks(X, nv, sum(dims), seed, chunk, plant, frames):
        store.add(buf)
    return store


# ---- C2: MSR-VTT full test shape, 59 800 captions x 2 990 videos -------------------------------------------
def test_c2_msrvtt_full_shape_cal_error_and_cal_perf(X):
    V, Q, vid, cap, _ = X.synth.msrvtt_like(1, 2990, 20, 1536, 14.0)
    V64, Q64 = V.astype(np.float64), Q.astype(np.float64)
    t0 = time.time()
    e_ref = linas.cal_error(V64, Q64)
    gt_ref = linas.get_gt(vid, cap)
    perf_ref = linas.cal_perf(e_ref, *gt_ref)
    t_ref = time.time() - t0
    torch.cuda.synchronize()
    t0 = time.time()
    e_gpu = X.evaluation.cal_error(torch.from_numpy(V64), torch.from_numpy(Q64))       # stays on the device
    v2t_gt, t2v_gt = X.metrics.get_gt(vid, cap)
    perf_gpu = X.validate.cal_perf(e_gpu, v2t_gt, t2v_gt)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    assert (v2t_gt, t2v_gt) == gt_ref
    assert perf_gpu == perf_ref                                   # R@1/5/10, MedR, MeanR, mAP both directions: ==
    assert 5.0 < perf_ref[1][0] < 95.0                            # non-degenerate (t2v R@1)
    assert float(np.abs(e_gpu.cpu().numpy() - e_ref).max()) <= 1e-13
    # float32 inputs: tensor-core path, north_star tolerance 1e-3 relative (floor 1e-2 on |s|)
    e32 = X.evaluation.cal_error(torch.from_numpy(V), torch.from_numpy(Q)).cpu().numpy()
    assert e32.dtype == np.float32
    assert float((np.abs(e32 - e_ref) / np.maximum(np.abs(e_ref), 1e-2)).max()) < 1e-3
    # the metrics of the reference matrix through the GPU kernels are identical as well
    assert X.validate.cal_perf(e_ref, v2t_gt, t2v_gt) == perf_ref
    print("C2 59800x2990: oracle (CPU) %.1f s, cal_error+get_gt+cal_perf on B200 %.2f s" % (t_ref, t_gpu))


def test_c2_multi_space_fused_search(X):
    """C2 with two embedding spaces (1536 + 512, w = 0.6 / 0.4): top-10 of every caption == fp64 oracle."""
    dims, w = (1536, 512), (0.6, 0.4)
    V, Q, _, _, _ = X.synth.msrvtt_like(2, 2990, 20, sum(dims), 14.0)
    store = X.engine.CorpusStore(len(V), dims).add(torch.from_numpy(V))
    s, i = store.search(torch.from_numpy(Q), 10, weights=w)
    V64, Q64 = V.astype(np.float64), Q.astype(np.float64)
    err = linas.fused_errors([V64[:, :1536], V64[:, 1536:]], [Q64[:, :1536], Q64[:, 1536:]], w)
    ref = np.argsort(err, axis=1, kind="stable")[:, :10]
    np.testing.assert_array_equal(i.cpu().numpy(), ref)
    np.testing.assert_allclose(s.cpu().numpy(), -np.take_along_axis(err, ref, axis=1), rtol=0, atol=1e-12)


# ---- C3: TRECVID AVS V3C1 shape, 1.08 M shots x 60 queries, top-1000 + AP@1000 --------------------------------
def test_c3_avs_shape_top1000_and_map(X):
    nv, nq, d, k = 1_080_000, 60, 2048, 1000
    Q = X.synth.device_gaussian(nq, d, 52, "cuda")
    g = torch.Generator(device="cuda").manual_seed(53)
    rows = torch.randperm(nv, device="cuda", generator=g)[: nq * 40]
    vecs = Q.repeat_interleave(40, 0) * 2.0 + 4.0 * X.synth.device_gaussian(nq * 40, d, 54, "cuda")
    store = _device_store(X, nv, (d,), 51, plant=(rows, vecs))
    t0 = time.time()
    s, i = X.avs.search_avs(store, Q, k)
    torch.cuda.synchronize()
    t_search = time.time() - t0
    ref_s, ref_i = _fp64_topk(X, Q, nv, (d,), (1.0,), k, 51, plant=(rows, vecs))
    assert torch.equal(i, ref_i)
    torch.testing.assert_close(s, ref_s, rtol=0, atol=1e-12)
    # relevant set: the 40 planted shots + 500 random ones per query; AP@1000 == the oracle's scorer on the ranks
    rel_rand = torch.randint(0, nv, (nq, 500), device="cuda", generator=g)
    relevant = [sorted(set(rows[q * 40:(q + 1) * 40].tolist() + rel_rand[q].tolist())) for q in range(nq)]
    ap, m = X.avs.ap_at_k(i, relevant, nv, k)
    i_host = ref_i.cpu().numpy()
    for q in range(nq):
        pos = {int(v): r + 1 for r, v in enumerate(i_host[q])}
        ranks = [pos.get(v, nv + 1) for v in relevant[q]]
        assert ap[q] == linas.ap_from_ranks(ranks, nr_relevant=len(relevant[q]), k=k, list_len=nv)
    assert m == np.mean(ap) and m > 0.02                            # the planted shots are retrieved
    print("C3 60 x 1.08M x 2048, top-1000: search %.1f ms" % (t_search * 1e3))


# ---- C4: MultiFusion composed retrieval, 4 096 queries x 1 M index items x 640, top-100 --------------------------
def test_c4_multifusion_shape(X):
    nv, nq, d, frames = 1_000_000, 4096, 640, 8
    g = torch.Generator(device="cuda").manual_seed(61)
    target = torch.randperm(nv, device="cuda", generator=g)[:nq]
    reference = (target + 1 + torch.randint(0, nv - 1, (nq,), device="cuda", generator=g)) % nv
    P = torch.nn.functional.normalize(X.synth.device_gaussian(nq, d, 62, "cuda"), dim=-1)
    tvec = P * 1.2 + 0.15 * X.synth.device_gaussian(nq, d, 63, "cuda")      # near the query after mean-pooling
    rvec = P * 3.0                                                        # the reference item scores highest
    plant = _dedupe_plant(torch.cat([target, reference]), torch.cat([tvec, rvec]))   # targets win over references
    store = _device_store(X, nv, (d,), 60, chunk=125000, frames=frames, norm_mode="eps", plant=plant)
    names = torch.randperm(10 * nv, device="cuda", generator=g)[:nv].cpu().numpy().astype(np.int64)
    t0 = time.time()
    metrics, top_names = X.multifusion.cirr_metrics_from_features(P, None, names, names[reference.cpu().numpy()],
                                                                names[target.cpu().numpy()], store=store)
    torch.cuda.synchronize()
    t_all = time.time() - t0
    sub = slice(0, 256)                                                    # checker on a slice of the queries
    ref_s, ref_i = _fp64_topk(X, P[sub], nv, (d,), (1.0,), 100, 60, chunk=125000, frames=frames, eps_norm=True,
                              plant=plant, exclude=reference[sub])
    np.testing.assert_array_equal(top_names[sub], names[ref_i.cpu().numpy()])
    assert metrics[:3] == (-1, -1, -1)
    labels = torch.from_numpy(top_names[:, :50]) == torch.from_numpy(names[target.cpu().numpy()])[:, None]
    for kk, got in zip((1, 5, 10, 50), metrics[3:]):
        assert got == float(np.float32(labels[:, :kk].sum().item()) / np.float32(nq)) * 100
    assert metrics[3] > 50.0                                               # planted targets rank first once the reference is dropped
    assert not (torch.from_numpy(top_names) == torch.from_numpy(names[reference.cpu().numpy()])[:, None]).any()
    print("C4 4096 x 1M x 640 top-100: %.1f ms (search + names + recalls)" % (t_all * 1e3))


# ---- C5: 10 M-video corpus x 8 192 queries, 2048-d in two spaces, top-100 ------------------------------------------
def test_c5_scale_sweep_full_shape(X):
    nv, nq, dims, w, k = 10_000_000, 8192, (1536, 512), (0.6, 0.4), 100
    d = sum(dims)
    Q = X.synth.device_gaussian(nq, d, 5, "cuda")
    g = torch.Generator(device="cuda").manual_seed(71)
    plant_rows = torch.randperm(nv, device="cuda", generator=g)[:nq]
    plant_vecs = Q * 1.5 + 0.2 * X.synth.device_gaussian(nq, d, 72, "cuda")
    store = _device_store(X, nv, dims, 4, plant=(plant_rows, plant_vecs))
    stats = {}
    t0 = time.time()
    s, i = store.search(Q, k, weights=w, stats=stats)
    torch.cuda.synchronize()
    t_search = time.time() - t0
    assert torch.equal(i[:, 0], plant_rows)                               # every planted near-duplicate is rank 1
    assert torch.all(s[:, :-1] >= s[:, 1:]) and torch.all(i >= 0)
    assert torch.all((s[:, :-1] > s[:, 1:]) | (i[:, :-1] < i[:, 1:]))     # ties (if any) by ascending index
    sub = torch.arange(0, nq, 64, device="cuda")                          # 128 queries against the fp64 checker
    ref_s, ref_i = _fp64_topk(X, Q[sub], nv, dims, w, k, 4, plant=(plant_rows, plant_vecs))
    assert torch.equal(i[sub], ref_i)
    torch.testing.assert_close(s[sub], ref_s, rtol=0, atol=1e-12)
    print("C5 8192 x 10M x 2048 top-100: %.1f ms, eps %.2e, %d rerun rows" % (t_search * 1e3, stats["eps"],
                                                                              stats.get("rerun_rows", 0)))
