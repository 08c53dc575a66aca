"""Parity of the drop-in API (B200 only) against the oracle and the committed golden fixtures.

Tolerances, as BASELINE.json's north_star states them: scores within 1e-3 relative of the reference's
fp32 path (we hold 1e-12 absolute against its fp64 path and ~1e-6 absolute on fp32 inputs); top-k index
lists identical except where adjacent score gaps fall below that tolerance (here: identical, exact ties
aside); R@K / MedR / MeanR / mAP identical (``==`` on the floats).
"""
import numpy as np
import pytest
import torch

from conftest import case_inputs, load_golden
from oracle import linas, multifusion as mf_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def X():
    import cross_modal_video_engine_b200 as pkg
    from cross_modal_video_engine_b200 import (_native, basic_metric, distributed, engine, evaluation, metrics,
                                               multifusion, synth, validate)
    _native.require_device()

    class NS:
        pass
    ns = NS()
    ns.__dict__.update(dict(native=_native, engine=engine, evaluation=evaluation, metrics=metrics, validate=validate,
                            multifusion=multifusion, synth=synth, basic_metric=basic_metric, distributed=distributed))
    return ns


# ---- evaluation.py -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny_ragged", "small_cpv20", "c1_1k"])
def test_cal_error_f64_matches_reference_golden(X, manifest, name):
    V, Q, vid_ids, cap_ids, _ = case_inputs(manifest, name)
    g = load_golden(name)
    errors = X.evaluation.cal_error(V.astype(np.float64), Q.astype(np.float64))
    assert errors.dtype == np.float64 and errors.shape == (len(Q), len(V))
    np.testing.assert_allclose(errors[:8, :8], g["errors_head_f64"], rtol=0, atol=1e-14)
    assert abs(errors.sum() - g["errors_sum_f64"]) < 1e-9
    # ranks and metrics from OUR matrix equal the reference's from ITS matrix
    v2t_gt, t2v_gt = X.metrics.get_gt(vid_ids, cap_ids)
    perf = np.array(X.validate.cal_perf(errors, v2t_gt, t2v_gt), dtype=np.float64)
    np.testing.assert_array_equal(perf, g["perf_f64"])
    top10 = np.argsort(errors[:64], axis=1, kind="stable")[:, :10]
    np.testing.assert_array_equal(top10, g["top10_f64"])


@pytest.mark.parametrize("name", ["tiny_ragged", "c1_1k"])
def test_cal_error_f32_within_tolerance(X, manifest, name):
    V, Q, _, _, _ = case_inputs(manifest, name)
    g = load_golden(name)
    errors = X.evaluation.cal_error(V, Q)                      # fp32 in -> tcgen05 split-bf16 path, fp32 out
    assert errors.dtype == np.float32
    ref = linas.cal_error(V.astype(np.float64), Q.astype(np.float64))
    assert np.abs(errors - ref).max() < 2e-5                    # split-bf16 bound: ~3 * 2^-18 * sum |q_i v_i|
    big = np.abs(ref) > 1e-2                                    # relative 1e-3 is ill-posed for scores near 0
    assert (np.abs(errors - ref)[big] / np.abs(ref)[big]).max() < 1e-3
    np.testing.assert_allclose(errors[:8, :8], g["errors_head_f32"], rtol=0, atol=2e-5)
    simi = X.evaluation.cal_simi(Q, V)                          # swapped argument order, + sign
    np.testing.assert_allclose(simi, -errors, rtol=0, atol=0)
    np.testing.assert_allclose(X.evaluation.cal_error_batch(V, Q), errors, rtol=0, atol=0)


def test_l2norm_and_norm_score(X, manifest):
    V, Q, _, _, _ = case_inputs(manifest, "tiny_ragged")
    g = load_golden("tiny_ragged")
    np.testing.assert_allclose(X.evaluation.l2norm(V.astype(np.float64)), g["l2norm_f64"], rtol=0, atol=3e-16)
    np.testing.assert_allclose(X.evaluation.l2norm(V), g["l2norm_f32"], rtol=0, atol=1e-7)
    for tag in ("f32", "f64"):
        out = X.validate.norm_score(g["errors_" + tag])
        assert out.dtype == g["norm_score_" + tag].dtype
        np.testing.assert_array_equal(out, g["norm_score_" + tag])          # same IEEE operations, same order
    with pytest.raises(ValueError):
        X.evaluation.cal_error(V, Q, "chebyshev")


@pytest.mark.parametrize("mode", ["weighted-cosine", "norm_score"])
def test_fused_errors_matrix(X, mode):
    """sum_s w_s * cal_error_s and sum_s w_s * norm_score(cal_error_s) against the oracle's
    composition of the reference functions; the weighted sum itself is bit-exact given the same per-space matrices."""
    dims, w = (96, 40), (0.6, 0.4)
    V, Q = X.synth.clustered(81, 300, sum(dims)), X.synth.clustered(82, 170, sum(dims))
    for cast, tol in ((np.float64, 2e-13), (np.float32, 3e-6)):
        Vs = [V[:, :96].astype(cast), V[:, 96:].astype(cast)]
        Qs = [Q[:, :96].astype(cast), Q[:, 96:].astype(cast)]
        got = X.evaluation.fused_errors(Vs, Qs, w, mode)
        ref = linas.fused_errors([v.astype(np.float64) for v in Vs], [q.astype(np.float64) for q in Qs], w, mode)
        assert got.dtype == cast and got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=0, atol=tol)
    # bit-exactness of the fusion arithmetic: feed the oracle's per-space matrices through the kernel
    import ctypes  # noqa: F401
    e = [linas.cal_error(V[:, :96].astype(np.float64), Q[:, :96].astype(np.float64)),
         linas.cal_error(V[:, 96:].astype(np.float64), Q[:, 96:].astype(np.float64))]
    acc = torch.empty(e[0].shape, dtype=torch.float64, device="cuda")
    for s_i, (m, ws) in enumerate(zip(e, w)):
        t = torch.from_numpy(m).cuda()
        X.native.call("xmve_fuse_accumulate", X.native.ptr(acc), acc.stride(0), X.native.ptr(t), t.stride(0), X.native.F64,
                      t.shape[0], t.shape[1], float(ws), 1 if s_i == 0 else 0, X.native.stream_ptr())
    np.testing.assert_array_equal(acc.cpu().numpy(), w[0] * e[0] + w[1] * e[1])
    # float64 inputs take the score kernel's FUSED epilogue (no stored per-space matrix): bit-identical to composing
    # the stored matrices of the same kernel
    if mode == "weighted-cosine":
        Vs = [V[:, :96].astype(np.float64), V[:, 96:].astype(np.float64)]
        Qs = [Q[:, :96].astype(np.float64), Q[:, 96:].astype(np.float64)]
        stored = [X.evaluation.cal_error(v, q) for v, q in zip(Vs, Qs)]
        np.testing.assert_array_equal(X.evaluation.fused_errors(Vs, Qs, w), w[0] * stored[0] + w[1] * stored[1])
    with pytest.raises(ValueError):
        X.evaluation.fused_errors([V], [Q], [1.0], "max")


def test_non_cosine_measures_match_reference_golden(X):
    """evaluation.py:22-35 (scipy cdist: float64 out) and loss.jaccard_sim (torch float32) on the reference's outputs."""
    g = load_golden("measures")
    V, Q = g["V"], g["Q"]
    for m in ("euclidean", "l1", "l2", "l1_norm", "l2_norm"):
        e = X.evaluation.cal_error(V, Q, m)
        assert e.dtype == np.float64 and e.shape == g["err_" + m].shape
        np.testing.assert_allclose(e, g["err_" + m], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(X.evaluation.cal_error(V.astype(np.float32), Q.astype(np.float32), m), g["err_" + m],
                                   rtol=1e-13, atol=1e-13)              # cdist widens float32 inputs to double
    e = X.evaluation.cal_error(V, Q, "jaccard")
    assert e.dtype == np.float32
    np.testing.assert_allclose(e, g["err_jaccard"], rtol=2e-6, atol=0)  # the reference sums in float32
    np.testing.assert_allclose(X.evaluation.cal_error_batch(V, Q, "jaccard", batch_size=20), g["batch_jaccard"],
                               rtol=2e-6, atol=0)
    np.testing.assert_allclose(X.evaluation.cal_simi(Q, V, "jaccard"), g["simi_jaccard"], rtol=2e-6, atol=0)
    t = X.evaluation.cal_error(torch.from_numpy(V), torch.from_numpy(Q), "l1")
    assert t.is_cuda and t.dtype == torch.float64                       # tensors in, device tensor out


# ---- metrics.py / validate.py: bit-exact given the same matrix ---------------------------------------
@pytest.mark.parametrize("name", ["tiny_ragged", "small_cpv20", "c1_1k"])
@pytest.mark.parametrize("tag,cast", [("f32", np.float32), ("f64", np.float64)])
def test_rank_metrics_bit_exact_on_reference_matrix(X, manifest, name, tag, cast):
    V, Q, vid_ids, cap_ids, _ = case_inputs(manifest, name)
    g = load_golden(name)
    errors = linas.cal_error(V.astype(cast), Q.astype(cast))                 # the reference's own matrix
    v2t_gt, t2v_gt = X.metrics.get_gt(vid_ids, cap_ids)
    assert (v2t_gt, t2v_gt) == linas.get_gt(vid_ids, cap_ids)
    r = X.metrics.RankResult(errors, t2v_gt)
    np.testing.assert_array_equal(r.best.cpu().numpy(), g["t2v_ranks_" + tag])
    r = X.metrics.RankResult(errors.T, v2t_gt)
    np.testing.assert_array_equal(r.best.cpu().numpy(), g["v2t_ranks_" + tag])
    perf = X.validate.cal_perf(errors, v2t_gt, t2v_gt)
    np.testing.assert_array_equal(np.array(perf, dtype=np.float64), g["perf_" + tag])
    assert X.metrics.eval_q2m(errors, t2v_gt) == linas.eval_q2m(errors, t2v_gt)
    assert X.metrics.eval_q2m(errors.T, v2t_gt) == linas.eval_q2m(errors.T, v2t_gt)
    assert X.metrics.t2v_map(errors, t2v_gt) == linas.t2v_map(errors, t2v_gt)
    assert X.metrics.v2t_map(errors, v2t_gt) == linas.v2t_map(errors, v2t_gt)


def test_rank_ties_are_stable_order(X):
    e = np.array([[0.5, 0.5, 0.1, 0.5], [0.2, 0.2, 0.2, 0.2]], dtype=np.float64)
    r = X.metrics.RankResult(e, [[0, 1, 3], [2]])
    assert r.ranks.cpu().numpy()[:4].tolist() == [2, 3, 4, 3]
    assert r.best.cpu().numpy().tolist() == [2, 3]


def test_legacy_metrics_and_apscorer(X, manifest):
    import json, os
    from conftest import GOLDEN
    rec = manifest["legacy"]
    V, Q, _, _, _ = X.synth.msrvtt_like(rec["seed"], 40, 5, 48, 2.0)
    errors = linas.cal_error(V.astype(np.float64), Q.astype(np.float64))
    assert list(X.metrics.t2v(errors, n_caption=5)) == rec["t2v"]
    assert list(X.metrics.v2t(errors, n_caption=5)) == rec["v2t"]
    assert float(X.metrics.t2v_inv_rank(errors, 5)) == rec["t2v_inv_rank"]
    assert float(X.metrics.v2t_inv_rank(errors, 5)) == rec["v2t_inv_rank"]
    assert [float(x) for x in X.metrics.v2t_inv_rank_multi(errors, 5)] == rec["v2t_inv_rank_multi"]
    with open(os.path.join(GOLDEN, "apscorer.json")) as f:
        for c in json.load(f):
            assert X.basic_metric.getScorer(c["scorer"]).score(c["labels"]) == c["score"], c


# ---- search: filter pass + exact rescore vs the fp64 oracle ------------------------------------------
def _oracle_topk(V, Q, k, weights=None, dims=None, exclude=None):
    V64, Q64 = V.astype(np.float64), Q.astype(np.float64)
    if dims is None:
        err = linas.cal_error(V64, Q64)
    else:
        offs = np.cumsum((0,) + tuple(dims))
        err = linas.fused_errors([V64[:, a:b] for a, b in zip(offs[:-1], offs[1:])],
                                 [Q64[:, a:b] for a, b in zip(offs[:-1], offs[1:])], weights)
    if exclude is not None:
        err = err.copy()
        err[np.arange(len(Q)), exclude] = np.inf
    idx = np.argsort(err, axis=1, kind="stable")[:, :k]
    return idx, -np.take_along_axis(err, idx, axis=1)


@pytest.mark.parametrize("gen,nv,nq,d,k", [("gaussian", 200000, 300, 256, 100), ("clustered", 120000, 200, 640, 100),
                                            ("gaussian", 60000, 60, 2048, 1000), ("gaussian", 3000, 500, 1536, 10)])
def test_search_topk_identical_to_fp64_oracle(X, gen, nv, nq, d, k):
    make = getattr(X.synth, gen)
    V, Q = make(100, nv, d), make(101, nq, d)
    store = X.engine.CorpusStore(nv, (d,))
    for lo in range(0, nv, 70000):                              # ragged batches, like encode_vid's loop
        store.add(torch.from_numpy(V[lo:lo + 70000]))
    stats = {}
    scores, idx = store.search(torch.from_numpy(Q), k, stats=stats)
    ref_idx, ref_s = _oracle_topk(V, Q, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx)
    np.testing.assert_allclose(scores.cpu().numpy(), ref_s, rtol=0, atol=1e-13)


def test_search_multi_space_fusion_and_exclude(X):
    dims, w = (1536, 512), (0.6, 0.4)
    nv, nq, k = 150000, 256, 100
    V, Q = X.synth.clustered(7, nv, sum(dims), n_centroid=200), X.synth.clustered(8, nq, sum(dims), n_centroid=200)
    excl = np.random.default_rng(0).integers(0, nv, size=nq)
    store = X.engine.CorpusStore(nv, dims).add(torch.from_numpy(V))
    scores, idx = store.search(torch.from_numpy(Q), k, weights=w, exclude=excl)
    ref_idx, ref_s = _oracle_topk(V, Q, k, weights=w, dims=dims, exclude=excl)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx)
    np.testing.assert_allclose(scores.cpu().numpy(), ref_s, rtol=0, atol=1e-13)


def test_search_recovers_when_threshold_is_wrong(X):
    """Adversarial order: each query's 30 near-duplicates sit exactly on the strided sampling grid, so the
    sampled threshold lands among them and fewer than k rows pass the first filter pass.  Certification must
    reject those rows and the re-run (lower threshold) must still return the exact top-k."""
    nv, nq, d, k = 100000, 64, 128, 50
    rng = np.random.default_rng(3)
    V = rng.standard_normal((nv, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    step = X.engine.CorpusStore(nv, (d,)).plan(k, n=nv)["step"]
    grid = rng.permutation(np.arange(0, nv, step))[: 30 * nq].reshape(nq, 30)
    for qi in range(nq):
        V[grid[qi]] = Q[qi] + 0.3 * rng.standard_normal((30, d)).astype(np.float32)
    store = X.engine.CorpusStore(nv, (d,)).add(torch.from_numpy(V))
    stats = {}
    scores, idx = store.search(torch.from_numpy(Q), k, stats=stats)
    ref_idx, ref_s = _oracle_topk(V, Q, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx)
    np.testing.assert_allclose(scores.cpu().numpy(), ref_s, rtol=0, atol=1e-13)
    assert stats.get("reruns", 0) >= 1


def test_eval_q2m_from_topk_lists(X):
    """R@1/5/10 and MedR from top-k lists == the reference's eval_q2m on the full matrix (planted captions)."""
    V, Q, vid, cap, _ = X.synth.msrvtt_like(5, 30000, 1, 128, 4.5)
    Q, cap = Q[:700], cap[:700]
    store = X.engine.CorpusStore(len(V), (128,)).add(torch.from_numpy(V))
    _, idx = store.search(torch.from_numpy(Q), 50)
    _, t2v_gt = linas.get_gt(vid, cap)
    gts = [t2v_gt[i] for i in range(len(Q))]
    r1, r5, r10, medr, meanr, n_found = X.metrics.eval_q2m_topk(idx, gts, len(V))
    ref = linas.eval_q2m(linas.cal_error(V.astype(np.float64), Q.astype(np.float64)), t2v_gt)
    assert (r1, r5, r10) == tuple(ref[:3]) and 5.0 < r1 < 95.0
    if ref[3] <= 50:
        assert medr == ref[3]
    else:
        assert medr == np.inf
    assert n_found < len(Q) and np.isnan(meanr)                 # some captions rank beyond the lists


def test_search_under_norm_score_fusion(X):
    """norm_score fusion at a size that takes the filter path: same top-k as the oracle's
    sum_s w_s * norm_score(cal_error_s) matrix, fused scores within 1e-12."""
    dims, w = (128, 64), (0.6, 0.4)
    nv, nq, k = 40000, 50, 20
    V, Q = X.synth.clustered(91, nv, sum(dims), n_centroid=300), X.synth.clustered(92, nq, sum(dims), n_centroid=300)
    store = X.engine.CorpusStore(nv, dims).add(torch.from_numpy(V))
    s, i = X.engine.search_norm_score(store, torch.from_numpy(Q), k, weights=w)
    V64, Q64 = V.astype(np.float64), Q.astype(np.float64)
    err = linas.fused_errors([V64[:, :128], V64[:, 128:]], [Q64[:, :128], Q64[:, 128:]], w, mode="norm_score")
    ref = np.argsort(err, axis=1, kind="stable")[:, :k]
    np.testing.assert_array_equal(i.cpu().numpy(), ref)
    np.testing.assert_allclose(s.cpu().numpy(), -np.take_along_axis(err, ref, axis=1), rtol=0, atol=1e-12)


def test_search_exact_ties_are_ordered_by_index(X):
    """Every corpus vector appears three times: the top-k is full of exact score ties, also across the k boundary.
    Engine order = (score desc, index asc) = what a stable argsort of the reference's error row gives."""
    nu, d, nq, k = 40000, 192, 80, 50
    U, Q = X.synth.gaussian(41, nu, d), X.synth.gaussian(42, nq, d)
    perm = np.random.default_rng(43).permutation(3 * nu)
    V = np.concatenate([U, U, U])[perm]
    store = X.engine.CorpusStore(len(V), (d,)).add(torch.from_numpy(V))
    s, i = store.search(torch.from_numpy(Q), k)
    ref_idx, ref_s = _oracle_topk(V, Q, k)
    np.testing.assert_array_equal(i.cpu().numpy(), ref_idx)
    np.testing.assert_allclose(s.cpu().numpy(), ref_s, rtol=0, atol=1e-13)
    assert (ref_s[:, 0] == ref_s[:, 2]).all()                   # the ties are really there


def test_search_survives_candidate_overflow(X):
    """Dense neighbourhoods: 30 000 rows sit within a few degrees of each of four query directions, far more than
    the candidate lists hold and invisible to the strided sample's order statistic for the other queries.  The
    overflowed rows must be re-run with a higher threshold (or longer lists) and still come back exact."""
    nv, nq, d, k = 200000, 64, 128, 100
    rng = np.random.default_rng(9)
    V = rng.standard_normal((nv, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    for qi in range(4):
        rows = rng.choice(nv, size=30000, replace=False)
        V[rows] = Q[qi] * 2.0 + 0.4 * rng.standard_normal((30000, d)).astype(np.float32)
    store = X.engine.CorpusStore(nv, (d,)).add(torch.from_numpy(V))
    stats = {}
    s, i = store.search(torch.from_numpy(Q), k, stats=stats)
    ref_idx, ref_s = _oracle_topk(V, Q, k)
    np.testing.assert_array_equal(i.cpu().numpy(), ref_idx)
    np.testing.assert_allclose(s.cpu().numpy(), ref_s, rtol=0, atol=1e-13)


def test_search_edge_cases(X):
    V, Q = X.synth.gaussian(1, 37, 70), X.synth.gaussian(2, 5, 70)
    store = X.engine.CorpusStore(64, (70,)).add(V[:20]).add(V[20:])
    s, i = store.search(Q, 50)                                  # k > corpus size: padded with -1 / -inf
    ref_idx, ref_s = _oracle_topk(V, Q, 37)
    np.testing.assert_array_equal(i.cpu().numpy()[:, :37], ref_idx)
    assert (i.cpu().numpy()[:, 37:] == -1).all() and np.isinf(s.cpu().numpy()[:, 37:]).all()
    s0, i0 = store.search(Q[:0], 5)                             # empty query batch
    assert s0.shape == (0, 5) and i0.shape == (0, 5)
    with pytest.raises(ValueError):
        X.engine.CorpusStore(4, (70,)).search(Q, 1)             # empty corpus
    with pytest.raises(ValueError):
        X.engine.CorpusStore(4, (70,)).add(V)                   # over capacity


# ---- MultiFusion composed retrieval ---------------------------------------------------------------
@pytest.mark.parametrize("n_index,n_query", [(2000, 100), (60000, 257)])
def test_multifusion_metrics_and_ranked_names(X, n_index, n_query):
    index, q, names, ref, tgt = X.synth.composed_retrieval(5, n_index, n_query)
    (m_ref, top_ref) = mf_oracle.compute_cirr_val_metrics(torch.from_numpy(q), torch.from_numpy(index), names, ref, tgt)
    (m_gpu, top_gpu) = X.multifusion.cirr_metrics_from_features(torch.from_numpy(q), torch.from_numpy(index), names, ref, tgt)
    assert m_gpu == m_ref
    assert m_ref[3] > 5.0                                       # the planted targets are actually retrieved
    # the oracle ranks in torch fp32, the engine in fp64: lists are identical except where two adjacent fp32
    # scores are closer than fp32 resolution (north_star: "identical except where adjacent score gaps fall
    # below tolerance")
    diff = top_gpu != top_ref
    assert diff.mean() < 2e-3
    if diff.any():
        pooled = torch.from_numpy(index).mean(dim=1)
        pooled = torch.nn.functional.normalize(pooled, dim=-1).double().numpy()
        name_to_row = {int(n): r for r, n in enumerate(names)}
        for r, c in zip(*np.nonzero(diff)):
            a, b = name_to_row[int(top_gpu[r, c])], name_to_row[int(top_ref[r, c])]
            assert abs(float(q[r].astype(np.float64) @ (pooled[a] - pooled[b]))) < 1e-6
    assert X.multifusion.top1_name(torch.from_numpy(q[0]), torch.from_numpy(index).mean(dim=1), list(names)) == \
        mf_oracle.top1_name(torch.from_numpy(q[:1]), torch.from_numpy(index).mean(dim=1), list(names))


# ---- size-independent properties at larger sizes ----------------------------------------------------
def test_sharded_search_equals_single_store(X):
    """Split the corpus into 3 'shards' on one GPU, search each, merge with K3: identical to one store."""
    nv, nq, d, k = 250000, 128, 512, 100
    V, Q = X.synth.gaussian(11, nv, d), X.synth.gaussian(12, nq, d)
    full = X.engine.CorpusStore(nv, (d,)).add(torch.from_numpy(V))
    s_full, i_full = full.search(torch.from_numpy(Q), k)
    parts_s, parts_i = [], []
    for r in range(3):
        lo, hi = X.distributed.shard_range(nv, 3, r)
        st = X.engine.CorpusStore(hi - lo, (d,), index_offset=lo).add(torch.from_numpy(V[lo:hi]))
        s, i = st.search(torch.from_numpy(Q), k)
        parts_s.append(s)
        parts_i.append(i)
    m_s, m_i = X.engine.merge_topk(torch.stack(parts_s), torch.stack(parts_i), k)
    assert torch.equal(m_i, i_full) and torch.equal(m_s, s_full)


def test_search_shards_one_global_threshold(X):
    """The shard-aware pipeline (3 shards of one process; the multi-GPU path with the gathers done in place):
    identical to one store, and every shard appends only about its share of the candidates."""
    nv, nq, k = 300000, 96, 100
    dims, w = (256, 64), (0.7, 0.3)
    V, Q = X.synth.gaussian(31, nv, sum(dims)), X.synth.gaussian(32, nq, sum(dims))
    full = X.engine.CorpusStore(nv, dims).add(torch.from_numpy(V))
    st_full = {}
    s_full, i_full = full.search(torch.from_numpy(Q), k, weights=w, stats=st_full)
    shards = []
    for r in range(3):
        lo, hi = X.distributed.shard_range(nv, 3, r)
        shards.append(X.engine.CorpusStore(hi - lo, dims, index_offset=lo).add(torch.from_numpy(V[lo:hi])))
    excl = np.where(np.arange(nq) % 3 == 0, i_full[:, 0].cpu().numpy(), -1)
    st_sh = {}
    s_sh, i_sh = X.engine.search_shards(shards, torch.from_numpy(Q), k, weights=w, stats=st_sh)
    assert torch.equal(i_sh, i_full) and torch.equal(s_sh, s_full)
    total_full = float(st_full["cand_count"][0].float().mean())
    total_sh = sum(float(c.float().mean()) for c in st_sh["cand_count"])
    assert total_sh < 1.6 * total_full                     # not 3x: the threshold is global, not per shard
    assert 5e-4 < st_sh["eps"] < 0.6 * X.engine.EPS_X1      # measured bound, well under the a-priori worst case
    # exclusion + an empty shard in the list
    shards.append(X.engine.CorpusStore(8, dims, index_offset=nv))
    s_e, i_e = X.engine.search_shards(shards, torch.from_numpy(Q), k, weights=w, exclude=excl)
    s_f, i_f = full.search(torch.from_numpy(Q), k, weights=w, exclude=excl)
    assert torch.equal(i_e, i_f) and torch.equal(s_e, s_f)
    # small-corpus branch (fp64 matrices per shard + merge): 3 shards of a 9 000-row corpus
    small = [X.engine.CorpusStore(3000, dims, index_offset=3000 * r).add(torch.from_numpy(V[3000 * r:3000 * (r + 1)]))
             for r in range(3)]
    s_a, i_a = X.engine.search_shards(small, torch.from_numpy(Q), 7, weights=w)
    one = X.engine.CorpusStore(9000, dims).add(torch.from_numpy(V[:9000]))
    s_b, i_b = one.search(torch.from_numpy(Q), 7, weights=w)
    assert torch.equal(i_a, i_b) and torch.allclose(s_a, s_b, rtol=0, atol=1e-14)


def test_planted_neighbours_found_at_1m(X):
    """1M-row corpus generated on the device: every query's planted near-duplicate must be rank 1 and the
    returned scores must be sorted; the corpus permuted gives the same answer (order independence)."""
    nv, nq, d, k = 1_000_000, 512, 256, 100
    V = X.synth.device_gaussian(nv, d, 21, "cuda")
    Q = X.synth.device_gaussian(nq, d, 22, "cuda")
    plant = torch.randperm(nv, device="cuda")[:nq]
    V[plant] = Q * 3.0 + 0.05 * X.synth.device_gaussian(nq, d, 23, "cuda")
    store = X.engine.CorpusStore(nv, (d,)).add(V)
    s, i = store.search(Q, k)
    assert torch.equal(i[:, 0], plant)
    assert torch.all(s[:, :-1] >= s[:, 1:])
    perm = torch.randperm(nv, device="cuda")
    store2 = X.engine.CorpusStore(nv, (d,)).add(V[perm])
    s2, i2 = store2.search(Q, k)
    assert torch.equal(perm[i2], i)
    torch.testing.assert_close(s2, s, rtol=0, atol=1e-14)


# ---- exact ground-truth ranks without the score matrix (engine.rank_of_gt) -------------------
def _planted_eval_set(X, nv=200_000, nq=600, d=128, sigma=2.2, seed=15):
    V, Q, vid, cap, _ = X.synth.msrvtt_like(seed, nv, 1, d, sigma)
    return V, Q[:nq], vid, cap[:nq]


def test_rank_metrics_against_a_resident_corpus_equal_the_reference(X):
    """eval_q2m / t2v_map of the reference on the FULL error matrix (oracle) == the same metrics computed against
    the resident store without ever forming the matrix: exact ranks from the tensor-core pass with a guard band."""
    V, Q, vid, cap = _planted_eval_set(X)
    store = X.engine.CorpusStore(len(V), (128,)).add(torch.from_numpy(V))
    _, t2v_gt = linas.get_gt(vid, cap)
    err = linas.cal_error(V.astype(np.float64), Q.astype(np.float64))
    ref = linas.eval_q2m(err, t2v_gt)
    stats = {}
    res = X.metrics.RankResult.from_store(store, torch.from_numpy(Q), t2v_gt, first_only=True, stats=stats)
    np.testing.assert_array_equal(res.ranks.cpu().numpy(), linas.gt_ranks(err, t2v_gt))       # every rank, exactly
    assert res.recall_medr_meanr() == ref and 2.0 < ref[0] < 98.0 and ref[4] > 10.0       # non-degenerate
    assert res.mean_ap() == linas.t2v_map(err, t2v_gt)
    assert X.metrics.eval_q2m_store(store, torch.from_numpy(Q), t2v_gt) == ref
    assert X.metrics.t2v_map_store(store, torch.from_numpy(Q), t2v_gt) == linas.t2v_map(err, t2v_gt)
    assert stats["deep_entries"] == 0


def test_rank_of_gt_deep_fallback_and_multi_gt(X):
    """A tiny candidate capacity forces the exact fp64 fallback for the deep ground truths; several ground truths
    per query (the v2t direction: AP over all of them) -- all ranks == the oracle's."""
    V, Q, vid, cap = _planted_eval_set(X, nv=60_000, nq=40, d=96, sigma=3.0, seed=16)
    rng = np.random.default_rng(3)
    gts = [[int(cap[i].split("#")[0][5:])] + rng.integers(0, len(V), 4).tolist() for i in range(len(Q))]
    store = X.engine.CorpusStore(len(V), (96,)).add(torch.from_numpy(V))
    err = linas.cal_error(V.astype(np.float64), Q.astype(np.float64))
    off = np.cumsum([0] + [len(g) for g in gts])
    flat = [x for g in gts for x in g]
    want = []
    for qi, gq in enumerate(gts):
        order = np.argsort(err[qi], kind="stable")
        pos = np.empty(len(V), np.int64)
        pos[order] = np.arange(len(V))
        want += [int(pos[x]) + 1 for x in gq]
    for cap_, deep in ((8192, False), (16, True)):
        stats = {}
        ranks = X.engine.rank_of_gt([store], torch.from_numpy(Q), off, flat, cap=cap_, stats=stats)
        assert ranks.cpu().tolist() == want
        assert (stats["deep_entries"] > 0) == deep
    res = X.metrics.RankResult.from_store(store, torch.from_numpy(Q), gts)
    assert res.mean_ap() == np.mean([linas.ap_from_ranks(want[off[i]:off[i + 1]], nr_relevant=5, k=0, list_len=len(V))
                                     for i in range(len(Q))])


def test_rank_of_gt_over_two_shards_and_two_spaces(X):
    dims, w = (96, 32), (0.7, 0.3)
    V, Q, vid, cap = _planted_eval_set(X, nv=150_000, nq=300, d=sum(dims), sigma=2.5, seed=17)
    cut = 80_001
    shards = [X.engine.CorpusStore(cut, dims).add(torch.from_numpy(V[:cut])),
              X.engine.CorpusStore(len(V) - cut, dims, index_offset=cut).add(torch.from_numpy(V[cut:]))]
    _, t2v_gt = linas.get_gt(vid, cap)
    V64, Q64 = V.astype(np.float64), Q.astype(np.float64)
    err = linas.fused_errors([V64[:, :96], V64[:, 96:]], [Q64[:, :96], Q64[:, 96:]], w)
    assert X.metrics.eval_q2m_store(shards, torch.from_numpy(Q), t2v_gt, weights=w) == linas.eval_q2m(err, t2v_gt)


# ---- CUDA-graph replay of a fixed-shape search ---------------------------------------------------------------------
def test_graph_search_replays_equal_the_eager_search(X):
    """GraphSearch captures the sync-free first pass once; replays on NEW query batches (and new exclusions) must equal
    the eager search, including a batch whose first pass misses certificates and is re-run eagerly."""
    nv, d, nq, k = 120_000, 128, 48, 20
    V = X.synth.gaussian(61, nv, d)
    store = X.engine.CorpusStore(nv, (d,)).add(torch.from_numpy(V))
    gs = X.engine.GraphSearch(store, nq, k, with_exclude=True)
    for seed in (62, 63, 64):
        Q = torch.from_numpy(X.synth.gaussian(seed, nq, d)).cuda()
        excl = torch.randint(0, nv, (nq,), generator=torch.Generator().manual_seed(seed))
        s_ref, i_ref = store.search(Q, k, exclude=excl)
        s, i = gs(Q, exclude=excl)
        assert torch.equal(i, i_ref) and torch.equal(s, s_ref)
    p = gs(Q, exclude=excl, defer=True)                              # deferred resolution works on a replay, too
    s, i = p.result()
    assert torch.equal(i, i_ref) and not p.reran
    # a batch that needs the re-run loop: near-duplicates of the queries sit on the sampling grid, so the sampled
    # threshold comes out too high for some rows (same construction as test_search_rerun_*)
    step = X.engine.plan(k, nv)["step"]
    V2 = V.copy()
    Qn = Q.cpu().numpy()
    grid = np.random.default_rng(5).permutation(np.arange(0, nv, step))[: 30 * nq].reshape(nq, 30)
    for qi in range(nq):
        V2[grid[qi]] = Qn[qi] + 0.3 * np.random.default_rng(qi).standard_normal((30, d)).astype(np.float32)
    store2 = X.engine.CorpusStore(nv, (d,)).add(torch.from_numpy(V2))
    gs2 = X.engine.GraphSearch(store2, nq, k)
    st = {}
    s_ref, i_ref = store2.search(Q, k, stats=st)
    p = gs2(Q, defer=True)
    s, i = p.result()
    assert torch.equal(i, i_ref) and torch.equal(s, s_ref)
    assert p.reran == (st.get("reruns", 0) >= 1)
    Q3 = torch.from_numpy(X.synth.gaussian(65, nq, d)).cuda()       # and the next replay is clean again
    s3, i3 = gs2(Q3)
    s_ref3, i_ref3 = store2.search(Q3, k)
    assert torch.equal(i3, i_ref3) and torch.equal(s3, s_ref3)


def test_head_stream_pipelines_batches_without_changing_results(X):
    """Steps 1-2 (K1, sampling, threshold) of batch i+1 on a second stream while batch i is still in flight: every
    batch must come back exactly as from the plain call -- the head's tensors cross streams (allocator hand-over) and
    several searches are pending at once."""
    nv, d, nq, k = 150_000, 192, 260, 40
    V = X.synth.gaussian(81, nv, d)
    store = X.engine.CorpusStore(nv, (d,)).add(torch.from_numpy(V))
    batches = [torch.from_numpy(X.synth.gaussian(90 + b, nq, d)).cuda() for b in range(5)]
    want = [store.search(q, k) for q in batches]
    head = torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(3):
        pend = []
        got = []
        for q in batches:
            pend.append(X.engine.search_shards([store], q, k, defer=True, head_stream=head))
            if len(pend) > 2:
                got.append(pend.pop(0).result())
            junk = torch.empty((nq, d), device="cuda").normal_()      # allocator churn between the calls
            del junk
        got += [p.result() for p in pend]
        for (s0, i0), (s1, i1) in zip(want, got):
            assert torch.equal(i0, i1) and torch.equal(s0, s1)
    # a host (pinned) batch goes up on the head stream itself
    qh = batches[0].cpu().pin_memory()
    s2, i2 = X.engine.search_shards([store], qh, k, head_stream=head)
    assert torch.equal(i2, want[0][1]) and torch.equal(s2, want[0][0])


@pytest.mark.parametrize("nv,k", [(22_000, 2000), (20_000, 8000)])
def test_deep_lists_over_two_shards(X, nv, k):
    """k = 2000 over a 22 k-row corpus cut into two shards (the sharded path asked row_topj for more
    than its 4096-value capacity and raised XMVE_ERR_LIMIT; lists deeper than 1000 take the one-round rescore), and
    k = 8000 of 20 k rows, where both the sampled threshold (j > 4096) and the rescore bound (k > 4096) come from
    per-shard order statistics joined by a maximum instead of the gathered top-j lists."""
    d, nq = 96, 9
    V, Q = X.synth.gaussian(71, nv, d), X.synth.gaussian(72, nq, d)
    cut = 10_001
    shards = [X.engine.CorpusStore(cut, (d,)).add(torch.from_numpy(V[:cut])),
              X.engine.CorpusStore(nv - cut, (d,), index_offset=cut).add(torch.from_numpy(V[cut:]))]
    s, i = X.engine.search_shards(shards, torch.from_numpy(Q), k)
    ref_idx, ref_s = _oracle_topk(V, Q, k)
    np.testing.assert_array_equal(i.cpu().numpy(), ref_idx)
    np.testing.assert_allclose(s.cpu().numpy(), ref_s, rtol=0, atol=1e-12)
    one = X.engine.CorpusStore(nv, (d,)).add(torch.from_numpy(V))
    s1, i1 = one.search(torch.from_numpy(Q), k)
    assert torch.equal(i1, i) and torch.equal(s1, s)


# ---- C1 with embeddings produced by the reference's random-init model ---------------------------------------------
def test_c1_random_init_model_embeddings(X):
    """BASELINE config 1 ("random-init model"): the embeddings come from the reference's own encoders
    (oracle/make_golden_c1_model.py), the metrics / ranks / top-10 from the reference's evaluation on them."""
    from test_oracle_golden import c1_model_inputs
    g, vid, cap, video_ids, caption_ids = c1_model_inputs()
    V64, Q64 = vid.astype(np.float64), cap.astype(np.float64)
    errors = X.evaluation.cal_error(V64, Q64)
    np.testing.assert_allclose(errors[:6, :6], g["errors_head"], rtol=0, atol=1e-14)
    v2t_gt, t2v_gt = X.metrics.get_gt(video_ids, caption_ids)
    np.testing.assert_array_equal(np.array(X.validate.cal_perf(errors, v2t_gt, t2v_gt), dtype=np.float64), g["perf"])
    store = X.engine.CorpusStore(len(vid), (1536,)).add(torch.from_numpy(vid))
    s, i = store.search(torch.from_numpy(cap), 10)
    np.testing.assert_array_equal(i.cpu().numpy(), g["top10"])
    res = X.metrics.RankResult.from_store(store, torch.from_numpy(cap), t2v_gt, first_only=True)
    np.testing.assert_array_equal(res.ranks.cpu().numpy(), g["t2v_ranks"])
    assert res.recall_medr_meanr() == tuple(g["perf"][1][:5])
