"""Mint golden vectors by running the UNMODIFIED reference on seeded inputs (build container only).

    HOME=/root python oracle/make_golden.py

imports ``evaluation``, ``validate``, ``util.metrics`` and ``basic.metric`` from
``/root/reference/LINAS-engine`` (they run on CPU under this image's NumPy / SciPy / torch) and
``combiner`` from ``/root/reference/MultiFusion/src``, and writes ``tests/golden/*.npz|json``.
The GPU box has no ``/root/reference``: tests read only the committed fixtures.  Inputs come from
``cross_modal_video_engine_b200.synth`` with the seeds recorded in each fixture, and an input
checksum is stored so generator drift is detected before any comparison.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
os.environ.setdefault("HOME", "/root")          # basic/constant.py:4 reads it


def checksum(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    sys.path.insert(0, os.path.join(REF, "LINAS-engine"))
    import evaluation as ref_eval                 # noqa: E402  (reference module)
    import validate as ref_validate               # noqa: E402
    import util.metrics as ref_metrics            # noqa: E402
    from basic.metric import getScorer            # noqa: E402
    from cross_modal_video_engine_b200 import synth

    os.makedirs(OUT, exist_ok=True)
    manifest = {}

    # ---- APScorer known answers (basic/metric.py:25-46; lists from the commented block :128-139)
    ap_cases = []
    for labels in ([1, 1, 0, 0, 0], [3, 2, 3, 0, 1, 2], [0, 0, 0], [0, 1, 0, 0, 1, 0, 0, 0, 1], [1]):
        for name in ("AP", "AP@1", "AP@2", "AP@3", "AP@1000"):
            ap_cases.append({"labels": labels, "scorer": name,
                             "score": float(getScorer(name).score(labels))})
    with open(os.path.join(OUT, "apscorer.json"), "w") as f:
        json.dump(ap_cases, f, indent=0)

    # ---- full-path cases: (name, seed, n_video, caps_per_video, dim, sigma, ragged, store_inputs)
    cases = [
        ("tiny_ragged", 11, 24, 4, 32, 1.5, True, True),
        ("small_cpv20", 12, 150, 20, 256, 6.0, False, False),
        ("c1_1k", 0, 1000, 1, 1536, 14.0, False, False),
    ]
    for name, seed, nv, cpv, dim, sigma, ragged, store in cases:
        V, Q, vid_ids, cap_ids, owner = synth.msrvtt_like(seed, nv, cpv, dim, sigma, ragged=ragged)
        rec = {"seed": seed, "nv": nv, "cpv": cpv, "dim": dim, "sigma": sigma, "ragged": ragged,
               "input_sha256": checksum(V, Q)}
        out = {}
        for tag, cast in (("f32", np.float32), ("f64", np.float64)):
            Vc, Qc = V.astype(cast), Q.astype(cast)
            errors = ref_eval.cal_error(Vc, Qc, "cosine")
            v2t_gt, t2v_gt = ref_metrics.get_gt(vid_ids, cap_ids)
            # the reference indexes t2v_gt by row: every caption here has a video
            perf = ref_validate.cal_perf(errors, v2t_gt, t2v_gt)
            out["perf_" + tag] = np.array(perf, dtype=np.float64)      # [2, 6]: (v2t, t2v)
            out["t2v_ranks_" + tag] = _ranks(errors, t2v_gt)
            out["v2t_ranks_" + tag] = _ranks(errors.T, v2t_gt)
            out["top10_" + tag] = np.stack([np.argsort(errors[i])[:10] for i in range(min(64, len(errors)))])
            out["errors_sum_" + tag] = np.array(errors.astype(np.float64).sum())
            out["errors_head_" + tag] = errors[:8, :8].copy()
            if store:
                out["errors_" + tag] = errors
                out["simi_" + tag] = ref_eval.cal_simi(Qc, Vc, "cosine")
                out["norm_score_" + tag] = ref_validate.norm_score(errors)
                out["l2norm_" + tag] = ref_eval.l2norm(Vc)
        if store:
            out["V"], out["Q"] = V, Q
            rec["video_ids"], rec["caption_ids"] = vid_ids, cap_ids
            rec["v2t_gt"] = v2t_gt
            rec["t2v_gt"] = {str(k): v for k, v in t2v_gt.items()}
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        manifest[name] = rec
        print(name, "perf_f64", out["perf_f64"].tolist())

    # ---- legacy fixed-n_caption metrics (util/metrics.py:5-57,161-218)
    V, Q, _, _, _ = synth.msrvtt_like(21, 40, 5, 48, 2.0)
    errors = ref_eval.cal_error(V.astype(np.float64), Q.astype(np.float64))
    legacy = {
        "t2v": [float(x) for x in ref_metrics.t2v(errors, n_caption=5)],
        "v2t": [float(x) for x in ref_metrics.v2t(errors, n_caption=5)],
        "t2v_inv_rank": float(ref_metrics.t2v_inv_rank(errors, n_caption=5)),
        "v2t_inv_rank": float(ref_metrics.v2t_inv_rank(errors, n_caption=5)),
        "v2t_inv_rank_multi": [float(x) for x in ref_metrics.v2t_inv_rank_multi(errors, n_caption=5)],
        "seed": 21, "input_sha256": checksum(V, Q),
    }
    manifest["legacy"] = legacy

    # ---- MultiFusion: only Combiner.time_process is runnable (combiner.py:140-143)
    try:
        import torch
        sys.path.insert(0, os.path.join(REF, "MultiFusion", "src"))
        from combiner import Combiner             # noqa: E402
        x = torch.from_numpy(synth.gaussian(31, 37 * 8, 640).reshape(37, 8, 640))
        pooled = Combiner.time_process(None, x)
        np.savez_compressed(os.path.join(OUT, "mf_time_process.npz"), pooled=pooled.numpy())
        manifest["mf_time_process"] = {"seed": 31, "shape": [37, 8, 640]}
    except Exception as exc:                       # pragma: no cover
        manifest["mf_time_process"] = {"unavailable": repr(exc)}

    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f)
    print("wrote", sorted(os.listdir(OUT)))


def _ranks(scores, gts):
    """gt_ranks as the reference's eval_q2m loop computes them (it does not return them)."""
    n_q, n_m = scores.shape
    out = np.zeros((n_q,), np.int32)
    for i in range(n_q):
        order = np.argsort(scores[i])
        rank = n_m + 1
        for k in gts[i]:
            rank = min(rank, int(np.where(order == k)[0][0]) + 1)
        out[i] = rank
    return out


if __name__ == "__main__":
    main()
