"""Mint the goldens of the per-epoch driver ``validate.validate`` (LINAS-engine/validate.py:58-90) and of
``evaluation.encode_vid`` / ``encode_text`` (evaluation.py:88-171) by running the UNMODIFIED reference on the toy
model and loaders of ``tests/toy_linas.py`` (build container only):

    HOME=/root python oracle/make_golden_validate.py

Writes ``tests/golden/validate_toy.json`` (``currscore`` and the tensorboard rows for every ``opt`` combination) and
``tests/golden/encode_toy.npz`` (the float64 embedding arrays and the id lists the reference's encoders return).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/LINAS-engine"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("HOME", "/root")


def main():
    sys.path.insert(0, REF)
    import evaluation as ref_eval                  # noqa: E402  (reference)
    import validate as ref_validate                # noqa: E402  (reference)
    import toy_linas as toy

    out = {"cases": []}
    for style, student, metric, direction in toy.OPTS:
        model, tb = toy.Model(48), toy.TbLogger()
        vid, txt = toy.loaders(style)
        score = ref_validate.validate(toy.Opt(style, student, metric, direction), tb, vid, txt, model)
        out["cases"].append({"opt": [style, student, metric, direction], "currscore": float(score), "tb": tb.rows,
                             "val_start_calls": model.started})
        print(style, student, metric, direction, float(score))
    try:
        vid, txt = toy.loaders("distill_from_best_model")
        ref_validate.validate(toy.Opt("other_style", "text", "recall", "all"), toy.TbLogger(), vid, txt, toy.Model(48))
        out["unknown_style_raises"] = None
    except Exception as exc:
        out["unknown_style_raises"] = type(exc).__name__
    print("unknown style ->", out["unknown_style_raises"])

    model = toy.Model(48)
    vid, txt = toy.loaders("distill_from_best_model")
    v_emb, v_ids = ref_eval.encode_vid(model.embed_vis, vid)
    t_emb, t_ids = ref_eval.encode_text(model.embed_txt_distill, txt, "distill_from_best_model")
    _, txt_gt = toy.loaders("GT")
    g_emb, g_ids = ref_eval.encode_text(model.embed_txt_GT, txt_gt, "GT")
    only = ref_eval.encode_vid(model.embed_vis_distill, vid, return_ids=False)
    assert v_emb.dtype == np.float64
    np.savez_compressed(os.path.join(OUT, "encode_toy.npz"), vid=v_emb, txt=t_emb, txt_gt=g_emb, vid_distill=only)
    out["encode"] = {"vid_ids": v_ids, "txt_ids": t_ids, "txt_gt_ids": g_ids,
                     "other_style_returns": repr(ref_eval.encode_text(model.embed_txt_distill, txt, "nope"))}
    with open(os.path.join(OUT, "validate_toy.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
