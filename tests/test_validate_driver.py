"""``evaluation.encode_vid`` / ``encode_text`` and the per-epoch driver ``validate.validate`` against goldens minted
by the UNMODIFIED reference functions on the toy model of tests/toy_linas.py (oracle/make_golden_validate.py)."""
import json
import os

import numpy as np
import pytest
import torch

import toy_linas as toy
from conftest import GOLDEN, load_golden


def _gold():
    with open(os.path.join(GOLDEN, "validate_toy.json")) as f:
        return json.load(f)


def _check_encoders(device):
    from cross_modal_video_engine_b200 import evaluation
    g, ids = load_golden("encode_toy"), _gold()["encode"]
    model = toy.Model(48, device)
    vid, txt = toy.loaders("distill_from_best_model", device)
    _, txt_gt = toy.loaders("GT", device)
    v_emb, v_ids = evaluation.encode_vid(model.embed_vis, vid)
    t_emb, t_ids = evaluation.encode_text(model.embed_txt_distill, txt, "distill_from_best_model")
    g_emb, g_ids = evaluation.encode_text(model.embed_txt_GT, txt_gt, "GT")
    only = evaluation.encode_vid(model.embed_vis_distill, vid, return_ids=False)
    for got, name in ((v_emb, "vid"), (t_emb, "txt"), (g_emb, "txt_gt"), (only, "vid_distill")):
        assert torch.is_tensor(got) and got.dtype == torch.float64 and got.device.type == device  # float64 like np.zeros
        np.testing.assert_array_equal(got.cpu().numpy(), g[name])
    assert (v_ids, t_ids, g_ids) == (ids["vid_ids"], ids["txt_ids"], ids["txt_gt_ids"])
    assert repr(evaluation.encode_text(model.embed_txt_distill, txt, "nope")) == ids["other_style_returns"]


def test_encoders_match_reference_cpu_tensors():
    """The scatter-by-dataset-index logic is device-agnostic torch: checked here on CPU tensors."""
    _check_encoders("cpu")


@pytest.mark.gpu
def test_encoders_match_reference_on_device():
    _check_encoders("cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(toy.OPTS)))
def test_validate_driver_matches_reference(case):
    from cross_modal_video_engine_b200 import validate
    rec = _gold()["cases"][case]
    style, student, metric, direction = rec["opt"]
    assert tuple(rec["opt"]) == toy.OPTS[case]
    model, tb = toy.Model(48, "cuda"), toy.TbLogger()
    vid, txt = toy.loaders(style, "cuda")
    score = validate.validate(toy.Opt(style, student, metric, direction), tb, vid, txt, model)
    assert float(score) == rec["currscore"]                     # bit-identical float
    assert tb.rows == rec["tb"]                                 # 12 cal_perf scalars + rsum, same keys, values, step
    assert model.started == rec["val_start_calls"]


@pytest.mark.gpu
def test_validate_driver_unknown_style_raises_like_the_reference():
    from cross_modal_video_engine_b200 import validate
    assert _gold()["unknown_style_raises"] == "UnboundLocalError"
    vid, txt = toy.loaders("distill_from_best_model", "cuda")
    with pytest.raises(UnboundLocalError):
        validate.validate(toy.Opt("other_style", "text", "recall", "all"), toy.TbLogger(), vid, txt,
                          toy.Model(48, "cuda"))
