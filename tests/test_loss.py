"""Forward value of the triplet ranking loss (cross_modal_video_engine_b200.loss) against goldens minted by the
UNMODIFIED ``LINAS-engine/loss.py`` (oracle/make_golden_loss.py) for every measure / max_violation / cost_style /
direction.  Tolerance: the reference sums fp32 hinge costs (torch CPU), the kernels accumulate in fp64 -- 2e-5 relative
(north_star allows 1e-3 on fp32 scores)."""
import json
import os

import pytest
import torch

import toy_linas as toy
from conftest import GOLDEN


def _gold():
    with open(os.path.join(GOLDEN, "triplet_loss.json")) as f:
        return json.load(f)


def test_loss_golden_is_complete():
    g = _gold()
    assert len(g["cases"]) == len(toy.LOSS_MEASURES) * 2 * 2 * 3
    assert all(c["loss"] > 0.0 for c in g["cases"])            # non-degenerate: margins are violated everywhere


@pytest.mark.gpu
def test_triplet_loss_matches_the_reference():
    from cross_modal_video_engine_b200 import loss
    g = _gold()
    s, im = toy.loss_batch()
    for c in g["cases"]:
        sm, imm = (s.abs(), im.abs()) if c["measure"] == "jaccard" else (s, im)
        crit = loss.TripletLoss(margin=g["margin"], measure=c["measure"], max_violation=c["max_violation"],
                                cost_style=c["cost_style"], direction=c["direction"])
        val = crit(sm.cuda(), imm.cuda())
        assert val.dtype == torch.float32 and val.is_cuda
        assert float(val) == pytest.approx(c["loss"], rel=2e-5, abs=1e-7), c


@pytest.mark.gpu
def test_loss_similarity_matrices_match_the_reference():
    from cross_modal_video_engine_b200 import loss
    g = _gold()
    s, im = toy.loss_batch()
    for name, rec in g["sims"].items():
        sm, imm = (s.abs(), im.abs()) if name == "jaccard" else (s, im)
        m = loss.get_sim(name)(imm.cuda(), sm.cuda())
        assert m.shape == (len(im), len(s)) and m.dtype == torch.float32
        assert float(m.double().sum()) == pytest.approx(rec["sum"], rel=1e-5)
        assert m[:2, :3].flatten().cpu().tolist() == pytest.approx(rec["corner"], rel=1e-5, abs=1e-6)
