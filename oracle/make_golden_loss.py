"""Mint the goldens of the triplet ranking loss by running the UNMODIFIED ``LINAS-engine/loss.py`` on CPU (build
container only):

    python oracle/make_golden_loss.py

``TripletLoss.forward`` (loss.py:112-153) calls ``.cuda()`` on the zero cost of an unused direction (:146-149); on
this GPU-less container ``torch.Tensor.cuda`` is patched to the identity for the run (the arithmetic is untouched).
Writes ``tests/golden/triplet_loss.json``: the loss of every (measure, max_violation, cost_style, direction)
combination on the seeded batch of ``tests/toy_linas.loss_batch``, plus the similarity matrices' checksums.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/LINAS-engine"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    sys.path.insert(0, REF)
    import loss as ref_loss                        # noqa: E402  (reference, unmodified)
    import toy_linas as toy
    torch.Tensor.cuda = lambda self, *a, **k: self  # no GPU here: keep the tensors where they are
    s, im = toy.loss_batch()
    cases = []
    for measure in toy.LOSS_MEASURES:
        sm, imm = (s.abs(), im.abs()) if measure == 'jaccard' else (s, im)
        for max_violation in (False, True):
            for cost_style in ('sum', 'mean'):
                for direction in ('all', 't2v', 'v2t'):
                    crit = ref_loss.TripletLoss(margin=0.2, measure=measure, max_violation=max_violation,
                                                cost_style=cost_style, direction=direction)
                    val = crit(sm, imm)
                    cases.append({"measure": measure, "max_violation": max_violation, "cost_style": cost_style,
                                  "direction": direction, "loss": float(val)})
    sims = {}
    for name in ('cosine', 'order', 'euclidean', 'jaccard'):
        sm, imm = (s.abs(), im.abs()) if name == 'jaccard' else (s, im)
        m = ref_loss.get_sim(name)(imm, sm)
        sims[name] = {"sum": float(m.double().sum()), "corner": [float(x) for x in m[:2, :3].flatten()]}
    with open(os.path.join(OUT, "triplet_loss.json"), "w") as f:
        json.dump({"margin": 0.2, "cases": cases, "sims": sims}, f, indent=0)
    print(len(cases), "cases;", cases[0], cases[-1])


if __name__ == "__main__":
    main()
