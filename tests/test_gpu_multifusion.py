"""MultiFusion composed retrieval on the GPU against goldens minted by the UNMODIFIED reference functions
(oracle/make_golden_mf.py: validate.compute_cirr_val_metrics, inference.compute_cirr_val_metrics)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, input_sha256, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mf():
    from cross_modal_video_engine_b200 import multifusion
    return multifusion


def _manifest():
    with open(os.path.join(GOLDEN, "mf_cirr.json")) as f:
        return json.load(f)


def _inputs(rec, n_query=None):
    from cross_modal_video_engine_b200 import synth
    return synth.composed_retrieval(rec["seed"], rec["n_index"], n_query or rec["n_query"],
                                    frames=rec.get("frames", 8), sigma=rec.get("sigma", 0.5))


def _assert_names_equal_up_to_fp32_ties(top_gpu, top_ref, index, P, names):
    """The reference ranks ``1 - P @ index.T`` in fp32, the engine ranks exact fp64 cosines: the lists are identical
    except where two adjacent scores are closer than fp32 resolution (north_star's tolerance rule)."""
    diff = top_gpu != top_ref
    if not diff.any():
        return
    pooled = torch.nn.functional.normalize(torch.from_numpy(index).mean(dim=1), dim=-1).double().numpy()
    row = {int(n): r for r, n in enumerate(names)}
    for r, c in zip(*np.nonzero(diff)):
        a, b = row[int(top_gpu[r, c])], row[int(top_ref[r, c])]
        assert abs(float(P[r].astype(np.float64) @ (pooled[a] - pooled[b]))) < 1e-6


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_scoring_matches_reference_validate(mf, name):
    rec = _manifest()["cases"][name]
    index, P, names, ref, tgt = _inputs(rec)
    assert input_sha256(index, P, names, ref, tgt) == rec["input_sha256"]
    metrics, top = mf.cirr_metrics_from_features(torch.from_numpy(P), torch.from_numpy(index), names, ref, tgt)
    gold = load_golden("mf_cirr_" + name)["top100"]
    _assert_names_equal_up_to_fp32_ties(top, gold, index, P, names)
    if np.array_equal(top[:, :50], gold[:, :50]):
        assert [float(x) for x in metrics] == rec["metrics"]              # bit-identical floats
    else:                                                                  # an fp32 tie inside the top-50
        np.testing.assert_allclose(metrics, rec["metrics"], atol=100.0 / rec["n_query"] + 1e-9)
    # the query's own reference item never appears (validate.py:76-83)
    assert not (top == np.asarray(ref)[:, None]).any()


class _ToyClip:
    """Stands for the CLIP text tower: deterministic features from the token ids."""
    class visual:
        output_dim = 640

    def eval(self):
        return self

    def float(self):
        return self

    def encode_text(self, tok):
        return tok.float()[:, :1].repeat(1, 640)


class _Relative(torch.utils.data.Dataset):
    """Items as data_utils.py:215 returns them in 'relative' / test mode."""

    def __init__(self, ref, tgt, n_skip=0):
        self.ref, self.tgt, self.n_skip = ref, tgt, n_skip

    def __len__(self):
        return len(self.ref) + self.n_skip

    def __getitem__(self, i):
        if i >= len(self.ref):
            return None                                                    # collate_fn drops these (utils.py:102)
        return int(self.ref[i]), int(self.tgt[i]), "caption %d" % i, [int(self.tgt[i])], np.full((2, 3), i, np.float32)


class _Classic(torch.utils.data.Dataset):
    def __init__(self, names, index):
        self.names, self.index = names, index

    def __len__(self):
        return len(self.names)

    def __getitem__(self, i):
        return int(self.names[i]), self.index[i]


def test_reference_signature_compute_cirr_val_metrics(mf, tmp_path):
    """validate.compute_cirr_val_metrics(relative_val_dataset, clip_model, index_features, index_names,
    combining_function, combiner) -- the call of validate.py:292 / combiner_train.py:398 -- against the golden of
    case 'a'.  The toy tokenizer carries the query number through the 'text tower' and the combining function looks
    the seeded predicted feature up, so that the scoring stage sees exactly the golden's inputs."""
    rec = _manifest()["cases"]["a"]
    index, P, names, ref, tgt = _inputs(rec)
    P_dev = torch.from_numpy(P).cuda()
    index_dev = torch.from_numpy(index).cuda()
    seen = {}

    def tokenize(captions):
        return torch.tensor([[int(c.split()[1])] for c in captions], dtype=torch.long)

    def combining_function(image_features, text_features):
        ref_feats, middle = image_features
        q = text_features[:, 0].long()
        seen["ref_ok"] = seen.get("ref_ok", True) and bool(
            torch.equal(ref_feats, index_dev[torch.from_numpy(mf.name_rows(names, ref)).cuda()[q]]))
        seen["mid_ok"] = seen.get("mid_ok", True) and bool((middle[:, 0, 0].long() == q).all())
        return P_dev[q] * 3.0                                              # un-normalised: :262 normalises

    out = tmp_path / "results_wo_attn"
    metrics = mf.validate.compute_cirr_val_metrics(_Relative(ref, tgt, n_skip=2), _ToyClip(), index_dev,
                                                   [np.int64(n) for n in names], combining_function, None,
                                                   results_path=str(out), tokenize=tokenize)
    assert seen == {"ref_ok": True, "mid_ok": True}
    top = np.load(str(out) + ".npy")
    gold = load_golden("mf_cirr_a")["top100"]
    _assert_names_equal_up_to_fp32_ties(top, gold, index, P, names)
    assert isinstance(metrics, tuple) and len(metrics) == 7
    assert [float(x) for x in metrics] == rec["metrics"]
    # cirr_val_retrieval = extract_index_features + the above (validate.py:275-293)
    m2 = mf.validate.cirr_val_retrieval(combining_function, _ToyClip(), None, None, None,
                                        datasets=(_Classic(names, index), _Relative(ref, tgt)),
                                        results_path=None, tokenize=tokenize)
    assert m2 == metrics


def test_multiple_of_32_queries_is_fine_here(mf):
    """The reference raises on this input (golden 'mult32'); the engine evaluates it and agrees with the oracle."""
    from oracle import multifusion as mf_oracle
    rec = _manifest()["mult32"]
    assert rec["raises"] == "RuntimeError"
    index, P, names, ref, tgt = _inputs(rec)
    metrics, top = mf.cirr_metrics_from_features(torch.from_numpy(P), torch.from_numpy(index), names, ref, tgt)
    m_ref, top_ref = mf_oracle.compute_cirr_val_metrics(torch.from_numpy(P), torch.from_numpy(index), names, ref, tgt)
    _assert_names_equal_up_to_fp32_ties(top, top_ref, index, P, names)
    assert metrics == m_ref


@pytest.mark.parametrize("name", ["i1", "i2"])
def test_reference_signature_inference_top1(mf, name):
    """inference.compute_cirr_val_metrics(ref_vdo_feature, mod_text, clip_model, index_features, index_names,
    combining_function, combiner) -> top-1 name (inference.py:26-66)."""
    rec = _manifest()["inference"][name]
    index, P, names, _, _ = _inputs(rec, n_query=3)
    assert input_sha256(index, P, names) == rec["input_sha256"]
    pooled = torch.from_numpy(index).mean(dim=1).cuda()
    tar_list = ["vid_%d.mp4" % int(x) for x in names]
    store = mf.build_index(pooled)
    got = []
    for qi in range(3):
        q = torch.from_numpy(P[qi:qi + 1]).cuda()
        high, middle = torch.zeros((2, 640)), torch.zeros((2, 18 * 18, 8))
        got.append(mf.inference.compute_cirr_val_metrics(
            (high.cuda(), middle), "mod text", _ToyClip(), pooled, tar_list, lambda img, txt, q=q: q, None,
            store=store if qi else None, tokenize=lambda t: torch.zeros((1, 77), dtype=torch.long)))
    assert got == rec["top1"]


def test_index_offset_names(mf):
    """A shard whose first global row is not 0 still maps rows to names correctly."""
    from cross_modal_video_engine_b200 import synth
    index, P, names, ref, tgt = synth.composed_retrieval(77, 300, 20)
    a = mf.cirr_metrics_from_features(torch.from_numpy(P), torch.from_numpy(index), names, ref, tgt)
    store = mf.build_index(torch.from_numpy(index), index_offset=1000)
    b = mf.cirr_metrics_from_features(torch.from_numpy(P), None, names, ref, tgt, store=store)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
