"""Mint the ingest-format goldens by running the UNMODIFIED reference (build container only):

    HOME=/root python oracle/make_golden_io.py

``util/txt2bin.process`` (LINAS-engine/util/txt2bin.py:21-75) converts a small text feature file (with a duplicated
name and a NaN row, both of which it drops) into a BigFile directory, committed as ``tests/golden/bigfile_toy/``;
``basic/bigfile.BigFile`` (basic/bigfile.py:4-56) then answers a few ``read`` / ``read_one`` requests, recorded in
``tests/golden/bigfile_toy.json``.  ``tests/test_corpus_io.py`` checks ``corpus_io`` against both.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/LINAS-engine"
OUT = os.path.join(ROOT, "tests", "golden")
os.environ.setdefault("HOME", "/root")


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "util"))
    from basic.bigfile import BigFile                  # noqa: E402  (reference)
    import txt2bin                                     # noqa: E402  (reference)

    rng = np.random.default_rng(41)
    names = ["video%d_%d" % (i // 3, i % 3) for i in range(14)]
    names[5] = names[2]                                # duplicated name: first occurrence wins
    feats = rng.standard_normal((14, 6)).astype(np.float32)
    feats[9, 3] = np.nan                               # NaN row: dropped
    txt = os.path.join(OUT, "bigfile_toy.txt")
    with open(txt, "w") as f:
        for n, v in zip(names, feats):
            f.write(n + " " + " ".join(repr(float(x)) for x in v) + "\n")
    d = os.path.join(OUT, "bigfile_toy")
    txt2bin.process(0, [txt], d, 1)
    bf = BigFile(d)
    requests = {
        "by_name": ["video3_1", "video0_0", "nope", "video0_2", "video0_0"],
        "by_index": [7, 0, 3, 3],
    }
    rec = {"names": names, "shape": bf.shape(), "requests": requests,
           "by_name": bf.read(requests["by_name"]),
           "by_index": bf.read(requests["by_index"], isname=False),
           "read_one": bf.read_one("video1_1"),
           "empty": bf.read(["nope"])}
    with open(os.path.join(OUT, "bigfile_toy.json"), "w") as f:
        json.dump(rec, f)
    print("wrote", d, rec["shape"])

    # ---- the non-cosine measures of cal_error / cal_error_batch / cal_simi (evaluation.py:22-35,46-72,80-83)
    import evaluation as ref_eval                      # noqa: E402  (reference)
    from cross_modal_video_engine_b200 import synth
    sys.path.insert(0, ROOT)
    V = np.abs(synth.gaussian(51, 70, 33)).astype(np.float64)      # non-negative, like the concept-space features
    Q = np.abs(synth.gaussian(52, 45, 33)).astype(np.float64)      # jaccard is meant for
    out = {"V": V, "Q": Q}
    for m in ("euclidean", "l1", "l2", "l1_norm", "l2_norm", "jaccard"):
        e = ref_eval.cal_error(V, Q, m)
        out["err_" + m] = e.numpy() if hasattr(e, "numpy") else np.asarray(e)
    out["batch_jaccard"] = np.asarray(ref_eval.cal_error_batch(V, Q, "jaccard", batch_size=20))
    out["simi_jaccard"] = ref_eval.cal_simi(Q, V, "jaccard").numpy()
    np.savez_compressed(os.path.join(OUT, "measures.npz"), **out)
    print("measures", {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
