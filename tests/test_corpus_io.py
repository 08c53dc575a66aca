"""Ingest formats against goldens minted by the reference's own ``util/txt2bin.py``
and ``basic/bigfile.py`` (oracle/make_golden_io.py).  Host logic: CPU only, except the two ``gpu`` tests."""
import filecmp
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from cross_modal_video_engine_b200 import corpus_io

TOY = os.path.join(GOLDEN, "bigfile_toy")


@pytest.fixture(scope="module")
def rec():
    with open(os.path.join(GOLDEN, "bigfile_toy.json")) as f:
        return json.load(f)


def _parse_txt():
    names, rows = [], []
    for line in open(os.path.join(GOLDEN, "bigfile_toy.txt")):
        e = line.split()
        names.append(e[0])
        rows.append([float(x) for x in e[1:]])
    return names, np.array(rows, dtype=np.float32)


def test_bigfile_reader_matches_reference(rec):
    bf = corpus_io.BigFile(TOY)
    assert bf.shape() == rec["shape"] and len(bf.names) == rec["shape"][0]
    names, vecs = bf.read(rec["requests"]["by_name"])
    assert [names, vecs] == rec["by_name"]                     # sorted by row, unknown + duplicate names dropped
    names, vecs = bf.read(rec["requests"]["by_index"], isname=False)
    assert [names, vecs] == rec["by_index"]
    assert bf.read_one("video1_1") == rec["read_one"]
    assert list(bf.read(["nope"])) == rec["empty"]
    assert bf.rows().shape == tuple(rec["shape"]) and bf.rows().dtype == np.float32


def test_bigfile_writer_matches_txt2bin(tmp_path, rec):
    names, feats = _parse_txt()
    assert names == rec["names"]
    n = corpus_io.write_bigfile(str(tmp_path / "bf"), names, feats)
    assert n == rec["shape"][0]                                 # the duplicated name and the NaN row are dropped
    for fn in ("shape.txt", "id.txt", "feature.bin"):
        assert filecmp.cmp(os.path.join(TOY, fn), str(tmp_path / "bf" / fn), shallow=False), fn


def test_video_data_round_trip_on_host(tmp_path):
    embs = np.random.default_rng(3).standard_normal((17, 9))
    ids = ["video%d" % i for i in range(17)]
    p = str(tmp_path / "video_data.pt")
    corpus_io.save_video_data(p, embs, ids)
    d = torch.load(p, weights_only=False)                       # what inference.py:58-60 reads back
    assert np.array_equal(d["video_embs"], embs) and d["video_ids"] == ids


@pytest.mark.gpu
def test_bigfile_to_store_and_search(tmp_path):
    from cross_modal_video_engine_b200 import synth
    from oracle import linas
    nv, d = 30000, 96
    V, Q = synth.clustered(61, nv, d), synth.clustered(62, 40, d)
    names = ["shot%07d" % i for i in range(nv)]
    corpus_io.write_bigfile(str(tmp_path / "bf"), names, V)
    store, ids = corpus_io.BigFile(str(tmp_path / "bf")).to_store(chunk_rows=7001)   # ragged chunks, staging reuse
    assert ids == names and store.n == nv
    assert torch.equal(store.raw[:nv, :d].cpu(), torch.from_numpy(V))
    s, i = store.search(torch.from_numpy(Q), 10)
    err = linas.cal_error(V.astype(np.float64), Q.astype(np.float64))
    ref = np.argsort(err, axis=1, kind="stable")[:, :10]
    np.testing.assert_array_equal(i.cpu().numpy(), ref)


@pytest.mark.gpu
def test_video_data_pt_to_store(tmp_path):
    from cross_modal_video_engine_b200 import synth
    from oracle import linas
    V = synth.gaussian(71, 5000, 64).astype(np.float64)         # encode_vid keeps float64 arrays (evaluation.py:102)
    ids = ["video%d" % i for i in range(len(V))]
    p = str(tmp_path / "video_data.pt")
    corpus_io.save_video_data(p, V, ids)
    store, got = corpus_io.load_video_data(p)
    assert got == ids and store.n == len(V)
    q = synth.gaussian(72, 3, 64).astype(np.float64)
    s, i = store.search(q, 5)
    err = linas.cal_error(V, q)
    np.testing.assert_array_equal(i.cpu().numpy(), np.argsort(err, axis=1, kind="stable")[:, :5])
    np.testing.assert_allclose(s.cpu().numpy(), -np.sort(err, axis=1)[:, :5], rtol=0, atol=1e-12)


def test_cli_arguments():
    from cross_modal_video_engine_b200 import cli
    a = cli.parse_args(["search", "--corpus", "c", "--queries", "q.npy", "--dims", "1536,512", "--weights", "0.6,0.4"])
    assert a.cmd == "search" and a.dims == (1536, 512) and a.weights == (0.6, 0.4) and a.topK == 10
    with pytest.raises(SystemExit):
        cli.parse_args(["search", "--corpus", "c", "--queries", "q.npy", "--dims", "8,8", "--weights", "1"])
    with pytest.raises(SystemExit):
        cli.parse_args(["eval", "--corpus", "c", "--queries", "q.npy"])          # --caption-ids is required


@pytest.mark.gpu
def test_cli_search_and_eval(tmp_path, capsys):
    """inference.py's tail (ids of the top-K) and tester.py's tail (cal_perf log lines) on files."""
    from cross_modal_video_engine_b200 import cli, synth
    from oracle import linas
    V, Q, vid, cap, _ = synth.msrvtt_like(91, 400, 3, 64, 2.0)
    corpus_io.write_bigfile(str(tmp_path / "bf"), vid, V)
    np.save(str(tmp_path / "q.npy"), Q)
    (tmp_path / "cap.txt").write_text(" ".join(cap))
    assert cli.main(["search", "--corpus", str(tmp_path / "bf"), "--queries", str(tmp_path / "q.npy"), "--topK", "5"]) == 0
    lines = capsys.readouterr().out.strip().splitlines()
    err = linas.cal_error(V.astype(np.float64), Q.astype(np.float64))
    assert len(lines) == len(Q)
    for r in (0, 7, len(Q) - 1):
        assert lines[r] == str([vid[i] for i in np.argsort(err[r], kind="stable")[:5]])
    run = tmp_path / "run.trec"
    cli.main(["search", "--corpus", str(tmp_path / "bf"), "--queries", str(tmp_path / "q.npy"), "--topK", "3",
              "--run-file", str(run)])
    first = run.read_text().splitlines()[0].split()
    assert first[:2] == ["q0", "Q0"] and first[2] == vid[int(np.argmin(err[0]))] and first[3] == "1"
    capsys.readouterr()
    cli.main(["eval", "--corpus", str(tmp_path / "bf"), "--queries", str(tmp_path / "q.npy"), "--caption-ids",
              str(tmp_path / "cap.txt"), "--save-topk", str(tmp_path / "topk.pt")])
    out = capsys.readouterr().out
    ref = linas.cal_perf(err, *linas.get_gt(vid, cap))
    for tup in ref:                                                              # validate.py:25-35: both directions
        assert " * r_1_5_10, medr, meanr: {}".format([round(x, 1) for x in tup[:5]]) in out
        assert " * mAP: {}".format(round(tup[5], 4)) in out
    saved = torch.load(str(tmp_path / "topk.pt"), weights_only=False)
    assert saved["idx"].shape == (len(Q), 10) and saved["video_ids"] == vid


@pytest.mark.gpu
def test_cli_eval_against_the_resident_store_when_spaces_are_fused(tmp_path, capsys):
    """--dims/--weights: no error matrix is formed; the t2v lines come from exact ranks against the store and equal
    the oracle's on the fused matrix."""
    from cross_modal_video_engine_b200 import cli, synth
    from oracle import linas
    V, Q, vid, cap, _ = synth.msrvtt_like(93, 30000, 1, 96, 2.2)
    Q, cap = Q[:200], cap[:200]
    corpus_io.write_bigfile(str(tmp_path / "bf"), vid, V)
    np.save(str(tmp_path / "q.npy"), Q)
    (tmp_path / "cap.txt").write_text(" ".join(cap))
    cli.main(["eval", "--corpus", str(tmp_path / "bf"), "--queries", str(tmp_path / "q.npy"), "--caption-ids",
              str(tmp_path / "cap.txt"), "--dims", "64,32", "--weights", "0.7,0.3"])
    out = capsys.readouterr().out
    V64, Q64 = V.astype(np.float64), Q.astype(np.float64)
    err = linas.fused_errors([V64[:, :64], V64[:, 64:]], [Q64[:, :64], Q64[:, 64:]], (0.7, 0.3))
    _, t2v_gt = linas.get_gt(vid, cap)
    ref = linas.eval_q2m(err, t2v_gt)
    assert " * r_1_5_10, medr, meanr: {}".format([round(x, 1) for x in ref]) in out
    assert " * mAP: {}".format(round(linas.t2v_map(err, t2v_gt), 4)) in out


@pytest.mark.gpu
def test_cli_composed_matches_the_reference_golden(tmp_path, capsys):
    """`cli composed` = MultiFusion validate.py's main tail on files: the printed recalls and results_wo_attn.npy of
    golden case 'a' (minted by the unmodified reference, oracle/make_golden_mf.py)."""
    import json
    from cross_modal_video_engine_b200 import cli, synth
    from conftest import GOLDEN, load_golden
    with open(os.path.join(GOLDEN, "mf_cirr.json")) as f:
        rec = json.load(f)["cases"]["a"]
    index, P, names, ref, tgt = synth.composed_retrieval(rec["seed"], rec["n_index"], rec["n_query"],
                                                         frames=rec["frames"], sigma=rec["sigma"])
    torch.save({"index_features": torch.from_numpy(index), "index_names": names}, str(tmp_path / "index.pt"))
    torch.save({"predicted_features": torch.from_numpy(P), "reference_names": ref, "target_names": tgt},
               str(tmp_path / "queries.pt"))
    assert cli.main(["composed", "--index", str(tmp_path / "index.pt"), "--queries", str(tmp_path / "queries.pt"),
                     "--out", str(tmp_path / "results_wo_attn")]) == 0
    out = capsys.readouterr().out
    m = rec["metrics"]
    for name, val in zip(("group_recall_at1", "group_recall_at2", "group_recall_at3"), (-1, -1, -1)):
        assert "%s = %r" % (name, val) in out
    for name, val in zip(("recall_at1", "recall_at5", "recall_at10", "recall_at50"), m[3:]):
        assert "%s = %r" % (name, val) in out
    top = np.load(str(tmp_path / "results_wo_attn.npy"))
    gold = load_golden("mf_cirr_a")["top100"]
    assert (top != gold).mean() < 2e-3                           # identical up to fp32 ties of the reference's ranking
