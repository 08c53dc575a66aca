"""Sharded search on REAL GPUs over NCCL: launches tools/check_sharded.py under torch.distributed.run with
min(4, visible GPUs) ranks (one process per GPU).  Every rank holds one shard; the sharded search must return, on
every rank, exactly what one store holding the whole corpus returns -- filtered path, small-corpus path,
exclusions, AVS AP, the re-run loop in lockstep, rank-of-ground-truth metrics, host upload.  Skipped on a box with
one GPU (the gloo world-size-2 tests cover the protocol on CPU)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_search_over_nccl_equals_one_store():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one visible GPU: the NCCL path needs at least two")
    ranks = min(4, n)
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "check_sharded.py")]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "sharded search on %d GPUs: OK" % ranks in out.stdout
    assert "MISMATCH" not in out.stdout


def test_two_stores_on_two_devices_in_one_process():
    """function attributes (opt-in shared memory) and the scheduler-counter symbol were cached per PROCESS,
    so the first launch on a second GPU of the same process failed.  A store on cuda:1, used while cuda:0 is the
    current device, must give the same lists as the store on cuda:0 -- alone and as two shards of one search."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    from cross_modal_video_engine_b200 import engine, synth
    torch.cuda.set_device(0)
    nv, d, nq, k = 90_000, 128, 300, 20
    V, Q = synth.gaussian(81, nv, d), synth.gaussian(82, nq, d)
    s0 = engine.CorpusStore(nv, (d,), device="cuda:0").add(torch.from_numpy(V))
    s1 = engine.CorpusStore(nv, (d,), device="cuda:1").add(torch.from_numpy(V))       # cuda:0 stays current
    a_s, a_i = s0.search(torch.from_numpy(Q), k)
    b_s, b_i = s1.search(torch.from_numpy(Q), k)
    assert b_i.device.index == 1
    assert torch.equal(a_i.cpu(), b_i.cpu()) and torch.equal(a_s.cpu(), b_s.cpu())
    gs = engine.GraphSearch(s1, nq, k)                                               # capture on the non-current device
    g_s, g_i = gs(torch.from_numpy(Q).to("cuda:1"))
    assert torch.equal(g_i.cpu(), a_i.cpu())
