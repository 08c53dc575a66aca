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
