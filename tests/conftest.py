import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:                      # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def input_sha256(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def case_inputs(manifest, name):
    """Regenerate a golden case's inputs from its recorded seed and check the stored checksum."""
    from cross_modal_video_engine_b200 import synth
    rec = manifest[name]
    V, Q, vid_ids, cap_ids, owner = synth.msrvtt_like(rec["seed"], rec["nv"], rec["cpv"], rec["dim"],
                                                      rec["sigma"], ragged=rec["ragged"])
    assert input_sha256(V, Q) == rec["input_sha256"], "synthetic generator drifted from the golden inputs"
    return V, Q, vid_ids, cap_ids, owner
