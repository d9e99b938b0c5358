import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

MODES = ["global", "local", "semiglobal_both", "semiglobal_one", "semiglobal_two"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_seq():
    with open(os.path.join(GOLDEN, "pairwise_seq.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_mats():
    return np.load(os.path.join(GOLDEN, "matrices.npz"))


@pytest.fixture(scope="session")
def golden_prof():
    d = np.load(os.path.join(GOLDEN, "pairwise_prof.npz"))
    out = []
    for k in range(int(d["n"])):
        pre = "c%d_" % k
        out.append({key[len(pre):]: d[key] for key in d.files if key.startswith(pre)})
    return out


@pytest.fixture(scope="session")
def golden_cells():
    d = np.load(os.path.join(GOLDEN, "fill_cells.npz"))
    out = []
    for k in range(int(d["n"])):
        pre = "c%d_" % k
        out.append({key[len(pre):]: d[key] for key in d.files if key.startswith(pre)})
    return out
