import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_json(name):
    path = os.path.join(GOLD, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


def unhex(v):
    return np.array([float.fromhex(a) for a in v], np.float64)


@pytest.fixture(scope="session")
def force_kats():
    return load_json("force_kats.json")


@pytest.fixture(scope="session")
def decay_tables():
    return load_json("decay_tables.json.gz")


@pytest.fixture(scope="session")
def decay_events():
    return load_json("decay_events.json.gz")


@pytest.fixture(scope="session")
def u238_traj():
    return dict(np.load(os.path.join(GOLD, "u238_traj.npz")))


@pytest.fixture(params=["ring", "block", "quad", "cluster"])
def ensemble_kernel(request, monkeypatch):
    """Pins pyqmd_ensemble_step's choice between the warp-local ring kernel, the block-wide ring (two
    nucleons per thread), the block-wide ring with four nucleons per thread ("quad") and the 8-CTA cluster
    kernel (csrc/ensemble.cu) so that ALL of them see the parity case, whatever the automatic dispatch would
    pick for its size ("cluster" applies to nuclei of 64..512 nucleons, "quad" to 125..1024; others take
    the automatic choice)."""
    monkeypatch.setenv("PYQMD_ENSEMBLE_KERNEL", request.param)
    return request.param
