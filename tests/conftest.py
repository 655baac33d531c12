import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "admm-quantization_b200")
GOLDEN = os.path.join(REPO, "tests", "golden")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(bytes(self.z["meta"]).decode())

    def __getitem__(self, key):
        return self.z[key]

    def case(self, name):
        return next(m for m in self.meta if m["name"] == name)


@pytest.fixture(scope="session")
def golden_projection():
    return Golden("projection")


@pytest.fixture(scope="session")
def golden_contractions():
    return Golden("contractions")


@pytest.fixture(scope="session")
def golden_admm():
    return Golden("admm_iteration")


@pytest.fixture(scope="session")
def golden_outer():
    return Golden("outer_loop")
