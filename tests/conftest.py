"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs on the CPU-only build container (oracle vs golden vectors, host logic,
C-ABI symbol export); ``-m gpu`` runs the parity tests proper on a B200 through the C-ABI.
Nothing here (or in any ``-m gpu`` test) reads /root/reference: it does not exist on the GPU box.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def pretrained_sd():
    return load_golden("weights_both_dtu_blended")
