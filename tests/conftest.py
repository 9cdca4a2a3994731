import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import diffab_pytorch_b200  # noqa: E402,F401  (registers the hyphenated package dir)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def checksum(t):
    t = t.detach().double().flatten().cpu()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 97 + 1
    return float((t * w).sum())


@pytest.fixture(scope="session")
def golden():
    return load_golden
