import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def ops_double():
    """Route the product's host logic to the CPU emulation of the C ABI (tests/ops_double.py) for one test."""
    import mvdfusion_b200.runtime as rt
    from ops_double import TorchOpsDouble
    dbl = TorchOpsDouble()
    real = rt.get_ops
    rt.get_ops = lambda dev: dbl   # test-side monkeypatch: the product has no dispatch seam of its own
    try:
        yield dbl
    finally:
        rt.get_ops = real
