import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import oracle as O
    if not O.available("port"):
        O.build(ref=False)
    return O.Oracle("port")


@pytest.fixture(scope="session")
def oracle_ref_exact():
    from oracle import oracle as O
    if not O.available("ref_exact"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return O.Oracle("ref_exact")


@pytest.fixture(scope="session")
def oracle_ref_native():
    from oracle import oracle as O
    if not O.available("ref_native"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return O.Oracle("ref_native")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return np.load(path)
