import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA (sm_100) device")


def pytest_sessionstart(session):
    """libecgmm.so is git-ignored: a checkout that has never been built gets built once here (nvcc cross-compiles
    without a GPU).  The product itself never builds or falls back on its own -- ecgmm.lib.load() raises."""
    so = os.path.join(ROOT, "ecg-multimodal-model_b200", "libecgmm.so")
    if not os.path.exists(so):
        import __graft_entry__ as g

        g.build()


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _fresh_shape_cache():
    """ecgmm.ops caches the library's pure shape queries (workspace sizes, partial-row counts) per shape.  Several tests
    switch kernels through ECGMM_* environment variables, which changes those answers: a cached value from another
    configuration would under- or over-size a buffer.  (Production processes never change the switches mid-run.)"""
    from ecgmm import ops

    ops._SHAPE_CACHE.clear()
    yield
    ops._SHAPE_CACHE.clear()
