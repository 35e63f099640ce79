import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box via gpurun)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _report_wait_watchdog(request):
    """If a kernel's mbarrier watchdog fired during a GPU test, say where (mpg_b200._lib.wait_debug)."""
    yield
    if 'gpu' not in request.keywords:
        return
    try:
        from mpg_b200 import _lib
        rec = _lib.wait_debug()
    except Exception:
        rec = None
    if rec:
        pytest.fail(f'mbarrier watchdog fired: {rec}', pytrace=False)
