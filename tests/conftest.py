import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path: sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


@pytest.fixture(autouse=True)
def _host_reference_arithmetic_for_cpu_tests(request):
    """The product refuses host tensors (no CPU fallback). Tests NOT marked `gpu` check naming / wiring / multi-process logic on
    the CPU, so they opt in to the package's host reference arithmetic; GPU tests run with it disabled, as in production."""
    from slowtv_monodepth_b200 import _lib
    _lib.host_test_mode('gpu' not in request.keywords)
    yield
    _lib.host_test_mode(False)


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available(): return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords: item.add_marker(skip)
