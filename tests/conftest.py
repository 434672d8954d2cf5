import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path: sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


@pytest.fixture(autouse=True)
def _host_reference_arithmetic_for_cpu_tests(request):
    """The product refuses host tensors (no CPU fallback, no host arithmetic in the package). Tests NOT marked `gpu` check naming /
    wiring / multi-process logic on the CPU with the reference implementations of tests/host_ref.py, registered for the duration
    of the test; GPU tests run without them, as in production."""
    from tests import host_ref
    if 'gpu' not in request.keywords: host_ref.install()
    yield
    host_ref.uninstall()


# Order of the GPU suite: the parity evidence of the hot path first (loss stack vs the float64 oracle, optimiser, networks, whole
# step, full-size cases, drop-in), then the kernel unit tests, the CUDA-graph runners last — so that a failure in one of the
# later groups cannot keep the parity tests from being collected and run.
_ORDER = ['test_loss_gpu', 'test_optim_gpu', 'test_nets_gpu', 'test_step_gpu', 'test_fullsize_gpu', 'test_plugin_gpu', 'test_gemm_gpu',
          'test_conv_gpu', 'test_aspect_gpu', 'test_graph_gpu']


def pytest_collection_modifyitems(config, items):
    import torch
    rank = {name: i for i, name in enumerate(_ORDER)}
    items.sort(key=lambda it: rank.get(Path(str(it.fspath)).stem, len(_ORDER)))  # stable: file order inside a group is kept
    if torch.cuda.is_available():
        # One red GPU test must not hide the others (round 1: `-x` stopped at a CUDA-graph test and 51 parity tests never ran on
        # the driver's box). The exit status is unchanged — any failure still fails the run — only the early stop is lifted.
        if 'gpu' in (config.getoption('-m') or '') and 'not gpu' not in (config.getoption('-m') or ''):
            config.option.maxfail = 0
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords: item.add_marker(skip)
