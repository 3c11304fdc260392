import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs at least two CUDA devices (spawns one process per GPU)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) on a machine without a CUDA device."""
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    no_gpu = pytest.mark.skip(reason="no CUDA device")
    one_gpu = pytest.mark.skip(reason="needs >= 2 CUDA devices")
    for it in items:
        if "gpu" in it.keywords and n == 0:
            it.add_marker(no_gpu)
        elif "multigpu" in it.keywords and n < 2:
            it.add_marker(one_gpu)


@pytest.fixture(scope="session")
def xarm():
    from easyhec_b200.scenes import load_xarm7
    return load_xarm7()


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from easyhec_b200._lib import Context
    ctx = Context("cuda:0")
    yield ctx
    ctx.close()
