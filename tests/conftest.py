import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")
    config.addinivalue_line("markers", "tf32: single-product tf32 tensor-core convolutions (default in tests: exact fp32)")
    config.addinivalue_line("markers", "tf32x3: 3-term tf32 forward, single tf32 backward")
    config.addinivalue_line("markers", "bf16x3: the library default: 3-term bf16-pair forward, single tf32 backward")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    has_ref = os.path.isdir("/root/reference/networks")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(autouse=True)
def _conv_precision(request):
    """Parity tests written against fp32 tolerances run the exact CUDA-core convolutions; tests marked ``tf32`` run the
    tensor-core path (the library default) and state their own tolerance."""
    import torch
    if not torch.cuda.is_available() or "gpu" not in request.keywords:
        yield
        return
    from viai_b200 import ops
    prev = ops.set_precision("bf16x3" if "bf16x3" in request.keywords else "tf32x3" if "tf32x3" in request.keywords
                             else "tf32" if "tf32" in request.keywords else "fp32")
    yield
    ops.set_precision(prev)
