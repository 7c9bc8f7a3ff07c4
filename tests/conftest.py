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
    config.addinivalue_line("markers", "fp32: the CUDA-core fp32 convolutions (the on-device validator), NOT the benched path")
    config.addinivalue_line("markers", "tf32: single-product tf32 tensor-core convolutions")
    config.addinivalue_line("markers", "tf32x3: 3-term tf32 forward / data gradient, single tf32 weight gradient")
    config.addinivalue_line("markers", "bf16x3: 3-term bf16-pair forward / data gradient, single tf32 weight gradient (round-1 forward)")
    config.addinivalue_line("markers", "fp16x3: the library default (what UNMARKED GPU tests run and what bench.py measures): 3-term "
                                       "fp16-pair forward, bf16-pair data gradient, tf32 weight gradient on rounded operands")


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
    """Unmarked GPU tests run the LIBRARY DEFAULT precision -- the path bench.py measures (tcgen05 convolutions: fp16x3 forward,
    bf16x3 data gradient, tf32 weight gradient on rounded operands).  ``fp32`` selects the CUDA-core validator, ``bf16x3`` /
    ``tf32`` / ``tf32x3`` the alternative tensor-core arithmetics; each test states its own tolerance."""
    import torch
    if not torch.cuda.is_available() or "gpu" not in request.keywords:
        yield
        return
    from viai_b200 import ops
    prev = ops.set_precision("fp32" if "fp32" in request.keywords else "tf32x3" if "tf32x3" in request.keywords
                             else "tf32" if "tf32" in request.keywords else "bf16x3" if "bf16x3" in request.keywords else "fp16x3")
    prev_d = ops.set_dgrad_x3(True)
    yield
    ops.set_precision(prev)
    ops.set_dgrad_x3(prev_d)
