"""Host logic of the ResNet-stem weight-gradient path (ops._stem_wgrad: im2col with channel-major columns + one (Cout, K)
GEMM that lands directly in the (Cout, Cin, kh, kw) layout), dry-run with torch stand-ins for the two kernels.  The kernels
themselves are checked by tests/test_fast_stem_gpu.py (on the B200)."""
import pytest
import torch
import torch.nn.functional as F

import torch_ops_shim as shim
import viai_test_helpers as H


@pytest.mark.parametrize("cfg", [(2, 3, 64, 7, 7, 2, 3, 30, 26), (1, 2, 16, 7, 7, 2, 3, 17, 23), (2, 4, 32, 3, 3, 1, 1, 9, 8)],
                         ids=lambda c: "N%d_Cin%d_Cout%d_%dx%d_s%d_p%d_%dx%d" % c)
@pytest.mark.parametrize("accumulate", [False, True])
def test_stem_wgrad_layout(cfg, accumulate):
    from viai_b200 import ops
    N, Cin, Cout, R, S, st, pd, Hh, W = cfg
    g = torch.Generator().manual_seed(sum(cfg))
    x = torch.randn(N, Cin, Hh, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, R, S, generator=g, dtype=torch.float64).requires_grad_(True)
    y = F.conv2d(x, w, None, st, pd)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    Ho, Wo = y.shape[2:]
    geom = ops._geom(N, Hh, W, Cin, Ho, Wo, Cout, R, S, (st, st), (pd, pd), 0)
    base = torch.randn(Cout, Cin, R, S, generator=g, dtype=torch.float64)
    dw = base.clone() if accumulate else torch.zeros(Cout, Cin, R, S, dtype=torch.float64)
    with shim.installed():
        ops._stem_wgrad(geom, x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), dw, accumulate)
    want = w.grad + (base if accumulate else 0)
    assert H.relerr(dw, want) < 1e-12


def test_fast_stem_is_on_by_default():
    """Validated on B200 in round 2 (tests/test_fast_stem_gpu.py) and switched on; VIAI_FAST_STEM=0 restores the old kernels."""
    import os
    from viai_b200 import ops
    assert ops._FAST_STEM == (os.environ.get("VIAI_FAST_STEM", "1") == "1")
