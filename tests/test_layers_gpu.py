"""Every convolution of the generator and the discriminator at its real C1 (B=2, 80x64 mel) geometry: forward,
data gradient and weight gradient against torch fp64 on the CPU.  Exercises multi-block grids and split-K."""
import math

import pytest
import torch
import torch.nn.functional as F

import viai_test_helpers as H

pytestmark = pytest.mark.gpu

# (name, transposed, Cin, Cout, kh, kw, stride, pad, H, W)
LAYERS = [
    ("E.conv1", False, 1, 32, 3, 3, (2, 2), (1, 1), 80, 64),
    ("E.conv2", False, 32, 64, 3, 3, (2, 1), (1, 1), 40, 32),
    ("E.conv3", False, 64, 128, 3, 3, (2, 2), (1, 1), 20, 32),
    ("E.conv4", False, 128, 256, 3, 3, (2, 2), (1, 1), 10, 16),
    ("E.conv5", False, 256, 256, 3, 3, (2, 2), (1, 1), 5, 8),
    ("G.deconv1_1", True, 256, 256, 3, 3, (1, 1), (0, 1), 1, 4),
    ("G.deconv1_1_1", True, 512, 256, 3, 3, (1, 1), (0, 1), 1, 4),
    ("G.block2_0", True, 256, 128, 3, 3, (1, 1), (1, 1), 5, 8),
    ("G.block3_0", True, 128, 64, 3, 3, (1, 1), (1, 1), 10, 16),
    ("G.block4_0", True, 128, 32, 3, 3, (1, 1), (1, 1), 20, 32),
    ("G.block5", True, 32, 32, 3, 3, (1, 1), (1, 1), 40, 32),
    ("G.conv6_1", True, 32, 32, 3, 3, (1, 1), (1, 1), 80, 64),
    ("G.conv6_2", True, 32, 1, 3, 3, (1, 1), (1, 1), 80, 64),
    ("D.conv1", False, 1, 64, 1, 4, (1, 2), (0, 1), 80, 64),
    ("D.conv2_1", False, 64, 128, 3, 3, (2, 2), (1, 1), 80, 32),
    ("D.conv2_2", False, 128, 256, 3, 3, (2, 2), (1, 1), 40, 16),
    ("D.conv3", False, 256, 512, 3, 3, (1, 1), (1, 1), 20, 8),
    ("D.conv4", False, 512, 1, 3, 3, (1, 1), (1, 1), 20, 8),
]


@pytest.mark.parametrize("layer", LAYERS, ids=[l[0] for l in LAYERS])
def test_layer_at_c1_geometry(layer):
    from viai_b200 import ops
    name, tr, Cin, Cout, kh, kw, stride, pad, Hh, W = layer
    N = 2
    g = torch.Generator().manual_seed(abs(hash(name)) % 100000)
    x = torch.randn(N, Cin, Hh, W, generator=g, dtype=torch.float64).float().double().requires_grad_(True)
    wshape = (Cin, Cout, kh, kw) if tr else (Cout, Cin, kh, kw)
    w = (torch.randn(wshape, generator=g, dtype=torch.float64) / math.sqrt(Cin * kh * kw)).float().double().requires_grad_(True)
    b = (torch.randn(Cout, generator=g, dtype=torch.float64) * 0.1).float().double().requires_grad_(True)
    y = F.conv_transpose2d(x, w, b, stride, pad) if tr else F.conv2d(x, w, b, stride, pad)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64).float().double()
    y.backward(dy)
    nhwc = lambda t: t.float().permute(0, 2, 3, 1).contiguous().cuda()
    xg = nhwc(x.detach()).requires_grad_(True)
    wg = w.detach().float().cuda().requires_grad_(True)
    bg = b.detach().float().cuda().requires_grad_(True)
    yg = ops.conv2d(xg, wg, bg, stride, pad, tr)
    yg.backward(nhwc(dy))
    errs = dict(fwd=H.relerr(yg.permute(0, 3, 1, 2), y), dgrad=H.relerr(xg.grad.permute(0, 3, 1, 2), x.grad),
                wgrad=H.relerr(wg.grad, w.grad), bgrad=H.relerr(bg.grad, b.grad))
    print(name, errs)
    for k, v in errs.items():
        assert v < 2e-5, (name, k, v)
