"""Every convolution of the generator and the discriminator at its real C1 (B=2, 80x64 mel) geometry and the dominant ones at
their C2 (256x256 mel) geometry: forward, data gradient and weight gradient against torch fp64 on the CPU.  Exercises
multi-block grids and split-K.

Tolerances (max-norm relative): the CUDA-core validator (``fp32``) 2e-5 everywhere; the library default (the benched path) 2e-5
for the forward (3-term fp16-pair products, ~2^-21 per product; the rest is the tensor core's truncating fp32 accumulation), 5e-5 for the data gradient (3-term bf16-pair products, ~2^-17
per product) and 1e-3 for the weight gradient (one tf32
product on operands ROUNDED to tf32: zero-mean 2^-12 per operand; with the random-sign inputs used here the sum and its error
are both random walks, ~2e-4 of the tensor's rms (3-5e-4 at the worst element of 10^5-10^6).  Without the rounding the hardware's truncation adds a one-sided ~1e-3
bias, see test_wgrad_rounding_removes_the_truncation_bias)."""
import math
import zlib

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


# C2 geometries (256 x 256 mel) of the layers that dominate the step, at B = 2 to keep the fp64 CPU side short
LAYERS_C2 = [
    ("E.conv2@c2", False, 32, 64, 3, 3, (2, 1), (1, 1), 128, 128),
    ("G.block4_0@c2", True, 128, 32, 3, 3, (1, 1), (1, 1), 64, 128),
    ("G.block5@c2", True, 32, 32, 3, 3, (1, 1), (1, 1), 128, 128),
    ("G.conv6_1@c2", True, 32, 32, 3, 3, (1, 1), (1, 1), 256, 256),
    ("D.conv2_1@c2", False, 64, 128, 3, 3, (2, 2), (1, 1), 256, 128),
    ("D.conv2_2@c2", False, 128, 256, 3, 3, (2, 2), (1, 1), 128, 64),
    ("D.conv3@c2", False, 256, 512, 3, 3, (1, 1), (1, 1), 64, 32),
]
TOL = {"fp32": dict(fwd=2e-5, dgrad=2e-5, wgrad=2e-5, bgrad=2e-5), "fp16x3": dict(fwd=2e-5, dgrad=5e-5, wgrad=1e-3, bgrad=2e-5)}
PREC = [pytest.param("fp16x3", id="default"), pytest.param("fp32", marks=pytest.mark.fp32, id="fp32")]


def _run_layer(layer, N=2):
    from viai_b200 import ops
    name, tr, Cin, Cout, kh, kw, stride, pad, Hh, W = layer
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 100000)
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
    print(name, ops.get_precision(), errs)
    return errs


@pytest.mark.parametrize("prec", PREC)
@pytest.mark.parametrize("layer", LAYERS, ids=[l[0] for l in LAYERS])
def test_layer_at_c1_geometry(layer, prec):
    from viai_b200 import ops
    assert ops.get_precision() == prec
    for k, v in _run_layer(layer).items():
        assert v < TOL[prec][k], (layer[0], k, v)


@pytest.mark.parametrize("layer", LAYERS_C2, ids=[l[0] for l in LAYERS_C2])
def test_layer_at_c2_geometry_default_precision(layer):
    """The benched arithmetic at the benched geometry (per image; B = 2)."""
    from viai_b200 import ops
    assert ops.get_precision() == "fp16x3"
    for k, v in _run_layer(layer).items():
        assert v < TOL["fp16x3"][k], (layer[0], k, v)
    assert ops.f16_overflow() == 0


def test_wgrad_rounding_removes_the_truncation_bias():
    """tcgen05 kind::tf32 TRUNCATES fp32 operands.  For same-sign operands (post-ReLU activations x a positive gradient) every
    product is then ~2^-10 too small and so is the weight gradient; rounding the operands in shared memory (the kernel's
    default) leaves a zero-mean error.  Measured through the C ABI at D.conv3's C2 geometry."""
    from viai_b200 import ops
    g = torch.Generator().manual_seed(5)
    N, Hh, W, Cin, Cout = 2, 64, 32, 256, 64
    x = torch.rand(N, Hh, W, Cin, generator=g).cuda()
    dy = torch.rand(N, Hh, W, Cout, generator=g).cuda()
    ref = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (Cout, Cin, 3, 3), dy.double().permute(0, 3, 1, 2), padding=1)
    w = torch.zeros(Cout, Cin, 3, 3, device="cuda", requires_grad=True)
    y = ops.conv2d(x, w, None, (1, 1), (1, 1), False)
    y.backward(dy)
    rel = ((w.grad.double() - ref) / ref).cpu()
    print("rounded-operand wgrad: mean rel err %.3e, max |rel err| %.3e" % (float(rel.mean()), float(rel.abs().max())))
    assert abs(float(rel.mean())) < 2e-5 and float(rel.abs().max()) < 1e-4
