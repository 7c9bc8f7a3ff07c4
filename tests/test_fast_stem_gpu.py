"""Fast paths of the ResNet stem (default on; VIAI_FAST_STEM=0 disables them and skips this file): viai_im2col, the im2col +
tensor-core 1x1 weight gradient of the 7x7 stem convolution, and the max-pool backward that reads the forward output."""
import ctypes
import math
import os

import pytest
import torch
import torch.nn.functional as F

import viai_test_helpers as H

pytestmark = pytest.mark.gpu
if os.environ.get("VIAI_FAST_STEM", "1") != "1":
    pytest.skip("the stem fast paths are disabled (VIAI_FAST_STEM=0)", allow_module_level=True)


@pytest.mark.parametrize("cfg", [(2, 3, 7, 7, 2, 3, 30, 26), (3, 2, 7, 7, 2, 3, 224, 224), (2, 4, 3, 3, 1, 1, 9, 8), (1, 3, 5, 3, 2, 1, 11, 13)])
def test_im2col_equals_unfold(cfg):
    from viai_b200 import ops
    N, Cin, R, S, st, pd, Hh, W = cfg
    x = torch.randn(N, Cin, Hh, W, generator=torch.Generator().manual_seed(sum(cfg)))
    Ho, Wo = (Hh + 2 * pd - R) // st + 1, (W + 2 * pd - S) // st + 1
    g = ops._geom(N, Hh, W, Cin, Ho, Wo, 16, R, S, (st, st), (pd, pd), 0)
    K = Cin * R * S
    Kpad = (K + 31) // 32 * 32
    got = ops.im2col(g, x.permute(0, 2, 3, 1).contiguous().cuda(), Kpad).cpu()
    want = F.unfold(x, (R, S), padding=pd, stride=st).transpose(1, 2).reshape(-1, K)
    assert torch.equal(got[:, :K], want) and float(got[:, K:].abs().max() if Kpad > K else 0.0) == 0.0


@pytest.mark.parametrize("shape", [(2, 12, 14, 8), (3, 112, 112, 64), (1, 7, 9, 4)])
def test_maxpool_backward_with_output_equals_torch_including_ties(shape):
    from viai_b200 import ops
    assert ops._FAST_STEM
    N, Hh, W, C = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(N, C, Hh, W, generator=g) * 2).round() / 2              # quantised: plenty of exact ties inside windows
    xr = x.clone().requires_grad_(True)
    y = F.max_pool2d(xr, 3, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xg = x.permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
    yg = ops.maxpool3s2(xg)
    assert torch.equal(yg.permute(0, 3, 1, 2).cpu(), y.detach())
    yg.backward(dy.permute(0, 2, 3, 1).contiguous().cuda())
    assert torch.equal(xg.grad.permute(0, 3, 1, 2).cpu(), xr.grad)


@pytest.mark.tf32
@pytest.mark.parametrize("cfg", [(4, 3, 64, 224), (3, 2, 64, 224), (2, 3, 32, 61)])
def test_stem_convolution_weight_gradient_on_tensor_cores(cfg):
    from viai_b200 import ops
    assert ops.get_precision() == "tf32" and ops._FAST_STEM
    N, Cin, Cout, S = cfg
    g = torch.Generator().manual_seed(sum(cfg))
    x = torch.randn(N, Cin, S, S, generator=g, dtype=torch.float64).float().double()
    w = (torch.randn(Cout, Cin, 7, 7, generator=g, dtype=torch.float64) / math.sqrt(Cin * 49)).float().double().requires_grad_(True)
    y = F.conv2d(x, w, None, 2, 3)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64).float().double()
    y.backward(dy)
    wg = w.detach().float().cuda().requires_grad_(True)
    yg = ops.conv2d(x.float().permute(0, 2, 3, 1).contiguous().cuda(), wg, None, (2, 2), (3, 3))
    yg.backward(dy.float().permute(0, 2, 3, 1).contiguous().cuda())
    assert H.relerr(wg.grad, w.grad) < 2e-3


@pytest.mark.bf16x3
def test_image_embedding_matches_reference_golden_with_fast_stem():
    from oracle import fixtures as FX
    from viai_b200.networks.Image_Embedding import ImageEmbedding
    g0 = H.load_golden("image_embedding.pt")
    M = ImageEmbedding()
    sd = M.state_dict()
    FX.deterministic_fill(sd)
    M.load_state_dict(sd)
    M = M.cuda()
    v = FX.normal("video", (1, 4, 3, 224, 224)).clamp(-1, 1).cuda()
    f = FX.normal("flow", (1, 4, 2, 224, 224)).clamp(-1, 1).cuda()
    out = M(v, f)
    assert H.relerr(out, g0["out"]) < 1e-3
    out.pow(2).sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in M.parameters() if p.grad is not None)
