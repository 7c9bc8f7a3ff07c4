"""Per-operator parity of the CUDA path (through the C ABI) against plain torch fp32 on the CPU.
Tolerance: fp32 operators must agree to 1e-4 normwise-relative (viai_test_helpers.relerr) -- one order tighter than
the 1e-3 the north star asks of whole spectrograms, so that error can accumulate over ~25 layers.  The convolution cases run
twice: on the library default (tensor cores; weight gradient = one tf32 product on rounded operands: 5e-4 with these
random-sign inputs, see tests/test_layers_gpu.py) and on the CUDA-core fp32 validator."""
import math
import zlib

import pytest
import torch
import torch.nn.functional as F

import viai_test_helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    from viai_b200 import ops as _ops, _lib
    _lib.lib()
    return _ops


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().cuda()


def nchw(t):
    return t.permute(0, 3, 1, 2).cpu()


CONV_CASES = [
    # (name, transposed, Cin, Cout, kh, kw, stride, pad, N, H, W, bias)
    ("enc.conv1", False, 1, 32, 3, 3, (2, 2), (1, 1), 2, 20, 16, False),
    ("enc.conv2", False, 32, 64, 3, 3, (2, 1), (1, 1), 2, 10, 8, True),
    ("enc.conv3", False, 64, 128, 3, 3, (2, 2), (1, 1), 2, 9, 7, False),
    ("enc.conv5", False, 256, 256, 3, 3, (2, 2), (1, 1), 1, 5, 8, False),
    ("dis.conv1", False, 1, 64, 1, 4, (1, 2), (0, 1), 2, 11, 18, False),
    ("dis.conv3", False, 256, 512, 3, 3, (1, 1), (1, 1), 1, 5, 6, True),
    ("dis.conv4", False, 512, 1, 3, 3, (1, 1), (1, 1), 2, 5, 6, False),
    ("resnet.conv1", False, 3, 64, 7, 7, (2, 2), (3, 3), 1, 30, 30, False),
    ("resnet.down", False, 64, 128, 1, 1, (2, 2), (0, 0), 2, 9, 9, False),
    ("dec.deconv1_1", True, 256, 256, 3, 3, (1, 1), (0, 1), 2, 1, 4, True),
    ("dec.block4_0", True, 128, 32, 3, 3, (1, 1), (1, 1), 1, 10, 12, False),
    ("dec.block5", True, 32, 32, 3, 3, (1, 1), (1, 1), 2, 9, 13, True),
    ("dec.conv6_2", True, 32, 1, 3, 3, (1, 1), (1, 1), 2, 12, 10, True),
    ("wavenet.upsample", True, 1, 1, 3, 4, (1, 4), (1, 0), 1, 8, 5, True),
    # one-channel layers over several (ragged) tiles of the TMA-tiled / multi-pixel kernels, several channel slabs
    ("dis.conv4.tiles", False, 64, 1, 3, 3, (1, 1), (1, 1), 2, 19, 70, True),
    ("dis.conv4.slabs", False, 512, 1, 3, 3, (1, 1), (1, 1), 1, 9, 40, False),
    ("dec.conv6_2.tiles", True, 32, 1, 3, 3, (1, 1), (1, 1), 2, 21, 45, True),
    ("dis.conv1.tiles", False, 1, 64, 1, 4, (1, 2), (0, 1), 2, 19, 150, False),
    ("enc.conv1.tiles", False, 1, 32, 3, 3, (2, 2), (1, 1), 2, 37, 70, True),
]


@pytest.mark.parametrize("prec", [pytest.param("fp16x3", id="default"), pytest.param("fp32", marks=pytest.mark.fp32, id="fp32")])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv2d_forward_backward(ops, case, prec):
    assert ops.get_precision() == prec
    name, tr, Cin, Cout, kh, kw, stride, pad, N, Hh, W, has_bias = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0xFFFF)
    x = torch.randn(N, Cin, Hh, W, generator=g, requires_grad=True)
    wshape = (Cin, Cout, kh, kw) if tr else (Cout, Cin, kh, kw)
    w = (torch.randn(wshape, generator=g) / math.sqrt(Cin * kh * kw)).requires_grad_(True)
    b = (torch.randn(Cout, generator=g) * 0.1).requires_grad_(True) if has_bias else None
    y = F.conv_transpose2d(x, w, b, stride, pad) if tr else F.conv2d(x, w, b, stride, pad)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xg = nhwc(x.detach()).requires_grad_(True)
    wg = w.detach().cuda().requires_grad_(True)
    bg = b.detach().cuda().requires_grad_(True) if has_bias else None
    yg = ops.conv2d(xg, wg, bg, stride, pad, tr)
    assert tuple(nchw(yg).shape) == tuple(y.shape)
    assert H.relerr(nchw(yg), y) < TOL
    yg.backward(nhwc(dy))
    assert H.relerr(nchw(xg.grad), x.grad) < TOL
    assert H.relerr(wg.grad, w.grad) < (TOL if prec == "fp32" else 1e-3)
    if has_bias:
        assert H.relerr(bg.grad, b.grad) < TOL


CIN1_STAT_CASES = [c for c in CONV_CASES if c[2] == 1 and c[3] >= 4] + [
    # more pixel groups than the grid has lane groups (the grid-stride loop runs twice for some lanes), ragged last group
    ("dis.conv1.large", False, 1, 64, 1, 4, (1, 2), (0, 1), 4, 128, 1278, True),
    ("enc.conv1.large", False, 1, 32, 3, 3, (2, 2), (1, 1), 3, 250, 1022, False),
    ("cin1.48ch", False, 1, 48, 3, 3, (1, 1), (1, 1), 2, 17, 23, True),
]


@pytest.mark.parametrize("case", CIN1_STAT_CASES, ids=[c[0] for c in CIN1_STAT_CASES])
def test_cin1_conv_fused_statistics(ops, case):
    """MelEncoder.conv1 / MelDiscriminator.conv1 in front of their BatchNorm: the per-channel sum and sum of squares come out of the
    convolution kernel (viai_conv2d_thin_stats) -- same numbers as viai_channel_stats over the output, and no launch of it."""
    import ctypes
    from viai_b200 import _lib
    name, tr, Cin, Cout, kh, kw, stride, pad, N, Hh, W, has_bias = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0xFFFF)
    x = torch.randn(N, Cin, Hh, W, generator=g)
    w = torch.randn((Cout, Cin, kh, kw), generator=g) / math.sqrt(Cin * kh * kw)
    b = torch.randn(Cout, generator=g) * 0.1 if has_bias else None
    y = F.conv2d(x.double(), w.double(), b.double() if has_bias else None, stride, pad)
    xg, wg, bg = nhwc(x), w.cuda(), (b.cuda() if has_bias else None)
    Ho, Wo = y.shape[2], y.shape[3]
    geom = ops._geom(N, Hh, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad, 0)
    assert _lib.lib().viai_conv2d_thin_stats_supported(ctypes.byref(geom)) == 1
    with torch.no_grad():
        ops.conv2d_stats(xg, wg, bg, stride, pad, False, 1)              # (packs the weight: not counted below)
        n0 = _lib.launch_count()
        yg, stats = ops.conv2d_stats(xg, wg, bg, stride, pad, False, 1)
        launches = _lib.launch_count() - n0
        yi, stats_in = ops.conv2d_stats(xg, wg, bg, stride, pad, False, N)   # InstanceNorm groups: the separate pass
    assert launches <= 2, launches                                        # (weight re-layout +) convolution: no statistics pass
    assert H.relerr(nchw(yg), y) < TOL
    assert stats.shape == (2, Cout) and stats.dtype == torch.float64
    assert H.relerr(stats[0], y.sum((0, 2, 3))) < 1e-5
    assert H.relerr(stats[1], (y * y).sum((0, 2, 3))) < 1e-5
    assert H.relerr(yi, yg) < 1e-6
    assert H.relerr(stats_in[0].view(N, Cout).sum(0), stats[0].view(-1)) < 1e-5


@pytest.mark.parametrize("norm", ["bn", "in", "none"])
@pytest.mark.parametrize("act", ["lrelu", "relu", "sigmoid", "none"])
@pytest.mark.parametrize("C", [1, 32, 6])
def test_norm_act_forward_backward(ops, norm, act, C):
    import torch.nn as nn
    g = torch.Generator().manual_seed(C * 7 + len(norm) + len(act))
    N, Hh, W = 3, 9, 7
    x = (torch.randn(N, C, Hh, W, generator=g) * 2 + 0.5).requires_grad_(True)
    mod = {"bn": nn.BatchNorm2d(C), "in": nn.InstanceNorm2d(C, affine=True), "none": None}[norm]
    if mod is not None:
        with torch.no_grad():
            mod.weight.copy_(torch.rand(C, generator=g) + 0.5)
            mod.bias.copy_(torch.randn(C, generator=g) * 0.2)
    import copy
    modg = copy.deepcopy(mod).cuda() if mod is not None else None
    z = mod(x) if mod is not None else x
    z = {"lrelu": lambda t: F.leaky_relu(t, 0.2), "relu": F.relu, "sigmoid": torch.sigmoid, "none": lambda t: t}[act](z)
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)
    code = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "lrelu": ops.ACT_LRELU, "sigmoid": ops.ACT_SIGMOID}[act]
    xg = nhwc(x.detach()).requires_grad_(True)
    zg = ops.norm_act(xg, modg, norm, code, 0.2)
    assert H.relerr(nchw(zg), z) < TOL
    zg.backward(nhwc(dz))
    assert H.relerr(nchw(xg.grad), x.grad) < 5 * TOL
    if mod is not None:
        assert H.relerr(modg.weight.grad, mod.weight.grad) < 5 * TOL
        assert H.relerr(modg.bias.grad, mod.bias.grad) < 5 * TOL
    if norm == "bn":
        assert H.relerr(modg.running_mean, mod.running_mean) < TOL
        assert H.relerr(modg.running_var, mod.running_var) < TOL
        assert int(modg.num_batches_tracked) == int(mod.num_batches_tracked) == 1
        mod.eval(); modg.eval()
        with torch.no_grad():
            ze = F.leaky_relu(mod(x), 0.2)
            zeg = ops.norm_act(xg.detach(), modg, "bn", ops.ACT_LRELU, 0.2)
        assert H.relerr(nchw(zeg), ze) < TOL


@pytest.mark.parametrize("norm", ["bn", "in", "none"])
@pytest.mark.parametrize("C", [32, 6, 512])
def test_norm_sweep_direction_does_not_change_results(ops, norm, C):
    """viai_norm_walk_mb: with the threshold at 0 every pass walks its tensors from where the previous launch ended (mirrored
    block order: apply back to front, bwd_reduce back to front, bwd_apply front to back); with a huge threshold everything walks
    front to back.  Same row ranges per block either way: the elementwise outputs are bit-identical, the per-channel sums equal
    up to the order of their double atomics.  Several blocks per group and a ragged last block."""
    import copy
    import torch.nn as nn
    from viai_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(C + len(norm))
    N, Hh, W = 3, 61, 47 if C < 512 else 13
    x = (torch.randn(N, Hh, W, C, generator=g) * 2 + 0.5).cuda()
    dz = torch.randn(N, Hh, W, C, generator=g).cuda()
    mod = {"bn": nn.BatchNorm2d(C), "in": nn.InstanceNorm2d(C, affine=True), "none": None}[norm]
    if mod is not None:
        with torch.no_grad():
            mod.weight.copy_(torch.rand(C, generator=g) + 0.5)
            mod.bias.copy_(torch.randn(C, generator=g) * 0.2)
    prev = L.viai_norm_walk_mb(-1)
    assert prev >= 0
    res = {}
    try:
        for name, mb in (("alternate", 0), ("forward", 1 << 20)):
            L.viai_norm_walk_mb(mb)
            assert L.viai_norm_walk_mb(-1) == mb
            m = copy.deepcopy(mod).cuda() if mod is not None else None
            xg = x.clone().requires_grad_(True)
            # a front-to-back launch first (as the producing convolution would be), then statistics -> apply, backward: reduce -> apply
            ops.mul(x, x)
            if norm == "none":
                z = ops.norm_act(xg, None, "none", ops.ACT_LRELU, 0.2)
            else:
                z = ops.norm_act(xg, m, norm, ops.ACT_LRELU, 0.2)
            ops.mul(x, x)
            z.backward(dz)
            res[name] = (z.detach().clone(), xg.grad.clone(), None if m is None else (m.weight.grad.clone(), m.bias.grad.clone()),
                         None if norm != "bn" else (m.running_mean.clone(), m.running_var.clone()))
    finally:
        L.viai_norm_walk_mb(prev)
    a, f = res["alternate"], res["forward"]
    if norm == "none":
        assert torch.equal(a[0], f[0]) and torch.equal(a[1], f[1])
    else:
        assert H.relerr(a[0], f[0]) < 1e-6 and H.relerr(a[1], f[1]) < 1e-6
        assert H.relerr(a[2][0], f[2][0]) < 1e-6 and H.relerr(a[2][1], f[2][1]) < 1e-6
    if norm == "bn":
        assert H.relerr(a[3][0], f[3][0]) < 1e-6 and H.relerr(a[3][1], f[3][1]) < 1e-6


BILINEAR_CASES = [(3, 16, 5, 32), (5, 32, 10, 64), (4, 16, 16, 32), (40, 128, 80, 256), (7, 9, 7, 9), (8, 8, 3, 5), (1, 4, 2, 8), (3, 3, 1, 1)]


@pytest.mark.parametrize("case", BILINEAR_CASES, ids=[str(c) for c in BILINEAR_CASES])
@pytest.mark.parametrize("C,Cs", [(8, 0), (4, 4), (3, 2)])
def test_bilinear_cat(ops, case, C, Cs):
    hi, wi, ho, wo = case
    g = torch.Generator().manual_seed(hi * 31 + wo)
    x = torch.randn(2, C, hi, wi, generator=g, requires_grad=True)
    skip = torch.randn(2, Cs, ho, wo, generator=g, requires_grad=True) if Cs else None
    y = F.interpolate(x, size=[ho, wo], mode="bilinear", align_corners=True)
    if Cs:
        y = torch.cat((y, skip), 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xg = nhwc(x.detach()).requires_grad_(True)
    sg = nhwc(skip.detach()).requires_grad_(True) if Cs else None
    yg = ops.bilinear_cat(xg, (ho, wo), sg)
    assert H.relerr(nchw(yg), y) < 1e-6
    yg.backward(nhwc(dy))
    assert H.relerr(nchw(xg.grad), x.grad) < 1e-5
    if Cs:
        assert torch.equal(nchw(sg.grad), skip.grad)


def test_cat_avgpool_maxpool_addact_mul(ops):
    g = torch.Generator().manual_seed(5)
    a = torch.randn(2, 6, 9, 5, generator=g, requires_grad=True)
    b = torch.randn(2, 3, 9, 5, generator=g, requires_grad=True)
    y = F.avg_pool2d(torch.cat((a, b), 1), (3, 1))
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    ag, bg = nhwc(a.detach()).requires_grad_(True), nhwc(b.detach()).requires_grad_(True)
    yg = ops.avgpool_h(ops.cat_channels(ag, bg), 3)
    assert H.relerr(nchw(yg), y) < 1e-6
    yg.backward(nhwc(dy))
    assert H.relerr(nchw(ag.grad), a.grad) < 1e-6 and H.relerr(nchw(bg.grad), b.grad) < 1e-6
    with pytest.raises(RuntimeError):
        ops.avgpool_h(torch.zeros(1, 2, 4, 3, device="cuda"), 3)           # H=2 < 3: the "64x64 mel" failure
    # max pool 3/2/1 incl. ties (quantised input) and odd sizes
    x = (torch.randn(2, 5, 11, 14, generator=g) * 2).round().requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xg = nhwc(x.detach()).requires_grad_(True)
    yg = ops.maxpool3s2(xg)
    assert torch.equal(nchw(yg), y)
    yg.backward(nhwc(dy))
    assert H.relerr(nchw(xg.grad), x.grad) < 1e-6
    # residual add + relu
    p = torch.randn(2, 4, 5, 6, generator=g, requires_grad=True)
    q = torch.randn(2, 4, 5, 6, generator=g, requires_grad=True)
    y = F.relu(p + q)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    pg, qg = nhwc(p.detach()).requires_grad_(True), nhwc(q.detach()).requires_grad_(True)
    yg = ops.add_act(pg, qg)
    assert torch.equal(nchw(yg), y)
    yg.backward(nhwc(dy))
    assert torch.equal(nchw(pg.grad), p.grad) and torch.equal(nchw(qg.grad), q.grad)
    # mask application is bit exact
    mel = torch.rand(2, 1, 80, 64, generator=g)
    mask = H.center_mask(mel.shape)
    assert torch.equal(ops.mul(mel.cuda(), mask.cuda()).cpu(), mel * mask)


def test_losses(ops):
    g = torch.Generator().manual_seed(9)
    p = torch.rand(3, 1, 7, 5, generator=g).clamp(1e-3, 1 - 1e-3).requires_grad_(True)
    q = torch.rand(3, 1, 7, 5, generator=g)
    for kind, t in (("mse", 1.0), ("mse", 0.0), ("bce", 1.0), ("bce", 0.0), ("bce", 0.93), ("l1", 0.0)):
        p.grad = None
        if kind == "mse":
            ref = F.mse_loss(p, torch.full_like(p, t))
        elif kind == "bce":
            ref = F.binary_cross_entropy(p, torch.full_like(p, t))
        else:
            ref = F.l1_loss(p, q)
        (ref * 3.0).backward()
        pg = p.detach().cuda().requires_grad_(True)
        got = {"mse": lambda: ops.mse_scalar(pg, t), "bce": lambda: ops.bce_scalar(pg, t), "l1": lambda: ops.l1_loss(pg, q.cuda())}[kind]()
        assert abs(float(got) - float(ref)) <= 1e-6 * abs(float(ref)) + 1e-9
        ops.lincomb2(got, 3.0).backward()
        assert H.relerr(pg.grad, p.grad) < 1e-6


def test_fused_adam_matches_torch_adam(ops):
    from viai_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(11)
    shapes = [(8, 4, 3, 3), (8,), (5, 7)]
    ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    mine = [torch.nn.Parameter(r.detach().clone().cuda()) for r in ref]
    o_ref = torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999))
    o_mine = FusedAdam(mine, lr=2e-4, betas=(0.5, 0.999))
    for it in range(5):
        o_mine.zero_grad()
        for r, m in zip(ref, mine):
            gr = torch.randn(r.shape, generator=g) * (10.0 ** (it - 2))
            r.grad = gr.clone()
            m._viai_grad.copy_(gr)
        o_ref.step()
        o_mine.step()
    for r, m in zip(ref, mine):
        assert H.relerr(m.data, r.data) < 1e-6
    sd = o_mine.state_dict()
    assert set(sd["state"][0].keys()) >= {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 5.0
    assert H.relerr(sd["state"][0]["exp_avg"], o_ref.state_dict()["state"][0]["exp_avg"]) < 1e-6


def test_errors_are_loud(ops):
    with pytest.raises(RuntimeError):
        ops.conv2d(torch.zeros(1, 4, 4, 2), torch.zeros(3, 2, 3, 3), None, (1, 1), (1, 1), False)    # CPU tensors: no fallback
    from viai_b200 import _lib
    L = _lib.lib()
    assert L.viai_fill(None, 0, 0.0, None) != 0 and b"viai_fill" in L.viai_last_error()
