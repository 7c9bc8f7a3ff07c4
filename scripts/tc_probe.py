#!/usr/bin/env python
"""GPU probe of the tcgen05 convolution: correctness against cuDNN fp32 (TF32 off) and timing, per variant flag.
Usage: python scripts/tc_probe.py [flags] [case ...]      (each case runs in this process; a trap kills it)"""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from viai_b200 import _lib
from viai_b200._lib import ConvGeom
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
L = _lib.lib()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

CASES = {
    # name: (N, H, W, Cin, Cout, R, S, stride, pad, mode)
    "s1_32_32_small": (2, 16, 8, 32, 32, 3, 3, (1, 1), (1, 1), 0),
    "s1_32_32_ragged": (2, 20, 13, 32, 32, 3, 3, (1, 1), (1, 1), 0),
    "s1_64_128": (2, 32, 32, 64, 128, 3, 3, (1, 1), (1, 1), 0),
    "s1_256_512": (2, 32, 16, 256, 512, 3, 3, (1, 1), (1, 1), 0),
    "s2_64_128": (2, 32, 32, 64, 128, 3, 3, (2, 2), (1, 1), 0),
    "s21_32_64": (2, 32, 32, 32, 64, 3, 3, (2, 1), (1, 1), 0),
    "t1_64_32": (2, 16, 16, 64, 32, 3, 3, (1, 1), (1, 1), 1),
    "t1_pad01": (2, 2, 16, 64, 64, 3, 3, (1, 1), (0, 1), 1),
    "t2_128_64": (2, 16, 16, 128, 64, 3, 3, (2, 2), (1, 1), 1),
    "t21_64_32": (2, 16, 16, 64, 32, 3, 3, (2, 1), (1, 1), 1),
    # timing shapes (C2)
    "T_conv6_1": (32, 256, 256, 32, 32, 3, 3, (1, 1), (1, 1), 0),
    "T_dconv3": (32, 64, 32, 256, 512, 3, 3, (1, 1), (1, 1), 0),
    "T_dconv2_2": (32, 128, 64, 128, 256, 3, 3, (2, 2), (1, 1), 0),
    "T_cb3": (32, 32, 64, 64, 64, 3, 3, (1, 1), (1, 1), 1),
    "T_dconv2_2_dgrad": (32, 64, 32, 256, 128, 3, 3, (2, 2), (1, 1), 1),
}


def run(name, flags):
    N, H, W, Ci, Co, R, Sx, st, pd, mode = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g)
    w = torch.randn(Co, R, Sx, Ci, device="cuda", generator=g) * (1.0 / (R * Sx * Ci) ** 0.5)      # logical wp[o][r][s][i]
    bias = torch.randn(Co, device="cuda", generator=g)
    if mode == 0:
        Ho, Wo = (H + 2 * pd[0] - R) // st[0] + 1, (W + 2 * pd[1] - Sx) // st[1] + 1
        ref = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), bias, st, pd).permute(0, 2, 3, 1).contiguous()
    else:
        Ho, Wo = (H - 1) * st[0] - 2 * pd[0] + R, (W - 1) * st[1] - 2 * pd[1] + Sx
        ref = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 0, 1, 2), bias, st, pd).permute(0, 2, 3, 1).contiguous()
    geom = ConvGeom(N, H, W, Ci, Ho, Wo, Co, R, Sx, st[0], st[1], pd[0], pd[1], mode)
    assert L.viai_conv2d_tc_supported(ctypes.byref(geom)), "unsupported"
    x3 = 1 if (flags & 4) else 0
    wp = torch.empty(L.viai_tc_packed_size(Co, Ci, R, Sx, x3), device="cuda")
    _lib.check(L.viai_pack_weight_tc(P(w), P(wp), Co, Ci, R, Sx, w.stride(0), w.stride(3), w.stride(1), w.stride(2), 0, x3, S()), "pack")
    out = torch.full((N, Ho, Wo, Co), float("nan"), device="cuda")
    ssum = torch.empty(N * Co, device="cuda", dtype=torch.float64)
    ssq = torch.empty(N * Co, device="cuda", dtype=torch.float64)
    groups = N if name.endswith("ragged") else 1
    _lib.check(L.viai_conv2d_tc(ctypes.byref(geom), P(x), P(wp), P(bias), P(out), P(ssum), P(ssq), groups, flags, S()), "conv2d_tc")
    torch.cuda.synchronize()
    err = float((out - ref).abs().max() / ref.abs().max())
    nan = int(torch.isnan(out).sum())
    if groups == 1:
        rs, rq = ref.double().sum((0, 1, 2)), (ref.double() ** 2).sum((0, 1, 2))
    else:
        rs, rq = ref.double().sum((1, 2)).reshape(-1), (ref.double() ** 2).sum((1, 2)).reshape(-1)
    k = groups * Co
    serr = float((ssum[:k] - rs).abs().max() / rs.abs().max())
    qerr = float((ssq[:k] - rq).abs().max() / rq.abs().max())
    msg = "%-18s flags=%d relerr=%.3e nan=%d stats_err=%.2e/%.2e" % (name, flags, err, nan, serr, qerr)
    if name.startswith("T_"):
        for _ in range(3):
            L.viai_conv2d_tc(ctypes.byref(geom), P(x), P(wp), P(bias), P(out), P(ssum), P(ssq), groups, flags, S())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 10
        e0.record()
        for _ in range(it):
            L.viai_conv2d_tc(ctypes.byref(geom), P(x), P(wp), P(bias), P(out), P(ssum), P(ssq), groups, flags, S())
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / it
        fl = 2.0 * N * (Ho * Wo if mode == 0 else H * W) * Co * Ci * R * Sx
        byt = 4.0 * (x.numel() + out.numel())
        msg += "  %.3f ms  %.1f TFLOP/s  %.0f GB/s(alg)" % (ms, fl / ms / 1e9, byt / ms / 1e6)
    print(msg, flush=True)


WCASES = {
    # name: (N, Hin, Win, Cin, Cout, R, S, stride, pad)
    "w_s1_32_32": (2, 16, 16, 32, 32, 3, 3, (1, 1), (1, 1)),
    "w_s1_ragged": (3, 20, 13, 64, 48, 3, 3, (1, 1), (1, 1)),
    "w_s1_256_512": (2, 16, 16, 256, 512, 3, 3, (1, 1), (1, 1)),
    "w_s2_64_128": (2, 32, 32, 64, 128, 3, 3, (2, 2), (1, 1)),
    "w_s21_32_64": (2, 32, 32, 32, 64, 3, 3, (2, 1), (1, 1)),
    "w_pad21": (2, 2, 16, 64, 64, 3, 3, (1, 1), (2, 1)),
    "WT_dconv3": (32, 64, 32, 256, 512, 3, 3, (1, 1), (1, 1)),
    "WT_conv6_1": (32, 256, 256, 32, 32, 3, 3, (1, 1), (1, 1)),
    "WT_dconv2_2": (32, 128, 64, 128, 256, 3, 3, (2, 2), (1, 1)),
}


def run_w(name):
    N, H, W, Ci, Co, R, Sx, st, pd = WCASES[name]
    g = torch.Generator(device="cuda").manual_seed(2)
    Ho, Wo = (H + 2 * pd[0] - R) // st[0] + 1, (W + 2 * pd[1] - Sx) // st[1] + 1
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g)
    dy = torch.randn(N, Ho, Wo, Co, device="cuda", generator=g)
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (Co, Ci, R, Sx), dy.permute(0, 3, 1, 2), st, pd)   # (Co,Ci,R,S)
    geom = ConvGeom(N, H, W, Ci, Ho, Wo, Co, R, Sx, st[0], st[1], pd[0], pd[1], 0)
    assert L.viai_conv2d_wgrad_tc_supported(ctypes.byref(geom)), "unsupported"
    ws = torch.empty(L.viai_wgrad_tc_workspace(ctypes.byref(geom)), device="cuda")
    dw = torch.full((Co, Ci, R, Sx), 1.0, device="cuda")
    call = lambda acc: _lib.check(L.viai_conv2d_wgrad_tc(ctypes.byref(geom), P(dy), P(x), P(dw), dw.stride(0), dw.stride(1),
                                                         dw.stride(2), dw.stride(3), acc, P(ws), S()), "wgrad_tc")
    call(0)
    torch.cuda.synchronize()
    err = float((dw - ref).abs().max() / ref.abs().max())
    call(1)
    torch.cuda.synchronize()
    err2 = float((dw - 2 * ref).abs().max() / ref.abs().max())
    msg = "%-18s relerr=%.3e accumulate=%.3e" % (name, err, err2)
    if name.startswith("WT_"):
        for _ in range(3):
            call(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call(0)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        msg += "  %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * N * Ho * Wo * Co * Ci * R * Sx / ms / 1e9)
    print(msg, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "w":
        for n in (sys.argv[2:] or list(WCASES)):
            try:
                run_w(n)
            except Exception as e:
                print("%-18s FAILED: %s" % (n, str(e)[:300]), flush=True)
                if "CUDA" in str(e) or "cuda" in str(e):
                    break
        sys.exit(0)
    flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    names = sys.argv[2:] or list(CASES)
    for n in names:
        try:
            run(n, flags)
        except Exception as e:
            print("%-18s flags=%d FAILED: %s" % (n, flags, str(e)[:300]), flush=True)
            if "CUDA" in str(e) or "cuda" in str(e):
                break
