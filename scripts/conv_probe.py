#!/usr/bin/env python
"""Per-layer timing of the tensor-core convolution (forward in the library's default precision, data gradient, weight
gradient) at the C2 shapes, L2 flushed between launches.  Tuning knobs are read from the environment by the library."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viai_b200 import ops
LAYERS = [  # name, transposed, Cin, Cout, stride, H, W (input), N
    ("D.conv3 256->512", False, 256, 512, (1, 1), 64, 32),
    ("D.conv2_2 128->256 s2", False, 128, 256, (2, 2), 128, 64),
    ("D.conv2_1 64->128 s2", False, 64, 128, (2, 2), 256, 128),
    ("G.conv6_1 32->32 T", True, 32, 32, (1, 1), 256, 256),
    ("G.block5 32->32 T", True, 32, 32, (1, 1), 128, 128),
    ("G.block4_0 128->32 T", True, 128, 32, (1, 1), 64, 128),
    ("G.block3 64->64 T", True, 64, 64, (1, 1), 32, 64),
    ("G.block2 128->128 T", True, 128, 128, (1, 1), 16, 32),
    ("G.conv3 64->128 s2", False, 64, 128, (2, 2), 64, 128),
]
flush = torch.empty(64 * 1024 * 1024, device="cuda")
def timeit(fn, it=5):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / it * 1e3
print("knobs", {k: v for k, v in os.environ.items() if k.startswith("VIAI_")})
res = {}
for name, tr, Cin, Cout, st, H, W in LAYERS:
    x = torch.randn(32, H, W, Cin, device="cuda", requires_grad=True)
    w = (torch.randn((Cin, Cout, 3, 3) if tr else (Cout, Cin, 3, 3), device="cuda") * 0.05).requires_grad_(True)
    with torch.no_grad():
        y = ops.conv2d(x, w, None, st, (1, 1), tr)
    gf = 2.0 * y.numel() / Cout * Cout * Cin * 9 / 1e9 if not tr else 2.0 * x.numel() / Cin * Cin * Cout * 9 / 1e9
    dy = torch.randn_like(y)
    def fwd():
        with torch.no_grad():
            ops.conv2d_stats(x, w, None, st, (1, 1), tr, 1)
    yy, _ = ops.conv2d_stats(x, w, None, st, (1, 1), tr, 0)
    t_f = timeit(fwd)
    def bwd():
        torch.autograd.grad(yy, (x, w), dy, retain_graph=True)
    def bwd_x():
        torch.autograd.grad(yy, (x,), dy, retain_graph=True)
    t_b, t_bx = timeit(bwd), timeit(bwd_x)
    print("%-24s %6.1f GF  fwd %7.1f us (%6.1f TF/s)  dgrad %7.1f us (%6.1f TF/s)  wgrad %7.1f us (%6.1f TF/s)" % (
        name, gf, t_f, gf / t_f * 1e-3, t_bx, gf / t_bx * 1e-3, t_b - t_bx, gf / max(t_b - t_bx, 1e-3) * 1e-3))
