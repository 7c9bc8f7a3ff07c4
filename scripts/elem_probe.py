#!/usr/bin/env python
"""Achieved HBM bandwidth of the elementwise / reduction kernels on the largest activation of C2, next to torch's own
copy / add on the same tensors (what the memory system can deliver for 1-, 2- and 3-stream passes)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viai_b200 import _lib, ops
L = _lib.lib()
N, H, W, C = 32, 256, 256, 32
if len(sys.argv) > 1:
    C = int(sys.argv[1]); H = W = int(sys.argv[2])
y = torch.randn(N, H, W, C, device="cuda"); dz = torch.randn_like(y); out = torch.empty_like(y)
mean = torch.zeros(C, device="cuda"); inv = torch.ones(C, device="cuda"); ga = torch.ones(C, device="cuda"); be = torch.zeros(C, device="cuda")
s = torch.zeros(2, C, device="cuda", dtype=torch.float64)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rows = N * H * W
nb = y.numel() * 4
def run(name, fn, nbytes, it=10):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / it
    print("%-28s %8.1f us  %6.2f TB/s" % (name, ms * 1e3, nbytes / ms / 1e9))
run("torch copy (1r+1w)", lambda: out.copy_(y), 2 * nb)
run("torch add (2r+1w)", lambda: torch.add(y, dz, out=out), 3 * nb)
run("torch sum (1r)", lambda: y.sum(), nb)
run("torch mul+sum (2r)", lambda: torch.dot(y.view(-1), dz.view(-1)), 2 * nb)
run("viai apply (1r+1w)", lambda: L.viai_norm_act_fwd(p(y), rows, 1, C, p(mean), p(inv), p(ga), p(be), 1, 0.0, p(out), st), 2 * nb)
run("viai stats (1r)", lambda: L.viai_channel_stats(p(y), rows, 1, C, p(s[0]), p(s[1]), st), nb)
run("viai bwd_reduce (2r)", lambda: L.viai_norm_act_bwd_reduce(p(dz), p(y), rows, 1, C, p(mean), p(inv), p(ga), p(be), 1, 0.0, p(s[0]), p(s[1]), st), 2 * nb)
run("viai bwd_apply (2r+1w)", lambda: L.viai_norm_act_bwd_apply(p(dz), p(y), rows, 1, C, p(mean), p(inv), p(ga), p(be), 1, 0.0, p(s[0]), p(s[1]), p(out), None, None, st), 3 * nb)
