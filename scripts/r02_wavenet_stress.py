"""Repeat the ws kernel many times and require bit-identical outputs run to run (a race shows up as a run that differs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402,F401
from viai_b200.wavenet_vocoder import WaveNet  # noqa: E402

torch.manual_seed(0)
m = WaveNet().cuda().eval()
m.make_generation_fast_()
os.environ["VIAI_WAVENET_KERNEL"] = "ws"
bad = 0
for B, T, reps in ((1, 1280, 25), (2, 640, 10), (3, 640, 10), (4, 640, 10)):
    c = torch.rand(B, 80, T // 160).cuda()
    u = torch.empty((T, B, 11), device="cuda").uniform_(1e-5, 1 - 1e-5)
    ref = None
    for r in range(reps):
        out, lg = m.incremental_forward(c=c, T=T, uniforms=u, return_logits=True)
        if ref is None:
            ref = (out.clone(), lg.clone())
        elif not (torch.equal(out, ref[0]) and torch.equal(lg, ref[1])):
            bad += 1
            print("B=%d run %d differs: max |dlogits| %.3e" % (B, r, float((lg - ref[1]).abs().max())), flush=True)
    print("B=%d T=%d: %d runs, finite %s" % (B, T, reps, bool(torch.isfinite(ref[1]).all())), flush=True)
print("differing runs:", bad)
