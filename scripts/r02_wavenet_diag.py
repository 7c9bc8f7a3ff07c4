"""Where (in time) does the ws kernel differ from the grid kernel under a given VIAI_WN3_EXP?"""
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
T = 640
c = torch.rand(1, 80, T // 160).cuda()
u = torch.empty((T, 1, 11), device="cuda").uniform_(1e-5, 1 - 1e-5)
ti = torch.rand(1, T, 1).cuda() * 2 - 1
os.environ["VIAI_WAVENET_KERNEL"] = "grid"
ref = m.incremental_forward(c=c, T=T, uniforms=u, test_inputs=ti, return_logits=True)[1]
os.environ["VIAI_WAVENET_KERNEL"] = "ws"
for rep in range(3):
    got = m.incremental_forward(c=c, T=T, uniforms=u, test_inputs=ti, return_logits=True)[1]
    d = (got - ref).abs().amax(dim=2)[0]           # per time step
    bad = (d > 1e-4 * ref.abs().max()).nonzero().flatten().tolist()
    print("run %d: %d bad steps of %d; first %s; max diff %.3e (ref max %.3e)" % (rep, len(bad), T, bad[:40], float(d.max()), float(ref.abs().max())))
