#!/usr/bin/env python
"""WaveNet synthesis rate of both kernels + their agreement on identical noise (teacher-forced logits and free-running samples)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viai_b200.wavenet_vocoder import WaveNet
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
torch.manual_seed(0)
m = WaveNet().cuda().eval()
m.make_generation_fast_()
c = torch.rand(1, 80, T // 160).cuda()
u = torch.empty((T, 1, 11), device="cuda").uniform_(1e-5, 1 - 1e-5)
res = {}
outs = {}
for kern in ("grid", "cluster"):
    os.environ["VIAI_WAVENET_KERNEL"] = kern
    m._packed = None
    m.incremental_forward(c=c[:, :, :5], T=800)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = m.incremental_forward(c=c, T=T, uniforms=u) if "uniforms" in m.incremental_forward.__code__.co_varnames else m.incremental_forward(c=c, T=T); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    outs[kern] = out
    res[kern] = dict(us_per_sample=ms * 1e3 / T, samples_per_s=T / ms * 1e3, finite=bool(torch.isfinite(out).all()))
res["max_abs_diff_first_2000"] = float((outs["grid"][..., :2000] - outs["cluster"][..., :2000]).abs().max())
print(json.dumps(res))
