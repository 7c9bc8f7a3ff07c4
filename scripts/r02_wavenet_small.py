"""Small WaveNet through the ws synthesis kernel (for compute-sanitizer runs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402,F401
from viai_b200.wavenet_vocoder import WaveNet  # noqa: E402

os.environ["VIAI_WAVENET_KERNEL"] = "ws"
kw = dict(layers=8, stacks=2, residual_channels=32, gate_channels=32, skip_out_channels=32, cin_channels=80, out_channels=30,
          upsample_scales=[2, 4], kernel_size=3)
torch.manual_seed(0)
m = WaveNet(dropout=0.0, **kw).cuda().eval()
m.make_generation_fast_()
T = int(os.environ.get("T", "24"))
c = torch.rand(1, 80, T // 8).cuda()
out = m.incremental_forward(c=c, T=T)
torch.cuda.synchronize()
print("done", float(out.abs().max()))
