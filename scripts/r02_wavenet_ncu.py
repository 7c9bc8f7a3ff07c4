"""One launch of the folded WaveNet synthesis kernel for an ncu capture (B = 1, T = 800)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402,F401
from viai_b200.wavenet_vocoder import WaveNet  # noqa: E402

os.environ["VIAI_WAVENET_KERNEL"] = os.environ.get("VIAI_WAVENET_KERNEL", "folded")
torch.manual_seed(0)
m = WaveNet().cuda().eval()
m.make_generation_fast_()
T = 800
c = torch.rand(1, 80, T // 160).cuda()
m.incremental_forward(c=c, T=T)
torch.cuda.synchronize()
print("done")
