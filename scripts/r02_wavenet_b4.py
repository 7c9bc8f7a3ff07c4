"""ws kernel throughput at B = 1, 2, 4 (T = 8000)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402,F401
from viai_b200.wavenet_vocoder import WaveNet  # noqa: E402

torch.manual_seed(0)
m = WaveNet().cuda().eval()
m.make_generation_fast_()
for B in (1, 2, 4):
    T = 8000
    c = torch.rand(B, 80, T // 160).cuda()
    m.incremental_forward(c=c[:, :, :10].contiguous(), T=1600)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.incremental_forward(c=c, T=T)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("B=%d %s: %.0f samples/s (%.1f us/step)" % (B, m.last_synthesis_kernel, B * T / dt, dt / T * 1e6), flush=True)
