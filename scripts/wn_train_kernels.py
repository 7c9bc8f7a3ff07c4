"""Per-kernel CUDA time of one eager WaveNet training step (torch.profiler / CUPTI; no kernel replay), aggregated by name."""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    from viai_b200.wavenet_step import WaveNetTrainer
    from viai_b200.wavenet_vocoder import WaveNet
    B, T = 4, 8000
    torch.manual_seed(0)
    tr = WaveNetTrainer(WaveNet().cuda().train())
    x = (torch.rand(B, 1, T) * 2 - 1).cuda()
    c = torch.rand(B, 80, T // 160).cuda()
    y, mask = x.transpose(1, 2).contiguous(), torch.ones(B, T, 1).cuda()
    for _ in range(2):
        tr.train_step(x, y, c, mask)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.train_step(x, y, c, mask)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "").replace("void ", ""))
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_us,share")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%s,%d,%.1f,%.4f" % (k.replace(",", ";"), a[0], a[1], a[1] / tot))
    print("TOTAL,%d,%.1f,1.0" % (sum(a[0] for a in agg.values()), tot))
