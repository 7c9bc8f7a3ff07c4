"""Per-kernel CUDA time (torch.profiler / CUPTI) of one eager vision-infused (C3) step, B = 4, 80 x 256 mel, T = 64 frames."""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    from viai_b200 import Options_inpainting as OI
    from viai_b200.networks.Image_Embedding import ImageEmbedding
    from viai_b200.step import GanTrainer
    B, W = 4, 256
    T = W // 4
    hp = OI.Inpainting_Config(cin_channels=80)
    torch.manual_seed(0)
    tr = GanTrainer(hp, "cuda", decoder="MelDecoderImage", video_encoder=ImageEmbedding(hp).cuda())
    mel = torch.rand(B, 1, 80, W).cuda()
    mask = torch.ones_like(mel)
    mask[..., W // 4:W // 4 + W // 2] = 0
    video = torch.randn(B, T, 3, 224, 224).clamp(-1, 1).cuda()
    flow = torch.randn(B, T, 2, 224, 224).clamp(-1, 1).cuda()
    for _ in range(2):
        tr.train_step(mel, mask, video, flow)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.train_step(mel, mask, video, flow)
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
