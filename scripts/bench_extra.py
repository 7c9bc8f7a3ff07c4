#!/usr/bin/env python
"""Timings of the non-GAN rows: WaveNet synthesis (C4), STFT->mel, ImageEmbedding forward/backward."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viai_b200.wavenet_vocoder import WaveNet
from viai_b200.utils import audio

def ev():
    return torch.cuda.Event(enable_timing=True)

res = {}
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
torch.manual_seed(0)
m = WaveNet().cuda().eval()
m.make_generation_fast_()
c = torch.rand(1, 80, T // 160).cuda()
m.incremental_forward(c=c[:, :, :5], T=800)
e0, e1 = ev(), ev()
e0.record(); out = m.incremental_forward(c=c, T=T); e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1)
res["wavenet"] = dict(T=T, ms=ms, samples_per_s=T / ms * 1e3, us_per_sample=ms * 1e3 / T, finite=bool(torch.isfinite(out).all()))
y = torch.randn(16000 * 600).cuda() * 0.1            # 10 minutes of audio
audio.melspectrogram_cuda(y[:160000])
e0, e1 = ev(), ev()
e0.record(); mel = audio.melspectrogram_cuda(y); e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1)
res["stft_mel"] = dict(samples=y.numel(), frames=mel.size(1), ms=ms, frames_per_s=mel.size(1) / ms * 1e3,
                       GBps_alg=(y.numel() * 4 + mel.numel() * 4) / ms / 1e6)
print(json.dumps(res))
