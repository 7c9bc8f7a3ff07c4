"""STFT -> mel kernel alone: one hour of 16 kHz audio, frames/s (same measurement as bench.py's `stft` sub-record)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from viai_b200.utils import audio
T = 16000 * 3600
y = torch.rand(T, device="cuda") * 2 - 1
for _ in range(2):
    mel = audio.melspectrogram_cuda(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    mel = audio.melspectrogram_cuda(y)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"mel_stage": os.environ.get("VIAI_STFT_MEL", "walk"), "frames": mel.size(1), "ms": ms, "frames_per_s": mel.size(1) / (ms * 1e-3),
                  "checksum": float(mel.double().sum())}))
