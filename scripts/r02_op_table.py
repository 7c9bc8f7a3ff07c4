"""Full per-operator table of one eager C2 step (ops.time_ops): family, geometry, launches, ms, TFLOP/s, GB/s, share."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from viai_b200 import Options_inpainting, ops
from viai_b200.step import GanTrainer
B, H, W = 32, 256, 256
hp = Options_inpainting.Inpainting_Config(cin_channels=H)
torch.manual_seed(0)
tr = GanTrainer(hp, "cuda")
mel = torch.rand(B, 1, H, W, device="cuda")
mask = torch.ones_like(mel); mask[..., W // 4:W // 4 + W // 2] = 0
for _ in range(3):
    tr.train_step(mel, mask)
torch.cuda.synchronize()
tr.segment_events = {}
with ops.time_ops() as log:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.train_step(mel, mask); e1.record()
seg = tr.segment_ms()
step = e0.elapsed_time(e1)
tab = ops.summarize_ops(log)
rows = sorted(tab.items(), key=lambda kv: -kv[1]["ms"])
tot = sum(d["ms"] for _, d in rows)
print("eager step %.3f ms, timed ops %.3f ms, segments %s" % (step, tot, seg))
print("%-13s %-52s %3s %8s %7s %8s %8s" % ("op", "geometry", "n", "ms", "share", "TFLOP/s", "GB/s"))
for (fam, key), d in rows:
    print("%-13s %-52s %3d %8.3f %7.4f %8.1f %8.1f" % (fam, key, d["launches"], d["ms"], d["ms"] / step, d["flops"] / (d["ms"] * 1e-3) / 1e12,
                                                      d["bytes"] / (d["ms"] * 1e-3) / 1e9))
fam = {}
for (f, _), d in rows:
    fam[f] = fam.get(f, 0.0) + d["ms"]
print({k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])})
