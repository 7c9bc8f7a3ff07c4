"""Which arithmetic dominates the SMOOTH gradient error (error vs the fp64 gradient at the CUDA path's own decisions)?
Runs the 2 x 128 x 128 step under the env-selected combination and prints the worst / median per-tensor error.
    VIAI_DGRAD={x3,tf32x3,tf32}  VIAI_WGRAD={,fp32}  VIAI_PRECISION={fp16x3,bf16x3,tf32x3}"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import viai_test_helpers as H
from oracle import viai_oracle as O
from viai_b200 import Options_inpainting as OI, ops
from viai_b200.step import GanTrainer

B, Hh, W = 2, 128, 128
hp = OI.Inpainting_Config(cin_channels=Hh)
torch.manual_seed(1234)
tr = GanTrainer(hp, "cuda")
cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
esd, gsd, dsd = cpu(tr.Mel_Encoder), cpu(tr.Mel_Decoder), cpu(tr.netD)
mel = torch.rand(B, 1, Hh, W)
mask = O.time_band_mask(mel.shape, W // 4, W // 2)
with ops.trace_activation_decisions() as trace:
    got = tr.train_step(mel.cuda(), mask.cuda())
want, want64, want64m, flips = H.matched_oracle(trace, got, esd, gsd, dsd, mel, mask, Hh)
out = {}
for mod, gk in ((tr.netD, "grads_D"), (tr.Mel_Encoder, "grads_E"), (tr.Mel_Decoder, "grads_Dec")):
    ps = dict(mod.named_parameters())
    rows = H.grad_table({k: ps[k]._viai_grad for k in want64m[gk]}, want64m[gk], gk, check=False)
    e = sorted(r[1] for r in rows)
    out[gk] = "med %.1e max %.1e" % (e[len(e) // 2], e[-1])
print("precision=%s dgrad=%s wgrad=%s | fake %.1e | flips cuda %d ref %d | %s" % (
    ops.get_precision(), os.environ.get("VIAI_DGRAD", "x3"), os.environ.get("VIAI_WGRAD", "tf32r"), H.relerr(got["fake"], want64["fake"]),
    flips[0], flips[1], out))
