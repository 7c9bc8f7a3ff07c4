"""Diagnostic (GPU box): end-to-end gradient / update errors of the GAN step vs the fp64 oracle, with the fp32 oracle's
own error (the 'envelope') beside it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn as nn
import viai_test_helpers as H
from oracle import viai_oracle as O, fixtures as FX
from viai_b200 import Options_inpainting, ops
from viai_b200.step import GanTrainer

for norm, B, Hh, W in (("bn", 1, 80, 64), ("in", 2, 96, 48), ("bn", 2, 128, 128), ("bn", 4, 256, 256)):
    nl = nn.BatchNorm2d if norm == "bn" else nn.InstanceNorm2d
    hp = Options_inpainting.Inpainting_Config(cin_channels=Hh, normlayer=nl)
    torch.manual_seed(1234)
    tr = GanTrainer(hp, "cuda", norm_layer_d=nl, norm_layer_e=nl)
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    esd, gsd, dsd = cpu(tr.Mel_Encoder), cpu(tr.Mel_Decoder), cpu(tr.netD)
    mel = torch.rand(B, 1, Hh, W)
    mask = O.time_band_mask(mel.shape, W // 4, W // 2)
    r32, r64 = H.oracle_pair(esd, gsd, dsd, mel, mask, Hh, norm, norm)
    got = tr.train_step(mel.cuda(), mask.cuda())
    print("=== %s B=%d %dx%d  fake err %.2e (oracle32 %.2e)  losses cuda %s oracle %s" % (
        norm, B, Hh, W, H.relerr(got["fake"], r64["fake"]), H.relerr(r32["fake"], r64["fake"]),
        ["%.6f" % float(got[k]) for k in ("loss_D", "loss_G_GAN", "loss_L1")], ["%.6f" % r64[k] for k in ("loss_D", "loss_G_GAN", "loss_L1")]))
    for mod, gk, opt in ((tr.netD, "grads_D", tr.optimizer_D), (tr.Mel_Encoder, "grads_E", tr.optimizer_G), (tr.Mel_Decoder, "grads_Dec", tr.optimizer_G)):
        gscale = max(float(g.abs().max()) for g in r64[gk].values())
        rows = []
        allg, all32, all64 = [], [], []
        for k, p in mod.named_parameters():
            if k not in r64[gk] or float(r64[gk][k].abs().max()) < 1e-7 * gscale:
                continue
            g = p._viai_grad.detach().cpu()
            rows.append((H.relerr(g, r64[gk][k]), H.relerr_l2(g, r64[gk][k]), H.relerr(r32[gk][k], r64[gk][k]), H.relerr_l2(r32[gk][k], r64[gk][k]), k))
            allg.append(g.flatten().double()); all32.append(r32[gk][k].flatten().double()); all64.append(r64[gk][k].flatten())
        rows.sort(reverse=True)
        A, B32, B64 = torch.cat(allg), torch.cat(all32), torch.cat(all64)
        cos = lambda a, b: float((a @ b) / (a.norm() * b.norm()))
        print("  %-9s whole-net L2 err: cuda %.3e  oracle32 %.3e | cosine cuda %.6f oracle32 %.6f" % (gk, float((A - B64).norm() / B64.norm()), float((B32 - B64).norm() / B64.norm()), cos(A, B64), cos(B32, B64)))
        for r in rows[:4]:
            print("      %-34s cuda max %.2e l2 %.2e | oracle32 max %.2e l2 %.2e" % (r[4], r[0], r[1], r[2], r[3]))
