"""Diagnostic (GPU box): per-parameter gradient error of D and of E+G against the fp64 oracle, next to the fp32 oracle's."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn as nn
import viai_test_helpers as H
from oracle import viai_oracle as O, fixtures as FX
from viai_b200 import Options_inpainting, ops
from viai_b200.networks import Discriminator_Networks as DN, Inpainting_Networks as IN, New_Inpainting_Networks as NN
from viai_b200.loss_functions import GANLoss

def leaf(sd, dt):
    return {k: (v.clone().to(dt).requires_grad_(True) if (v.is_floating_point() and "running" not in k) else (v.clone().to(dt) if v.is_floating_point() else v.clone())) for k, v in sd.items()}

for norm in ("in", "bn"):
    nl = nn.BatchNorm2d if norm == "bn" else nn.InstanceNorm2d
    B, Hh, W = 1, 80, 64
    x = FX.uniform("diagx", (B, 1, Hh, W))
    dsd = H.filled(H.discriminator_sd(norm))
    res = {}
    for dt in (torch.float32, torch.float64):
        d = leaf(dsd, dt)
        p = O.mel_discriminator_forward(d, x.to(dt), norm)
        O.gan_loss(p, True).backward()
        res[dt] = (p.detach(), {k: v.grad for k, v in d.items() if v.requires_grad})
    D = DN.MelDiscriminator(norm_layer=nl); D.load_state_dict(dsd); D.cuda()
    pg = D(x.cuda())
    GANLoss(True).cuda()(pg, True).backward()
    print("== D", norm, "pred err", H.relerr(pg, res[torch.float64][0]), "oracle32", H.relerr(res[torch.float32][0], res[torch.float64][0]))
    for k, p in D.named_parameters():
        r64 = res[torch.float64][1][k]; r32 = res[torch.float32][1][k]
        print("   %-18s cuda %.3e   oracle32 %.3e   |g| %.3e" % (k, H.relerr(p.grad, r64), H.relerr(r32, r64), float(r64.abs().max())))

print("==== generator forward, per stage, vs fp64 oracle")
for norm in ("bn", "in"):
    nl = nn.BatchNorm2d if norm == "bn" else nn.InstanceNorm2d
    B, Hh, W = 1, 80, 64
    mel = FX.uniform("melc1", (B, 1, Hh, W))
    mask = H.center_mask(mel.shape)
    esd, gsd = H.filled(H.encoder_sd(norm)), H.filled(H.decoder_sd(norm))
    hp = Options_inpainting.Inpainting_Config(cin_channels=Hh, normlayer=nl)
    E = IN.MelEncoder(hp, norm_layer=nl); E.load_state_dict(esd); E.cuda()
    G = NN.MelDecoder(hp, norm_layer=nl); G.load_state_dict(gsd); G.cuda()
    out = {}
    for dt in (torch.float32, torch.float64):
        e, g = H.to_dtype(esd, dt), H.to_dtype(gsd, dt)
        f = O.mel_encoder_forward(e, (mel * mask).to(dt), Hh, norm)
        out[dt] = (f, O.mel_decoder_forward(g, f, mel.shape, norm))
    with torch.no_grad():
        fg = E((mel * mask).cuda())
        fk = G(fg, mel.shape)
    for i in range(5):
        print(norm, "feat%d cuda %.3e oracle32 %.3e" % (i, H.relerr(fg[i], out[torch.float64][0][i]), H.relerr(out[torch.float32][0][i], out[torch.float64][0][i])))
    print(norm, "fake  cuda %.3e oracle32 %.3e" % (H.relerr(fk, out[torch.float64][1]), H.relerr(out[torch.float32][1], out[torch.float64][1])))
    # decoder stage by stage from the oracle's fp64 features (isolates the decoder)
    with torch.no_grad():
        f64 = [t.float().cuda() for t in out[torch.float64][0]]
        fk2 = G(f64, mel.shape)
    print(norm, "fake from exact feats cuda %.3e" % H.relerr(fk2, out[torch.float64][1]))

print("==== D phase bisect (IN): single passes vs accumulated two passes")
norm = "in"; nl = nn.InstanceNorm2d
B, Hh, W = 1, 80, 64
mel = FX.uniform("melc1", (B, 1, Hh, W)); mask = H.center_mask(mel.shape)
esd, gsd, dsd = H.filled(H.encoder_sd(norm)), H.filled(H.decoder_sd(norm)), H.filled(H.discriminator_sd(norm))
f = O.mel_encoder_forward(dict(esd), mel * mask, Hh, norm)
fake = O.mel_decoder_forward(dict(gsd), f, mel.shape, norm).detach()
def oracle_d(inputs, dt):
    d = leaf(dsd, dt)
    loss = 0
    for x, real in inputs:
        loss = loss + 0.5 * O.gan_loss(O.mel_discriminator_forward(d, x.to(dt), norm), real)
    loss.backward()
    return {k: v.grad for k, v in d.items() if v.requires_grad}
def cuda_d(inputs):
    D = DN.MelDiscriminator(norm_layer=nl); D.load_state_dict(dsd); D.cuda()
    gl = GANLoss(True).cuda()
    losses = [gl(D(x.cuda()), real) for x, real in inputs]
    loss = ops.lincomb2(losses[0], 0.5, losses[1], 0.5) if len(losses) == 2 else ops.lincomb2(losses[0], 0.5)
    loss.backward()
    return {k: p.grad for k, p in D.named_parameters()}
for name, inputs in (("fake/False", [(fake, False)]), ("real/True", [(mel, True)]), ("both", [(fake, False), (mel, True)])):
    r64, r32, gc = oracle_d(inputs, torch.float64), oracle_d(inputs, torch.float32), cuda_d(inputs)
    for k in ("conv1.weight", "conv2_1.weight", "conv3.weight", "conv4.weight"):
        print("  %-10s %-16s cuda %.3e oracle32 %.3e" % (name, k, H.relerr(gc[k], r64[k]), H.relerr(r32[k], r64[k])))

print("==== full D phase with GPU generator (IN)")
hp = Options_inpainting.Inpainting_Config(cin_channels=Hh, normlayer=nl)
E = IN.MelEncoder(hp, norm_layer=nl); E.load_state_dict(esd); E.cuda()
G = NN.MelDecoder(hp, norm_layer=nl); G.load_state_dict(gsd); G.cuda()
D = DN.MelDiscriminator(norm_layer=nl); D.load_state_dict(dsd); D.cuda()
gl = GANLoss(True).cuda()
melg = mel.cuda()
masked = ops.mul(melg, mask.cuda())
fake_g = G(E(masked), mel.shape)
print("fake gpu vs oracle32", H.relerr(fake_g, fake))
pf = D(fake_g.detach()); pr = D(melg)
loss = ops.lincomb2(gl(pf, False), 0.5, gl(pr, True), 0.5)
loss.backward()
gc = {k: p.grad.clone() for k, p in D.named_parameters()}
inputs_o = [(fake_g.detach().cpu(), False), (mel, True)]
r64o = oracle_d(inputs_o, torch.float64)          # oracle D on the GPU's fake
inputs_p = [(fake, False), (mel, True)]
r64p = oracle_d(inputs_p, torch.float64)          # oracle D on the oracle's fake
for k in ("conv1.weight", "conv2_1.weight", "conv3.weight", "conv4.weight"):
    print("  %-16s cuda-vs-oracle(same fake) %.3e   oracle(gpu fake)-vs-oracle(cpu fake) %.3e" % (k, H.relerr(gc[k], r64o[k]), H.relerr(r64o[k], r64p[k])))
fc = fake_g.detach().cpu()
print("fake_g.cpu strides", fc.stride(), fc.is_contiguous(), "max abs diff", float((fc - fake).abs().max()), "mean abs diff", float((fc - fake).abs().mean()))
for k in ("conv1.weight", "conv3.weight"):
    print("  %-16s cuda-vs-oracle(cpu fake) %.3e" % (k, H.relerr(gc[k], r64p[k])))
r64o2 = oracle_d([(fc.contiguous().clone(), False), (mel, True)], torch.float64)
r32o2 = oracle_d([(fc.contiguous().clone(), False), (mel, True)], torch.float32)
pert = fake + 1e-5 * torch.randn_like(fake)
r64q = oracle_d([(pert, False), (mel, True)], torch.float64)
for k in ("conv1.weight", "conv3.weight"):
    print("  %-16s oracle64(gpu fake, contiguous clone) vs oracle64(cpu fake) %.3e ; oracle32(gpu fake) vs same %.3e; oracle64(cpu fake + 1e-5 noise) vs oracle64(cpu fake) %.3e" % (
        k, H.relerr(r64o2[k], r64p[k]), H.relerr(r32o2[k], r64p[k]), H.relerr(r64q[k], r64p[k])))
