"""Generates tests/golden/*.pt by running the UNMODIFIED reference classes (build container only).

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  Each fixture stores the outputs of the reference itself for weights/inputs that are
a pure function of names and shapes (oracle/fixtures.py), so the tests can rebuild the inputs anywhere.
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures as FX  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _fill(m, salt=0):
    sd = m.state_dict()
    FX.deterministic_fill(sd, salt)
    m.load_state_dict(sd)
    return m


def gan_fixture(ref, norm, B, H, W, tag):
    nl = nn.BatchNorm2d if norm == "bn" else nn.InstanceNorm2d
    ref.Options_inpainting.Inpainting_Config.cin_channels = H
    E = _fill(ref.Inpainting_Networks.MelEncoder(norm_layer=nl))
    G = _fill(ref.New_Inpainting_Networks.MelDecoder(norm_layer=nl))
    D = _fill(ref.Discriminator_Networks.MelDiscriminator(norm_layer=nl))
    E.hparams.cin_channels = H
    gl = ref.loss_functions.GANLoss(use_lsgan=True, device=torch.device("cpu"))
    mel = FX.uniform("mel%s" % tag, (B, 1, H, W))
    mask = torch.ones_like(mel)
    mask[..., W // 4:W // 4 + W // 2] = 0
    feats = E(mel * mask)
    fake = G(feats, mel.shape)
    # D phase
    pred_fake_d = D(fake.detach())
    pred_real = D(mel)
    loss_D = 0.5 * (gl(pred_fake_d, False) + gl(pred_real, True))
    loss_D.backward()
    gD = {k: p.grad.clone() for k, p in D.named_parameters()}
    D.zero_grad()
    for p in D.parameters():
        p.requires_grad_(False)
    pred_fake_g = D(fake)
    loss_gan = gl(pred_fake_g, True)
    loss_l1 = nn.L1Loss()(fake, mel)
    (loss_gan + 100.0 * loss_l1).backward()
    gE = {k: p.grad.clone() for k, p in E.named_parameters()}
    gG = {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None}
    small = lambda d: {k: v for k, v in d.items() if v.numel() <= 4608}
    norms = lambda d: {k: float(v.norm()) for k, v in d.items()}
    fx = dict(norm=norm, B=B, H=H, W=W, tag=tag,
              feats=[f.detach() if f.numel() < 70000 else None for f in feats],
              feat_sums=[float(f.double().sum()) for f in feats], feat_abs=[float(f.double().abs().sum()) for f in feats],
              fake=fake.detach(), pred_fake_d=pred_fake_d.detach(), pred_real=pred_real.detach(),
              pred_fake_g=pred_fake_g.detach(), loss_D=float(loss_D), loss_G_GAN=float(loss_gan), loss_L1=float(loss_l1),
              grad_E_small=small(gE), grad_G_small=small(gG), grad_D_small=small(gD),
              grad_E_norm=norms(gE), grad_G_norm=norms(gG), grad_D_norm=norms(gD),
              dead=[k for k, p in G.named_parameters() if p.grad is None])
    if norm == "bn":
        fx["running"] = {"E." + k: v.clone() for k, v in E.state_dict().items() if "running" in k and v.numel() <= 64}
        fx["running"].update({"D." + k: v.clone() for k, v in D.state_dict().items() if "running" in k and v.numel() <= 64})
        fx["nbt_D"] = int(D.state_dict()["bn1.num_batches_tracked"])
    torch.save(fx, os.path.join(OUT, "gan_%s_%s.pt" % (norm, tag)))
    print("gan", norm, tag, "loss_D %.6f loss_GAN %.6f L1 %.6f" % (fx["loss_D"], fx["loss_G_GAN"], fx["loss_L1"]))


def decoder_image_fixture(ref):
    ref.Options_inpainting.Inpainting_Config.cin_channels = 80
    E = _fill(ref.Inpainting_Networks.MelEncoder())
    E.hparams.cin_channels = 80
    mel = FX.uniform("melimg", (2, 1, 80, 64))
    feats = E(mel)
    video = FX.normal("video_net", (2, 256, 1, 4))
    out = {}
    for variant in ("MelDecoderImage", "MelDecoderImage2", "MelDecoder_old"):
        G = _fill(getattr(ref.New_Inpainting_Networks, variant)())
        args = (feats, mel.shape) + ((video,) if "Image" in variant else ())
        out[variant] = G(*args).detach()
    torch.save(out, os.path.join(OUT, "decoder_variants.pt"))
    print("decoder variants ok")


def wavenet_fixture(ref, tag, T, **kw):
    W = _fill(ref.wavenet.WaveNet(**kw)).eval()
    cin = kw.get("cin_channels", 80)
    hop = 1
    for s in kw["upsample_scales"]:
        hop *= s
    x = FX.uniform("wav_x" + tag, (1, 1, T), -1.0, 1.0)
    c = FX.uniform("wav_c" + tag, (1, cin, T // hop))
    with torch.no_grad():
        yb = W(x, c)
    nr_mix = kw.get("out_channels", 30) // 3
    u = FX.uniform("wav_u" + tag, (T, 1, nr_mix + 1), 1e-5, 1.0 - 1e-5)
    # drive the reference sampler with the stored uniforms by monkey-patching Tensor.uniform_ order:
    # mixture.py:136 draws (B,T=1,nr_mix) then :148 draws (B,T=1)
    calls = {"i": 0}
    orig = torch.Tensor.uniform_

    def fake_uniform(self, a=0.0, b=1.0, **k):
        t, which = divmod(calls["i"], 2)
        calls["i"] += 1
        if which == 0:
            self.copy_(u[t, :, :nr_mix].view_as(self))
        else:
            self.copy_(u[t, :, nr_mix].view_as(self))
        return self
    torch.Tensor.uniform_ = fake_uniform
    try:
        with torch.no_grad():
            out = W.incremental_forward(c=c, T=T, log_scale_min=-7.0)
    finally:
        torch.Tensor.uniform_ = orig
    torch.save(dict(kw=kw, T=T, tag=tag, logits=yb, samples=out), os.path.join(OUT, "wavenet_%s.pt" % tag))
    print("wavenet", tag, tuple(yb.shape), float(out.abs().mean()))


def image_embedding_fixture(ref):
    if ref.Image_Embedding is None:
        print("skip image embedding:", ref.Image_Embedding_error)
        return
    M = _fill(ref.Image_Embedding.ImageEmbedding())
    v = FX.normal("video", (1, 4, 3, 224, 224)).clamp(-1, 1)
    f = FX.normal("flow", (1, 4, 2, 224, 224)).clamp(-1, 1)
    out = M(v, f)
    sd = M.state_dict()
    torch.save(dict(out=out.detach(), bn_1_running_mean=sd["bn_1.running_mean"].clone(),
                    img_bn1_running_mean=sd["image_single_model.bn1.running_mean"].clone()),
               os.path.join(OUT, "image_embedding.pt"))
    print("image embedding", tuple(out.shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load(normlayer=nn.BatchNorm2d, cin_channels=80)
    torch.manual_seed(0)
    gan_fixture(ref, "bn", 1, 80, 64, "c1")
    gan_fixture(ref, "in", 1, 80, 64, "c1")
    gan_fixture(ref, "bn", 2, 128, 128, "s128")
    ref.Options_inpainting.Inpainting_Config.normlayer = nn.BatchNorm2d
    decoder_image_fixture(ref)
    wavenet_fixture(ref, "small", 48, layers=8, stacks=2, residual_channels=32, gate_channels=32,
                    skip_out_channels=16, cin_channels=8, upsample_scales=[2, 4])
    wavenet_fixture(ref, "full", 320, layers=24, stacks=4, residual_channels=512, gate_channels=512,
                    skip_out_channels=256, cin_channels=80, upsample_scales=[4, 4, 10])
    image_embedding_fixture(ref)


if __name__ == "__main__":
    main()
