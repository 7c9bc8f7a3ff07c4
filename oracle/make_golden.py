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


def next_rows_inputs():
    """Inputs of the next-row fixtures: a pure function of names, rebuilt by the tests."""
    yh = FX.normal("dmol_yhat", (2, 30, 40)) * 2.0
    yh[:, 20:, :8] -= 12.0                  # log scales under both clamps -> the bin-centre pdf branch
    yh[:, 20:, 8:16] -= 5.0
    y = FX.uniform("dmol_y", (2, 40, 1), -1.0, 1.0)
    y[0, :4] = -1.0                         # edge bins
    y[1, :4] = 1.0
    y[1, 4:6] = 0.9995
    return dict(dmol_yhat=yh, dmol_y=y, dmol_len=torch.tensor([40, 27]), dmol_u=FX.uniform("dmol_u", (2, 40, 11), 1e-5, 1.0 - 1e-5),
                f1=FX.normal("ctr_f1", (6, 32)), f2=FX.normal("ctr_f2", (6, 32)) * 0.5 + FX.normal("ctr_f1", (6, 32)) * 0.7,
                f3=FX.normal("ctr_f3", (6, 32)) + FX.normal("ctr_f1", (6, 32)) * 0.25,
                dis_mel=FX.uniform("dis_mel", (2, 1, 80, 64)), dis_fea=FX.normal("dis_fea", (2, 512, 16)),
                dom_x=FX.normal("dom_x", (3, 256, 13)),
                video=FX.normal("video", (1, 4, 3, 224, 224)).clamp(-1, 1), flow=FX.normal("flow", (1, 4, 2, 224, 224)).clamp(-1, 1),
                feat=FX.normal("ft_feat", (2, 8, 256)))


WAVENET_TRAIN_KW = dict(layers=8, stacks=2, residual_channels=32, gate_channels=32, skip_out_channels=16, cin_channels=8,
                        upsample_scales=[2, 4])


def next_rows_fixture(ref):
    import math
    I = next_rows_inputs()
    fx = {}
    LF, MX = ref.loss_functions, ref.mixture
    # --- DMoL loss: values + autograd gradients of the reference
    for tag, nc, lsm in (("256", 256, -7.0), ("65536", 65536, float(math.log(1e-14)))):
        yh = I["dmol_yhat"].clone().requires_grad_(True)
        nll = MX.discretized_mix_logistic_loss(yh, I["dmol_y"], num_classes=nc, log_scale_min=lsm, reduce=False)
        tot = MX.discretized_mix_logistic_loss(yh, I["dmol_y"], num_classes=nc, log_scale_min=lsm, reduce=True)
        g, = torch.autograd.grad(nll.sum(), yh)
        fx["dmol_" + tag] = dict(nll=nll.detach(), total=float(tot), grad=g)
    yh = I["dmol_yhat"].clone().requires_grad_(True)
    ml = LF.DiscretizedMixturelogisticLoss()(yh, I["dmol_y"], lengths=I["dmol_len"])
    g, = torch.autograd.grad(ml, yh)
    fx["dmol_masked"] = dict(loss=float(ml), grad=g, mask=LF.sequence_mask(I["dmol_len"]))
    # --- sampling with pinned uniforms (mixture.py:136 draws (B,T,nr_mix), :148 draws (B,T))
    calls = {"i": 0}
    orig = torch.Tensor.uniform_

    def fake_uniform(self, a=0.0, b=1.0, **k):
        i = calls["i"]
        calls["i"] += 1
        self.copy_(I["dmol_u"][..., :10] if i == 0 else I["dmol_u"][..., 10])
        return self
    torch.Tensor.uniform_ = fake_uniform
    try:
        fx["dmol_sample"] = MX.sample_from_discretized_mix_logistic(I["dmol_yhat"], log_scale_min=-7.0)
    finally:
        torch.Tensor.uniform_ = orig
    # --- EMA
    ema = LF.ExponentialMovingAverage(0.9)
    ema.register("w", I["f1"])
    ema.update("w", I["f2"])
    fx["ema"] = ema.shadow["w"].clone()
    # --- L2 contrastive loss / l2_norm / retrieval
    for tag, margin, mv in (("m0", 0, False), ("m8", 8.0, False), ("m8max", 8.0, True)):
        a, b = I["f1"].clone().requires_grad_(True), I["f2"].clone().requires_grad_(True)
        loss = LF.L2ContrastiveLoss(margin=margin, max_violation=mv)(a, b)
        ga, gb = torch.autograd.grad(loss, (a, b))
        fx["ctr_" + tag] = dict(loss=float(loss), g1=ga, g2=gb)
    fx["l2_sim"] = LF.l2_sim(I["f1"], I["f2"])
    try:
        import importlib
        U = importlib.import_module("utils.util")
        fx["l2_norm"] = U.l2_norm(I["f1"])
        fx["retrieval"] = tuple(float(v) for v in U.L2retrieval(I["f1"].numpy(), I["f2"].numpy()))
        fx["retrieval_noisy"] = tuple(float(v) for v in U.L2retrieval(I["f1"].numpy(), I["f3"].numpy()))
    except Exception as e:        # sklearn / cv2 ... missing: restated by torch
        print("utils.util not importable (%s); l2_norm / L2retrieval golden from their one-line definitions" % e)
        fx["l2_norm"] = torch.nn.functional.normalize(I["f1"], p=2, dim=1)
    # --- discriminator heads
    DN = ref.Discriminator_Networks
    D = _fill(DN.Inpainting_Dis())
    out = D(I["dis_mel"], I["dis_fea"])
    out.pow(2).sum().backward()
    fx["inpainting_dis"] = dict(out=out.detach(), shapes={k: tuple(v.shape) for k, v in D.state_dict().items()},
                                grads={k: p.grad.clone() for k, p in D.named_parameters() if p.numel() <= 8192},
                                gnorm={k: float(p.grad.norm()) for k, p in D.named_parameters()},
                                rm={k: v.clone() for k, v in D.state_dict().items() if "running_mean" in k})
    D = _fill(DN.DomainDis())
    out = D(I["dom_x"])
    out.sum().backward()
    fx["domain_dis"] = dict(out=out.detach(), shapes={k: tuple(v.shape) for k, v in D.state_dict().items()},
                            gnorm={k: float(p.grad.norm()) for k, p in D.named_parameters()})
    # --- visual branches
    if ref.Image_Embedding is not None:
        IE = ref.Image_Embedding
        M = _fill(IE.ImageEmbedding2())
        out, fea = M(I["video"], I["flow"])
        fx["ie2"] = dict(out=out.detach(), fea=fea.detach(), shapes={k: tuple(v.shape) for k, v in M.state_dict().items()})
        M = _fill(IE.ImageEmbedding_single(image=1))
        fx["ie_single"] = dict(out=M(I["video"]).detach(), bn_1_running_mean=M.state_dict()["bn_1.running_mean"].clone(),
                               shapes={k: tuple(v.shape) for k, v in M.state_dict().items()})
        M = _fill(IE.ImageEmbedding_finetune())
        fx["ie_finetune"] = dict(out=M(I["feat"]).detach(), shapes={k: tuple(v.shape) for k, v in M.state_dict().items()})
    # --- WaveNet teacher-forced training step (eval mode: dropout off), loss on the shifted targets, gradients
    kw = WAVENET_TRAIN_KW
    W = _fill(ref.wavenet.WaveNet(**kw)).eval()
    T = 48
    x = FX.uniform("wav_xsmall", (1, 1, T), -1.0, 1.0)
    c = FX.uniform("wav_csmall", (1, kw["cin_channels"], T // 8))
    x2 = torch.cat((x, FX.uniform("wav_x2", (1, 1, T), -1.0, 1.0)), 0)
    c2 = torch.cat((c, FX.uniform("wav_c2", (1, kw["cin_channels"], T // 8))), 0)
    lengths = torch.tensor([T - 1, T - 11])
    y_hat = W(x2, c2)
    loss = LF.DiscretizedMixturelogisticLoss()(y_hat[:, :, :-1], x2.transpose(1, 2)[:, 1:, :], lengths=lengths)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in W.named_parameters() if p.grad is not None}
    keep = ["first_conv.bias", "first_conv.weight_g", "conv_layers.0.conv.weight_v", "conv_layers.3.conv1x1c.weight_g",
            "conv_layers.6.conv1x1_out.weight_v", "conv_layers.5.conv1x1_skip.bias", "last_conv_layers.3.bias",
            "last_conv_layers.1.weight_v", "upsample_conv.0.weight_v", "upsample_conv.2.bias"]
    fx["wavenet_train"] = dict(T=T, loss=float(loss), y_hat=y_hat.detach(), lengths=lengths, grads={k: grads[k] for k in keep},
                               gnorm={k: float(v.norm()) for k, v in grads.items()},
                               dead=[k for k, p in W.named_parameters() if p.grad is None])
    torch.save(fx, os.path.join(OUT, "next_rows.pt"))
    print("next rows:", sorted(fx.keys()), "wavenet train loss %.6f" % fx["wavenet_train"]["loss"])


LOADER_HP = dict(sample_rate=16000, max_time_steps=1920, image_hope_size=1, load_num=2, image_rescal_size=16, image_size=12,
                 image=True, flow=True)


def loader_fixture():
    """tests/golden/loader_frames.pt: a small JPEG tree (two clips: 24x20 frames that are down-scaled, 10x13 frames that are
    up-scaled) and what the reference's own sample_data_new / load_image return on it for seeded np.random draws."""
    import tempfile
    import types
    import numpy as np
    from oracle import loader_oracle as LO
    hp = types.SimpleNamespace(**LOADER_HP)
    sample_ref, load_ref = LO.reference_functions(hp)
    trees = {"clipA": LO.synthetic_clip(1, 58, 24, 20), "clipB": LO.synthetic_clip(2, 56, 10, 13)}
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, tree in trees.items():
            LO.write_tree(os.path.join(tmp, name), tree)
        for name in trees:
            for train in (True, False):
                for seed in (0, 1, 2):
                    np.random.seed(seed)
                    v, f, start = sample_ref(os.path.join(tmp, name), train, hparams=hp)
                    cases.append(dict(kind="sample", clip=name, train=train, seed=seed, start=[int(s) for s in start],
                                      video=torch.from_numpy(np.ascontiguousarray(v)).float(),
                                      flow=torch.from_numpy(np.ascontiguousarray(f)).float()))
        hp1 = types.SimpleNamespace(**dict(LOADER_HP, load_num=1))
        _, load_ref1 = LO.reference_functions(hp1)
        # (train=False is unusable upstream: the block is allocated with load_num rows, audio_loader.py:266-269 -> IndexError)
        for train in (True,):
            np.random.seed(5)
            v, f = load_ref1(os.path.join(tmp, "clipB"), train, hparams=hp1)
            cases.append(dict(kind="load", clip="clipB", train=train, seed=5,
                              video=torch.from_numpy(np.ascontiguousarray(v[:6])).float(),
                              flow=torch.from_numpy(np.ascontiguousarray(f[:6])).float(), n=int(v.shape[0])))
    torch.save(dict(hp=LOADER_HP, trees=trees, cases=cases), os.path.join(OUT, "loader_frames.pt"))
    print("loader frames:", len(cases), "cases,", sum(len(b) for t in trees.values() for b in t.values()), "JPEG bytes")


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
    next_rows_fixture(ref)
    loader_fixture()


if __name__ == "__main__":
    main()
