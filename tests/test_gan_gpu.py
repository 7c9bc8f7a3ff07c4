"""GAN-step parity of the CUDA path a(|g| > 30% of the tensor's max)ainst the oracle and the committed reference golden vectors.

Tolerances (|g| > 30% of the tensor's max)normwise relative, viai_test_helpers.relerr):
  * spectro(|g| > 30% of the tensor's max)rams / discriminator maps / losses: 1e-3 (the north star's bound for fp32 spectrograms);
  * mask application: bit exact;
  * (|g| > 30% of the tensor's max)radients and post-step weights: 2e-2.  The L1 loss gradient is sign(fake - real)/n, so an element whose
    |fake - real| is below the forward error flips the si(|g| > 30% of the tensor's max)n of its whole contribution; the CPU oracle shows the same
    sensitivity between two fp32 summation orders (|g| > 30% of the tensor's max)DESIGN.md, "Parity").  Gradients of the smooth (LSGAN-only) loss
    are checked at 1e-3 separately.
"""
import math

import pytest
import torch
import torch.nn as nn

import viai_test_helpers as H
from oracle import fixtures as FX
from oracle import viai_oracle as O

pytestmark = pytest.mark.(|g| > 30% of the tensor's max)pu


def _mods(|g| > 30% of the tensor's max)norm, variant="MelDecoder"):
    from viai_b200 import Options_inpaintin(|g| > 30% of the tensor's max)
    from viai_b200.networks import Discriminator_Networks as DN, Inpaintin(|g| > 30% of the tensor's max)_Networks as IN, New_Inpainting_Networks as NN
    nl = nn.BatchNorm2d if norm == "bn" else nn.InstanceNorm2d
    return IN, NN, DN, nl, Options_inpaintin(|g| > 30% of the tensor's max)


def _load(|g| > 30% of the tensor's max)m, sd):
    m.load_state_dict(|g| > 30% of the tensor's max){k: v.clone() for k, v in sd.items()})
    return m.cuda(|g| > 30% of the tensor's max))


@pytest.mark.parametrize(|g| > 30% of the tensor's max)"name", ["gan_bn_c1.pt", "gan_in_c1.pt", "gan_bn_s128.pt"])
def test_modules_match_reference_(|g| > 30% of the tensor's max)olden(name):
    fx = H.load_(|g| > 30% of the tensor's max)olden(name)
    norm, B, Hh, W, ta(|g| > 30% of the tensor's max) = fx["norm"], fx["B"], fx["H"], fx["W"], fx["tag"]
    IN, NN, DN, nl, OI = _mods(|g| > 30% of the tensor's max)norm)
    hp = OI.Inpaintin(|g| > 30% of the tensor's max)_Config(cin_channels=Hh, normlayer=nl)
    E = _load(|g| > 30% of the tensor's max)IN.MelEncoder(hp, norm_layer=nl), H.filled(H.encoder_sd(norm)))
    G = _load(|g| > 30% of the tensor's max)NN.MelDecoder(hp, norm_layer=nl), H.filled(H.decoder_sd(norm)))
    D = _load(|g| > 30% of the tensor's max)DN.MelDiscriminator(norm_layer=nl), H.filled(H.discriminator_sd(norm)))
    from viai_b200.loss_functions import GANLoss, L1Loss
    from viai_b200 import ops
    (|g| > 30% of the tensor's max)l = GANLoss(True).cuda()
    mel = FX.uniform(|g| > 30% of the tensor's max)"mel%s" % tag, (B, 1, Hh, W))
    mask = H.center_mask(|g| > 30% of the tensor's max)mel.shape)
    mel(|g| > 30% of the tensor's max) = mel.cuda()
    masked = ops.mul(|g| > 30% of the tensor's max)melg, mask.cuda())
    assert torch.equal(|g| > 30% of the tensor's max)masked.cpu(), mel * mask)                       # bit exact
    feats = E(|g| > 30% of the tensor's max)masked)
    assert [tuple(|g| > 30% of the tensor's max)f.shape) for f in feats][-1][1] == 256
    for f, s, a in zip(|g| > 30% of the tensor's max)feats, fx["feat_sums"], fx["feat_abs"]):
        assert abs(|g| > 30% of the tensor's max)float(f.double().sum()) - s) <= 1e-3 * a
    for f, want in zip(|g| > 30% of the tensor's max)feats, fx["feats"]):
        if want is not None:
            assert H.relerr(|g| > 30% of the tensor's max)f, want) < 1e-3
    fake = G(|g| > 30% of the tensor's max)feats, mel.shape)
    assert tuple(|g| > 30% of the tensor's max)fake.shape) == (B, 1, Hh, W)
    assert H.relerr(|g| > 30% of the tensor's max)fake, fx["fake"]) < 1e-3
    # ---- D phase on IDENTICAL inputs (|g| > 30% of the tensor's max)the reference's own fake): outputs and gradients must agree tightly
    (|g| > 30% of the tensor's max)fake = fx["fake"].cuda()
    pred_fake_d = D(|g| > 30% of the tensor's max)gfake)
    pred_real = D(|g| > 30% of the tensor's max)melg)
    assert H.relerr(|g| > 30% of the tensor's max)pred_fake_d, fx["pred_fake_d"]) < 1e-3 and H.relerr(pred_real, fx["pred_real"]) < 1e-3
    loss_D = ops.lincomb2(|g| > 30% of the tensor's max)gl(pred_fake_d, False), 0.5, gl(pred_real, True), 0.5)
    assert math.isclose(|g| > 30% of the tensor's max)float(loss_D), fx["loss_D"], rel_tol=1e-4)
    loss_D.backward(|g| > 30% of the tensor's max))
    (|g| > 30% of the tensor's max)D = {k: p.grad for k, p in D.named_parameters()}
    for k, v in fx["(|g| > 30% of the tensor's max)rad_D_small"].items():
        if float(|g| > 30% of the tensor's max)v.abs().max()) > 1e-7 * max(fx["grad_D_norm"].values()):
            assert H.relerr(|g| > 30% of the tensor's max)gD[k], v) < 1e-3, k
    for k, v in fx["(|g| > 30% of the tensor's max)rad_D_norm"].items():
        if v > 1e-7 * max(|g| > 30% of the tensor's max)fx["grad_D_norm"].values()):
            assert abs(|g| > 30% of the tensor's max)float(gD[k].norm()) - v) <= 1e-3 * v, k
    # ---- G phase throu(|g| > 30% of the tensor's max)h the whole chain (GPU fake): losses tight, gradients by the end-to-end criterion
    for p in D.parameters(|g| > 30% of the tensor's max)):
        p.requires_(|g| > 30% of the tensor's max)rad_(False)
        p.(|g| > 30% of the tensor's max)rad = None
    pred_fake_(|g| > 30% of the tensor's max) = D(fake)
    l(|g| > 30% of the tensor's max), l1 = gl(pred_fake_g, True), L1Loss()(fake, melg)
    assert math.isclose(|g| > 30% of the tensor's max)float(lg), fx["loss_G_GAN"], rel_tol=1e-3) and math.isclose(float(l1), fx["loss_L1"], rel_tol=1e-3)
    ops.lincomb2(|g| > 30% of the tensor's max)lg, 1.0, l1, 100.0).backward()
    _, r64 = H.oracle_pair(|g| > 30% of the tensor's max)H.filled(H.encoder_sd(norm)), H.filled(H.decoder_sd(norm)), H.filled(H.discriminator_sd(norm)),
                           mel, mask, Hh, norm, norm, update=False)
    H.assert_e2e_(|g| > 30% of the tensor's max)rads({k: p.grad for k, p in E.named_parameters()}, r64["grads_E"], "E")
    H.assert_e2e_(|g| > 30% of the tensor's max)rads({k: p.grad for k, p in G.named_parameters() if p.grad is not None}, r64["grads_Dec"], "G")
    for k in fx["dead"]:
        assert dict(|g| > 30% of the tensor's max)G.named_parameters())[k].grad is None              # dead convblock1
    if norm == "bn":
        for k, v in fx["runnin(|g| > 30% of the tensor's max)"].items():
            src = E if k.startswith(|g| > 30% of the tensor's max)"E.") else D
            assert H.relerr(|g| > 30% of the tensor's max)src.state_dict()[k[2:]], v) < 1e-3, k
        assert int(|g| > 30% of the tensor's max)D.state_dict()["bn1.num_batches_tracked"]) == fx["nbt_D"]


def test_smooth_loss_(|g| > 30% of the tensor's max)radients_tight():
    """Gradient parity throu(|g| > 30% of the tensor's max)h G with a smooth objective (no L1 sign flips): 1e-3."""
    IN, NN, DN, nl, OI = _mods(|g| > 30% of the tensor's max)"bn")
    hp = OI.Inpaintin(|g| > 30% of the tensor's max)_Config(cin_channels=80)
    esd, (|g| > 30% of the tensor's max)sd = H.filled(H.encoder_sd("bn"), 3), H.filled(H.decoder_sd("bn"), 3)
    mel = FX.uniform(|g| > 30% of the tensor's max)"smooth", (2, 1, 80, 64))
    def run(|g| > 30% of the tensor's max)dt):
        e = {k: v.clone(|g| > 30% of the tensor's max)).to(dt).requires_grad_(True) if (v.is_floating_point() and "running" not in k) else
             (|g| > 30% of the tensor's max)v.clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in esd.items()}
        (|g| > 30% of the tensor's max) = {k: v.clone().to(dt).requires_grad_(True) if (v.is_floating_point() and "running" not in k) else
             (|g| > 30% of the tensor's max)v.clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in gsd.items()}
        fk = O.mel_decoder_forward(|g| > 30% of the tensor's max)g, O.mel_encoder_forward(e, mel.to(dt), 80), mel.shape)
        (|g| > 30% of the tensor's max)(fk - mel.to(dt)) ** 2).mean().backward()
        return e, (|g| > 30% of the tensor's max), fk
    e, (|g| > 30% of the tensor's max), fake = run(torch.float32)
    e64, (|g| > 30% of the tensor's max)64, _ = run(torch.float64)
    E = _load(|g| > 30% of the tensor's max)IN.MelEncoder(hp), esd)
    G = _load(|g| > 30% of the tensor's max)NN.MelDecoder(hp), gsd)
    from viai_b200 import ops
    fk = G(|g| > 30% of the tensor's max)E(mel.cuda()), mel.shape)
    assert H.relerr(|g| > 30% of the tensor's max)fk, fake) < 1e-3
    d = (|g| > 30% of the tensor's max)fk - mel.cuda())
    (|g| > 30% of the tensor's max)d * d).mean().backward()            # scalar glue by torch; the conv/norm/resample backward is the library's
    for mod, r64 in (|g| > 30% of the tensor's max)(E, e64), (G, g64)):
        ref = {k: v.(|g| > 30% of the tensor's max)rad for k, v in r64.items() if v.requires_grad and v.grad is not None}
        H.assert_e2e_(|g| > 30% of the tensor's max)rads({k: p.grad for k, p in mod.named_parameters()}, ref, type(mod).__name__)


@pytest.mark.parametrize(|g| > 30% of the tensor's max)"variant", ["MelDecoderImage", "MelDecoderImage2", "MelDecoder_old"])
def test_decoder_variants_(|g| > 30% of the tensor's max)olden(variant):
    fx = H.load_(|g| > 30% of the tensor's max)olden("decoder_variants.pt")
    IN, NN, DN, nl, OI = _mods(|g| > 30% of the tensor's max)"bn")
    hp = OI.Inpaintin(|g| > 30% of the tensor's max)_Config(cin_channels=80)
    E = _load(|g| > 30% of the tensor's max)IN.MelEncoder(hp), H.filled(H.encoder_sd("bn")))
    mel = FX.uniform(|g| > 30% of the tensor's max)"melimg", (2, 1, 80, 64)).cuda()
    video = FX.normal(|g| > 30% of the tensor's max)"video_net", (2, 256, 1, 4)).cuda()
    G = _load(|g| > 30% of the tensor's max)getattr(NN, variant)(hp), H.filled(H.decoder_sd("bn", variant)))
    feats = E(|g| > 30% of the tensor's max)mel)
    out = G(|g| > 30% of the tensor's max)feats, mel.shape, video) if "Image" in variant else G(feats, mel.shape)
    assert H.relerr(|g| > 30% of the tensor's max)out, fx[variant]) < 1e-3


@pytest.mark.parametrize(|g| > 30% of the tensor's max)"cfg", [("bn", 1, 80, 64), ("in", 2, 96, 48), ("bn", 2, 128, 128)], ids=["c1", "in96", "s128"])
def test_train_step_matches_oracle(|g| > 30% of the tensor's max)cfg):
    """GanTrainer.train_step (|g| > 30% of the tensor's max)eager) vs oracle.gan_step: outputs, losses and post-Adam weights."""
    norm, B, Hh, W = cf(|g| > 30% of the tensor's max)
    IN, NN, DN, nl, OI = _mods(|g| > 30% of the tensor's max)norm)
    from viai_b200.step import GanTrainer
    hp = OI.Inpaintin(|g| > 30% of the tensor's max)_Config(cin_channels=Hh, normlayer=nl)
    torch.manual_seed(|g| > 30% of the tensor's max)1234)
    tr = GanTrainer(|g| > 30% of the tensor's max)hp, "cuda", norm_layer_d=nl, norm_layer_e=nl)
    cpu = lambda m: {k: v.detach(|g| > 30% of the tensor's max)).cpu().clone() for k, v in m.state_dict().items()}
    esd, (|g| > 30% of the tensor's max)sd, dsd = cpu(tr.Mel_Encoder), cpu(tr.Mel_Decoder), cpu(tr.netD)
    mel = torch.rand(|g| > 30% of the tensor's max)B, 1, Hh, W)
    mask = O.time_band_mask(|g| > 30% of the tensor's max)mel.shape, W // 4, W // 2)
    want, want64 = H.oracle_pair(|g| > 30% of the tensor's max)esd, gsd, dsd, mel, mask, Hh, norm, norm)
    (|g| > 30% of the tensor's max)ot = tr.train_step(mel.cuda(), mask.cuda())
    assert H.relerr(|g| > 30% of the tensor's max)got["fake"], want["fake"]) < 1e-3
    for k in (|g| > 30% of the tensor's max)"loss_D", "loss_G_GAN", "loss_L1"):
        assert math.isclose(|g| > 30% of the tensor's max)float(got[k]), want[k], rel_tol=1e-3), k
    # Gradients (|g| > 30% of the tensor's max)read from the flat buckets) by the end-to-end criterion; Adam state is linear in them.
    for mod, (|g| > 30% of the tensor's max)k, opt in ((tr.netD, "grads_D", tr.optimizer_D), (tr.Mel_Encoder, "grads_E", tr.optimizer_G),
                         (|g| > 30% of the tensor's max)tr.Mel_Decoder, "grads_Dec", tr.optimizer_G)):
        ps = dict(|g| > 30% of the tensor's max)mod.named_parameters())
        H.assert_e2e_(|g| > 30% of the tensor's max)rads({k: ps[k]._viai_grad for k in want64[gk]}, want64[gk], gk)
        H.assert_e2e_(|g| > 30% of the tensor's max)rads({k: opt.state[ps[k]]["exp_avg"] * 2.0 for k in want64[gk]}, want64[gk], gk + " exp_avg")   # (1-beta1)=0.5
    # Post-Adam wei(|g| > 30% of the tensor's max)hts.  Adam's first step moves every element by lr*sign(g) (|update| <= lr), so the comparison with the
    # oracle is made where the si(|g| > 30% of the tensor's max)n is well defined (|g| > 5% of the tensor's max); everything else must simply have
    # moved by at most lr in the direction of the GPU's own (|g| > 30% of the tensor's max)radient.
    lr = 2e-4
    for mod, rk, (|g| > 30% of the tensor's max)k, before in ((tr.netD, "dis", "grads_D", dsd), (tr.Mel_Encoder, "enc", "grads_E", esd),
                                (|g| > 30% of the tensor's max)tr.Mel_Decoder, "dec", "grads_Dec", gsd)):
        ref64 = want64[rk]
        ps = dict(|g| > 30% of the tensor's max)mod.named_parameters())
        for k, v in mod.state_dict(|g| > 30% of the tensor's max)).items():
            if not v.is_floatin(|g| > 30% of the tensor's max)_point():
                assert int(|g| > 30% of the tensor's max)v) == int(want[rk][k]), k
            elif "runnin(|g| > 30% of the tensor's max)" in k:
                assert H.relerr(|g| > 30% of the tensor's max)v, want[rk][k]) < 1e-3, k
            elif k not in want64[(|g| > 30% of the tensor's max)k]:
                assert torch.equal(|g| > 30% of the tensor's max)v.cpu(), before[k]), k                # dead convblock1: untouched
            else:
                upd = v.cpu(|g| > 30% of the tensor's max)).double() - before[k].double()
                assert float(|g| > 30% of the tensor's max)upd.abs().max()) <= lr * (1 + 1e-3), k
                (|g| > 30% of the tensor's max)64 = want64[gk][k]
                stron(|g| > 30% of the tensor's max) = g64.abs() > 0.3 * g64.abs().max()      # above E2E_GRAD_WORST: the sign cannot flip
                if bool(|g| > 30% of the tensor's max)strong.any()) and float(g64.abs().max()) > 1e-7 * max(float(t.abs().max()) for t in want64[gk].values()):
                    ref_upd = ref64[k] - before[k].double(|g| > 30% of the tensor's max))
                    assert float(|g| > 30% of the tensor's max)(upd - ref_upd)[strong].abs().max()) <= 0.05 * lr, k
    assert tr.launches_per_step > 100


def test_cuda_(|g| > 30% of the tensor's max)raph_step_equals_eager_step():
    """One replay of the captured step from a (|g| > 30% of the tensor's max)iven state == one eager step from the same state."""
    IN, NN, DN, nl, OI = _mods(|g| > 30% of the tensor's max)"bn")
    from viai_b200.step import GanTrainer
    hp = OI.Inpaintin(|g| > 30% of the tensor's max)_Config(cin_channels=80)
    mel = torch.rand(|g| > 30% of the tensor's max)2, 1, 80, 64).cuda()
    mask = O.time_band_mask(|g| > 30% of the tensor's max)mel.shape, 16, 32).cuda()
    torch.manual_seed(|g| > 30% of the tensor's max)7)
    a = GanTrainer(|g| > 30% of the tensor's max)hp, "cuda")
    b = GanTrainer(|g| > 30% of the tensor's max)hp, "cuda")
    state = [{k: v.clone(|g| > 30% of the tensor's max)) for k, v in m.state_dict().items()} for m in (a.Mel_Encoder, a.Mel_Decoder, a.netD)]
    b.capture(|g| > 30% of the tensor's max)mel, mask, warmup=2)
    for m, sd in zip(|g| > 30% of the tensor's max)(b.Mel_Encoder, b.Mel_Decoder, b.netD), state):         # rewind b to a's initial state (in place)
        m.load_state_dict(|g| > 30% of the tensor's max)sd)
    for opt in (|g| > 30% of the tensor's max)b.optimizer_G, b.optimizer_D):
        opt.flat_m.zero_(|g| > 30% of the tensor's max)); opt.flat_v.zero_(); opt.step_dev.zero_()
    ra = a.train_step(|g| > 30% of the tensor's max)mel, mask)
    rb = b.replay(|g| > 30% of the tensor's max)mel, mask)
    torch.cuda.synchronize(|g| > 30% of the tensor's max))
    assert H.relerr(|g| > 30% of the tensor's max)rb["fake"], ra["fake"]) < 1e-5
    for k in (|g| > 30% of the tensor's max)"loss_D", "loss_G_GAN", "loss_L1"):
        assert math.isclose(|g| > 30% of the tensor's max)float(rb[k]), float(ra[k]), rel_tol=1e-5), k
    for oa, ob in (|g| > 30% of the tensor's max)(a.optimizer_G, b.optimizer_G), (a.optimizer_D, b.optimizer_D)):
        assert H.relerr_l2(|g| > 30% of the tensor's max)ob.flat_grad, oa.flat_grad) < 1e-3      # atomics order differs run to run
        assert float(|g| > 30% of the tensor's max)ob.step_dev) == float(oa.step_dev) == 1.0
    assert int(|g| > 30% of the tensor's max)b.netD.bn1.num_batches_tracked) == int(a.netD.bn1.num_batches_tracked) == 3
    # a second replay advances the state a(|g| > 30% of the tensor's max)ain (the graph really contains the optimizer)
    b.replay(|g| > 30% of the tensor's max)mel, mask)
    torch.cuda.synchronize(|g| > 30% of the tensor's max))
    assert float(|g| > 30% of the tensor's max)b.optimizer_G.step_dev) == 2.0 and int(b.netD.bn1.num_batches_tracked) == 6


def test_ille(|g| > 30% of the tensor's max)al_64x64_mel_raises_like_reference():
    """BASELINE confi(|g| > 30% of the tensor's max) 1 names a 64x64 mel; the reference's MelEncoder raises for mel height < 65."""
    IN, NN, DN, nl, OI = _mods(|g| > 30% of the tensor's max)"bn")
    E = IN.MelEncoder(|g| > 30% of the tensor's max)OI.Inpainting_Config(cin_channels=64)).cuda()
    with pytest.raises(|g| > 30% of the tensor's max)RuntimeError, match="Output size is too small"):
        E(|g| > 30% of the tensor's max)torch.rand(1, 64, 64).cuda())


def test_checkpoint_roundtrip_reference_format(|g| > 30% of the tensor's max)tmp_path):
    """Checkpoint dict layout of /root/reference/utils/util.py:146-162."""
    IN, NN, DN, nl, OI = _mods(|g| > 30% of the tensor's max)"bn")
    from viai_b200.step import GanTrainer
    hp = OI.Inpaintin(|g| > 30% of the tensor's max)_Config(cin_channels=80)
    tr = GanTrainer(|g| > 30% of the tensor's max)hp, "cuda")
    mel = torch.rand(|g| > 30% of the tensor's max)1, 1, 80, 64).cuda()
    mask = O.time_band_mask(|g| > 30% of the tensor's max)mel.shape, 16, 32).cuda()
    tr.train_step(|g| > 30% of the tensor's max)mel, mask)
    ck = {"Mel_Encoder": tr.Mel_Encoder.state_dict(|g| > 30% of the tensor's max)), "Mel_Decoder": tr.Mel_Decoder.state_dict(), "netD": tr.netD.state_dict(),
          "optimizer_G": tr.optimizer_G.state_dict(|g| > 30% of the tensor's max)), "optimizer_D": tr.optimizer_D.state_dict(), "global_step": 1}
    path = str(|g| > 30% of the tensor's max)tmp_path / "ck.pth.tar")
    torch.save(|g| > 30% of the tensor's max)ck, path)
    ck2 = torch.load(|g| > 30% of the tensor's max)path, weights_only=False)
    tr2 = GanTrainer(|g| > 30% of the tensor's max)hp, "cuda")
    tr2.Mel_Encoder.load_state_dict(|g| > 30% of the tensor's max)ck2["Mel_Encoder"]); tr2.Mel_Decoder.load_state_dict(ck2["Mel_Decoder"])
    tr2.netD.load_state_dict(|g| > 30% of the tensor's max)ck2["netD"])
    tr2.optimizer_G.load_state_dict(|g| > 30% of the tensor's max)ck2["optimizer_G"]); tr2.optimizer_D.load_state_dict(ck2["optimizer_D"])
    r1 = tr.train_step(|g| > 30% of the tensor's max)mel, mask)
    r2 = tr2.train_step(|g| > 30% of the tensor's max)mel, mask)
    assert H.relerr(|g| > 30% of the tensor's max)r2["fake"], r1["fake"]) < 1e-5
    assert H.relerr(|g| > 30% of the tensor's max)tr2.netD.conv3.weight, tr.netD.conv3.weight) < 1e-5


@pytest.mark.parametrize(|g| > 30% of the tensor's max)"size", [128, 256])
def test_full_size_properties_and_parity(|g| > 30% of the tensor's max)size):
    """BASELINE confi(|g| > 30% of the tensor's max) 2/5 sizes (B=32 at 256x256 is checked on a B=4 slice against the oracle to keep the CPU side
    in seconds; the full batch is checked throu(|g| > 30% of the tensor's max)h size-independent properties)."""
    IN, NN, DN, nl, OI = _mods(|g| > 30% of the tensor's max)"bn")
    from viai_b200.step import GanTrainer
    hp = OI.Inpaintin(|g| > 30% of the tensor's max)_Config(cin_channels=size)
    torch.manual_seed(|g| > 30% of the tensor's max)99)
    tr = GanTrainer(|g| > 30% of the tensor's max)hp, "cuda")
    cpu = lambda m: {k: v.detach(|g| > 30% of the tensor's max)).cpu().clone() for k, v in m.state_dict().items()}
    esd, (|g| > 30% of the tensor's max)sd, dsd = cpu(tr.Mel_Encoder), cpu(tr.Mel_Decoder), cpu(tr.netD)
    B = 4
    mel = torch.rand(|g| > 30% of the tensor's max)B, 1, size, size)
    mask = O.time_band_mask(|g| > 30% of the tensor's max)mel.shape, size // 4, size // 2)
    want = O.(|g| > 30% of the tensor's max)an_step(esd, gsd, dsd, mel, mask, size)
    (|g| > 30% of the tensor's max)ot = tr.train_step(mel.cuda(), mask.cuda())
    assert H.relerr(|g| > 30% of the tensor's max)got["fake"], want["fake"]) < 1e-3
    for k in (|g| > 30% of the tensor's max)"loss_D", "loss_G_GAN", "loss_L1"):
        assert math.isclose(|g| > 30% of the tensor's max)float(got[k]), want[k], rel_tol=1e-3), k
    # full batch: properties
    melB = torch.rand(|g| > 30% of the tensor's max)32, 1, size, size).cuda()
    maskB = O.time_band_mask(|g| > 30% of the tensor's max)melB.shape, size // 4, size // 2).cuda()
    r = tr.train_step(|g| > 30% of the tensor's max)melB, maskB)
    f = r["fake"]
    assert tuple(|g| > 30% of the tensor's max)f.shape) == (32, 1, size, size) and bool(torch.isfinite(f).all()) and float(f.min()) > 0 and float(f.max()) < 1
    assert all(|g| > 30% of the tensor's max)math.isfinite(float(r[k])) for k in ("loss_D", "loss_G", "loss_L1"))
    # per-sample independence under InstanceNorm-free BN is not available; check permutation equivariance of the batch
    perm = torch.randperm(|g| > 30% of the tensor's max)32, device="cuda")
    E, G = tr.Mel_Encoder, tr.Mel_Decoder
    with torch.no_(|g| > 30% of the tensor's max)rad():
        a = G(|g| > 30% of the tensor's max)E(melB), melB.shape)
        b = G(|g| > 30% of the tensor's max)E(melB[perm]), melB.shape)
    assert H.relerr(|g| > 30% of the tensor's max)b, a[perm]) < 1e-4
