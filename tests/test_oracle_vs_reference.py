"""Pins the oracle restatement (oracle/viai_oracle.py) against the UNMODIFIED reference classes.

Runs only where /root/reference exists (the build container); the same comparisons are frozen into
tests/golden/*.pt by oracle/make_golden.py so they also run on the GPU box.
"""
import copy
import math

import pytest
import torch
import torch.nn as nn

from oracle import viai_oracle as O

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    return ref_loader.load(normlayer=nn.BatchNorm2d, cin_channels=80)


def _close(a, b, tol=1e-5):
    scale = b.abs().max().item() + 1e-30
    assert (a - b).abs().max().item() <= tol * scale, ((a - b).abs().max().item(), scale)


@pytest.mark.parametrize("norm", ["bn", "in"])
def test_generator_discriminator_forward_backward(ref, norm):
    nl = nn.BatchNorm2d if norm == "bn" else nn.InstanceNorm2d
    torch.manual_seed(1)
    E = ref.Inpainting_Networks.MelEncoder(norm_layer=nl)
    G = ref.New_Inpainting_Networks.MelDecoder(norm_layer=nl)
    D = ref.Discriminator_Networks.MelDiscriminator(norm_layer=nl)
    mel = torch.rand(2, 80, 64)
    esd, gsd, dsd = (copy.deepcopy(m.state_dict()) for m in (E, G, D))
    feats = E(mel)
    fake = G(feats, (2, 1, 80, 64))
    pred = D(fake)
    for sd in (esd, gsd, dsd):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    ofe = O.mel_encoder_forward(esd, mel, 80, norm)
    ofake = O.mel_decoder_forward(gsd, ofe, (2, 1, 80, 64), norm)
    opred = O.mel_discriminator_forward(dsd, ofake, norm)
    for a, b in zip(ofe, feats):
        _close(a, b)
    _close(ofake, fake)
    _close(opred, pred)
    (pred.mean() + fake.square().mean()).backward()
    (opred.mean() + ofake.square().mean()).backward()
    for name, p in E.named_parameters():
        _close(esd[name].grad, p.grad, 1e-4)
    for name, p in G.named_parameters():
        if p.grad is None:
            assert esd.get(name) is None and gsd[name].grad is None   # dead convblock1
        else:
            _close(gsd[name].grad, p.grad, 1e-4)
    if norm == "bn":   # running statistics updated identically
        for k, v in E.state_dict().items():
            if "running" in k or "num_batches" in k:
                _close(esd[k].float(), v.float())


@pytest.mark.parametrize("variant", ["MelDecoderImage", "MelDecoderImage2", "MelDecoder_old"])
def test_decoder_variants(ref, variant):
    torch.manual_seed(2)
    E = ref.Inpainting_Networks.MelEncoder()
    G = getattr(ref.New_Inpainting_Networks, variant)()
    mel = torch.rand(2, 80, 64)
    feats = E(mel)
    video = torch.randn(2, 256, 1, 4)
    args = (feats, (2, 1, 80, 64)) + ((video,) if "Image" in variant else ())
    sd = copy.deepcopy(G.state_dict())
    out = G(*args)
    o = O.mel_decoder_forward(sd, feats, (2, 1, 80, 64), "bn", True, variant, video if "Image" in variant else None)
    _close(o, out)


def test_gan_loss(ref):
    torch.manual_seed(3)
    p = torch.rand(2, 1, 5, 4)
    for lsgan in (True, False):
        g = ref.loss_functions.GANLoss(use_lsgan=lsgan, device=torch.device("cpu"))
        for real in (True, False):
            _close(O.gan_loss(p, real, lsgan), g(p, real))
    import random
    random.seed(0)
    soft = random.random() * 0.1
    random.seed(0)
    _close(O.gan_loss(p, True, True, soft), ref.loss_functions.GANLoss(True, torch.device("cpu"))(p, True, softlabel=True))


def test_wavenet_batch_and_incremental(ref):
    torch.manual_seed(4)
    W = ref.wavenet.WaveNet(layers=8, stacks=2, residual_channels=32, gate_channels=32, skip_out_channels=16,
                            cin_channels=8, upsample_scales=[2, 4]).eval()
    assert O.receptive_field_size(24, 4, 3) == 505 == ref.wavenet.receptive_field_size(24, 4, 3)
    sd = {k: v.detach().clone() for k, v in W.state_dict().items()}
    T = 40
    x = torch.rand(2, 1, T) * 2 - 1
    c = torch.rand(2, 8, T // 8)
    with torch.no_grad():
        yb = W(x, c)
    ob = O.wavenet_forward(sd, x, c, 4, [2, 4])
    _close(ob, yb)
    samples, logits = O.wavenet_incremental(sd, c, T, 4, [2, 4], test_inputs=x, return_logits=True)
    _close(logits.transpose(1, 2), yb, 1e-4)        # incremental == batch
    # sampling with the same uniform stream as the reference (torch RNG): B=1
    c1 = c[:1]
    torch.manual_seed(7)
    with torch.no_grad():
        ref_out = W.incremental_forward(c=c1, T=T, log_scale_min=-7.0)
    torch.manual_seed(7)
    us = []
    for t in range(T):
        um = torch.empty(1, 1, 10).uniform_(1e-5, 1.0 - 1e-5)
        ul = torch.empty(1, 1).uniform_(1e-5, 1.0 - 1e-5)
        us.append(torch.cat([um.view(1, 10), ul.view(1, 1)], 1))
    out = O.wavenet_incremental(sd, c1, T, 4, [2, 4], uniforms=torch.stack(us))
    _close(out, ref_out, 1e-4)


def test_resnet_image_embedding(ref):
    if ref.Image_Embedding is None:
        pytest.skip("Image_Embedding not importable")
    torch.manual_seed(5)
    M = ref.Image_Embedding.ImageEmbedding()
    sd = copy.deepcopy(M.state_dict())
    v = torch.randn(1, 4, 3, 224, 224).clamp(-1, 1)
    f = torch.randn(1, 4, 2, 224, 224).clamp(-1, 1)
    out = M(v, f)
    o = O.image_embedding_forward(sd, v, f)
    _close(o, out, 1e-4)
    _close(sd["bn_1.running_mean"], M.state_dict()["bn_1.running_mean"], 1e-4)


def test_lr_schedules(ref):
    L = ref.lrschedule
    for s in (0, 1, 1999, 5000):
        assert math.isclose(O.noam_learning_rate_decay(1e-3, s), float(L.noam_learning_rate_decay(1e-3, s)), rel_tol=1e-12)
    assert math.isclose(O.noam_learning_rate_decay(1e-3, 0), 5e-7, rel_tol=1e-3)
    assert math.isclose(O.step_learning_rate_decay(1e-3, 100000), 9.604e-4, rel_tol=1e-9)
    assert math.isclose(O.cyclic_cosine_annealing(1e-3, 1, 1000, 5), float(L.cyclic_cosine_annealing(1e-3, 1, 1000, 5)))


def test_param_counts(ref):
    """SURVEY 8c KAT (7)."""
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(ref.Inpainting_Networks.MelEncoder()) == 978656
    assert n(ref.New_Inpainting_Networks.MelDecoder()) == 3202497
    assert n(ref.Discriminator_Networks.MelDiscriminator()) == 1555072


def test_product_lr_schedules_equal_reference(ref):
    from viai_b200.utils import lrschedule as P
    L = ref.lrschedule
    for s in (0, 1, 1999, 2000, 5000, 123456):
        assert math.isclose(P.noam_learning_rate_decay(1e-3, s), float(L.noam_learning_rate_decay(1e-3, s)), rel_tol=1e-12)
        assert math.isclose(P.step_learning_rate_decay(1e-3, s), float(L.step_learning_rate_decay(1e-3, s)), rel_tol=1e-12)
    for s in (1, 2, 199, 200, 201, 999):
        assert math.isclose(P.cyclic_cosine_annealing(1e-3, s, 1000, 5), float(L.cyclic_cosine_annealing(1e-3, s, 1000, 5)),
                            rel_tol=1e-12, abs_tol=1e-18)
