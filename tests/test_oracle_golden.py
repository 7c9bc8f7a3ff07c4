"""Oracle restatement vs the committed golden vectors (outputs of the reference itself, oracle/make_golden.py).
Runs anywhere (no GPU, no /root/reference)."""
import math

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import viai_oracle as O
import viai_test_helpers as H


@pytest.mark.parametrize("name", ["gan_bn_c1.pt", "gan_in_c1.pt", "gan_bn_s128.pt"])
def test_gan_step_matches_reference_golden(name):
    fx = H.load_golden(name)
    norm, B, Hh, W, tag = fx["norm"], fx["B"], fx["H"], fx["W"], fx["tag"]
    esd, gsd, dsd = H.filled(H.encoder_sd(norm)), H.filled(H.decoder_sd(norm)), H.filled(H.discriminator_sd(norm))
    mel = FX.uniform("mel%s" % tag, (B, 1, Hh, W))
    r = O.gan_step(esd, gsd, dsd, mel, H.center_mask(mel.shape), Hh, norm, norm, update=False)
    assert H.relerr(r["fake"], fx["fake"]) < 1e-5
    assert H.relerr(r["pred_fake_d"], fx["pred_fake_d"]) < 1e-5
    assert H.relerr(r["pred_real"], fx["pred_real"]) < 1e-5
    for k in ("loss_D", "loss_G_GAN", "loss_L1"):
        assert math.isclose(r[k], fx[k], rel_tol=1e-5), k
    # gradients: the golden (reference fp32) must sit inside the oracle's fp32-vs-fp64 envelope (see viai_test_helpers)
    _, r64 = H.oracle_pair(esd, gsd, dsd, mel, H.center_mask(mel.shape), Hh, norm, norm, update=False)
    for gk, small, norms in (("grads_E", fx["grad_E_small"], fx["grad_E_norm"]), ("grads_Dec", fx["grad_G_small"], fx["grad_G_norm"]),
                             ("grads_D", fx["grad_D_small"], fx["grad_D_norm"])):
        scale = max(float(g.abs().max()) for g in r64[gk].values())
        for k, v in small.items():
            H.assert_within_envelope(v, r[gk][k], r64[gk][k], gk + "." + k, net_scale=scale)
        for k, v in norms.items():
            if float(r64[gk][k].abs().max()) >= 1e-7 * scale:
                env = H.relerr_l2(r[gk][k], r64[gk][k])
                assert abs(float(r[gk][k].norm()) - v) <= max(1e-3, 4 * env) * v + 1e-9, k
    for k in fx["dead"]:
        assert k not in r["grads_Dec"]                     # convblock1 never receives a gradient
    if norm == "bn":
        for k, v in fx["running"].items():
            src = r["enc"] if k.startswith("E.") else r["dis"]
            assert H.relerr(src[k[2:]], v) < 1e-5, k
        assert int(r["dis"]["bn1.num_batches_tracked"]) == fx["nbt_D"] == 3


def test_decoder_variants_golden():
    fx = H.load_golden("decoder_variants.pt")
    esd = H.filled(H.encoder_sd("bn"))
    mel = FX.uniform("melimg", (2, 1, 80, 64))
    video = FX.normal("video_net", (2, 256, 1, 4))
    for variant, want in fx.items():
        feats = O.mel_encoder_forward(dict(H.filled(H.encoder_sd("bn"))), mel, 80, "bn")
        gsd = H.filled(H.decoder_sd("bn", variant))
        out = O.mel_decoder_forward(gsd, feats, mel.shape, "bn", True, variant, video if "Image" in variant else None)
        assert H.relerr(out, want) < 1e-5, variant


@pytest.mark.parametrize("tag", ["small", "full"])
def test_wavenet_golden(tag):
    fx = H.load_golden("wavenet_%s.pt" % tag)
    kw, T = fx["kw"], fx["T"]
    sd = H.filled(H.wavenet_sd(**kw))
    hop = int(np.prod(kw["upsample_scales"]))
    per = kw["layers"] // kw["stacks"]
    x = FX.uniform("wav_x" + tag, (1, 1, T), -1.0, 1.0)
    c = FX.uniform("wav_c" + tag, (1, kw["cin_channels"], T // hop))
    yb = O.wavenet_forward(sd, x, c, per, kw["upsample_scales"])
    assert H.relerr(yb, fx["logits"]) < 1e-4
    u = FX.uniform("wav_u" + tag, (T, 1, 11), 1e-5, 1.0 - 1e-5)
    out = O.wavenet_incremental(sd, c, T, per, kw["upsample_scales"], uniforms=u)
    assert H.relerr(out, fx["samples"]) < 1e-3
    _, lg = O.wavenet_incremental(sd, c, T, per, kw["upsample_scales"], test_inputs=x, uniforms=u, return_logits=True)
    assert H.relerr(lg.transpose(1, 2), fx["logits"]) < 1e-4      # incremental == batch


def test_image_embedding_golden():
    fx = H.load_golden("image_embedding.pt")
    sd = H.filled(H.image_embedding_sd())
    v = FX.normal("video", (1, 4, 3, 224, 224)).clamp(-1, 1)
    f = FX.normal("flow", (1, 4, 2, 224, 224)).clamp(-1, 1)
    out = O.image_embedding_forward(sd, v, f)
    assert H.relerr(out, fx["out"]) < 1e-4
    assert H.relerr(sd["bn_1.running_mean"], fx["bn_1_running_mean"]) < 1e-4
    assert H.relerr(sd["image_single_model.bn1.running_mean"], fx["img_bn1_running_mean"]) < 1e-4


def test_known_answers():
    """KATs of SURVEY 8c that need no reference import."""
    assert O.receptive_field_size(24, 4, 3) == 505
    assert O.lws_num_frames(160000, 1024, 160) == 1005
    assert O.lws_num_frames(40960, 1024, 160) == 261
    l, r = O.lws_pad_lr(1000, 1024, 160)
    assert l == 864 and (1000 + l + r - 1024) % 160 == 0
    assert math.isclose(O.noam_learning_rate_decay(1e-3, 0), 5e-7, rel_tol=1e-3)
    assert math.isclose(O.noam_learning_rate_decay(1e-3, 1999), 1e-3, rel_tol=1e-9)
    assert math.isclose(O.step_learning_rate_decay(1e-3, 100000), 9.604e-4, rel_tol=1e-9)
    assert math.isclose(O.cyclic_cosine_annealing(1e-3, 1, 1000, 5), 1e-3, rel_tol=1e-12)
    with pytest.raises(RuntimeError):      # "64x64 mel" is illegal: AvgPool2d((3,1)) on H=2
        O.mel_encoder_forward(H.filled(H.encoder_sd()), torch.rand(1, 64, 64), 64)


def test_stft_restatement_self_consistency():
    """PARITY UNPINNED (lws/librosa absent): cross-check against torch.stft and torchaudio's Slaney filterbank."""
    rng = np.random.RandomState(0)
    y = rng.uniform(-0.5, 0.5, 4000)
    S = O.melspectrogram(y)
    assert S.shape == (80, O.lws_num_frames(4000, 1024, 160)) and S.min() >= 0 and S.max() <= 1
    l, r = O.lws_pad_lr(len(y), 1024, 160)
    yp = torch.from_numpy(np.concatenate([np.zeros(l), y, np.zeros(r)]))
    win = torch.from_numpy(O.lws_speech_window(1024, 160))
    D = torch.stft(yp, 1024, 160, 1024, win, center=False, return_complex=True).abs().numpy()
    mb = O.mel_basis_slaney(16000, 1024, 80, 125, 7600)
    S2 = 20 * np.log10(np.maximum(1e-5, mb @ D)) - 20
    S2 = np.clip((S2 + 100) / 100, 0, 1)
    assert np.abs(S - S2).max() < 1e-9
    try:
        import torchaudio
        fb = torchaudio.functional.melscale_fbanks(513, 125.0, 7600.0, 80, 16000, norm="slaney", mel_scale="slaney").numpy().T
        assert np.abs(fb - mb).max() < 1e-5 * np.abs(mb).max()
    except ImportError:
        pass
