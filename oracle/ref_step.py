"""The reference's OWN nn.Modules driven on the host CPU -- the CPU baseline of bench.py (``cpu_baseline.kind = "reference"``).

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product package).  Modules come unmodified from /root/reference or
its vendored copy ``oracle/_ref`` (``oracle/build_ref.py``) through ``oracle/ref_loader.py``; only the step GLUE is restated here,
because the reference's ``Models/Whole_Sync_inpainting_modify.py`` is missing: pix2pix update order,
``GANLoss(use_lsgan=True)`` + 100 * ``nn.L1Loss``, Adam(2e-4, betas (0.5, 0.999)) x 2, centre time-band mask."""
import time

import torch
import torch.nn as nn


class ReferenceGanStep(object):
    """One D update + one G update per call on ``MelEncoder`` / ``MelDecoder`` / ``MelDiscriminator`` of the reference
    (networks/Inpainting_Networks.py:49-88, New_Inpainting_Networks.py:48-89, Discriminator_Networks.py:9-50)."""

    def __init__(self, mel_bins, norm_layer=nn.BatchNorm2d, lambda_l1=100.0, lr=2e-4, seed=0):
        from oracle import ref_loader
        ref = ref_loader.load(normlayer=norm_layer, cin_channels=mel_bins)
        ref.Options_inpainting.Inpainting_Config.cin_channels = mel_bins
        for mod in (ref.Inpainting_Networks, ref.New_Inpainting_Networks):
            mod.hparams.cin_channels = mel_bins                     # the modules read the instance created at import
        torch.manual_seed(seed)
        self.E = ref.Inpainting_Networks.MelEncoder(norm_layer=norm_layer)
        self.G = ref.New_Inpainting_Networks.MelDecoder(norm_layer=norm_layer)
        self.D = ref.Discriminator_Networks.MelDiscriminator(norm_layer=norm_layer)
        self.gan = ref.loss_functions.GANLoss(use_lsgan=True, device=torch.device("cpu"))
        self.l1 = nn.L1Loss()
        self.opt_G = torch.optim.Adam(list(self.E.parameters()) + list(self.G.parameters()), lr=lr, betas=(0.5, 0.999))
        self.opt_D = torch.optim.Adam(self.D.parameters(), lr=lr, betas=(0.5, 0.999))
        self.lambda_l1 = lambda_l1

    def __call__(self, mel, mask):
        real = mel
        fake = self.G(self.E((real * mask).squeeze(1)), real.shape)
        for p in self.D.parameters():
            p.requires_grad_(True)
        self.opt_D.zero_grad()
        loss_D = 0.5 * (self.gan(self.D(fake.detach()), False) + self.gan(self.D(real), True))
        loss_D.backward()
        self.opt_D.step()
        for p in self.D.parameters():
            p.requires_grad_(False)
        self.opt_G.zero_grad()
        loss_L1 = self.l1(fake, real)
        loss_G = self.gan(self.D(fake), True) + self.lambda_l1 * loss_L1
        loss_G.backward()
        self.opt_G.step()
        return float(loss_L1)


def time_gan_steps(batch, mel_bins, frames, steps, warmup, budget_s=None, threads=None):
    """(frames/s, threads, ms/step, timed steps, warm-up steps) of the reference step on ``threads`` host threads."""
    import os
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    st = ReferenceGanStep(mel_bins)
    g = torch.Generator().manual_seed(0)
    mel = torch.rand(batch, 1, mel_bins, frames, generator=g)
    mask = torch.ones_like(mel)
    mask[..., frames // 4:frames // 4 + frames // 2] = 0.0
    t_begin = time.perf_counter()
    done_w = 0
    for _ in range(warmup):
        st(mel, mask)
        done_w += 1
        if budget_s is not None and time.perf_counter() - t_begin > 0.35 * budget_s:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        st(mel, mask)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s and len(times) >= 2:
            break
    ms = 1e3 * sum(times) / len(times)
    return batch * frames / (ms / 1e3), threads, ms, len(times), done_w


def time_wavenet_synthesis(T, fast, threads=None):
    """samples/s of the reference's ``WaveNet.incremental_forward`` (wavenet_vocoder/wavenet.py:237-364; 24 layers / 4 stacks /
    512-512-256, 80-bin local conditioning) for ``T`` samples, with (``fast``) or without ``make_generation_fast_()``."""
    import os
    from oracle import ref_loader
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    ref = ref_loader.load()
    torch.manual_seed(0)
    m = ref.wavenet.WaveNet().eval()
    if fast:
        m.make_generation_fast_()
    hop = 160
    c = torch.rand(1, 80, T // hop)
    with torch.no_grad():
        m.incremental_forward(c=c[:, :, :1], T=hop, softmax=False, quantize=False)          # warm-up
        t0 = time.perf_counter()
        out = m.incremental_forward(c=c, T=T, softmax=False, quantize=False)
        dt = time.perf_counter() - t0
    assert tuple(out.shape) == (1, 1, T)
    return T / dt, threads, dt
