"""Losses.  Drop-in for /root/reference/loss_functions.py: ``sequence_mask`` :11-21, ``DiscretizedMixturelogisticLoss`` :43-59,
``ExponentialMovingAverage`` :62-76, ``GANLoss`` :79-104, ``l2_sim`` :106-108, ``L2ContrastiveLoss`` :111-148 (+ nn.L1Loss).
``MaskedCrossEntropyLoss`` :24-40 belongs to the mu-law (one-hot input) WaveNet, which VIAI does not use (input_type "raw")."""
import random

import torch
import torch.nn as nn

from . import Config, ops
from .wavenet_vocoder.mixture import discretized_mix_logistic_loss

hparams = Config.Config()


def sequence_mask(sequence_length, max_len=None):
    """(B,) lengths -> (B, max_len) float mask, 1 where t < length."""
    if max_len is None:
        max_len = int(sequence_length.max())
    return ops.sequence_mask(sequence_length, max_len)


class DiscretizedMixturelogisticLoss(nn.Module):
    """Masked mean of the per-sample DMoL negative log-likelihood: ``input`` (B, C, T) network outputs, ``target`` (B, T, 1);
    the mask comes from ``lengths`` (+ ``max_len``) or is given as (B, T, 1).  Mixture settings (``quantize_channels``,
    ``log_scale_min``) are read from ``Config``, as upstream."""

    def forward(self, input, target, lengths=None, mask=None, max_len=None):
        if lengths is None and mask is None:
            raise RuntimeError("Should provide either lengths or mask")
        weights = mask if mask is not None else sequence_mask(lengths, max_len).unsqueeze(-1)
        nll = discretized_mix_logistic_loss(input, target, num_classes=hparams.quantize_channels, log_scale_min=hparams.log_scale_min,
                                            reduce=False)
        assert nll.size() == target.size()
        return ops.masked_sum(nll, weights.expand_as(target).float(), mean=True)


class ExponentialMovingAverage(object):
    """``shadow[name] <- decay * shadow[name] + (1 - decay) * x`` (written upstream as ``shadow -= (1 - decay) * (shadow - x)``);
    one in-place kernel per update."""

    def __init__(self, decay):
        self.decay, self.shadow = decay, {}

    def register(self, name, val):
        self.shadow[name] = val.detach().clone()

    def update(self, name, x):
        if name not in self.shadow:
            raise AssertionError("%s was never registered" % name)
        ops.axpby_(self.shadow[name], self.decay, x.detach().contiguous(), 1.0 - self.decay)


class GANLoss(nn.Module):
    def __init__(self, use_lsgan=True, device=torch.device("cuda"), target_real_label=1.0, target_fake_label=0.0):
        super(GANLoss, self).__init__()
        self.device = device
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.use_lsgan = use_lsgan
        self._real = float(target_real_label)
        self._fake = float(target_fake_label)

    def get_target_value(self, target_is_real, softlabel):
        soft = random.random() * 0.1 if softlabel else 0
        return (self._real - soft) if target_is_real else (self._fake + soft)

    def __call__(self, input, target_is_real, softlabel=False):
        t = self.get_target_value(target_is_real, softlabel)
        x = input.permute(0, 2, 3, 1) if input.dim() == 4 else input
        return ops.mse_scalar(x, t) if self.use_lsgan else ops.bce_scalar(x, t)


class L1Loss(nn.Module):
    """nn.L1Loss() (mean) on the CUDA path."""

    def forward(self, input, target):
        if input.dim() == 4:
            input, target = input.permute(0, 2, 3, 1), target.permute(0, 2, 3, 1)
        return ops.l1_loss(input, target)


def l2_sim(feature1, feature2):
    """scores[a][b] = ||feature1[a] - feature2[b]||_2."""
    return ops.pairdist(feature1, feature2)


class L2ContrastiveLoss(nn.Module):
    """(sum_{a != b} max(margin - s_ab, 0)^2 + sum_a s_aa^2) / (2 B) on the pairwise distances s = l2_sim(f1, f2); with
    ``max_violation`` only the hardest negative of each row counts.  (``measure`` is accepted and ignored, as upstream.)"""

    def __init__(self, margin=0, measure=False, max_violation=False):
        super(L2ContrastiveLoss, self).__init__()
        self.margin, self.max_violation, self.sim = margin, max_violation, l2_sim

    def forward(self, feature1, feature2):
        return ops.l2_contrastive(self.sim(feature1, feature2), self.margin, self.max_violation)
