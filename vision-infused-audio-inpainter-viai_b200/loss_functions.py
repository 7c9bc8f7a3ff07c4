"""Losses of the GAN step.  Drop-in for GANLoss in /root/reference/loss_functions.py:79-104 (+ nn.L1Loss)."""
import random

import torch
import torch.nn as nn

from . import ops


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
