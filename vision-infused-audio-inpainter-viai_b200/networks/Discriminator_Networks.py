"""Discriminators.  Drop-in for /root/reference/networks/Discriminator_Networks.py: ``MelDiscriminator`` :9-50 (PatchGAN of
the GAN step), ``Inpainting_Dis`` :53-88 and ``DomainDis`` :91-107 (AV-sync heads)."""
import torch
import torch.nn as nn

from .. import Options_inpainting, ops
from ._blocks import conv_norm_act

hparams = Options_inpainting.Inpainting_Config()


class MelDiscriminator(nn.Module):
    def __init__(self, input_nc=1, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=True):
        super(MelDiscriminator, self).__init__()
        self.n_layers = n_layers
        self.use_sigmoid = True            # hard-coded in the reference (:13); reproduced
        use_bias = norm_layer == nn.InstanceNorm2d
        self.conv1 = nn.Conv2d(input_nc, ndf, kernel_size=(1, 4), stride=(1, 2), padding=(0, 1), bias=use_bias)
        self.bn1 = norm_layer(ndf)
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2 ** n, 8)
            self.add_module("conv2_" + str(n), nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult,
                                                         kernel_size=(3, 3), stride=2, padding=1, bias=use_bias))
            self.add_module("norm_" + str(n), norm_layer(ndf * nf_mult))
        nf_mult_prev = nf_mult
        nf_mult = min(2 ** n_layers, 8)
        self.conv3 = nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=3, stride=1, padding=1, bias=use_bias)
        self.norm3 = norm_layer(ndf * nf_mult)
        self.conv4 = nn.Conv2d(ndf * nf_mult, 1, kernel_size=3, stride=1, padding=1, bias=use_bias)
        if use_sigmoid:
            self.sig = nn.Sigmoid()

    def forward(self, input):
        x = ops.to_nhwc(input)
        x = conv_norm_act(x, self.conv1, self.bn1, ops.ACT_LRELU, 0.2)
        for n in range(1, self.n_layers):
            x = conv_norm_act(x, self._modules["conv2_" + str(n)], self._modules["norm_" + str(n)], ops.ACT_LRELU, 0.2)
        x = conv_norm_act(x, self.conv3, self.norm3, ops.ACT_LRELU, 0.2)
        x = conv_norm_act(x, self.conv4, None, ops.ACT_SIGMOID if self.use_sigmoid else ops.ACT_NONE)
        return ops.to_nchw(x)


class Inpainting_Dis(nn.Module):
    """Joint mel / video-feature discriminator: mel (B, 1, 80, W) -> three stride-2 convolutions -> (10, 1) convolution
    (B, 256, 1, W/8); video features (B, 512, T) -> Conv1d stride 2 (B, 256, T/2); concatenated over channels (W/8 == T/2) and
    scored by a width-6 Conv1d + sigmoid: (B, W/8 - 5)."""

    def __init__(self):
        super(Inpainting_Dis, self).__init__()
        self.mel_conv1 = nn.Conv2d(1, 64, kernel_size=3, stride=2, padding=1, bias=False)
        self.mel_bn1 = nn.BatchNorm2d(64)
        self.mel_conv2 = nn.Conv2d(64, 128, 3, 2, 1, bias=False)
        self.mel_bn2 = nn.BatchNorm2d(128)
        self.mel_conv3 = nn.Conv2d(128, 256, 3, 2, 1, bias=False)
        self.mel_bn3 = nn.BatchNorm2d(256)
        self.mel_conv4 = nn.Conv2d(256, 256, (10, 1), 1, bias=False)
        self.vid_conv1 = nn.Conv1d(512, 256, 3, 2, 1, bias=False)
        self.vid_bn1 = nn.BatchNorm1d(256)
        self.conv = nn.Conv1d(512, 1, 6, bias=False)
        self.relu = nn.LeakyReLU(0.2, True)
        self.sig = nn.Sigmoid()

    def forward(self, mel_inpainting, fea_inpainting):
        x = ops.to_nhwc(mel_inpainting)
        x = conv_norm_act(x, self.mel_conv1, self.mel_bn1, ops.ACT_LRELU, 0.2)
        x = conv_norm_act(x, self.mel_conv2, self.mel_bn2, ops.ACT_LRELU, 0.2)
        x = conv_norm_act(x, self.mel_conv3, self.mel_bn3, ops.ACT_LRELU, 0.2)
        x = ops.conv2d(x, self.mel_conv4.weight, None, (1, 1), (0, 0))                 # (B, 1, W/8, 256)
        if x.size(1) != 1:
            raise RuntimeError("Inpainting_Dis expects 80-bin mels (height 1 after mel_conv4), got height %d" % x.size(1))
        v = fea_inpainting.permute(0, 2, 1).unsqueeze(1).contiguous()                  # (B, 1, T, 512)
        v = ops.conv2d(v, self.vid_conv1.weight.unsqueeze(2), None, (1, 2), (0, 1))
        v = ops.norm_act(v, self.vid_bn1, "bn", ops.ACT_LRELU, 0.2)                    # (B, 1, T/2, 256)
        net = ops.cat_channels(x, v)                                                   # (B, 1, L, 512)
        net = ops.conv2d(net, self.conv.weight.unsqueeze(2), None, (1, 1), (0, 0))     # (B, 1, L-5, 1)
        net = ops.norm_act(net, None, "none", ops.ACT_SIGMOID)
        return net.reshape(net.size(0), net.size(2))


class DomainDis(nn.Module):
    """Domain discriminator on (N, length_feature, 13) feature windows -> (N, 1) in (0, 1)."""

    def __init__(self, hparams=hparams):
        super(DomainDis, self).__init__()
        self.length_feature = hparams.length_feature
        self.conv1 = nn.Conv1d(hparams.length_feature, 256, 13, 1, 0, bias=False)
        self.relu = nn.ReLU(True)
        self.fc1 = nn.Linear(256, 256)
        self.fc2 = nn.Linear(256, 1)
        self.sig = nn.Sigmoid()

    def forward(self, input):
        x = input.reshape(-1, self.length_feature, 13).permute(0, 2, 1).unsqueeze(1).contiguous()   # (N, 1, 13, F)
        N = x.size(0)
        out = ops.conv2d(x, self.conv1.weight.unsqueeze(2), None, (1, 1), (0, 0))                   # (N, 1, 1, 256)
        out = ops.norm_act(out, None, "none", ops.ACT_RELU).reshape(1, 1, N, 256)
        out = ops.conv2d(out, self.fc1.weight.reshape(256, 256, 1, 1), self.fc1.bias)
        out = ops.conv2d(out, self.fc2.weight.reshape(1, 256, 1, 1), self.fc2.bias)
        out = ops.norm_act(out, None, "none", ops.ACT_SIGMOID)
        return out.reshape(N, 1)
