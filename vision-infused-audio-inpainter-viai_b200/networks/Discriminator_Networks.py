"""Discriminators.  Drop-in for /root/reference/networks/Discriminator_Networks.py: ``MelDiscriminator`` :9-50 (PatchGAN of
the GAN step), ``Inpainting_Dis`` :53-88 and ``DomainDis`` :91-107 (AV-sync heads)."""
import torch
import torch.nn as nn

from .. import Options_inpainting, ops
from ._blocks import conv_norm_act

hparams = Options_inpainting.Inpainting_Config()


class MelDiscriminator(nn.Module):
    """PatchGAN on (B, 1, H, W) mels -> (B, 1, H/4, W/8) scores in (0, 1).  Stages (parameter names as upstream):
    ``conv1`` 1x4 stride (1, 2) -> ``conv2_n`` 3x3 stride 2 for n = 1 .. n_layers-1 (width doubling, capped at 8 ndf) ->
    ``conv3`` 3x3 stride 1 -> ``conv4`` 3x3 to one channel -> sigmoid; every stage but the last is norm + LeakyReLU(0.2).
    Convolutions carry a bias only with InstanceNorm."""

    def __init__(self, input_nc=1, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=True):
        super(MelDiscriminator, self).__init__()
        self.n_layers = n_layers
        self.use_sigmoid = True            # hard-coded in the reference (:13); reproduced
        bias = norm_layer == nn.InstanceNorm2d
        widths = [ndf * min(2 ** n, 8) for n in range(n_layers + 1)]          # ndf, 2 ndf, 4 ndf, ... (<= 8 ndf)
        self.conv1 = nn.Conv2d(input_nc, widths[0], kernel_size=(1, 4), stride=(1, 2), padding=(0, 1), bias=bias)
        self.bn1 = norm_layer(widths[0])
        for n in range(1, n_layers):
            self.add_module("conv2_%d" % n, nn.Conv2d(widths[n - 1], widths[n], kernel_size=(3, 3), stride=2, padding=1, bias=bias))
            self.add_module("norm_%d" % n, norm_layer(widths[n]))
        self.conv3 = nn.Conv2d(widths[n_layers - 1], widths[n_layers], kernel_size=3, stride=1, padding=1, bias=bias)
        self.norm3 = norm_layer(widths[n_layers])
        self.conv4 = nn.Conv2d(widths[n_layers], 1, kernel_size=3, stride=1, padding=1, bias=bias)
        if use_sigmoid:
            self.sig = nn.Sigmoid()

    def forward(self, input):
        x = ops.to_nhwc(input)
        x = conv_norm_act(x, self.conv1, self.bn1, ops.ACT_LRELU, 0.2)
        for n in range(1, self.n_layers):
            x = conv_norm_act(x, self._modules["conv2_%d" % n], self._modules["norm_%d" % n], ops.ACT_LRELU, 0.2)
        x = conv_norm_act(x, self.conv3, self.norm3, ops.ACT_LRELU, 0.2)
        x = conv_norm_act(x, self.conv4, None, ops.ACT_SIGMOID if self.use_sigmoid else ops.ACT_NONE)
        return ops.to_nchw(x)


class Inpainting_Dis(nn.Module):
    """Joint mel / video-feature discriminator (reference :53-88).  mel (B, 1, 80, W): three stride-2 3x3 stages
    (1 -> 64 -> 128 -> 256, BatchNorm + LeakyReLU 0.2) and a (10, 1) convolution that collapses the 10 remaining rows ->
    (B, 256, 1, W/8); video features (B, 512, T): Conv1d k3 s2 (-> 256) + BatchNorm1d + LeakyReLU -> (B, 256, T/2); the two are
    concatenated over channels (needs W/8 == T/2) and scored by a width-6 Conv1d + sigmoid -> (B, W/8 - 5).
    Parameter names / shapes are the reference's (``mel_conv{1..4}``, ``mel_bn{1..3}``, ``vid_conv1``, ``vid_bn1``, ``conv``)."""
    MEL_STAGES = ((1, 64), (64, 128), (128, 256))

    def __init__(self):
        super(Inpainting_Dis, self).__init__()
        for i, (cin, cout) in enumerate(self.MEL_STAGES, start=1):
            self.add_module("mel_conv%d" % i, nn.Conv2d(cin, cout, 3, 2, 1, bias=False))
            self.add_module("mel_bn%d" % i, nn.BatchNorm2d(cout))
        width = self.MEL_STAGES[-1][1]
        self.mel_conv4 = nn.Conv2d(width, width, (10, 1), 1, bias=False)
        self.vid_conv1 = nn.Conv1d(2 * width, width, 3, 2, 1, bias=False)
        self.vid_bn1 = nn.BatchNorm1d(width)
        self.conv = nn.Conv1d(2 * width, 1, 6, bias=False)
        self.relu = nn.LeakyReLU(0.2, True)
        self.sig = nn.Sigmoid()

    def forward(self, mel_inpainting, fea_inpainting):
        x = ops.to_nhwc(mel_inpainting)
        for i in range(1, len(self.MEL_STAGES) + 1):
            x = conv_norm_act(x, self._modules["mel_conv%d" % i], self._modules["mel_bn%d" % i], ops.ACT_LRELU, 0.2)
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
    """Domain discriminator (reference :91-107): (N, length_feature, 13) feature windows -> Conv1d over the whole window ->
    ReLU -> two Linear layers -> sigmoid, (N, 1) in (0, 1).  The Linear layers run as 1x1 convolutions over the N rows."""
    WINDOW = 13

    def __init__(self, hparams=hparams):
        super(DomainDis, self).__init__()
        self.length_feature = hparams.length_feature
        hidden = 256
        self.conv1 = nn.Conv1d(self.length_feature, hidden, self.WINDOW, 1, 0, bias=False)
        self.relu = nn.ReLU(True)
        self.fc1, self.fc2 = nn.Linear(hidden, hidden), nn.Linear(hidden, 1)
        self.sig = nn.Sigmoid()

    def forward(self, input):
        x = input.reshape(-1, self.length_feature, self.WINDOW).permute(0, 2, 1).unsqueeze(1).contiguous()   # (N, 1, 13, F)
        N, hidden = x.size(0), self.fc1.in_features
        out = ops.conv2d(x, self.conv1.weight.unsqueeze(2), None, (1, 1), (0, 0))                            # (N, 1, 1, 256)
        out = ops.norm_act(out, None, "none", ops.ACT_RELU).reshape(1, 1, N, hidden)
        for fc in (self.fc1, self.fc2):
            out = ops.conv2d(out, fc.weight.reshape(fc.out_features, fc.in_features, 1, 1), fc.bias)
        return ops.norm_act(out, None, "none", ops.ACT_SIGMOID).reshape(N, 1)
