"""PatchGAN discriminator.  Drop-in for /root/reference/networks/Discriminator_Networks.py:9-50."""
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
