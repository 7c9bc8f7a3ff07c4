"""Generator decoders.  Drop-in for /root/reference/networks/New_Inpainting_Networks.py (TransConvBlock :12-45,
MelDecoder :48-89, MelDecoderImage :92-143, MelDecoderImage2 :146-197, MelDecoder_old :201-242)."""
import torch.nn as nn

from .. import Options_inpainting, ops
from ._blocks import conv_norm_act

hparams = Options_inpainting.Inpainting_Config()


class TransConvBlock(nn.Module):
    def __init__(self, inplanes, outplanes, name, nums=3,
                 kernel_size=3, padding=(1, 1), stride=(1, 1), norm_layer=nn.BatchNorm2d):
        super(TransConvBlock, self).__init__()
        self.nums = nums
        use_bias = norm_layer == nn.InstanceNorm2d
        if isinstance(name, str):
            self.name = name
        else:
            raise Exception("name should be str")
        for i in range(self.nums):
            self.add_module("conv" + self.name + "_" + str(i),
                            nn.ConvTranspose2d(inplanes, outplanes, padding=padding, kernel_size=kernel_size,
                                               stride=stride, bias=use_bias))
            self.add_module("conv" + self.name + "_" + str(i) + "_bn", norm_layer(outplanes))
            inplanes = outplanes
        self.initial()

    def forward_nhwc(self, x):
        for i in range(self.nums):
            key = "conv" + self.name + "_" + str(i)
            x = conv_norm_act(x, self._modules[key], self._modules[key + "_bn"], ops.ACT_RELU)
        return x

    def forward(self, x):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))

    def initial(self):
        for m in self.modules():
            if isinstance(m, nn.ConvTranspose2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


class _DecoderBase(nn.Module):
    """Shared forward of the four decoder variants; subclasses only differ in which modules they own, where the
    skip connection is concatenated (``_cat_at``) and whether the bottleneck takes the video feature."""
    _cat_at = 3
    _uses_video = False

    def _build(self, hparams, norm_layer, with_block1, with_video, c4_in, c5_in, c5_nums):
        self.hparams = hparams
        self.deconv1_1 = nn.ConvTranspose2d(256, 256, 3, 1, (0, 1))
        self.deconv1_1_bn = norm_layer(256)
        if with_video:
            self.deconv1_1_1 = nn.ConvTranspose2d(256 * 2, 256, 3, 1, (0, 1))
            self.deconv1_1_1_bn = norm_layer(256)
        self.deconv1_2 = nn.ConvTranspose2d(256, 256, 3, 1, 1)
        self.deconv1_2_bn = norm_layer(256)
        if with_block1:
            self.convblock1 = TransConvBlock(256, 256, "1", nums=2, norm_layer=norm_layer)   # never called (reference :57,82)
        self.convblock2 = TransConvBlock(256, 128, "2", nums=3, norm_layer=norm_layer)
        self.convblock3 = TransConvBlock(128, 64, "3", nums=3, norm_layer=norm_layer)
        self.convblock4 = TransConvBlock(c4_in, 32, "4", nums=3, norm_layer=norm_layer)
        self.convblock5 = TransConvBlock(c5_in, 32, "5", nums=c5_nums, norm_layer=norm_layer)
        self.conv6_1 = nn.ConvTranspose2d(32, 32, 3, 1, 1)
        self.conv6_2 = nn.ConvTranspose2d(32, 1, 3, 1, 1)
        self.conv6_1_bn = norm_layer(32)
        self.relu = nn.ReLU(True)
        self.sig = nn.Sigmoid()
        self.orig_size = [hparams.cin_channels, hparams.max_mel_lengths]
        self.upsample_mode = "bilinear"

    def _forward(self, net, x_size, video_net=None):
        feats = [ops.to_nhwc(f) for f in net]
        if self._uses_video:
            b = feats[-1]
            v = video_net.view(b.size(0), -1, b.size(1), b.size(2))           # reference :119 (NCHW view)
            x = ops.cat_channels(b, ops.to_nhwc(v))
            x = conv_norm_act(x, self.deconv1_1_1, self.deconv1_1_1_bn, ops.ACT_RELU)
        else:
            x = conv_norm_act(feats[-1], self.deconv1_1, self.deconv1_1_bn, ops.ACT_RELU)
        x = conv_norm_act(x, self.deconv1_2, self.deconv1_2_bn, ops.ACT_RELU)
        for i in range(1, len(feats)):
            tgt = feats[-1 - i]
            x = ops.bilinear_cat(x, (tgt.size(1), tgt.size(2)), tgt if i == self._cat_at else None)
            x = self._modules["convblock" + str(i + 1)].forward_nhwc(x)
        x = ops.bilinear_cat(x, (x_size[2], x_size[3]))
        x = conv_norm_act(x, self.conv6_1, self.conv6_1_bn, ops.ACT_RELU)
        x = conv_norm_act(x, self.conv6_2, None, ops.ACT_SIGMOID)
        return ops.to_nchw(x)

    def init_deconv_1_1_1(self):
        deconv1weight = self.deconv1_1.weight.unsqueeze(0)
        deconv1weight = deconv1weight.expand(2, 256, 256, 3, 3).contiguous()
        self.deconv1_1_1.weight.data = deconv1weight.view(512, 256, 3, 3)


class MelDecoder(_DecoderBase):
    def __init__(self, hparams=hparams, norm_layer=hparams.normlayer):
        super(MelDecoder, self).__init__()
        self._build(hparams, norm_layer, True, False, 64 * 2, 32, 4)

    def forward(self, net, x_size):
        return self._forward(net, x_size)


class MelDecoderImage(_DecoderBase):
    _uses_video = True

    def __init__(self, hparams=hparams, norm_layer=hparams.normlayer):
        super(MelDecoderImage, self).__init__()
        self._build(hparams, norm_layer, False, True, 64 * 2, 32, 4)

    def forward(self, net, x_size, video_net=None):
        return self._forward(net, x_size, video_net)


class MelDecoderImage2(_DecoderBase):
    _uses_video = True
    _cat_at = 4

    def __init__(self, hparams=hparams, norm_layer=hparams.normlayer):
        super(MelDecoderImage2, self).__init__()
        self._build(hparams, norm_layer, False, True, 64, 32 * 2, 2)

    def forward(self, net, x_size, video_net=None):
        return self._forward(net, x_size, video_net)


class MelDecoder_old(_DecoderBase):
    _cat_at = 4

    def __init__(self, hparams=hparams, norm_layer=hparams.normlayer):
        super(MelDecoder_old, self).__init__()
        self._build(hparams, norm_layer, True, False, 64, 32 * 2, 2)

    def forward(self, net, x_size):
        return self._forward(net, x_size)
