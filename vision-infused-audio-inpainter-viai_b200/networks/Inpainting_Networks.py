"""Generator encoder.  Drop-in for /root/reference/networks/Inpainting_Networks.py: same class names, constructor
signatures, ``state_dict()`` keys/shapes and ``forward`` contract; the arithmetic runs in libviai_b200.so."""
import torch.nn as nn

from .. import Options_inpainting, ops
from ._blocks import conv_norm_act
from .New_Inpainting_Networks import TransConvBlock  # noqa: F401  (the reference defines a copy here, :13-46)

hparams = Options_inpainting.Inpainting_Config()


class MelEncoder(nn.Module):
    """5 x [Conv2d 3x3 -> norm(affine) -> LeakyReLU(0.2)], strides (2,2),(2,1),(2,2),(2,2),(2,2), then
    AvgPool2d((3,1)) on the last map; returns the list of five maps (reference :49-88)."""

    def __init__(self, hparams=hparams, norm_layer=nn.BatchNorm2d):
        super(MelEncoder, self).__init__()
        use_bias = norm_layer == nn.InstanceNorm2d
        self.hparams = hparams
        self.conv1 = nn.Conv2d(1, 32, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1), bias=use_bias)
        self.bn1 = norm_layer(32, affine=True)
        self.conv2 = nn.Conv2d(32, 64, (3, 3), stride=(2, 1), padding=(1, 1), bias=use_bias)
        self.bn2 = norm_layer(64, affine=True)
        self.conv3 = nn.Conv2d(64, 128, (3, 3), (2, 2), (1, 1), bias=use_bias)
        self.bn3 = norm_layer(128, affine=True)
        self.conv4 = nn.Conv2d(128, 256, (3, 3), (2, 2), (1, 1), bias=use_bias)
        self.bn4 = norm_layer(256, affine=True)
        self.conv5 = nn.Conv2d(256, 256, (3, 3), (2, 2), (1, 1), bias=use_bias)
        self.bn5 = norm_layer(256, affine=True)
        self.avgpool = nn.AvgPool2d((3, 1))
        self.initial()

    def forward(self, c):
        B = c.size(0)
        x = c.reshape(B, self.hparams.cin_channels, -1, 1)       # NHWC with C == 1 (same memory as (B,1,H,W))
        feats = []
        for i in range(1, 6):
            x = conv_norm_act(x, self._modules["conv%d" % i], self._modules["bn%d" % i], ops.ACT_LRELU, 0.2)
            feats.append(x)
        feats[-1] = ops.avgpool_h(feats[-1], 3)
        return [ops.to_nchw(f) for f in feats]

    def initial(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
