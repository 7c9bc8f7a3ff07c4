"""Generator encoder.  Drop-in for /root/reference/networks/Inpainting_Networks.py: same class names, constructor
signatures, ``state_dict()`` keys/shapes and ``forward`` contract; the arithmetic runs in libviai_b200.so."""
import torch.nn as nn

from .. import Options_inpainting, ops
from ._blocks import conv_norm_act
from .New_Inpainting_Networks import TransConvBlock  # noqa: F401  (the reference defines a copy here, :13-46)

hparams = Options_inpainting.Inpainting_Config()


class MelEncoder(nn.Module):
    """Five 3x3 stages -- (channels, stride): (32, (2,2)), (64, (2,1)), (128, (2,2)), (256, (2,2)), (256, (2,2)) -- each
    Conv2d -> norm(affine) -> LeakyReLU(0.2); AvgPool2d((3,1)) on the last map; returns the list of five maps
    (reference :49-88; parameters ``conv{i}`` / ``bn{i}``, a conv bias only with InstanceNorm)."""
    STAGES = ((32, (2, 2)), (64, (2, 1)), (128, (2, 2)), (256, (2, 2)), (256, (2, 2)))

    def __init__(self, hparams=hparams, norm_layer=nn.BatchNorm2d):
        super(MelEncoder, self).__init__()
        self.hparams = hparams
        cin = 1
        for i, (cout, stride) in enumerate(self.STAGES, start=1):
            self.add_module("conv%d" % i, nn.Conv2d(cin, cout, 3, stride, 1, bias=norm_layer == nn.InstanceNorm2d))
            self.add_module("bn%d" % i, norm_layer(cout, affine=True))
            cin = cout
        self.avgpool = nn.AvgPool2d((3, 1))
        self.initial()

    def forward(self, c):
        B = c.size(0)
        x = c.reshape(B, self.hparams.cin_channels, -1, 1)       # NHWC with C == 1 (same memory as (B,1,H,W))
        feats = []
        for i in range(1, len(self.STAGES) + 1):
            x = conv_norm_act(x, self._modules["conv%d" % i], self._modules["bn%d" % i], ops.ACT_LRELU, 0.2)
            feats.append(x)
        feats[-1] = ops.avgpool_h(feats[-1], 3)
        return [ops.to_nchw(f) for f in feats]

    def initial(self):
        """kaiming-normal (fan_out) convolution weights, zero biases, unit BatchNorm scales -- drawn in module order, as upstream."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1.0)
                m.bias.data.zero_()
