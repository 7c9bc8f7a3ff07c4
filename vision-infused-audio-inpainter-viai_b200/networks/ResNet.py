"""ResNet building blocks.  Drop-in for /root/reference/networks/ResNet.py ``conv3x3`` :20-23 and ``BasicBlock`` :26-55
(the only parts VIAI uses; ``Bottleneck`` and the model-zoo constructors :58-214 are out of scope).
The arithmetic runs in libviai_b200.so on NHWC tensors."""
import torch.nn as nn

from .. import ops
from ._blocks import conv_norm_act


def conv3x3(in_planes, out_planes, stride=1):
    """3x3 convolution with padding"""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward_nhwc(self, x):
        residual = x
        out = conv_norm_act(x, self.conv1, self.bn1, ops.ACT_RELU)
        out = conv_norm_act(out, self.conv2, self.bn2, ops.ACT_NONE)
        if self.downsample is not None:
            residual = conv_norm_act(x, self.downsample[0], self.downsample[1], ops.ACT_NONE)
        return ops.add_act(out, residual, ops.ACT_RELU)

    def forward(self, x):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))
