"""Visual encoder.  Drop-in for /root/reference/networks/Image_Embedding.py: ``ResNet`` :13-71, ``ImageResnet18`` :74-84,
``FlowResnet18`` :87-97, ``ImageEmbedding`` :100-126 -- same constructors, ``state_dict()`` keys/shapes and forward contract.
``ImageEmbedding_single`` :128-148, ``ImageEmbedding_finetune`` :150-171 and ``ImageEmbedding2`` :174-200 are the visual
branches of the AV-sync heads."""
import math

import torch
import torch.nn as nn

from .. import Options_inpainting, ops
from ._blocks import conv_norm_act
from .ResNet import BasicBlock

hparams = Options_inpainting.Inpainting_Config()


def copy_state_dict(state_dict, model, strip=None):
    """Shape-tolerant loader of /root/reference/utils/util.py:124-144 (used for the ImageNet pre-train)."""
    tgt = model.state_dict()
    for name, param in state_dict.items():
        if strip is not None and name.startswith(strip):
            name = name[len(strip):]
        if name not in tgt or tgt[name].size() != param.size():
            continue
        tgt[name].copy_(param.data if isinstance(param, nn.Parameter) else param)
    return model


class ResNet(nn.Module):
    def __init__(self, block, layers, channel_size=3, length_feature=hparams.length_feature):
        self.inplanes = 64
        super(ResNet, self).__init__()
        self.conv1 = nn.Conv2d(channel_size, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AvgPool2d(7, stride=1)
        self.fc = nn.Linear(512 * block.expansion, length_feature)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for i in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        """x: (N, channel_size, S, S) with S such that the last map is 7x7 (S = 224).  Returns (N, length_feature)."""
        x = ops.to_nhwc(x)
        x = conv_norm_act(x, self.conv1, self.bn1, ops.ACT_RELU)
        x = ops.maxpool3s2(x)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x = blk.forward_nhwc(x)
        N, H, W, C = x.shape
        if H != 7 or W != 7:
            raise RuntimeError("ResNet expects a 7x7 final map (image_size 224), got %dx%d" % (H, W))
        x = ops.avgpool_h(x, 7)                                   # AvgPool2d(7): rows ...
        x = ops.avgpool_h(x.reshape(N, 7, 1, C), 7)               # ... then columns
        x = ops.conv2d(x.reshape(1, 1, N, C), self.fc.weight.reshape(self.fc.out_features, C, 1, 1), self.fc.bias)
        return x.reshape(N, self.fc.out_features)


def ImageResnet18(hparams=hparams):
    """Constructs a ResNet-18 model."""
    model = ResNet(BasicBlock, [2, 2, 2, 2], length_feature=hparams.length_feature)
    if hparams.resnet_pretrain:
        pretrain = torch.load(hparams.resnet_pretrain_path)
        copy_state_dict(pretrain, model)
    return model


def FlowResnet18(hparams=hparams):
    model = ResNet(BasicBlock, [2, 2, 2, 2], channel_size=2, length_feature=hparams.length_feature)
    if hparams.resnet_pretrain:
        pretrain = torch.load(hparams.resnet_pretrain_path)
        model = copy_state_dict(pretrain, model)
        conv1_weight = pretrain["conv1.weight"].data
        flow_conv1_weight = conv1_weight.mean(1).unsqueeze(1).expand([64, 2, 7, 7]).contiguous()
        model.conv1.weight.data = flow_conv1_weight
    return model


class ImageEmbedding(nn.Module):
    """Frames + optical flow -> two ResNet-18 streams -> per-frame features concatenated over channels -> temporal head
    (two stride-2 Conv1d) -> (B, length_feature, 1, T/4) (reference :100-126)."""

    def __init__(self, hparams=hparams):
        super(ImageEmbedding, self).__init__()
        self.hparams = hparams
        self.image_single_model, self.flow_single_model = ImageResnet18(hparams), FlowResnet18(hparams)
        _add_temporal_head(self, 2 * hparams.length_feature, 2 * hparams.length_feature, hparams.length_feature)

    @staticmethod
    def _conv1d(x, conv):
        """x: (B, 1, T, C) NHWC; nn.Conv1d (k=3, s=2, p=1) as a 1x3 convolution."""
        return ops.conv2d(x, conv.weight.unsqueeze(2), conv.bias, (1, conv.stride[0]), (0, conv.padding[0]))

    def forward(self, video_block, flow_block):
        fea_cat = ops.cat_channels(_per_frame_features(self.image_single_model, video_block, 3, self.hparams),
                                   _per_frame_features(self.flow_single_model, flow_block, 2, self.hparams))      # (B, 1, T, 2F)
        # the reference evaluates relu(bn_1(out)) and discards it (:123): only bn_1's running statistics change
        return _temporal_convs(self, fea_cat, True).permute(0, 3, 1, 2)                                            # (B, F, 1, T/4)


def _add_temporal_head(mod, cin, mid, cout):
    """Registers the reference's temporal head on ``mod`` under its names: conv_1 (k3 s2 p1) / bn_1 / conv_2 (k3 s2 p1) / bn_2 / relu."""
    for idx, (a, b) in enumerate(((cin, mid), (mid, cout)), start=1):
        mod.add_module("conv_%d" % idx, torch.nn.Conv1d(a, b, 3, 2, 1, bias=False))
        mod.add_module("bn_%d" % idx, nn.BatchNorm1d(b))
    mod.relu = nn.ReLU(True)


def _temporal_convs(mod, fea, discard_bn):
    """conv_1 -> [relu(bn_1(.)) evaluated and discarded, as the reference does] -> conv_2 on (B, 1, T, C) rows."""
    out = ImageEmbedding._conv1d(fea, mod.conv_1)
    if discard_bn and mod.bn_1.training:
        with torch.no_grad():
            ops.norm_act(out.detach(), mod.bn_1, "bn", ops.ACT_RELU)        # only bn_1's running statistics change
    return ImageEmbedding._conv1d(out, mod.conv_2)


def _per_frame_features(net, block, channels, hp):
    """(B, T, channels, S, S) frames -> (B, 1, T, length_feature) rows through one ResNet-18 stream."""
    B = block.size(0)
    return net(block.reshape(-1, channels, hp.image_size, hp.image_size)).reshape(B, 1, -1, hp.length_feature)


class ImageEmbedding_single(nn.Module):
    """One visual stream (frames when ``image``, else flow) + the temporal head; (B, T, c, S, S) -> (B, length_feature, T/4)
    (reference :128-148)."""

    def __init__(self, hparams=hparams, image=1):
        super(ImageEmbedding_single, self).__init__()
        self.image, self.hparams = image, hparams
        self.image_single_model = (ImageResnet18 if image else FlowResnet18)(hparams)
        _add_temporal_head(self, hparams.length_feature, hparams.length_feature, hparams.length_feature)

    def forward(self, video_block):
        fea = _per_frame_features(self.image_single_model, video_block, 3 if self.image else 2, self.hparams)
        return _temporal_convs(self, fea, True).squeeze(1).permute(0, 2, 1)


class ImageEmbedding_finetune(nn.Module):
    """The temporal head alone on precomputed per-frame features (B, T, length_feature) -> (B, F, 1, T/4) (reference :150-171)."""

    def __init__(self, hparams=hparams, image=1):
        super(ImageEmbedding_finetune, self).__init__()
        self.image, self.hparams = image, hparams
        _add_temporal_head(self, hparams.length_feature, hparams.length_feature, hparams.length_feature)

    def forward(self, image_out):
        return _temporal_convs(self, image_out.unsqueeze(1), True).permute(0, 3, 1, 2)


class ImageEmbedding2(nn.Module):
    """ImageEmbedding that also hands back the concatenated per-frame features: (out (B, F, 1, T/4), fea_cat (B, 2F, T))
    (reference :174-200; unlike ImageEmbedding.forward it does not evaluate bn_1 -- that call is commented out at :195)."""

    def __init__(self, hparams=hparams):
        super(ImageEmbedding2, self).__init__()
        self.hparams = hparams
        self.image_single_model, self.flow_single_model = ImageResnet18(hparams), FlowResnet18(hparams)
        _add_temporal_head(self, 2 * hparams.length_feature, 2 * hparams.length_feature, hparams.length_feature)

    def forward(self, video_block, flow_block):
        fea_cat = ops.cat_channels(_per_frame_features(self.image_single_model, video_block, 3, self.hparams),
                                   _per_frame_features(self.flow_single_model, flow_block, 2, self.hparams))      # (B, 1, T, 2F)
        out = _temporal_convs(self, fea_cat, False)
        return out.permute(0, 3, 1, 2), fea_cat.squeeze(1).permute(0, 2, 1)
