"""Shared conv + norm + activation step on NHWC tensors (host glue; the arithmetic is in libviai_b200.so)."""
import torch.nn as nn

from .. import ops


def norm_kind(norm_layer):
    if norm_layer is nn.BatchNorm2d or isinstance(norm_layer, nn.BatchNorm2d):
        return "bn"
    if norm_layer is nn.InstanceNorm2d or isinstance(norm_layer, nn.InstanceNorm2d):
        return "in"
    raise ValueError("VIAI-B200 supports nn.BatchNorm2d and nn.InstanceNorm2d as norm_layer, got %r" % (norm_layer,))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv_norm_act(x, conv, norm, act, slope=0.0):
    """x: NHWC.  ``conv`` is an nn.Conv2d / nn.ConvTranspose2d used as a parameter container (its own forward is never
    called); ``norm`` an nn.BatchNorm2d / nn.InstanceNorm2d container or None."""
    transposed = isinstance(conv, nn.ConvTranspose2d)
    if transposed and (_pair(conv.output_padding) != (0, 0) or _pair(conv.dilation) != (1, 1) or conv.groups != 1):
        raise RuntimeError("unsupported ConvTranspose2d configuration on the VIAI hot path")
    if not transposed and (_pair(conv.dilation) != (1, 1) or conv.groups != 1):
        raise RuntimeError("unsupported Conv2d configuration on the VIAI hot path")
    if norm is None:
        y = ops.conv2d(x, conv.weight, conv.bias, _pair(conv.stride), _pair(conv.padding), transposed)
        return ops.norm_act(y, None, "none", act, slope)
    kind = norm_kind(norm)
    y, stats = ops.conv2d_stats(x, conv.weight, conv.bias, _pair(conv.stride), _pair(conv.padding), transposed,
                                ops.stat_groups_for(norm, kind, x.size(0)))
    return ops.norm_act(y, norm, kind, act, slope, pre_stats=stats if stats.numel() else None)
