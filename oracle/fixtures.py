"""Deterministic weights / inputs shared by oracle/make_golden.py and the tests.

TEST INFRASTRUCTURE ONLY.  Weights are a pure function of (key name, shape) so that golden vectors can be
committed without committing multi-megabyte state dicts: the golden script fills the *reference* modules'
state dicts with these values, the tests fill the oracle's / the CUDA modules' state dicts with the same.
"""
import math
import zlib

import torch


def _gen(name, salt=0):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) + 7919 * salt) & 0x7FFFFFFF)
    return g


def deterministic_fill(sd, salt=0):
    """Fills a state dict in place.  conv / linear weights ~ N(0, 2/fan_in) (so activations stay O(1)),
    norm scales in [0.5,1.5], biases / norm shifts in [-0.1,0.1], running_var in [0.5,1.5],
    weight_g in [0.5, 1.5] * ||v||-ish scale."""
    for k in sorted(sd.keys()):
        v = sd[k]
        if not v.is_floating_point():
            v.zero_()
            continue
        g = _gen(k, salt)
        leaf = k.split(".")[-1]
        if leaf in ("weight", "weight_v") and v.dim() >= 2:
            fan_in = v[0].numel() if v.dim() > 1 else v.numel()
            if "deconv" in k or "convblock" in k or "conv6" in k or "upsample_conv" in k:
                fan_in = v.size(0) * v[0][0].numel()       # ConvTranspose2d: (Cin, Cout, kh, kw)
            v.copy_(torch.randn(v.shape, generator=g) * math.sqrt(2.0 / max(fan_in, 1)))
            if "upsample_conv" in k:
                v.abs_()
        elif leaf == "weight_g":
            v.copy_(0.5 + torch.rand(v.shape, generator=g))
        elif leaf == "weight":                               # 1-d: norm scale
            v.copy_(0.5 + torch.rand(v.shape, generator=g))
        elif leaf == "running_var":
            v.copy_(0.5 + torch.rand(v.shape, generator=g))
        else:                                                # bias, running_mean
            v.copy_((torch.rand(v.shape, generator=g) - 0.5) * 0.2)
    return sd


def uniform(name, shape, lo=0.0, hi=1.0):
    return lo + (hi - lo) * torch.rand(shape, generator=_gen("input:" + name))


def normal(name, shape):
    return torch.randn(shape, generator=_gen("input:" + name))
