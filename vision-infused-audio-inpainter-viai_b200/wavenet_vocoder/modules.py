"""Parameter containers of the WaveNet layers.  Drop-in for /root/reference/wavenet_vocoder/modules.py ``Conv1d`` :22-32,
``ConvTranspose2d`` :41-51, ``Conv1d1x1`` :54-63 and ``ResidualConv1dGLU`` :84-216: same constructors, weight-norm
parametrisation (``weight_g`` / ``weight_v``) and ``state_dict()`` layout.  The synthesis arithmetic itself lives in
csrc/wavenet_synth.cu and is driven by ``WaveNet.incremental_forward``; the modules keep no incremental buffers (the kernel
owns the ring buffers for the duration of one call), so ``clear_buffer`` is a no-op kept for API compatibility.  The
teacher-forced training path (``forward``) runs per layer: csrc/wavenet_train.cu + the implicit-GEMM convolution kernels."""
import math

import torch
from torch import nn

from .. import ops


def Conv1d(in_channels, out_channels, kernel_size=1, padding=0, dilation=1, bias=True, weight_normalization=True,
           dropout=0, std_mul=1.0, **kwargs):
    m = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, padding=padding, dilation=dilation, bias=bias, **kwargs)
    if weight_normalization:
        assert bias
        std = math.sqrt((std_mul * (1.0 - dropout)) / (m.kernel_size[0] * in_channels))
        m.weight.data.normal_(mean=0, std=std)
        m.bias.data.zero_()
        return nn.utils.weight_norm(m)
    return m


def ConvTranspose2d(in_channels, out_channels, kernel_size, weight_normalization=True, **kwargs):
    freq_axis_kernel_size = kernel_size[0]
    m = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, **kwargs)
    m.weight.data.fill_(1.0 / freq_axis_kernel_size)
    m.bias.data.zero_()
    if weight_normalization:
        return nn.utils.weight_norm(m)
    return m


def Conv1d1x1(in_channels, out_channels, bias=True, weight_normalization=True):
    """1-by-1 convolution layer"""
    if weight_normalization:
        assert bias
        return Conv1d(in_channels, out_channels, kernel_size=1, padding=0, dilation=1, bias=bias, std_mul=1.0)
    return nn.Conv1d(in_channels, out_channels, kernel_size=1, padding=0, dilation=1, bias=bias)


def effective_weight(m):
    """The weight a (possibly weight-normalised) module applies: g * v / ||v|| per slice of dim 0 -- what
    ``make_generation_fast_`` (wavenet.py:387-393) folds once."""
    if hasattr(m, "weight_g"):
        return ops.weight_norm(m.weight_v, m.weight_g)
    return m.weight


class ResidualConv1dGLU(nn.Module):
    """Residual dilated conv1d + gated linear unit (reference :84-216)."""

    def __init__(self, residual_channels, gate_channels, kernel_size, skip_out_channels=None, cin_channels=-1, gin_channels=-1,
                 dropout=1 - 0.95, padding=None, dilation=1, causal=True, bias=True, weight_normalization=True, *args, **kwargs):
        super(ResidualConv1dGLU, self).__init__()
        self.dropout = dropout
        if skip_out_channels is None:
            skip_out_channels = residual_channels
        if padding is None:
            padding = (kernel_size - 1) * dilation if causal else (kernel_size - 1) // 2 * dilation
        self.causal = causal
        self.dilation = dilation
        if weight_normalization:
            assert bias
            self.conv = Conv1d(residual_channels, gate_channels, kernel_size, dropout=dropout, padding=padding,
                               dilation=dilation, bias=bias, std_mul=1.0, *args, **kwargs)
        else:
            self.conv = nn.Conv1d(residual_channels, gate_channels, kernel_size, padding=padding, dilation=dilation, bias=bias,
                                  *args, **kwargs)
        self.conv1x1c = Conv1d1x1(cin_channels, gate_channels, bias=bias,
                                  weight_normalization=weight_normalization) if cin_channels > 0 else None
        if gin_channels > 0:
            raise NotImplementedError("global conditioning (gin_channels > 0) is outside the VIAI hot path")
        self.conv1x1g = None
        gate_out_channels = gate_channels // 2
        self.conv1x1_out = Conv1d1x1(gate_out_channels, residual_channels, bias=bias, weight_normalization=weight_normalization)
        self.conv1x1_skip = Conv1d1x1(gate_out_channels, skip_out_channels, bias=bias, weight_normalization=weight_normalization)

    # ---- teacher-forced (T-parallel) path: reference ``_forward(x, c, g, is_incremental=False)`` :162-210 ------------------
    def forward_rows(self, x, c=None):
        """x (B, T, R), c (B, T, Cc) or None, channels innermost -> (x_out (B, T, R), skip (B, T, S)).
        The dilated causal convolution and the conditioning 1x1 are ONE GEMM over the B*T rows: the operand is
        [x(t-2d) | x(t-d) | x(t) | c(t)] (ops.shiftcat) and the weight the tap-major linearised one of conv.py:51-62."""
        if not self.causal:
            raise NotImplementedError("non-causal ResidualConv1dGLU is outside the VIAI hot path")
        if (c is None) != (self.conv1x1c is None):
            raise RuntimeError("local conditioning features and conv1x1c go together (modules.py:184)")
        residual = x
        mask, scale = None, 1.0
        if self.training and self.dropout > 0:                      # F.dropout(x, p, training) :173, applied while gathering
            keep = 1.0 - self.dropout
            mask, scale = torch.empty_like(x).bernoulli_(keep), 1.0 / keep
        K = self.conv.kernel_size[0]
        G, R = self.conv.out_channels, self.conv.in_channels
        Cc = c.size(2) if c is not None else 0
        Kpad = (K * R + Cc + 31) // 32 * 32
        X = ops.shiftcat(x, c, K, self.dilation, Kpad, mask, scale)
        lin = effective_weight(self.conv).permute(0, 2, 1).reshape(G, K * R)
        b = self.conv.bias
        if c is not None:
            lin = torch.cat((lin, effective_weight(self.conv1x1c).reshape(G, Cc)), 1)
            b = b + self.conv1x1c.bias
        if Kpad > K * R + Cc:
            lin = torch.cat((lin, lin.new_zeros(G, Kpad - K * R - Cc)), 1)
        z = ops.glu_tanh_sigmoid(rows_linear(X, lin, b))            # tanh(a) * sigmoid(b) :196
        s = rows_linear(z, effective_weight(self.conv1x1_skip).reshape(self.conv1x1_skip.out_channels, -1), self.conv1x1_skip.bias)
        o = rows_linear(z, effective_weight(self.conv1x1_out).reshape(R, -1), self.conv1x1_out.bias)
        return ops.axpby(o, math.sqrt(0.5), residual, math.sqrt(0.5)), s

    def forward(self, x, c=None, g=None):
        """Reference layout: x (B, R, T), c (B, Cc, T) -> (x (B, R, T), s (B, S, T))."""
        if g is not None:
            raise NotImplementedError("global conditioning is outside the VIAI hot path")
        xo, s = self.forward_rows(x.transpose(1, 2).contiguous(), None if c is None else c.transpose(1, 2).contiguous())
        return xo.transpose(1, 2), s.transpose(1, 2)

    def incremental_forward(self, x, c=None, g=None):
        raise RuntimeError("ResidualConv1dGLU's incremental path runs inside the fused synthesis kernel "
                           "(WaveNet.incremental_forward); there is no per-layer incremental path")

    def clear_buffer(self):
        pass


def _fold(P):
    """Width of the 2-D view of P rows handed to the implicit-GEMM kernels (their M tile is 16 x 8 pixels)."""
    for w in (128, 64, 32, 16, 8, 4, 2):
        if P % w == 0:
            return w
    return 1


def rows_linear(x, weight, bias=None):
    """x (B, T, Cin) @ weight (Cout, Cin)^T + bias as a 1x1 convolution over the B*T rows (F.conv1d with kernel 1)."""
    B, T, C = x.shape
    W = _fold(B * T)
    y = ops.conv2d(x.reshape(1, (B * T) // W, W, C), weight.reshape(weight.size(0), C, 1, 1), bias)
    return y.reshape(B, T, weight.size(0))
