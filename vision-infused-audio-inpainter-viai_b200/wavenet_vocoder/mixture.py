"""Discretized mixture of logistics.  Drop-in for /root/reference/wavenet_vocoder/mixture.py: ``discretized_mix_logistic_loss``
:25-105 (value and gradient from one CUDA kernel, csrc/wavenet_train.cu ``viai_dmol_nll``) and
``sample_from_discretized_mix_logistic`` :117-153 (``viai_dmol_sample``; the synthesis kernel csrc/wavenet_synth.cu has the same
arithmetic inline for the autoregressive loop)."""
import torch

from .. import ops


def discretized_mix_logistic_loss(y_hat, y, num_classes=256, log_scale_min=-7.0, reduce=True):
    """y_hat (B, C, T) with C = 3 * nr_mix, y (B, T, 1) in [-1, 1].  reduce=True: -sum(log-likelihood) (scalar);
    reduce=False: per-sample losses (B, T, 1)."""
    assert y_hat.dim() == 3
    assert y_hat.size(1) % 3 == 0
    rows = y_hat.transpose(1, 2)                                  # (B, T, C); free when y_hat came out of WaveNet.forward
    nll = ops.dmol_nll(rows, y.reshape(rows.shape[0], rows.shape[1]), num_classes, log_scale_min)
    if reduce:
        return ops.masked_sum(nll, None, mean=False)
    return nll.unsqueeze(-1)


def sample_from_discretized_mix_logistic(y, log_scale_min=-7.0, uniforms=None):
    """y (B, C, T) -> samples (B, T) in [-1, 1].  ``uniforms`` (B, T, nr_mix + 1) supplies the draws of :136,148 (default: fresh
    uniform(1e-5, 1 - 1e-5) draws, as in the reference)."""
    assert y.size(1) % 3 == 0
    nr_mix = y.size(1) // 3
    rows = y.transpose(1, 2)
    if uniforms is None:
        uniforms = torch.empty(rows.shape[:2] + (nr_mix + 1,), device=y.device).uniform_(1e-5, 1.0 - 1e-5)
    return ops.dmol_sample(rows, uniforms, log_scale_min)
