"""Discretized mixture of logistics.  Drop-in for ``discretized_mix_logistic_loss`` of
/root/reference/wavenet_vocoder/mixture.py:25-105 (forward value and gradient come from one CUDA kernel, csrc/wavenet_train.cu
``viai_dmol_nll``).  Sampling (mixture.py:117-153) runs inside the synthesis kernel (csrc/wavenet_synth.cu)."""
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
