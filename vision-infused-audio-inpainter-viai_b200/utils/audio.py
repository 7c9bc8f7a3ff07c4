"""STFT -> mel front end.  Drop-in for the feature path of /root/reference/utils/audio.py (``melspectrogram`` :70-75,
``get_hop_size`` :78-83, ``lws_num_frames`` :90-98, ``lws_pad_lr`` :101-108, ``_build_mel_basis`` :123-127, ``_amp_to_db``
:130-132, ``_db_to_amp`` :135-136, ``_normalize`` :139-140, ``_denormalize`` :143-144).

The reference computes the STFT with the un-vendored ``lws`` C extension and the filterbank with ``librosa``; neither is
available, so the analysis window and the Slaney filterbank are built here explicitly (host side, float64,
once per configuration) and the whole chain frame -> window -> rFFT -> |.| -> mel -> dB -> normalise runs in ONE CUDA kernel
(csrc/stft_mel.cu).  File I/O, trimming and ``adjust_time_resolution`` are out of scope."""
import ctypes
import math

import numpy as np
import torch

from .. import Config, _lib

hparams = Config.Config()
_cache = {}


def get_hop_size():
    hop_size = hparams.hop_size
    if hop_size is None:
        assert hparams.frame_shift_ms is not None
        hop_size = int(hparams.frame_shift_ms / 1000 * hparams.sample_rate)
    return hop_size


def lws_num_frames(length, fsize, fshift):
    """Compute number of time frames of lws spectrogram (reference :90-98)."""
    pad = (fsize - fshift)
    if length % fshift == 0:
        M = (length + pad * 2 - fsize) // fshift + 1
    else:
        M = (length + pad * 2 - fsize) // fshift + 2
    return M


def lws_pad_lr(x, fsize, fshift):
    """Compute left and right padding lws internally uses (reference :101-108)."""
    M = lws_num_frames(len(x), fsize, fshift)
    pad = (fsize - fshift)
    T = len(x) + 2 * pad
    r = (M - 1) * fshift + fsize - T
    return pad, pad + r


def _lws_window(fsize, fshift):
    """Analysis window of ``lws.lws(fsize, fshift, mode='speech')``: sqrt(periodic Hann * 2*fshift/fsize)."""
    n = np.arange(fsize, dtype=np.float64)
    return np.sqrt(0.5 * (1.0 - np.cos(2.0 * np.pi * n / fsize)) * 2.0 * fshift / fsize)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, np.log(6.4) / 27.0
    return np.where(m >= min_log_hz / f_sp, min_log_hz * np.exp(logstep * (m - min_log_hz / f_sp)), f_sp * m)


def _build_mel_basis():
    """librosa.filters.mel(sr, n_fft, fmin=, fmax=, n_mels=) with its defaults (Slaney scale, Slaney area norm)."""
    assert hparams.fmax <= hparams.sample_rate // 2
    sr, n_fft, n_mels = hparams.sample_rate, hparams.fft_size, hparams.num_mels
    fftfreqs = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(hparams.fmin), _hz_to_mel(hparams.fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.maximum(0, np.minimum(-ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]))
    return w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]


def _amp_to_db(x):
    min_level = np.exp(hparams.min_level_db / 20 * np.log(10))
    return 20 * np.log10(np.maximum(min_level, x))


def _db_to_amp(x):
    return np.power(10.0, x * 0.05)


def _normalize(S):
    return np.clip((S - hparams.min_level_db) / -hparams.min_level_db, 0, 1)


def _denormalize(S):
    return (np.clip(S, 0, 1) * -hparams.min_level_db) + hparams.min_level_db


def _device_tables(device):
    key = (str(device), hparams.sample_rate, hparams.fft_size, get_hop_size(), hparams.num_mels, hparams.fmin, hparams.fmax)
    t = _cache.get(key)
    if t is None:
        basis = _build_mel_basis()
        span = np.zeros((basis.shape[0], 2), dtype=np.int32)
        for r in range(basis.shape[0]):
            nz = np.nonzero(basis[r])[0]
            span[r] = (nz[0], nz[-1] + 1) if len(nz) else (0, 0)
        t = (torch.from_numpy(_lws_window(hparams.fft_size, get_hop_size()).astype(np.float32)).to(device),
             torch.from_numpy(basis.astype(np.float32)).contiguous().to(device), torch.from_numpy(span).contiguous().to(device))
        _cache[key] = t
    return t


def melspectrogram_cuda(y, return_magnitude=False):
    """y: 1-D float32 CUDA tensor (one waveform).  Returns the normalised mel spectrogram (num_mels, M) as a CUDA tensor
    [and |STFT| (fft_size/2+1, M)]: the device-resident form the GAN step consumes."""
    if not (torch.is_tensor(y) and y.is_cuda and y.dtype == torch.float32 and y.dim() == 1):
        raise RuntimeError("melspectrogram_cuda needs a 1-D float32 CUDA tensor; there is no CPU path")
    y = y.contiguous()
    L = _lib.lib()
    fsize, hop = hparams.fft_size, get_hop_size()
    T = y.numel()
    M = lws_num_frames(T, fsize, hop)
    win, basis, span = _device_tables(y.device)
    out = torch.empty((hparams.num_mels, M), device=y.device, dtype=torch.float32)
    mag = torch.empty((fsize // 2 + 1, M), device=y.device, dtype=torch.float32) if return_magnitude else None
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    _lib.check(L.viai_stft_mel(P(y), T, fsize, hop, fsize - hop, M, P(win), P(basis), P(span), hparams.num_mels,
                               float(hparams.min_level_db), float(hparams.ref_level_db), P(out), P(mag),
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "stft_mel")
    return (out, mag) if return_magnitude else out


def melspectrogram(y):
    """Reference signature (:70): numpy waveform in, numpy (num_mels, M) in [0,1] out -- computed on the GPU."""
    yt = torch.as_tensor(np.asarray(y, dtype=np.float32)).cuda()
    S = melspectrogram_cuda(yt)
    if not hparams.allow_clipping_in_normalization:
        assert float(S.max()) <= 1 and float(S.min()) >= 0
    return S.cpu().numpy()
