"""STFT -> mel kernel against the oracle's float64 restatement of utils/audio.py:70-144 (parity of the STFT itself is
UNPINNED against the reference: lws / librosa are unavailable; the framing KATs are the reference's own)."""
import numpy as np
import pytest
import torch

from oracle import viai_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T", [160000, 40960, 12345, 1024, 160, 1])
def test_melspectrogram_matches_oracle(T):
    from viai_b200.utils import audio
    rng = np.random.RandomState(T)
    t = np.arange(T) / 16000.0
    y = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 3000 * t * (1 + 0.1 * t)) + 0.05 * rng.randn(T)).astype(np.float32)
    want = O.melspectrogram(y.astype(np.float64))
    mel, mag = audio.melspectrogram_cuda(torch.from_numpy(y).cuda(), return_magnitude=True)
    assert tuple(mel.shape) == want.shape == (80, O.lws_num_frames(T, 1024, 160))
    got = mel.cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() <= 1e-3 * max(want.max(), 1e-6)        # north star: 1e-3 rel for fp32 spectrograms
    assert got.min() >= 0.0 and got.max() <= 1.0
    # |STFT| against numpy's rfft of the same frames
    l, r = O.lws_pad_lr(T, 1024, 160)
    yp = np.concatenate([np.zeros(l), y.astype(np.float64), np.zeros(r)])
    M = want.shape[1]
    idx = np.arange(1024)[None, :] + 160 * np.arange(M)[:, None]
    D = np.abs(np.fft.rfft(yp[idx] * O.lws_speech_window(1024, 160)[None, :], axis=1)).T
    assert np.abs(mag.cpu().numpy() - D).max() <= 2e-5 * max(D.max(), 1e-6)
    # numpy front door (reference signature)
    assert np.abs(audio.melspectrogram(y) - want).max() <= 1e-3


def test_framing_kats_and_silence():
    from viai_b200.utils import audio
    assert audio.lws_num_frames(160000, 1024, 160) == 1005 and audio.lws_num_frames(40960, 1024, 160) == 261   # SURVEY 8c KAT (3)
    assert audio.lws_pad_lr(np.zeros(160000), 1024, 160) == O.lws_pad_lr(160000, 1024, 160)
    mel = audio.melspectrogram_cuda(torch.zeros(16000, device="cuda"))
    assert float(mel.abs().max()) == 0.0                 # silence -> floor (-100 dB) -> 0 after normalisation
    with pytest.raises(RuntimeError, match="no CPU path"):
        audio.melspectrogram_cuda(torch.zeros(100))


def _stft_mel_with_basis(y, basis, n_mels):
    """viai_stft_mel through the C ABI with a caller-supplied filterbank (float32 (n_mels, 513)); returns (mel, |STFT|)."""
    import ctypes
    from viai_b200 import _lib
    from viai_b200.utils import audio
    L = _lib.lib()
    T = y.numel()
    M = audio.lws_num_frames(T, 1024, 160)
    win = torch.from_numpy(O.lws_speech_window(1024, 160).astype(np.float32)).cuda()
    span = np.zeros((n_mels, 2), dtype=np.int32)
    for r in range(n_mels):
        nz = np.nonzero(basis[r])[0]
        span[r] = (nz[0], nz[-1] + 1) if len(nz) else (0, 0)
    bt, st = torch.from_numpy(basis).contiguous().cuda(), torch.from_numpy(span).contiguous().cuda()
    out = torch.empty((n_mels, M), device="cuda")
    mag = torch.empty((513, M), device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.viai_stft_mel(P(y), T, 1024, 160, 1024 - 160, M, P(win), P(bt), P(st), n_mels, -100.0, 20.0, P(out), P(mag),
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "stft_mel")
    return out.cpu().numpy().astype(np.float64), mag.cpu().numpy().astype(np.float64)


def _mel_from_mag(mag, basis, min_level_db=-100.0, ref_level_db=20.0):
    S = basis.astype(np.float64) @ mag
    db = 20 * np.log10(np.maximum(np.exp(min_level_db / 20 * np.log(10)), S)) - ref_level_db
    return np.clip((db - min_level_db) / -min_level_db, 0, 1)


@pytest.mark.parametrize("kind", ["adjacent", "adjacent_gaps", "dense", "three_per_bin", "repeated_filter"])
@pytest.mark.parametrize("n_mels", [80, 37, 128])
def test_mel_stage_with_other_filterbanks(kind, n_mels):
    """viai_stft_mel takes ANY (n_mels, 513) filterbank with its per-filter non-zero span: triangular-like banks with random
    weights (two adjacent filters per bin, empty bins), a dense random matrix, three filters per bin, a filter whose support comes
    back after another filter's.  Against basis @ |STFT| of the kernel's own magnitudes in float64.  (Written for a bin-major mel
    stage -- per-bin weight tables + a segmented shuffle scan -- that was measured slower than the filter-per-lane walk and is not
    in the tree: 244 vs 262 M frames/s; the cases stay as coverage of the filterbank argument.)"""
    rng = np.random.RandomState(n_mels * 7 + len(kind))
    T = 16000
    y = torch.from_numpy((0.4 * rng.randn(T)).astype(np.float32)).cuda()
    B = np.zeros((n_mels, 513), dtype=np.float32)
    if kind in ("adjacent", "adjacent_gaps"):
        edges = np.sort(rng.choice(np.arange(5, 500), size=n_mels, replace=False))       # bin where filter r starts to be the lower one
        edges = np.concatenate([edges, [506]])
        for r in range(n_mels):
            for k in range(edges[r], edges[r + 1]):
                if kind == "adjacent_gaps" and rng.rand() < 0.15:
                    continue                                                            # a bin no filter looks at
                B[r, k] = rng.rand() + 0.1
                if r + 1 < n_mels and rng.rand() < 0.8:
                    B[r + 1, k] = rng.rand() + 0.1
    elif kind == "dense":
        B[:] = rng.rand(n_mels, 513) * (rng.rand(n_mels, 513) < 0.3)
    elif kind == "three_per_bin":
        for r in range(n_mels):
            lo = 3 + 3 * r
            B[r, lo:lo + 9] = rng.rand(9) + 0.1                                          # bins see filters r, r-1, r-2
    else:                                                                                # filter 0 returns after filter 1 in one round
        B[0, 10:14] = 0.5; B[1, 14:18] = 0.7; B[0, 18:22] = 0.9
        for r in range(2, n_mels):
            B[r, 30 + 3 * r:33 + 3 * r] = rng.rand(3) + 0.1
    got, mag = _stft_mel_with_basis(y, B, n_mels)
    want = _mel_from_mag(mag, B)
    assert np.abs(got - want).max() <= 2e-5, (kind, n_mels, np.abs(got - want).max())
    assert want.max() > 0.2                                                              # (not everything at the floor)
