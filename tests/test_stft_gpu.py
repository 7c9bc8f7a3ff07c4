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
