"""Writer side of the loader's on-disk format (utils/librivox.py, reference :17-120): chunking into 8-second utterances, lws padding
and the N * hop cut, file naming and metadata tuples -- host logic, with the oracle's float64 mel as the mel function (the CUDA mel
kernel has its own parity tests); the GPU test (tests/test_librivox_gpu.py) runs the same through the real kernel."""
import os

import numpy as np
import pytest

from oracle import viai_oracle as O


def _signal(seconds, sr=16000, seed=0):
    t = np.arange(int(seconds * sr)) / sr
    rng = np.random.RandomState(seed)
    return (0.4 * np.sin(2 * np.pi * 220 * t) + 0.1 * rng.randn(t.size)).astype(np.float32)


def test_process_utterance_writes_the_loader_format(tmp_path):
    from viai_b200.utils import audio, librivox
    wav = _signal(19.3)                                   # 2 chunks: 8 s and 11.3 s (the last chunk runs to the end)
    rows = librivox._process_utterance(str(tmp_path), 7, "x.wav", "x.wav", mel_fn=lambda w: O.melspectrogram(w.astype(np.float64)),
                                       wav_whole=wav.copy())
    assert [r[0] for r in rows] == ["librivox-audio-0007-00000.npy", "librivox-audio-0007-00001.npy"]
    assert [r[1] for r in rows] == ["librivox-mel-0007-00000.npy", "librivox-mel-0007-00001.npy"]
    assert rows[1][3] == "x.wav - 00001"
    scaled = wav / np.abs(wav).max() * 0.999
    hop = audio.get_hop_size()
    for (af, mf, timesteps, _), (b, e) in zip(rows, ((0, 128000), (128000, None))):
        a, m = np.load(tmp_path / af), np.load(tmp_path / mf)
        chunk = scaled[b:e]
        N = O.lws_num_frames(len(chunk), 1024, hop)
        assert m.shape == (N, 80) and m.dtype == np.float32 and a.dtype == np.float32
        assert a.shape == (N * hop,) and timesteps == N * hop
        left, _ = O.lws_pad_lr(len(chunk), 1024, hop)
        assert np.all(a[:left] == 0) and np.array_equal(a[left:left + 1000], chunk[:1000])      # lws's left padding, then the signal
        assert np.abs(m - O.melspectrogram(chunk.astype(np.float64)).T).max() < 1e-6


def test_mulaw_restatement_and_file_type_guard(tmp_path):
    from viai_b200.utils import librivox
    x = np.linspace(-1, 1, 11)
    y = librivox.mulaw(x, 255)
    assert y[0] == -1 and y[-1] == 1 and abs(y[5]) == 0 and np.all(np.diff(y) > 0)
    q = librivox.mulaw_quantize(x, 255)
    assert q[0] == 0 and q[-1] == 255 and q[5] == 127          # the reference's silence level (|q - 127| > threshold)
    with pytest.raises(RuntimeError, match="only .wav"):
        librivox.load_wav("clip.mp3")
