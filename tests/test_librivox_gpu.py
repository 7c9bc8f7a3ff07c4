"""utils/librivox.build_from_path end to end on the GPU: wav files in, the loader's ``librivox-{audio,mel}-*.npy`` files out, mel
spectrograms from the fused CUDA kernel (<= 1e-3 from the oracle), readable by the product's own loader sources."""
import os

import numpy as np
import pytest

from oracle import viai_oracle as O

pytestmark = pytest.mark.gpu


def test_build_from_path_writes_files_the_loader_reads(tmp_path):
    from scipy.io import wavfile
    from viai_b200.utils import audio, librivox
    sr = 16000
    t = np.arange(int(17.0 * sr)) / sr
    wav = (0.5 * np.sin(2 * np.pi * 330 * t) * (1 + 0.3 * np.sin(2 * np.pi * 2 * t))).astype(np.float32)
    in_dir, out_dir = tmp_path / "in", tmp_path / "out"
    in_dir.mkdir()
    wavfile.write(str(in_dir / "a.wav"), sr, (wav * 32767).astype(np.int16))
    (in_dir / "notes.txt").write_text("ignored")
    rows = librivox.build_from_path(str(in_dir), str(out_dir))
    assert len(rows) == 2 and rows[0][0] == "librivox-audio-0001-00000.npy" and rows[1][1] == "librivox-mel-0001-00001.npy"
    hop = audio.get_hop_size()
    decoded = (wav * 32767).astype(np.int16).astype(np.float32) / 32768.0
    scaled = decoded / np.abs(decoded).max() * 0.999
    for (af, mf, timesteps, _), (b, e) in zip(rows, ((0, 128000), (128000, None))):
        a, m = np.load(out_dir / af), np.load(out_dir / mf)
        want = O.melspectrogram(scaled[b:e].astype(np.float64)).T
        assert m.shape == want.shape and a.shape == (m.shape[0] * hop,) and timesteps == a.shape[0]
        assert np.abs(m - want).max() <= 1e-3
        assert m.min() >= 0 and m.max() <= 1
