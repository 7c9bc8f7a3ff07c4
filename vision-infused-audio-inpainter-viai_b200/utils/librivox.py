"""Offline preprocessing that WRITES the loader's on-disk format: drop-in for /root/reference/utils/librivox.py
(``build_from_path`` :17-41, ``_process_utterance`` :44-120).

A long recording is cut into 8-second utterances (the last one runs to the end of the file); every utterance becomes
``librivox-audio-%04d-%05d.npy`` (the waveform, padded by lws's left / right padding and cut to N * hop samples so that the WaveNet's
transposed convolutions can upsample N mel frames to it) and ``librivox-mel-%04d-%05d.npy`` ((N, num_mels) float32 in [0, 1]) --
exactly the files ``Data_loaders/audio_loader.py``'s ``_NPYDataSource`` reads.

What differs from the reference, on purpose:
  * the mel spectrogram of every utterance is computed on the GPU by the fused STFT -> mel kernel (``utils/audio.melspectrogram_cuda``);
    the reference's ProcessPoolExecutor over CPU lws / librosa calls is therefore not needed (``num_workers`` is accepted and ignored);
  * wav files are read with scipy (the reference's ``audio.load_wav`` is librosa, un-vendored); .ogg / .mp3 are refused loudly;
  * mu-law companding (the reference calls the un-vendored ``nnmnkwii.preprocessing.mulaw[_quantize]``) is restated from its
    published definition: y = sign(x) log(1 + mu |x|) / log(1 + mu), quantised as floor((y + 1) / 2 * mu), with mu =
    ``quantize_channels`` exactly as the reference passes it."""
import os

import numpy as np

from .. import Config
from . import audio

hparams = Config.Config()


def is_mulaw_quantize(s):
    return s == "mulaw-quantize"


def is_mulaw(s):
    return s == "mulaw"


def is_raw(s):
    return s == "raw"


def mulaw(x, mu=256):
    x = np.asarray(x, dtype=np.float64)
    return np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)


def mulaw_quantize(x, mu=256):
    return ((mulaw(x, mu) + 1) / 2 * mu).astype(np.int64)


def start_and_end_indices(quantized, silence_threshold=2):
    """utils/audio.py:39-48 of the reference (trimming of a mu-law-quantised signal)."""
    start = end = 0
    for start in range(quantized.size):
        if abs(quantized[start] - 127) > silence_threshold:
            break
    for end in range(quantized.size - 1, 1, -1):
        if abs(quantized[end] - 127) > silence_threshold:
            break
    assert abs(quantized[start] - 127) > silence_threshold
    assert abs(quantized[end] - 127) > silence_threshold
    return start, end


def load_wav(path):
    """float32 mono waveform in [-1, 1] at ``hparams.sample_rate`` (the file must already have that rate)."""
    if not path.lower().endswith(".wav"):
        raise RuntimeError("only .wav input is supported here (the reference decodes .ogg / .mp3 through librosa): %s" % path)
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if sr != hparams.sample_rate:
        raise RuntimeError("%s is sampled at %d Hz, hparams.sample_rate is %d (resample it first)" % (path, sr, hparams.sample_rate))
    if data.ndim > 1:
        data = data.mean(axis=1)
    if np.issubdtype(data.dtype, np.integer):
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    return data.astype(np.float32)


def _gpu_mel(wav):
    import torch
    return audio.melspectrogram_cuda(torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32)).cuda()).cpu().numpy()


def _process_utterance(out_dir, index, audio_filepath, text, mel_fn=None, wav_whole=None):
    """Reference :44-120.  ``mel_fn`` (waveform -> (num_mels, N) array in [0, 1]) defaults to the CUDA kernel; ``wav_whole`` lets a
    caller hand in an already decoded waveform."""
    mel_fn = mel_fn if mel_fn is not None else _gpu_mel
    if wav_whole is None:
        wav_whole = load_wav(audio_filepath)
    if hparams.rescaling:
        wav_whole = wav_whole / np.abs(wav_whole).max() * hparams.rescaling_max
    hop = audio.get_hop_size()
    n_samples = int(8.0 * hparams.sample_rate)            # all 8-second utterances
    n_chunks = wav_whole.shape[0] // n_samples
    results = []
    for chunk_idx in range(n_chunks):
        begin = chunk_idx * n_samples
        end = None if chunk_idx == n_chunks - 1 else begin + n_samples      # the last chunk runs to the end of the file
        wav = wav_whole[begin:end]
        if is_mulaw_quantize(hparams.input_type):
            out = mulaw_quantize(wav, hparams.quantize_channels)
            s, e = start_and_end_indices(out, hparams.silence_threshold)
            wav, out = wav[s:e], out[s:e]
            constant_values, out_dtype = mulaw_quantize(0, hparams.quantize_channels), np.int16
        elif is_mulaw(hparams.input_type):
            out = mulaw(wav, hparams.quantize_channels)
            constant_values, out_dtype = mulaw(0.0, hparams.quantize_channels), np.float32
        else:
            out, constant_values, out_dtype = wav, 0.0, np.float32
        mel = np.asarray(mel_fn(wav), dtype=np.float32).T                    # (N, num_mels)
        left, right = audio.lws_pad_lr(wav, hparams.fft_size, hop)           # the zeros lws pads internally
        out = np.pad(out, (left, right), mode="constant", constant_values=constant_values)
        N = mel.shape[0]
        assert len(out) >= N * hop
        out = out[:N * hop]                                                  # audio length = N * hop: upsampling by transposed convs
        assert len(out) % hop == 0
        audio_filename = "librivox-audio-%04d-%05d.npy" % (index, chunk_idx)
        mel_filename = "librivox-mel-%04d-%05d.npy" % (index, chunk_idx)
        np.save(os.path.join(out_dir, audio_filename), out.astype(out_dtype), allow_pickle=False)
        np.save(os.path.join(out_dir, mel_filename), mel, allow_pickle=False)
        results.append((audio_filename, mel_filename, len(out), "%s - %05d" % (text, chunk_idx)))
    return results


def build_from_path(in_dir, out_dir, num_workers=1, tqdm=lambda x: x):
    """Reference :17-41: every .wav of ``in_dir`` (sorted) -> utterance files in ``out_dir``; returns the metadata tuples
    (audio file, mel file, timesteps, text)."""
    os.makedirs(out_dir, exist_ok=True)
    files = [f for f in sorted(os.listdir(in_dir)) if any(f.endswith(ext) for ext in (".ogg", ".wav", ".mp3"))]
    out = []
    for index, f in enumerate(tqdm(files), start=1):
        path = os.path.join(in_dir, f)
        out.extend(_process_utterance(out_dir, index, path, path))
    return out
