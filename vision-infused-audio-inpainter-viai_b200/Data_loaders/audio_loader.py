"""On-disk formats and the video-frame path of the loader.  Mirrors the parts of
/root/reference/Data_loaders/audio_loader.py that define WHAT is on disk and what a batch looks like:

  * ``{train,test}<new_split_name>`` metadata: one ``|``-separated line per clip -- ``image-dir | mel.npy | audio.npy | speaker |
    length`` (4 or 5 columns; the last is the length, audio lengths are ``length * 1280``) -- ``_NPYDataSource.collect_files``
    :63-107, ``ImageDataSource.collect_files`` :127-149;
  * frame folders ``image_crop/ flow_x_crop/ flow_y_crop/ {n}.jpg`` -- ``sample_data_new`` :155-246, ``load_image`` :249-311;
  * padding helpers ``_pad``, ``_pad_2d``, ``ensure_divisible`` :24-43 and the raw-audio collate that builds the 8-tuple
    ``(video, flow, c, x, y, g, input_lengths, paths)`` :478-532.

JPEG decoding stays on the host (cv2, as in the reference); everything after the decode -- resize, colour order, flip,
normalisation, crop, layout -- is ONE kernel launch per stream on the GPU (csrc/frames.cu), bit-exact with the reference's
cv2 / numpy sequence, and lands directly in the NHWC layout the ResNet stem reads.  The random choices (window start, crop,
flip) consume ``np.random`` in the reference's order, so a seeded run picks the same frames."""
import glob
import os
from os.path import join

import numpy as np
import torch

from .. import Options_inpainting, ops

hparams = Options_inpainting.Inpainting_Config()


def _pad(seq, max_len, constant_values=0):
    return np.pad(seq, (0, max_len - len(seq)), mode="constant", constant_values=constant_values)


def _pad_2d(x, max_len, b_pad=0):
    return np.pad(x, [(b_pad, max_len - len(x) - b_pad), (0, 0)], mode="constant", constant_values=0)


def ensure_divisible(length, divisible_by=256, lower=True):
    if length % divisible_by == 0:
        return length
    if lower:
        return length - length % divisible_by
    return length + (divisible_by - length % divisible_by)


class _SplitFileSource(object):
    """Common part of the reference's two FileDataSource classes: parse the split file, keep lengths / speaker ids."""
    length_scale = 1

    def __init__(self, data_root, col, speaker_id=None, train=True, test_size=0.05, test_num_samples=None, random_state=1234,
                 hparams=hparams):
        self.data_root = data_root
        self.col = col
        self.lengths = []
        self.speaker_id = speaker_id
        self.multi_speaker = True
        self.speaker_ids = None
        self.train = train
        self.test_size = test_size
        self.test_num_samples = test_num_samples
        self.random_state = random_state
        self.hparams = hparams

    def _lines(self):
        meta = join(self.data_root, ("train" if self.train else "test") + self.hparams.new_split_name)
        with open(meta, "rb") as f:
            lines = [l.decode("utf-8").rstrip("\n").split("|") for l in f.readlines()]
        assert len(lines[0]) == 4 or len(lines[0]) == 5
        return lines


class _NPYDataSource(_SplitFileSource):
    """mel / audio ``.npy`` columns (reference :46-110).  Lengths are in samples (``length * 1280``)."""

    def collect_files(self):
        lines = self._lines()
        self.lengths = [int(l[-1]) * 1280 for l in lines]
        paths = [join(self.data_root, l[self.col]) for l in lines]
        speaker_ids = [int(l[-2]) for l in lines]
        self.speaker_ids = speaker_ids
        if self.speaker_id is not None:                    # a multi-speaker set used as a single-speaker one
            keep = [i for i, s in enumerate(speaker_ids) if s == self.speaker_id]
            self.lengths = [int(self.lengths[i]) for i in keep]
            self.multi_speaker = False
            return [paths[i] for i in keep]
        assert len(paths) == len(self.speaker_ids)
        return sorted(paths)

    def collect_features(self, path):
        return np.load(path)


class ImageDataSource(_SplitFileSource):
    """frame-folder column (reference :113-153); ``collect_features`` samples a window of frames on ``self.device`` (the GPU)."""
    device = "cuda"

    def collect_files(self):
        lines = self._lines()
        self.lengths = [int(l[-1]) for l in lines]
        return sorted(join(self.data_root, l[self.col]) for l in lines)

    def collect_features(self, path):
        video_block, flow_block, start = sample_data_new(path, self.train, hparams=self.hparams, device=self.device)
        return video_block, flow_block, start, path


class RawAudioDataSource(_NPYDataSource):
    def __init__(self, data_root, **kwargs):
        super(RawAudioDataSource, self).__init__(data_root, 2, **kwargs)


class MelSpecDataSource(_NPYDataSource):
    def __init__(self, data_root, **kwargs):
        super(MelSpecDataSource, self).__init__(data_root, 1, **kwargs)


class ImageSpecDataSource(ImageDataSource):
    def __init__(self, data_root, **kwargs):
        super(ImageSpecDataSource, self).__init__(data_root, 0, **kwargs)


# ---- frames ------------------------------------------------------------------------------------------------------------------
def _decode(paths, flags):
    """cv2.imread of a list of files into one pinned uint8 batch (n, h, w[, 3]); every frame of a clip has the same size."""
    import cv2
    first = cv2.imread(paths[0], flags)
    if first is None:
        raise FileNotFoundError(paths[0])
    batch = torch.empty((len(paths),) + first.shape, dtype=torch.uint8)
    if torch.cuda.is_available():
        batch = batch.pin_memory()
    view = batch.numpy()
    view[0] = first
    for i, p in enumerate(paths[1:], 1):
        img = cv2.imread(p, flags)
        if img is None:
            raise FileNotFoundError(p)
        if img.shape != first.shape:
            raise RuntimeError("frames of one clip must share a size: %s is %s, expected %s" % (p, img.shape, first.shape))
        view[i] = img
    return batch


def _frames_to_blocks(data_path, items, train, flip, crop_x, crop_y, hparams, device):
    """items: list (one per loaded window) of lists of 1-based frame numbers.  Returns (video, flow) CUDA blocks shaped like the
    reference's ((load_num, T, 3|2, S, S)); memory is NHWC per frame (the permute at the end is a view)."""
    S = hparams.image_size
    R = hparams.image_rescal_size if train else S
    L, T = len(items), len(items[0])
    flat = [n for it in items for n in it]
    video = torch.zeros((L * T, S, S, 3), device=device)
    flow = torch.zeros((L * T, S, S, 2), device=device)
    if hparams.image:
        src = _decode([join(data_path, "image_crop", "%d.jpg" % n) for n in flat], 1).to(device, non_blocking=True)
        ops.frames_preprocess(src, video, 0, (R, R), flip, (crop_x, crop_y), swap_rb=True)
    if hparams.flow:
        for c, sub in enumerate(("flow_x_crop", "flow_y_crop")):
            src = _decode([join(data_path, sub, "%d.jpg" % n) for n in flat], 0).to(device, non_blocking=True)
            ops.frames_preprocess(src, flow, c, (R, R), flip, (crop_x, crop_y), swap_rb=False)
    return (video.view(L, T, S, S, 3).permute(0, 1, 4, 2, 3), flow.view(L, T, S, S, 2).permute(0, 1, 4, 2, 3))


def sample_data_new(data_path, train=True, hparams=hparams, device="cuda"):
    """Reference :155-246: a random window of ``use_image_num`` frames (+ ``load_num - 1`` further windows at least 25 frames
    away), one random crop / flip per call in training.  Returns (video_block, flow_block, start) with CUDA float blocks."""
    num_images = len(glob.glob(join(data_path, "flow_x_crop", "*.jpg")))
    max_time_second = hparams.max_time_steps / hparams.sample_rate
    use_image_num = int(np.floor(max_time_second / (0.04 * hparams.image_hope_size)))
    image_start = np.random.randint(25, num_images - use_image_num - 25 + 1)
    start = [image_start]
    for ln in range(1, hparams.load_num):
        random1 = np.random.randint(0, image_start - 25 + 1)
        random2 = np.random.randint(image_start + 25, num_images - use_image_num + 1)
        if np.random.randint(0, 2) == 1:
            start.append(random1 if random1 - start[-1] > 10 else random2)
        else:
            start.append(random2 if random2 - start[-1] > 10 else random1)
    crop_x = crop_y = flip = 0
    if train:
        crop_x = np.random.randint(0, hparams.image_rescal_size - hparams.image_size)
        crop_y = np.random.randint(0, hparams.image_rescal_size - hparams.image_size)
        flip = np.random.randint(0, 2)
    items = [[item + 1 for item in range(s, use_image_num + s)] for s in start]
    video, flow = _frames_to_blocks(data_path, items, train, flip, crop_x, crop_y, hparams, device)
    return video, flow, start


def load_image(path, train, hparams=hparams, device="cuda"):
    """Reference :249-311: every frame of the clip -> ((n, 3, S, S), (n, 2, S, S))."""
    n = len(glob.glob(join(path, "flow_x_crop", "*.jpg")))
    crop_x = crop_y = flip = 0
    if train:
        crop_x = np.random.randint(0, hparams.image_rescal_size - hparams.image_size)
        crop_y = np.random.randint(0, hparams.image_rescal_size - hparams.image_size)
        flip = np.random.randint(0, 2)
    video, flow = _frames_to_blocks(path, [list(range(1, n + 1))], train, flip, crop_x, crop_y, hparams, device)
    return video[0], flow[0]


# ---- batch assembly (raw-audio input type) -------------------------------------------------------------------------------------
def collate_raw(batch, video_block, flow_block, local_conditioning=True):
    """The tail of the reference's collate_fn (:478-532) for ``input_type == "raw"``: ``batch`` is a list of
    ``(x (T,), c (frames, n_mel), g, path)``; returns the 8-tuple the training loop unpacks."""
    input_lengths = [len(x[0]) for x in batch]
    max_input_len = max(input_lengths)
    x_batch = np.array([_pad_2d(x[0].reshape(-1, 1), max_input_len) for x in batch], dtype=np.float32)
    y_batch = np.array([_pad(x[0], max_input_len) for x in batch], dtype=np.float32)
    c_batch = None
    if local_conditioning:
        max_len = max(len(x[1]) for x in batch)
        c_batch = torch.FloatTensor(np.array([_pad_2d(x[1], max_len) for x in batch], dtype=np.float32)).transpose(1, 2).contiguous()
    g = [x[2] for x in batch]
    g_batch = torch.LongTensor(g) if all(v is not None for v in g) else None
    path_batch = [x[3] for x in batch]
    video_batch = torch.cat(list(video_block), 0) if isinstance(video_block, (list, tuple)) else video_block
    flow_batch = torch.cat(list(flow_block), 0) if isinstance(flow_block, (list, tuple)) else flow_block
    x_batch = torch.FloatTensor(x_batch).transpose(1, 2).contiguous()
    y_batch = torch.FloatTensor(y_batch).unsqueeze(-1).contiguous()
    return video_batch, flow_batch, c_batch, x_batch, y_batch, g_batch, torch.LongTensor(input_lengths), path_batch


def assert_ready_for_upsampling(x, c, hop_size):
    assert len(x) % len(c) == 0 and len(x) // len(c) == hop_size


def collate_fn(batch, hparams=hparams):
    """Reference collate_fn (:431-532) for the raw-audio input type.  ``batch``: list of
    ``(x (T,), c (frames, n_mel), video, flow, start, g, path)`` as ``PyTorchImageDataset`` yields them; every clip longer than the
    window contributes ``load_num`` (audio, mel) windows aligned with its frame windows (4 mel frames per video frame, offset 3),
    shorter clips are dropped (as upstream).  Returns the 8-tuple ``(video, flow, c, x, y, g, input_lengths, paths)``."""
    local_conditioning = len(batch[0]) >= 2 and hparams.cin_channels > 0
    if getattr(hparams, "max_time_sec", None) is not None:
        max_time_steps = int(hparams.max_time_sec * hparams.sample_rate)
    else:
        max_time_steps = hparams.max_time_steps
    use_image_num = int(np.floor((max_time_steps / hparams.sample_rate) / (0.04 * hparams.image_hope_size)))
    video_block, flow_block = [], []
    if local_conditioning:
        new_batch = []
        max_steps = ensure_divisible(max_time_steps, hparams.hop_size, True)
        for x, c, video, flow, start, g, path in batch:
            assert_ready_for_upsampling(x, c, hparams.hop_size)
            if len(x) > max_steps:
                for ln in range(hparams.load_num):
                    x1, c1 = slice_clip(x, c, start[ln], use_image_num, hparams.hop_size)
                    new_batch.append((x1, c1, g, os.path.join(path, str(start[ln]))))
                video_block.append(torch.as_tensor(video).float())
                flow_block.append(torch.as_tensor(flow).float())
        batch = new_batch
    return collate_raw(batch, video_block, flow_block, local_conditioning)


class FileSourceDataset(object):
    """The slice of nnmnkwii.datasets.FileSourceDataset the reference uses: files collected once, features loaded per item."""

    def __init__(self, file_data_source):
        self.file_data_source = file_data_source
        self.collected_files = file_data_source.collect_files()

    def __getitem__(self, idx):
        return self.file_data_source.collect_features(self.collected_files[idx])

    def __len__(self):
        return len(self.collected_files)


class PyTorchImageDataset(object):
    """(raw_audio, mel, video_block, flow_block, start, speaker_id, path) per clip (reference :393-419)."""

    def __init__(self, X, Mel, Image):
        self.X, self.Mel, self.Image = X, Mel, Image
        self.multi_speaker = X.file_data_source.multi_speaker

    def __getitem__(self, idx):
        mel = None if self.Mel is None else self.Mel[idx]
        raw_audio = self.X[idx]
        video_block, flow_block, start, path = self.Image[idx]
        speaker_id = self.X.file_data_source.speaker_ids[idx] if self.multi_speaker else None
        return raw_audio, mel, video_block, flow_block, start, speaker_id, path

    def __len__(self):
        return len(self.X)


def get_data_loaders(data_root, speaker_id=None, test_shuffle=True, hparams=hparams):
    """{"train": DataLoader, "test": DataLoader} over the split files (reference :536-587).  The frame path runs on the GPU inside
    ``__getitem__``, so the loaders use the main process (``num_workers = 0``): CUDA work does not belong in forked workers."""
    from torch.utils import data as data_utils
    loaders = {}
    for phase in ("train", "test"):
        train = phase == "train"
        kw = dict(speaker_id=speaker_id, train=train, hparams=hparams)
        X = FileSourceDataset(RawAudioDataSource(data_root, **kw))
        Image = FileSourceDataset(ImageSpecDataSource(data_root, **kw))
        Mel = FileSourceDataset(MelSpecDataSource(data_root, **kw)) if hparams.cin_channels > 0 else None
        assert Mel is None or len(X) == len(Mel)
        dataset = PyTorchImageDataset(X, Mel, Image)
        loaders[phase] = data_utils.DataLoader(dataset, batch_size=hparams.batch_size, num_workers=0,
                                               shuffle=True if train else test_shuffle,
                                               collate_fn=lambda b: collate_fn(b, hparams))
    return loaders


def slice_clip(x, c, start, use_image_num, hop_size):
    """The audio / mel window that goes with a frame window (reference :470-474): 4 mel frames per video frame, offset 3."""
    mel_start = 3 + 4 * start
    return x[mel_start * hop_size:(mel_start + use_image_num * 4) * hop_size], c[mel_start:mel_start + use_image_num * 4]
