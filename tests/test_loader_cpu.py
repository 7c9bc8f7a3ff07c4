"""Loader row on the CPU: the restated 8-bit cv2.resize against cv2 itself (bit exact), the oracle's frame path
against what the reference's OWN sample_data_new / load_image returned on the committed JPEG tree (tests/golden/
loader_frames.pt; and, in the build container, against those functions run live), the split-file / padding / collate host code,
and a dry run of the product's frame-window logic (np.random draw order, frame numbering, crop / flip bookkeeping) with the
torch stand-in for the preprocessing kernel."""
import os
import types

import numpy as np
import pytest
import torch

import torch_ops_shim as shim
import viai_test_helpers as H
from oracle import loader_oracle as LO

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def gold(tmp_path_factory):
    g = H.load_golden("loader_frames.pt")
    root = str(tmp_path_factory.mktemp("frames"))
    for name, tree in g["trees"].items():
        LO.write_tree(os.path.join(root, name), tree)
    g["root"] = root
    return g


@pytest.mark.parametrize("shape", [(256, 256, 224, 224, 3), (97, 131, 256, 256, 3), (128, 128, 256, 256, 1), (24, 20, 16, 16, 3),
                                   (10, 13, 16, 16, 1), (512, 448, 256, 224, 3), (33, 47, 12, 12, 3), (16, 16, 16, 16, 3),
                                   # exact 2x down-scaling in BOTH directions: OpenCV switches INTER_LINEAR to its INTER_AREA fast path
                                   # there ((a + b + c + d + 2) >> 2); the 11-bit bilinear arithmetic gives the same bytes (both
                                   # coefficients are exactly 1024), so no special case is needed -- pinned here; and 2x in one
                                   # direction only (stays bilinear)
                                   (512, 512, 256, 256, 3), (64, 48, 32, 24, 1), (64, 48, 32, 16, 3), (64, 48, 24, 24, 3)])
def test_resize_restatement_is_bit_exact_with_cv2(shape):
    sh, sw, dh, dw, cn = shape
    rng = np.random.default_rng(sh * 1000 + sw)
    src = rng.integers(0, 256, (sh, sw, cn) if cn > 1 else (sh, sw), dtype=np.uint8)
    assert np.array_equal(LO.resize_linear_u8(src, dw, dh), cv2.resize(src, (dw, dh)))


def _hp(g, **over):
    return types.SimpleNamespace(**dict(g["hp"], **over))


def test_oracle_frames_equal_reference_golden(gold):
    for case in gold["cases"]:
        np.random.seed(case["seed"])
        path = os.path.join(gold["root"], case["clip"])
        if case["kind"] == "sample":
            v, f, start = LO.sample_frames(path, case["train"], _hp(gold))
            assert [int(s) for s in start] == case["start"]
        else:
            v, f = LO.load_frames(path, case["train"], _hp(gold, load_num=1))
            assert v.shape[0] == case["n"]
            v, f = v[:6], f[:6]
        assert torch.equal(torch.from_numpy(np.ascontiguousarray(v)).float(), case["video"]), case["clip"]
        assert torch.equal(torch.from_numpy(np.ascontiguousarray(f)).float(), case["flow"]), case["clip"]


@pytest.mark.reference
def test_oracle_frames_equal_reference_functions_live(gold):
    hp = _hp(gold)
    sample_ref, load_ref = LO.reference_functions(hp)
    for clip in gold["trees"]:
        for train in (True, False):
            path = os.path.join(gold["root"], clip)
            np.random.seed(11)
            v0, f0, s0 = sample_ref(path, train, hparams=hp)
            np.random.seed(11)
            v1, f1, s1 = LO.sample_frames(path, train, hp)
            assert list(s0) == list(s1) and np.array_equal(v0, v1) and np.array_equal(f0, f1)


def test_product_frame_window_logic_dry_run(gold):
    """Same draws, same frames, same crop / flip as the reference (the kernel is replaced by its torch stand-in here)."""
    from viai_b200.Data_loaders import audio_loader as AL
    with shim.installed():
        for case in gold["cases"]:
            np.random.seed(case["seed"])
            path = os.path.join(gold["root"], case["clip"])
            if case["kind"] == "sample":
                v, f, start = AL.sample_data_new(path, case["train"], hparams=_hp(gold), device="cpu")
                assert [int(s) for s in start] == case["start"]
            else:
                v, f = AL.load_image(path, case["train"], hparams=_hp(gold, load_num=1), device="cpu")
                v, f = v[:6], f[:6]
            assert tuple(v.shape) == tuple(case["video"].shape) and torch.equal(v, case["video"])
            assert tuple(f.shape) == tuple(case["flow"].shape) and torch.equal(f, case["flow"])


def test_product_frames_have_no_cpu_path():
    from viai_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.frames_preprocess(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), torch.zeros(1, 2, 2, 3), 0, (4, 4), 0, (0, 0), True)


def test_split_file_sources_padding_and_collate(tmp_path):
    from viai_b200.Data_loaders import audio_loader as AL
    root = str(tmp_path)
    lines = ["spk2/clip9|spk2/clip9-mel.npy|spk2/clip9-audio.npy|2|7", "spk1/clip3|spk1/clip3-mel.npy|spk1/clip3-audio.npy|1|5",
             "spk2/clip1|spk2/clip1-mel.npy|spk2/clip1-audio.npy|2|6"]
    hp = types.SimpleNamespace(new_split_name="_new_split.txt")
    for phase in ("train", "test"):
        with open(os.path.join(root, phase + hp.new_split_name), "w") as f:
            f.write("\n".join(lines) + "\n")
    mel = AL.MelSpecDataSource(root, train=True, hparams=hp)
    paths = mel.collect_files()
    assert paths == sorted(os.path.join(root, l.split("|")[1]) for l in lines)
    assert mel.lengths == [7 * 1280, 5 * 1280, 6 * 1280] and mel.speaker_ids == [2, 1, 2]      # file order, as upstream (:74,:104)
    aud = AL.RawAudioDataSource(root, train=False, speaker_id=2, hparams=hp)
    assert aud.collect_files() == [os.path.join(root, lines[0].split("|")[2]), os.path.join(root, lines[2].split("|")[2])]
    assert aud.lengths == [7 * 1280, 6 * 1280] and not aud.multi_speaker
    img = AL.ImageSpecDataSource(root, train=True, hparams=hp)
    assert img.collect_files() == sorted(os.path.join(root, l.split("|")[0]) for l in lines) and img.lengths == [7, 5, 6]
    os.makedirs(os.path.join(root, "spk1"))
    np.save(os.path.join(root, "spk1/clip3-mel.npy"), np.arange(12, dtype=np.float32).reshape(4, 3))
    assert mel.collect_features(os.path.join(root, "spk1/clip3-mel.npy")).shape == (4, 3)
    # padding helpers and the raw-audio collate (audio_loader.py:24-43,478-532)
    assert AL.ensure_divisible(1000) == 768 and AL.ensure_divisible(1000, 256, lower=False) == 1024 and AL.ensure_divisible(512) == 512
    assert AL._pad(np.ones(3), 5).tolist() == [1, 1, 1, 0, 0] and AL._pad_2d(np.ones((2, 3)), 4).shape == (4, 3)
    x0, x1 = np.linspace(-1, 1, 640 * 3).astype(np.float32), np.linspace(-1, 1, 640 * 2).astype(np.float32)
    c0, c1 = np.random.rand(12, 80).astype(np.float32), np.random.rand(8, 80).astype(np.float32)
    video, flow = torch.zeros(2, 3, 3, 12, 12), torch.zeros(2, 3, 2, 12, 12)
    out = AL.collate_raw([(x0, c0, None, "a"), (x1, c1, None, "b")], video, flow)
    v, f, c, x, y, g, lengths, paths = out
    assert tuple(x.shape) == (2, 1, 1920) and tuple(y.shape) == (2, 1920, 1) and tuple(c.shape) == (2, 80, 12) and g is None
    assert lengths.tolist() == [1920, 1280] and paths == ["a", "b"] and float(x[1, 0, 1280:].abs().max()) == 0.0
    assert torch.equal(x[0, 0], y[0, :, 0]) and torch.equal(c[1, :, :8], torch.from_numpy(c1).t())
    xs, cs = AL.slice_clip(np.arange(160 * 200), np.arange(200 * 80).reshape(200, 80), start=5, use_image_num=3, hop_size=160)
    assert len(xs) == 12 * 160 and xs[0] == 23 * 160 and cs.shape == (12, 80) and cs[0, 0] == 23 * 80


def _clips(hp, seed=0):
    rng = np.random.default_rng(seed)
    T = 3                                                      # use_image_num for max_time_steps = 1920
    batch = []
    for i, frames in enumerate((40, 1, 64)):                   # the 1-frame clip (1920 samples) is not longer than the window: dropped
        mel_frames = frames * 4 + 8
        x = rng.uniform(-1, 1, mel_frames * hp.hop_size).astype(np.float32)
        c = rng.uniform(0, 1, (mel_frames, hp.cin_channels)).astype(np.float32)
        start = [int(s) for s in rng.integers(0, max(frames - T - 1, 1), hp.load_num)]
        video = rng.normal(size=(hp.load_num, T, 3, 12, 12))
        flow = rng.normal(size=(hp.load_num, T, 2, 12, 12))
        batch.append((x, c, video, flow, start, None, "clip%d" % i))
    return batch


def _collate_hp():
    return types.SimpleNamespace(cin_channels=80, file_channel=-1, max_time_sec=None, max_time_steps=1920, sample_rate=16000,
                                 image_hope_size=1, upsample_conditional_features=True, load_num=2, hop_size=160, input_type="raw")


@pytest.mark.reference
def test_collate_fn_equals_reference_collate(tmp_path):
    from viai_b200.Data_loaders import audio_loader as AL
    hp = _collate_hp()
    want = LO.reference_collate(hp)(_clips(hp))
    got = AL.collate_fn(_clips(hp), hparams=hp)
    assert len(want) == len(got) == 8
    for w, g in zip(want[:5], got[:5]):
        assert tuple(w.shape) == tuple(g.shape) and torch.equal(w, g)
    assert want[5] is None and got[5] is None and torch.equal(want[6], got[6]) and want[7] == got[7]


def test_collate_fn_shapes_and_alignment():
    from viai_b200.Data_loaders import audio_loader as AL
    hp = _collate_hp()
    batch = _clips(hp)
    v, f, c, x, y, g, lengths, paths = AL.collate_fn(batch, hparams=hp)
    assert tuple(v.shape) == (4, 3, 3, 12, 12) and tuple(f.shape) == (4, 3, 2, 12, 12)          # 2 clips x load_num windows
    assert tuple(x.shape) == (4, 1, 1920) and tuple(y.shape) == (4, 1920, 1) and tuple(c.shape) == (4, 80, 12)
    assert lengths.tolist() == [1920] * 4 and paths[0] == os.path.join("clip0", str(batch[0][4][0])) and paths[2].startswith("clip2")
    m0 = 3 + 4 * batch[0][4][0]
    assert torch.equal(x[0, 0], torch.from_numpy(batch[0][0][m0 * 160:(m0 + 12) * 160]))
    assert torch.equal(c[0], torch.from_numpy(batch[0][1][m0:m0 + 12]).t())


def test_file_source_dataset_and_image_dataset(tmp_path, gold):
    from viai_b200.Data_loaders import audio_loader as AL
    root = str(tmp_path)
    hp = types.SimpleNamespace(**dict(gold["hp"], new_split_name="_new_split.txt", cin_channels=80, batch_size=2, hop_size=160))
    for name, tree in gold["trees"].items():
        LO.write_tree(os.path.join(root, name), tree)
    for phase in ("train", "test"):
        with open(os.path.join(root, phase + hp.new_split_name), "w") as fh:
            for name in gold["trees"]:
                fh.write("%s|%s-mel.npy|%s-audio.npy|1|6\n" % (name, name, name))
    for name in gold["trees"]:
        np.save(os.path.join(root, name + "-mel.npy"), np.zeros((6 * 8, 80), np.float32))
        np.save(os.path.join(root, name + "-audio.npy"), np.zeros(6 * 1280, np.float32))
    X = AL.FileSourceDataset(AL.RawAudioDataSource(root, train=True, hparams=hp))
    Mel = AL.FileSourceDataset(AL.MelSpecDataSource(root, train=True, hparams=hp))
    assert len(X) == len(Mel) == 2 and X[0].shape == (7680,) and Mel[1].shape == (48, 80)
    assert X.file_data_source.lengths == [7680, 7680] and X.file_data_source.speaker_ids == [1, 1]
    with shim.installed():                                      # frame kernel replaced by its torch stand-in (CPU)
        Image = AL.FileSourceDataset(AL.ImageSpecDataSource(root, train=True, hparams=hp))
        Image.file_data_source.device = "cpu"
        ds = AL.PyTorchImageDataset(X, Mel, Image)
        np.random.seed(3)
        raw, mel, video, flow, start, spk, path = ds[0]
    assert raw.shape == (7680,) and mel.shape == (48, 80) and tuple(video.shape) == (2, 3, 3, 12, 12) and tuple(flow.shape) == (2, 3, 2, 12, 12)
    assert spk == 1 and len(start) == 2 and path.endswith("clipA")
