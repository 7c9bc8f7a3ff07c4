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
                                   (10, 13, 16, 16, 1), (512, 448, 256, 224, 3), (33, 47, 12, 12, 3), (16, 16, 16, 16, 3)])
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
