"""Loader row on the B200: the frame-preprocessing kernel (csrc/frames.cu through the C ABI) is BIT EXACT with the
reference's cv2 / numpy sequence -- against cv2 on random frames (down / up-scaling, non-square, 1 and 3 channels, flip, crop,
channel offset) and against what the reference's own sample_data_new / load_image returned on the committed JPEG tree."""
import os
import types

import numpy as np
import pytest
import torch

import viai_test_helpers as H
from oracle import loader_oracle as LO

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("cfg", [(5, 256, 256, 3, 256, 224, 1, 17, 30), (3, 97, 131, 3, 256, 224, 0, 0, 32), (4, 128, 128, 1, 256, 224, 1, 32, 0),
                                 (2, 480, 640, 1, 224, 224, 0, 0, 0), (3, 224, 224, 3, 224, 224, 1, 0, 0), (6, 24, 20, 3, 16, 12, 1, 3, 1),
                                 (2, 10, 13, 1, 16, 12, 0, 2, 4), (2, 512, 512, 3, 256, 224, 1, 5, 9), (3, 448, 448, 1, 224, 224, 0, 0, 0)], ids=lambda c: "n%d_%dx%dx%d_to%d_crop%d" % c[:6])
def test_frames_kernel_bit_exact_with_cv2(cfg):
    from viai_b200 import ops
    n, sh, sw, cn, R, S, flip, cr, cc = cfg
    rng = np.random.default_rng(sum(cfg))
    src = rng.integers(0, 256, (n, sh, sw, cn) if cn > 1 else (n, sh, sw), dtype=np.uint8)
    out_c, c_off = (3, 0) if cn == 3 else (2, 1)
    out = torch.zeros(n, S, S, out_c, device="cuda")
    ops.frames_preprocess(torch.from_numpy(src).cuda(), out, c_off, (R, R), flip, (cr, cc), swap_rb=(cn == 3))
    want = np.zeros((n, S, S, out_c), np.float32)
    for i in range(n):
        img = cv2.resize(src[i], (R, R))
        if cn == 3:
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        if flip:
            img = np.fliplr(img)
        img = ((img - 127.) / 128.)[cr:cr + S, cc:cc + S]
        want[i, :, :, c_off:c_off + cn] = img.reshape(S, S, cn)
    assert torch.equal(out.cpu(), torch.from_numpy(want))


def test_frames_kernel_argument_checks():
    from viai_b200 import ops
    src = torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="outside the 8x8 resized frame"):
        ops.frames_preprocess(src, torch.zeros(1, 6, 6, 3, device="cuda"), 0, (8, 8), 0, (3, 0), True)
    with pytest.raises(RuntimeError, match="channels"):
        ops.frames_preprocess(src, torch.zeros(1, 6, 6, 3, device="cuda"), 1, (8, 8), 0, (0, 0), True)


def test_loader_frame_path_equals_reference_golden(tmp_path):
    from viai_b200.Data_loaders import audio_loader as AL
    g = H.load_golden("loader_frames.pt")
    for name, tree in g["trees"].items():
        LO.write_tree(os.path.join(str(tmp_path), name), tree)
    for case in g["cases"]:
        np.random.seed(case["seed"])
        path = os.path.join(str(tmp_path), case["clip"])
        if case["kind"] == "sample":
            v, f, start = AL.sample_data_new(path, case["train"], hparams=types.SimpleNamespace(**g["hp"]))
            assert [int(s) for s in start] == case["start"]
        else:
            v, f = AL.load_image(path, case["train"], hparams=types.SimpleNamespace(**dict(g["hp"], load_num=1)))
            assert v.size(0) == case["n"]
            v, f = v[:6], f[:6]
        assert v.is_cuda and tuple(v.shape) == tuple(case["video"].shape)
        assert torch.equal(v.cpu(), case["video"]) and torch.equal(f.cpu(), case["flow"]), (case["clip"], case["train"], case["seed"])
    # the blocks are NHWC in memory: handing them to the ResNet stem costs no layout pass
    assert v.permute(0, 1, 3, 4, 2).is_contiguous() if v.dim() == 5 else v.permute(0, 2, 3, 1).is_contiguous()
