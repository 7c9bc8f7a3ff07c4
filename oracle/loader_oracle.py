"""CPU oracle for the loader's video-frame path -- TEST INFRASTRUCTURE ONLY.

``sample_frames`` / ``load_frames`` restate /root/reference/Data_loaders/audio_loader.py:155-246 and :249-311 with the same
cv2 / numpy calls in the same order (cv2.imread -> cvtColor -> cv2.resize -> np.fliplr -> (v - 127.) / 128. -> crop ->
transpose) and the same ``np.random`` draws.  ``resize_linear_u8`` restates what cv2.resize does for 8-bit INTER_LINEAR
(OpenCV imgproc/resize.cpp: 11-bit fixed-point coefficients, HResizeLinear, VResizeLinear<uchar,int,short>) -- the arithmetic
csrc/frames.cu implements; it is pinned here against cv2 itself.

Pinning: ``reference_functions()`` compiles the UNMODIFIED ``sample_data_new`` / ``load_image`` function bodies out of the
reference file (the module itself cannot be imported: nnmnkwii / keras are absent) -- build container only; their outputs on
a committed JPEG tree are the golden vectors in tests/golden/loader_frames.pt (oracle/make_golden.py ``loader_fixture``)."""
import ast
import glob
import os

import numpy as np


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh)) for uint8, INTER_LINEAR (bit exact; see tests/test_loader_cpu.py)."""
    sh, sw = src.shape[:2]
    if (sh, sw) == (dh, dw):
        return src.copy()
    cn = 1 if src.ndim == 2 else src.shape[2]
    s = src.reshape(sh, sw, cn).astype(np.int32)

    def coeffs(dn, sn, is_col):
        i0, i1, a = np.zeros(dn, np.int32), np.zeros(dn, np.int32), np.zeros((dn, 2), np.int32)
        for d in range(dn):
            f = np.float32((d + 0.5) * (sn / dn) - 0.5)
            i = int(np.floor(f))
            f = np.float32(f - np.float32(i))
            if is_col:                                   # columns: coefficient reset at the borders
                if i < 0:
                    i, f = 0, np.float32(0)
                if i >= sn - 1:
                    i, f = sn - 1, np.float32(0)
                i0[d], i1[d] = i, min(i + 1, sn - 1)
            else:                                        # rows: coefficients kept, source rows clipped
                i0[d], i1[d] = min(max(i, 0), sn - 1), min(max(i + 1, 0), sn - 1)
            a[d, 0] = int(np.rint(np.float32(1.0 - f) * np.float32(2048)))
            a[d, 1] = int(np.rint(f * np.float32(2048)))
        return i0, i1, a
    x0, x1, xa = coeffs(dw, sw, True)
    y0, y1, ya = coeffs(dh, sh, False)
    hr = s[:, x0, :] * xa[:, 0][None, :, None] + s[:, x1, :] * xa[:, 1][None, :, None]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (hr[y0] >> 4)) >> 16) + ((b1 * (hr[y1] >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8).reshape((dh, dw) if src.ndim == 2 else (dh, dw, cn))


def _frame_blocks(data_path, items, train, flip, crop_x, crop_y, hp):
    import cv2
    R = hp.image_rescal_size if train else hp.image_size
    L, T = len(items), len(items[0])
    video = np.zeros((L, T, R, R, 3))
    flow = np.zeros((L, T, R, R, 2))
    for ln in range(L):
        for i, n in enumerate(items[ln]):
            if hp.image:
                img = cv2.cvtColor(cv2.imread(os.path.join(data_path, "image_crop", "%d.jpg" % n)), cv2.COLOR_BGR2RGB)
                img = cv2.resize(img, (R, R))
                if train and flip:
                    img = np.fliplr(img)
                video[ln, i] = (img - 127.) / 128.
            if hp.flow:
                for c, sub in enumerate(("flow_x_crop", "flow_y_crop")):
                    f = cv2.resize(cv2.imread(os.path.join(data_path, sub, "%d.jpg" % n), 0), (R, R))
                    if train and flip:
                        f = np.fliplr(f)
                    flow[ln, i, :, :, c] = (f - 127.) / 128.
    S = hp.image_size
    video = video[:, :, crop_x:crop_x + S, crop_y:crop_y + S]
    flow = flow[:, :, crop_x:crop_x + S, crop_y:crop_y + S]
    return video.transpose((0, 1, 4, 2, 3)), flow.transpose((0, 1, 4, 2, 3))


def sample_frames(data_path, train, hp):
    """audio_loader.py:155-246 sample_data_new."""
    num_images = len(glob.glob(os.path.join(data_path, "flow_x_crop", "*.jpg")))
    use_image_num = int(np.floor((hp.max_time_steps / hp.sample_rate) / (0.04 * hp.image_hope_size)))
    image_start = np.random.randint(25, num_images - use_image_num - 25 + 1)
    start = [image_start]
    for ln in range(1, hp.load_num):
        random1 = np.random.randint(0, image_start - 25 + 1)
        random2 = np.random.randint(image_start + 25, num_images - use_image_num + 1)
        if np.random.randint(0, 2) == 1:
            start.append(random1 if random1 - start[-1] > 10 else random2)
        else:
            start.append(random2 if random2 - start[-1] > 10 else random1)
    crop_x = crop_y = flip = 0
    if train:
        crop_x = np.random.randint(0, hp.image_rescal_size - hp.image_size)
        crop_y = np.random.randint(0, hp.image_rescal_size - hp.image_size)
        flip = np.random.randint(0, 2)
    items = [[item + 1 for item in range(s, use_image_num + s)] for s in start]
    video, flow = _frame_blocks(data_path, items, train, flip, crop_x, crop_y, hp)
    return video, flow, start


def load_frames(path, train, hp):
    """audio_loader.py:249-311 load_image."""
    n = len(glob.glob(os.path.join(path, "flow_x_crop", "*.jpg")))
    crop_x = crop_y = flip = 0
    if train:
        crop_x = np.random.randint(0, hp.image_rescal_size - hp.image_size)
        crop_y = np.random.randint(0, hp.image_rescal_size - hp.image_size)
        flip = np.random.randint(0, 2)
    video, flow = _frame_blocks(path, [list(range(1, n + 1))], train, flip, crop_x, crop_y, hp)
    return video[0], flow[0]


def reference_functions(hp, reference_root="/root/reference"):
    """The reference's own ``sample_data_new`` and ``load_image`` (unmodified source, compiled out of the module that cannot be
    imported as a whole).  Build container only."""
    import cv2
    path = os.path.join(reference_root, "Data_loaders", "audio_loader.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("sample_data_new", "load_image")]
    ns = dict(os=os, glob=glob, np=np, cv2=cv2, hparams=hp)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns["sample_data_new"], ns["load_image"]


def reference_collate(hp, reference_root="/root/reference"):
    """The reference's own ``collate_fn`` (:431-532) with its helpers, compiled out of the module (build container only)."""
    import types
    import torch
    path = os.path.join(reference_root, "Data_loaders", "audio_loader.py")
    tree = ast.parse(open(path).read())
    names = ("collate_fn", "_pad", "_pad_2d", "ensure_divisible", "assert_ready_for_upsampling")
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    audio = types.SimpleNamespace(get_hop_size=lambda: hp.hop_size)
    ns = dict(os=os, np=np, torch=torch, hparams=hp, audio=audio, is_mulaw_quantize=lambda s: s == "mulaw-quantize")
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns["collate_fn"]


def write_tree(root, tree):
    """tree: {relative path: bytes} -> files under root."""
    for rel, blob in tree.items():
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "wb") as f:
            f.write(bytes(blob))
    return root


def synthetic_clip(seed, n_frames, h, w):
    """{relative path: JPEG bytes} of one clip: smooth random colour frames + two gray flow streams."""
    import cv2
    rng = np.random.default_rng(seed)
    tree = {}
    yy, xx = np.mgrid[0:h, 0:w]
    for n in range(1, n_frames + 1):
        for sub, cn in (("image_crop", 3), ("flow_x_crop", 1), ("flow_y_crop", 1)):
            ph = rng.uniform(0, 6.28, (cn,))
            fr = rng.uniform(0.1, 0.9, (cn, 2))
            img = 127 + 100 * np.sin(fr[:, 0, None, None] * yy + fr[:, 1, None, None] * xx + ph[:, None, None]) + rng.normal(0, 6, (cn, h, w))
            img = np.clip(img, 0, 255).astype(np.uint8).transpose(1, 2, 0)
            ok, buf = cv2.imencode(".jpg", img if cn == 3 else img[:, :, 0], [cv2.IMWRITE_JPEG_QUALITY, 80])
            assert ok
            tree["%s/%d.jpg" % (sub, n)] = buf.tobytes()
    return tree
