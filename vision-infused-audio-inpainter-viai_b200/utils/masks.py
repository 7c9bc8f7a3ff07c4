"""Inpainting masks of the GAN step: {0,1} float tensors (B, 1, H, W), 0 = the region to inpaint.

The reference's mask logic lives in the missing ``Models/Whole_Sync_inpainting_modify.py`` (contract: train_whole_sync.py:49-50,92;
picture misc/pipeline2.png: a contiguous TIME band over all mel bins).  The free-form variant (BASELINE config 5) has no
reference definition at all; ours is a seeded random walk of square brush strokes driven by a 64-bit LCG in pure integer
arithmetic, so that it can be restated anywhere bit for bit (tests/test_masks_cpu.py pins it against the oracle's restatement)."""
import torch

_M64 = (1 << 64) - 1


def time_band_mask(shape, t0, blank_length):
    B, _, H, W = shape
    m = torch.ones((B, 1, H, W), dtype=torch.float32)
    m[..., max(0, t0):max(0, min(W, t0 + blank_length))] = 0.0
    return m


def freeform_mask(shape, seed, strokes=6, max_width=12):
    """``strokes`` random walks per sample, each 16-63 steps of a (2 .. max_width + 1)-pixel square brush."""
    B, _, H, W = shape
    m = torch.ones((B, 1, H, W), dtype=torch.float32)
    state = (seed * 2654435761 + 1442695040888963407) & _M64

    def nxt():
        nonlocal state
        state = (state * 6364136223846793005 + 1442695040888963407) & _M64
        return state >> 33

    for b in range(B):
        for _ in range(strokes):
            y, x = nxt() % H, nxt() % W
            wdt = 2 + nxt() % max_width
            for _ in range(16 + nxt() % 48):
                dy, dx = (nxt() % 7) - 3, (nxt() % 15) - 3
                y = min(max(y + dy, 0), H - 1)
                x = min(max(x + dx, 0), W - 1)
                m[b, 0, max(0, y - wdt // 2):y + wdt // 2 + 1, max(0, x - wdt // 2):x + wdt // 2 + 1] = 0.0
    return m
