"""Throughput of the frame-preprocessing kernel: n decoded 256x256 BGR frames -> 224x224 crop of the (256 -> 256 copy | 320 -> 256
resize) -> float NHWC.  Algorithmic bytes per frame: source bytes read + 224*224*3*4 written."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viai_b200 import ops  # noqa: E402

if __name__ == "__main__":
    for n, s, R in ((4096, 256, 256), (4096, 320, 256), (4096, 128, 256)):
        src = torch.randint(0, 256, (n, s, s, 3), dtype=torch.uint8, device="cuda")
        out = torch.empty(n, 224, 224, 3, device="cuda")
        for _ in range(3):
            ops.frames_preprocess(src, out, 0, (R, R), 1, (16, 16), True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.frames_preprocess(src, out, 0, (R, R), 1, (16, 16), True)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        # bytes: the source region the crop maps to + the float block
        read = n * (224 * s / R) ** 2 * 3
        wr = n * 224 * 224 * 3 * 4
        print(json.dumps({"frames": n, "src": s, "resize": R, "ms": ms, "frames_per_s": n / ms * 1e3, "GB_per_s": (read + wr) / ms / 1e6}))
