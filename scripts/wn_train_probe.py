"""WaveNet teacher-forced training step timing (bench.time_wavenet_train) at a few (B, T)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    for B, T, graph in ((4, 8000, False), (4, 8000, True), (8, 8000, True)):
        try:
            print(json.dumps(bench.time_wavenet_train(torch, B=B, T=T, steps=3, warmup=2, graph=graph)), flush=True)
        except Exception as e:
            print("B=%d T=%d failed: %r" % (B, T, e), flush=True)
        torch.cuda.empty_cache()
