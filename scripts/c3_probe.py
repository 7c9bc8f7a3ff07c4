"""BASELINE config 3 (vision-infused step: ResNet-18 x2 ImageEmbedding fused at the generator bottleneck), per-GPU share of the
global batch 32 on 8 GPUs (B = 4), native 80 x 256 mels, T = 64 frames of 224 x 224 per clip: ms/step of the eager step."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    from viai_b200 import Options_inpainting as OI, _lib
    from viai_b200.networks.Image_Embedding import ImageEmbedding
    from viai_b200.step import GanTrainer
    B, W = int(os.environ.get("C3_B", 4)), 256
    T = W // 4
    hp = OI.Inpainting_Config(cin_channels=80)
    torch.manual_seed(0)
    ve = ImageEmbedding(hp).cuda()
    tr = GanTrainer(hp, "cuda", decoder="MelDecoderImage", video_encoder=ve)
    mel = torch.rand(B, 1, 80, W).cuda()
    mask = torch.ones_like(mel)
    mask[..., W // 4:W // 4 + W // 2] = 0
    video = torch.randn(B, T, 3, 224, 224).clamp(-1, 1).cuda()
    flow = torch.randn(B, T, 2, 224, 224).clamp(-1, 1).cuda()
    for _ in range(2):
        out = tr.train_step(mel, mask, video, flow)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 4
    e0.record()
    for _ in range(steps):
        out = tr.train_step(mel, mask, video, flow)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # forward FLOPs: 7.25 GFLOP per video frame pair + G/D at 80x256; x3 for fwd + dgrad + wgrad of the visual encoder
    print(json.dumps({"config": "C3 native: B=%d, 80x256 mel, T=%d frames 224x224 (RGB + flow), MelDecoderImage, eager" % (B, T),
                      "ms_per_step": ms, "mel_frames_per_s": B * W / ms * 1e3, "video_frames_per_s": B * T / ms * 1e3,
                      "visual_encoder_algorithmic_tflops": 3 * 7.25e9 * B * T / (ms * 1e-3) / 1e12,
                      "launches_per_step": int(tr.launches_per_step), "loss_L1": float(out["loss_L1"]),
                      "max_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}))
