"""Runs one layer (forward + data gradient + weight gradient) a few times at its C2 geometry: the target of ncu --set full."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from viai_b200 import ops
LAYERS = {  # name: (transposed, Cin, Cout, k, stride, pad, H, W)
    "block5": (True, 32, 32, 3, (1, 1), (1, 1), 128, 128),
    "conv6_1": (True, 32, 32, 3, (1, 1), (1, 1), 256, 256),
    "d_conv2_1": (False, 64, 128, 3, (2, 2), (1, 1), 256, 128),
    "d_conv2_2": (False, 128, 256, 3, (2, 2), (1, 1), 128, 64),
    "d_conv3": (False, 256, 512, 3, (1, 1), (1, 1), 64, 32),
}
name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tr, Cin, Cout, k, st, pd, H, W = LAYERS[name]
torch.manual_seed(0)
x = torch.randn(32, H, W, Cin, device="cuda", requires_grad=True)
w = (torch.randn((Cin, Cout, k, k) if tr else (Cout, Cin, k, k), device="cuda") * 0.05).requires_grad_(True)
for _ in range(reps):
    y, stats = ops.conv2d_stats(x, w, None, st, pd, tr, 1)
    y.backward(torch.randn_like(y))
    x.grad = None; w.grad = None
torch.cuda.synchronize()
print("done", name, tuple(y.shape))
