"""Per-phase clocks of the warp-specialised WaveNet synthesis kernel (VIAI_WN3_PROF=1), CTA 0."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402,F401
from viai_b200 import _lib  # noqa: E402
from viai_b200.wavenet_vocoder import WaveNet  # noqa: E402

D = ["rest", "stage h", "group barrier", "weight wait", "dots", "P' wait", "gate+publish", "last skip+fence", "barrier A", "head",
     "barrier B", "sampler+barrier C"]
I = ["rest", "B: x + h gather", "B: group barrier", "A: weight wait", "A: operand wait", "A: taps dot", "B: dot + barrier", "P hand-over", "barrier A",
     "head..C"]


def main():
    os.environ["VIAI_WN3_PROF"] = "1"
    os.environ["VIAI_WAVENET_KERNEL"] = "ws"
    torch.manual_seed(0)
    m = WaveNet().cuda().eval()
    m.make_generation_fast_()
    for B in (1, 4):
        T = 3200
        c = torch.rand(B, 80, T // 160).cuda()
        m.incremental_forward(c=c, T=T)
        buf = (ctypes.c_longlong * 32)()
        _lib.check(_lib.lib().viai_wavenet3_profile(buf), "profile")
        print("B=%d T=%d  gate warp: %.0f clocks/step" % (B, T, sum(buf[:12]) / T))
        for n, v in zip(D, buf[:12]):
            print("   D %-18s %8.0f clocks/step" % (n, v / T))
        for n, v in zip(["dots only (12)", "select (13)", "P' loads + add (14)"], buf[12:15]):
            print("   D   %-18s %8.0f clocks/step" % (n, v / T))
        print("   D   P' not ready at first look: %.2f of 24 slots per step" % (buf[15] / T))
        print("   independent group: %.0f clocks/step" % (sum(buf[16:26]) / T))
        for n, v in zip(I, buf[16:26]):
            print("   I %-18s %8.0f clocks/step" % (n, v / T))


if __name__ == "__main__":
    main()
