"""Per-phase clocks of the folded WaveNet synthesis kernel (VIAI_WN2_PROF=1), CTA 0, B = 1 and 4."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402,F401
from viai_b200 import _lib  # noqa: E402
from viai_b200.wavenet_vocoder import WaveNet  # noqa: E402

NAMES = ["rest", "operand wait", "wait h", "dep matvec", "gate+publish", "prefetch issue", "x read", "indep matvec",
         "last skip", "wait skips", "head1+wait", "head2+sample", "indep: dot+warp sum", "indep: issue weights", "indep: barrier"]


def main():
    os.environ["VIAI_WN2_PROF"] = "1"
    os.environ["VIAI_WAVENET_KERNEL"] = "folded"
    torch.manual_seed(0)
    m = WaveNet().cuda().eval()
    m.make_generation_fast_()
    for B in (1,):
        T = 3200
        c = torch.rand(B, 80, T // 160).cuda()
        m.incremental_forward(c=c, T=T)
        buf = (ctypes.c_longlong * 16)()
        _lib.check(_lib.lib().viai_wavenet2_profile(buf), "profile")
        tot = sum(buf[:15])
        print("B=%d T=%d: %.0f clocks/step total" % (B, T, tot / T))
        for n, v in zip(NAMES, buf):
            print("   %-22s %8.0f clocks/step  %5.1f %%" % (n, v / T, 100.0 * v / tot))


if __name__ == "__main__":
    main()
