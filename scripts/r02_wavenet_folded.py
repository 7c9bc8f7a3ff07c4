"""Round-2 check of the folded WaveNet synthesis kernel: small model vs the oracle, full model vs the grid kernel (teacher
forced), then samples/s of both at T = 8000 (run under `timeout`)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__  # noqa: E402,F401
import viai_test_helpers as H  # noqa: E402
from oracle import viai_oracle as O  # noqa: E402
from viai_b200.wavenet_vocoder import WaveNet  # noqa: E402


def run(m, kern, **kw):
    os.environ["VIAI_WAVENET_KERNEL"] = kern
    return m.incremental_forward(**kw)


def main():
    kw = dict(layers=8, stacks=2, residual_channels=32, gate_channels=32, skip_out_channels=32, cin_channels=80, out_channels=30,
              upsample_scales=[2, 4], kernel_size=3)
    torch.manual_seed(0)
    m = WaveNet(dropout=0.0, **kw).cuda().eval()
    m.make_generation_fast_()
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    T, B = 64, 3
    c = torch.rand(B, 80, T // 8)
    u = torch.rand(T, B, 11) * (1 - 2e-5) + 1e-5
    want, wlg = O.wavenet_incremental(sd, c, T, 4, [2, 4], uniforms=u, return_logits=True)
    for kern in ("grid", "folded", "ws"):
        out, lg = run(m, kern, c=c.cuda(), T=T, uniforms=u.cuda(), return_logits=True)
        torch.cuda.synchronize()
        print("small %-6s logits relerr %.2e samples relerr %.2e" % (kern, H.relerr(lg, wlg), H.relerr(out, want)), flush=True)
    torch.manual_seed(0)
    m = WaveNet().cuda().eval()
    m.make_generation_fast_()
    T = 640
    c = torch.rand(1, 80, T // 160).cuda()
    u = torch.empty((T, 1, 11), device="cuda").uniform_(1e-5, 1 - 1e-5)
    ti = torch.rand(1, T, 1).cuda() * 2 - 1
    res = {}
    for kern in ("grid", "folded", "ws"):
        res[kern] = run(m, kern, c=c, T=T, uniforms=u, test_inputs=ti, return_logits=True)
        torch.cuda.synchronize()
    print("full teacher-forced T=%d: folded vs grid logits relerr %.2e, ws vs grid %.2e" % (
        T, H.relerr(res["folded"][1], res["grid"][1]), H.relerr(res["ws"][1], res["grid"][1])), flush=True)
    o1 = run(m, "grid", c=c, T=T, uniforms=u)
    o2 = run(m, "ws", c=c, T=T, uniforms=u)
    print("full free-running T=%d: ws vs grid samples relerr %.2e" % (T, H.relerr(o2, o1)), flush=True)
    for Bn in (1, 4):
        T = 8000
        c = torch.rand(Bn, 80, T // 160).cuda()
        for kern in ("grid", "folded", "ws"):
            run(m, kern, c=c[:, :, :10].contiguous(), T=1600)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(m, kern, c=c, T=T)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print("B=%d %-6s T=%d: %.3f s  %.0f samples/s  (%.1f us/step)" % (Bn, kern, T, dt, Bn * T / dt, dt / T * 1e6), flush=True)


if __name__ == "__main__":
    main()
