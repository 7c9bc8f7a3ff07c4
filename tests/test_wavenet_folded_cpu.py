"""Host logic of the folded WaveNet synthesis kernel (csrc/wavenet_synth2.cu): ``WaveNet.pack_for_synthesis_folded`` and the slot
schedule the kernel runs, emulated here CTA by CTA from the PACKED blocks (same buffers, same slot order, same exchange
parities) and compared with the oracle's sample-by-sample synthesis (oracle/viai_oracle.py wavenet_incremental, which follows
wavenet.py:237-364).  The algebra is exact; the only differences are fp32 re-association, so the bound is 1e-5."""
import math

import pytest
import torch

import viai_test_helpers as H
from oracle import viai_oracle as O


def emulate_folded(pk, dims, lps, cond, uniforms, test_inputs=None, log_scale_min=-7.0):
    """Mirror of wavenet_synth2_kernel; vectorised over CTAs (dim 0) but otherwise the kernel's data flow."""
    L, R, G, S, C, K, Oc = dims
    nC, Kn, rowsC, rows1 = pk["nC"], pk["Kn"], pk["rowsC"], pk["rows1"]
    pad4 = lambda n: (n + 3) // 4 * 4
    pairs, srows, orows, hrows = (G // 2) // nC, S // nC, R // nC, S // nC
    K2 = G // 2
    B, T = cond.size(0), cond.size(1)
    r2 = math.sqrt(0.5)
    W = pk["layers"]
    o0 = rowsC * K2
    Wc = W[:, :, :o0].reshape(L, nC, rowsC, K2)
    bC = W[:, :, o0:o0 + rowsC]
    o1 = o0 + pad4(rowsC)
    Wn = W[:, :, o1:o1 + rows1 * Kn].reshape(L, nC, rows1, Kn)
    cN = W[:, :, o1 + rows1 * Kn:o1 + rows1 * Kn + rows1]
    uc = W[:, :, o1 + rows1 * Kn + pad4(rows1):o1 + rows1 * Kn + pad4(rows1) + rows1]
    fw, fb = pk["first"][:R], pk["first"][R:]
    wlast = pk["last"][:, :srows * K2].reshape(nC, srows, K2)
    blast = pk["last"][:, srows * K2:srows * K2 + srows]
    h1w = pk["head1"][:, :hrows * S].reshape(nC, hrows, S)
    h1b = pk["head1"][:, hrows * S:hrows * S + hrows]
    h2w, h2b = pk["head2"][:Oc * S].reshape(Oc, S), pk["head2"][Oc * S:]
    ring = [torch.zeros((K - 1) * 2 ** (l % lps) + 1, B, R) for l in range(L)]
    gbuf = [torch.zeros(B, K2), torch.zeros(B, K2)]
    xnew = [torch.zeros(B, R) for _ in range(3)]
    Hc = max(K - 1, 1)
    curh = torch.zeros(Hc, B)
    cur = torch.zeros(B)
    xin = torch.zeros(B, Kn)                      # every CTA holds the same input vector; one copy is enough here
    Pn = torch.zeros(nC, rows1, B)
    skips = torch.zeros(nC, srows, B)
    xown = torch.zeros(nC, orows, B)
    out, logits = torch.zeros(B, T), torch.zeros(B, T, Oc)

    def taps_into(x, layer, t):
        d = 2 ** (layer % lps)
        rl = (K - 1) * d + 1
        for j in range(K - 1):
            back = (K - 1 - j) * d
            if layer == 0:                        # layer 0's history is rebuilt from the input samples, not read from a ring
                tau = t - back
                x[:, K2 + R + j * R:K2 + R + (j + 1) * R] = (curh[tau % Hc][:, None] * fw[None] + fb[None]) if tau >= 0 else 0.0
            else:
                x[:, K2 + R + j * R:K2 + R + (j + 1) * R] = ring[layer][(t - back) % rl]
        if C:
            x[:, K2 + R + (K - 1) * R:] = cond[:, t]

    def gate(z):                                   # z: (nC, rows1, B) with (a, b) rows adjacent -> (B, K2)
        a, g = z[:, 0::2], z[:, 1::2]
        return (torch.tanh(a) * torch.sigmoid(g)).reshape(nC * pairs, B).t().contiguous()

    # prologue: P'_0 for t = 0 from block L-1
    taps_into(xin, 0, 0)
    Pn = torch.einsum("crk,bk->crb", Wn[L - 1], xin) + cN[L - 1][:, :, None]
    for t in range(T):
        if test_inputs is not None and t < test_inputs.size(1):
            cur = test_inputs[:, t].clone()
        curh[t % Hc] = cur
        x0 = cur[:, None] * fw[None] + fb[None]                              # (B, R)
        for l in range(L):
            if l == 0:
                z = Pn + uc[0][:, :, None] * cur[None, None, :]
            else:
                xin[:, :K2] = gbuf[(l - 1) & 1]
                res = torch.einsum("crk,bk->crb", Wc[l], xin[:, :K2])
                z = Pn + res[:, srows + orows:]
            gbuf[l & 1] = gate(z)
            if l >= 1:
                v = res[:, :srows] + bC[l][:, :srows, None]
                skips = v if l == 1 else (skips + v) * r2
                xprev = x0.t().reshape(nC, orows, B) if l == 1 else xown
                xown = (res[:, srows:srows + orows] + bC[l][:, srows:srows + orows, None] + xprev) * r2
                rl = (K - 1) * 2 ** (l % lps) + 1
                ring[l][t % rl] = xown.reshape(R, B).t()
                if l <= L - 3:
                    xnew[l % 3] = xown.reshape(R, B).t().clone()
            nl, nt = (l + 1) % L, t + (1 if l + 1 == L else 0)
            if nt < T:
                if l <= 1:
                    xin[:, K2:K2 + R] = x0
                elif l <= L - 2:
                    xin[:, K2:K2 + R] = xnew[(l - 1) % 3]
                taps_into(xin, nl, nt)
                Pn = torch.einsum("crk,bk->crb", Wn[l], xin) + cN[l][:, :, None]
        hl = gbuf[(L - 1) & 1]
        v = torch.einsum("crk,bk->crb", wlast, hl) + blast[:, :, None]
        skips = v if L == 1 else (skips + v) * r2
        sv = torch.relu(skips).reshape(S, B).t()                              # (B, S)
        hv = torch.relu(torch.einsum("crk,bk->crb", h1w, sv) + h1b[:, :, None]).reshape(S, B).t()
        y = hv @ h2w.t() + h2b
        logits[:, t] = y
        nm = Oc // 3
        cur = O.sample_dmol(y, uniforms[t][:, :nm], uniforms[t][:, nm], log_scale_min)
        out[:, t] = cur
    return out, logits


@pytest.mark.parametrize("kw,nC,B", [
    (dict(layers=8, stacks=2, residual_channels=32, gate_channels=32, skip_out_channels=16, cin_channels=80, out_channels=30,
          upsample_scales=[2, 4], kernel_size=3), 4, 2),
    (dict(layers=6, stacks=1, residual_channels=16, gate_channels=32, skip_out_channels=16, cin_channels=80, out_channels=30,
          upsample_scales=[8], kernel_size=3), 8, 1),
    (dict(layers=4, stacks=4, residual_channels=16, gate_channels=16, skip_out_channels=8, cin_channels=80, out_channels=30,
          upsample_scales=[8], kernel_size=2), 2, 3),
])
def test_folded_schedule_from_packed_blocks_matches_oracle(kw, nC, B):
    from viai_b200.wavenet_vocoder import WaveNet
    torch.manual_seed(3)
    m = WaveNet(dropout=0.0, **kw).eval()
    m.make_generation_fast_()                                                 # plain weights: packing then needs no CUDA op
    with torch.no_grad():
        for p in m.parameters():
            p.uniform_(-0.3, 0.3)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    T = 48
    hop = 8
    g = torch.Generator().manual_seed(11)
    c = torch.rand(B, 80, T // hop, generator=g)
    u = torch.rand(T, B, 11, generator=g) * (1 - 2e-5) + 1e-5
    ti = torch.rand(B, 1, T // 2, generator=g) * 2 - 1                        # half teacher-forced, half free-running
    lps = kw["layers"] // kw["stacks"]
    want, wlg = O.wavenet_incremental(sd, c, T, lps, kw["upsample_scales"], test_inputs=ti, uniforms=u, return_logits=True,
                                      kernel_size=kw["kernel_size"])
    cond = O.wavenet_upsample(O.wavenet_dims(sd)[0], c, kw["upsample_scales"]).transpose(1, 2).contiguous()   # (B, T, C)
    pk = m.pack_for_synthesis_folded(nC)
    got, glg = emulate_folded(pk, m._dims(), lps, cond, u, test_inputs=ti.reshape(B, -1))
    assert H.relerr(glg, wlg) < 1e-5 and H.relerr(got, want.reshape(B, T)) < 1e-5
