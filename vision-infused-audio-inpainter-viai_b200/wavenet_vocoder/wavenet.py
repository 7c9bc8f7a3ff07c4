"""WaveNet vocoder.  Drop-in for /root/reference/wavenet_vocoder/wavenet.py ``WaveNet`` :62-393 on the synthesis path:
same constructor (17 keyword arguments, defaults from ``Config``), ``state_dict()`` layout, ``incremental_forward`` /
``forward`` signatures, ``clear_buffer``, ``make_generation_fast_``, ``receptive_field``.  The T-step autoregressive loop is
one persistent CUDA kernel (csrc/wavenet_synth.cu)."""
import ctypes
import math

import torch
from torch import nn

from .. import Config, _lib, ops
from .modules import Conv1d1x1, ConvTranspose2d, ResidualConv1dGLU, effective_weight, rows_linear

hparams = Config.Config()


def receptive_field_size(total_layers, num_cycles, kernel_size, dilation=lambda x: 2 ** x):
    """Compute receptive field size (reference :41-59)."""
    assert total_layers % num_cycles == 0
    layers_per_cycle = total_layers // num_cycles
    dilations = [dilation(i % layers_per_cycle) for i in range(total_layers)]
    return (kernel_size - 1) * sum(dilations) + 1


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class WaveNet(nn.Module):
    def __init__(self, out_channels=hparams.out_channels, layers=hparams.decode_layers, stacks=hparams.decode_stacks,
                 residual_channels=hparams.residual_channels,
                 gate_channels=hparams.gate_channels,
                 skip_out_channels=hparams.skip_out_channels,
                 kernel_size=hparams.kernel_size, dropout=hparams.dropout,
                 cin_channels=hparams.cin_channels, gin_channels=hparams.gin_channels, n_speakers=hparams.n_speakers,
                 weight_normalization=hparams.weight_normalization,
                 upsample_conditional_features=hparams.upsample_conditional_features,
                 upsample_scales=hparams.upsample_scales,
                 freq_axis_kernel_size=hparams.freq_axis_kernel_size,
                 scalar_input=True,
                 use_speaker_embedding=True,
                 ):
        super(WaveNet, self).__init__()
        if not scalar_input:
            raise NotImplementedError("one-hot (mu-law) input is outside the VIAI hot path (input_type 'raw', SURVEY Appendix A)")
        if gin_channels > 0:
            raise NotImplementedError("global conditioning is outside the VIAI hot path")
        self.scalar_input = scalar_input
        self.out_channels = out_channels
        self.cin_channels = cin_channels
        assert layers % stacks == 0
        layers_per_stack = layers // stacks
        self.layers_per_stack = layers_per_stack
        self.kernel_size = kernel_size
        self.first_conv = Conv1d1x1(1, residual_channels)
        self.conv_layers = nn.ModuleList()
        for layer in range(layers):
            dilation = 2 ** (layer % layers_per_stack)
            self.conv_layers.append(ResidualConv1dGLU(
                residual_channels, gate_channels, kernel_size=kernel_size, skip_out_channels=skip_out_channels, bias=True,
                dilation=dilation, dropout=dropout, cin_channels=cin_channels, gin_channels=gin_channels,
                weight_normalization=weight_normalization))
        self.last_conv_layers = nn.ModuleList([
            nn.ReLU(inplace=True),
            Conv1d1x1(skip_out_channels, skip_out_channels, weight_normalization=weight_normalization),
            nn.ReLU(inplace=True),
            Conv1d1x1(skip_out_channels, out_channels, weight_normalization=weight_normalization),
        ])
        self.embed_speakers = None
        if upsample_conditional_features:
            self.upsample_conv = nn.ModuleList()
            for s in upsample_scales:
                freq_axis_padding = (freq_axis_kernel_size - 1) // 2
                self.upsample_conv.append(ConvTranspose2d(1, 1, (freq_axis_kernel_size, s), padding=(freq_axis_padding, 0),
                                                          dilation=1, stride=(1, s), weight_normalization=weight_normalization))
                self.upsample_conv.append(nn.ReLU(inplace=True))
        else:
            self.upsample_conv = None
        self.receptive_field = receptive_field_size(layers, stacks, kernel_size)
        self.softmax = nn.Softmax()
        self._packed = None

    def has_speaker_embedding(self):
        return self.embed_speakers is not None

    def local_conditioning_enabled(self):
        return self.cin_channels > 0

    def clear_buffer(self):
        """The kernel owns its ring buffers per call; nothing persists between calls (reference :377-385)."""
        self._packed = None

    def make_generation_fast_(self):
        def remove_weight_norm(m):
            try:
                nn.utils.remove_weight_norm(m)
            except ValueError:
                return
        self.apply(remove_weight_norm)
        self._packed = None

    # ---- host-side preparation -----------------------------------------------------------------------------------
    def _dims(self):
        l0 = self.conv_layers[0]
        R = l0.conv.in_channels
        G = l0.conv.out_channels
        S = l0.conv1x1_skip.out_channels
        C = l0.conv1x1c.in_channels if l0.conv1x1c is not None else 0
        return len(self.conv_layers), R, G, S, C, self.kernel_size, self.out_channels

    @torch.no_grad()
    def pack_for_synthesis(self, nC):
        """Re-lays the (weight-norm folded) parameters into the per-(layer, CTA) blocks the kernel streams:
        [2*pairs rows of (tap-major dilated conv | conditioning 1x1)] [their biases, padded to 4] [skip rows; residual rows]
        [their biases, padded to 4].  Row pairs (i, i + gate/2) are adjacent so that tanh(a) * sigmoid(b) is CTA-local."""
        L, R, G, S, C, K, O = self._dims()
        dev = self.first_conv.bias.device
        pairs, srows, orows = (G // 2) // nC, S // nC, R // nC
        pad4 = lambda n: (n + 3) // 4 * 4
        K1, K2 = K * R + C, G // 2
        rows1, rows2 = 2 * pairs, srows + orows
        stride = rows1 * K1 + pad4(rows1) + rows2 * K2 + pad4(rows2)
        out = torch.zeros((L, nC, stride), device=dev, dtype=torch.float32)
        i1 = torch.arange(nC * pairs, device=dev).view(nC, pairs)
        idx1 = torch.stack((i1, i1 + G // 2), dim=2).reshape(nC, rows1)                   # a_i, b_i interleaved
        idx_s = torch.arange(S, device=dev).view(nC, srows)
        idx_o = torch.arange(R, device=dev).view(nC, orows)
        for l, f in enumerate(self.conv_layers):
            w = effective_weight(f.conv).float()                                          # (G, R, K)
            lin = w.permute(0, 2, 1).reshape(G, K * R)                                    # tap-major (conv.py:56-61)
            b1 = f.conv.bias.float().clone()
            if f.conv1x1c is not None:
                lin = torch.cat((lin, effective_weight(f.conv1x1c).float().reshape(G, C)), 1)
                b1 += f.conv1x1c.bias.float()
            ws, wo = effective_weight(f.conv1x1_skip).float().reshape(S, K2), effective_weight(f.conv1x1_out).float().reshape(R, K2)
            o = 0
            out[l, :, o:o + rows1 * K1] = lin[idx1].reshape(nC, -1); o += rows1 * K1
            out[l, :, o:o + rows1] = b1[idx1]; o += pad4(rows1)
            out[l, :, o:o + rows2 * K2] = torch.cat((ws[idx_s], wo[idx_o]), 1).reshape(nC, -1); o += rows2 * K2
            out[l, :, o:o + rows2] = torch.cat((f.conv1x1_skip.bias.float()[idx_s], f.conv1x1_out.bias.float()[idx_o]), 1)
        first = torch.cat((effective_weight(self.first_conv).float().reshape(-1), self.first_conv.bias.float()))
        hrows = S // nC
        h1 = torch.zeros((nC, hrows * S + pad4(hrows)), device=dev)
        w1 = effective_weight(self.last_conv_layers[1]).float().reshape(S, S)
        h1[:, :hrows * S] = w1.view(nC, hrows * S)
        h1[:, hrows * S:hrows * S + hrows] = self.last_conv_layers[1].bias.float().view(nC, hrows)
        h2 = torch.cat((effective_weight(self.last_conv_layers[3]).float().reshape(-1), self.last_conv_layers[3].bias.float()))
        return dict(layers=out.contiguous(), first=first.contiguous(), head1=h1.contiguous(), head2=h2.contiguous(), nC=nC)

    @torch.no_grad()
    def pack_for_synthesis_folded(self, nC):
        """Parameter blocks of the FOLDED synthesis kernel (csrc/wavenet_synth2.cu), which needs one dependent cross-CTA exchange
        per layer instead of two.  With A_l the current-time tap of layer l's dilated convolution, Wo/bo its residual 1x1 and
        x_l = r (Wo_{l-1} h_{l-1} + bo_{l-1} + x_{l-1}), r = sqrt(0.5) (modules.py:196-207), the gate pre-activation is

            z_l = [old taps + conditioning + bias]_l + A_l x_l
                = P'_l + M_l h_{l-1},      M_l = r A_l Wo_{l-1}
            P'_l = [old taps + conditioning]_l + N_l h_{l-2} + T_l x_{l-2} + const_l,   N_l = r^2 A_l Wo_{l-2},  T_l = r^2 A_l

        P'_l only needs vectors that were exchanged one slot earlier, so it is evaluated off the critical path while h_{l-1} is
        in flight; only M_l h_{l-1} (fused with the skip / residual rows of layer l-1, which read the same h_{l-1}) is dependent.
        Layers 0 and 1 use x_0 = fw * sample + fb directly (T_1 = r A_1, N_1 = 0; layer 0: rank-1 term uc = A_0 fw).
        Block l (per CTA):  [skip rows l-1 | residual rows l-1 | M_l gate rows][G/2]  [their biases, padded to 4]
                            [gate rows of layer nl = (l+1) % L: N | T | old taps | conditioning]  [const, padded]  [uc, padded]
        The products are formed in fp64 and rounded once to fp32."""
        L, R, G, S, C, K, O = self._dims()
        dev = self.first_conv.bias.device
        pairs, srows, orows = (G // 2) // nC, S // nC, R // nC
        pad4 = lambda n: (n + 3) // 4 * 4
        K2 = G // 2
        Kn = K2 + R + (K - 1) * R + C
        rows1, rowsC = 2 * pairs, srows + orows + 2 * pairs
        stride = rowsC * K2 + pad4(rowsC) + rows1 * Kn + 2 * pad4(rows1)
        out = torch.zeros((L, nC, stride), device=dev, dtype=torch.float32)
        i1 = torch.arange(nC * pairs, device=dev).view(nC, pairs)
        idx1 = torch.stack((i1, i1 + G // 2), dim=2).reshape(nC, rows1)
        idx_s = torch.arange(S, device=dev).view(nC, srows)
        idx_o = torch.arange(R, device=dev).view(nC, orows)
        r2 = math.sqrt(0.5)
        A, Wt, Wc, b, Wo, bo, Ws, bs = [], [], [], [], [], [], [], []
        for f in self.conv_layers:
            w = effective_weight(f.conv).double()                                         # (G, R, K)
            lin = w.permute(0, 2, 1).reshape(G, K * R)                                    # tap-major (conv.py:56-61)
            A.append(lin[:, (K - 1) * R:])
            Wt.append(lin[:, :(K - 1) * R])
            b1 = f.conv.bias.double().clone()
            if f.conv1x1c is not None:
                Wc.append(effective_weight(f.conv1x1c).double().reshape(G, C))
                b1 += f.conv1x1c.bias.double()
            else:
                Wc.append(torch.zeros((G, C), device=dev, dtype=torch.float64))
            b.append(b1)
            Ws.append(effective_weight(f.conv1x1_skip).double().reshape(S, K2)); bs.append(f.conv1x1_skip.bias.double())
            Wo.append(effective_weight(f.conv1x1_out).double().reshape(R, K2)); bo.append(f.conv1x1_out.bias.double())
        fw, fb = effective_weight(self.first_conv).double().reshape(-1), self.first_conv.bias.double()
        for l in range(L):
            if l >= 1:
                M = r2 * (A[l] @ Wo[l - 1])
                blk = torch.cat((Ws[l - 1][idx_s], Wo[l - 1][idx_o], M[idx1]), 1)          # (nC, rowsC, K2)
                out[l, :, :rowsC * K2] = blk.reshape(nC, -1).float()
                out[l, :, rowsC * K2:rowsC * K2 + srows + orows] = torch.cat((bs[l - 1][idx_s], bo[l - 1][idx_o]), 1).float()
            o = rowsC * K2 + pad4(rowsC)
            nl = (l + 1) % L
            N = torch.zeros((G, K2), device=dev, dtype=torch.float64)
            Tm = torch.zeros((G, R), device=dev, dtype=torch.float64)
            cN = b[nl].clone()
            if nl == 0:
                cN += A[0] @ fb
            elif nl == 1:
                Tm = r2 * A[1]
                cN += r2 * (A[1] @ bo[0])
            else:
                N = 0.5 * (A[nl] @ Wo[nl - 2])
                Tm = 0.5 * A[nl]
                cN += r2 * (A[nl] @ bo[nl - 1]) + 0.5 * (A[nl] @ bo[nl - 2])
            Wn = torch.cat((N, Tm, Wt[nl], Wc[nl]), 1)                                     # (G, Kn)
            out[l, :, o:o + rows1 * Kn] = Wn[idx1].reshape(nC, -1).float(); o += rows1 * Kn
            out[l, :, o:o + rows1] = cN[idx1].float(); o += pad4(rows1)
            if l == 0:
                out[0, :, o:o + rows1] = (A[0] @ fw)[idx1].float()
        last = torch.zeros((nC, srows * K2 + pad4(srows)), device=dev)
        last[:, :srows * K2] = Ws[L - 1][idx_s].reshape(nC, -1).float()
        last[:, srows * K2:srows * K2 + srows] = bs[L - 1][idx_s].float()
        first = torch.cat((fw, fb)).float()
        hrows = S // nC
        h1 = torch.zeros((nC, hrows * S + pad4(hrows)), device=dev)
        w1 = effective_weight(self.last_conv_layers[1]).float().reshape(S, S)
        h1[:, :hrows * S] = w1.view(nC, hrows * S)
        h1[:, hrows * S:hrows * S + hrows] = self.last_conv_layers[1].bias.float().view(nC, hrows)
        h2 = torch.cat((effective_weight(self.last_conv_layers[3]).float().reshape(-1), self.last_conv_layers[3].bias.float()))
        return dict(layers=out.contiguous(), last=last.contiguous(), first=first.contiguous(), head1=h1.contiguous(),
                    head2=h2.contiguous(), nC=nC, Kn=Kn, rowsC=rowsC, rows1=rows1, stride=stride)

    def _upsample(self, c):
        """(B, cin, Tc) -> (B, T, cin): ConvTranspose2d(1,1,(f,s), stride (1,s)) + ReLU per scale (reference :294-304)."""
        if self.upsample_conv is None:
            return c.transpose(1, 2).contiguous()
        x = c.unsqueeze(3).contiguous()                                    # NHWC with one channel: (B, cin, Tc, 1)
        for f in self.upsample_conv:
            if isinstance(f, nn.ReLU):
                x = ops.norm_act(x, None, "none", ops.ACT_RELU)
            else:
                x = ops.conv2d(x, effective_weight(f), f.bias, f.stride, f.padding, True)
        return x.squeeze(3).transpose(1, 2).contiguous()

    # ---- synthesis -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def incremental_forward(self, initial_input=None, c=None, g=None, T=100, test_inputs=None, tqdm=lambda x: x, softmax=True,
                            quantize=True, log_scale_min=-7.0, uniforms=None, return_logits=False):
        """Reference signature (:237-240) plus two keyword extras used by the parity tests: ``uniforms`` (T, B, nr_mix + 1)
        supplies the draws of mixture.py:136,148 (default: fresh uniform(1e-5, 1 - 1e-5) draws); ``return_logits`` also
        returns the (B, T, out_channels) pre-sampling outputs.  Returns (B, 1, T)."""
        if g is not None:
            raise NotImplementedError("global conditioning is outside the VIAI hot path")
        if initial_input is not None and float(initial_input.abs().max()) != 0.0:
            raise NotImplementedError("a non-zero initial_input is not supported by the fused synthesis kernel")
        Lh = _lib.lib()
        L, R, G, S, C, K, O = self._dims()
        dev = self.first_conv.bias.device
        if dev.type != "cuda":
            raise RuntimeError("WaveNet synthesis needs CUDA parameters; there is no CPU path")
        B = 1
        if test_inputs is not None:
            if test_inputs.size(1) == 1:
                test_inputs = test_inputs.transpose(1, 2).contiguous()     # (B, T', 1)
            B = test_inputs.size(0)
            T = test_inputs.size(1) if T is None else max(T, test_inputs.size(1))
        elif c is not None:
            B = c.size(0)
        T = int(T)
        if c is None:
            if C > 0:
                raise RuntimeError("local conditioning features are required (cin_channels = %d)" % C)
            cond = torch.zeros((B, T, 4), device=dev)
        else:
            cond = self._upsample(c.to(dev).float())
            assert cond.size(1) == T, "upsampled conditioning covers %d steps, T = %d" % (cond.size(1), T)
        # Kernel choice (VIAI_WAVENET_KERNEL; default "auto" = the fastest one that supports the configuration).  Samples/s of the
        # C4 network on one B200 at B = 1, T = 8000 (scripts/r02_wavenet_folded.py):
        #   ws      22.4 k  csrc/wavenet_synth3.cu  folded schedule (one dependent exchange per layer), warp-specialised
        #   folded   9.5 k  csrc/wavenet_synth2.cu  folded schedule, one instruction stream
        #   grid     7.1 k  csrc/wavenet_synth.cu   two exchanges per layer (round 1)
        #   cluster  5.4 k  csrc/wavenet_synth_cluster.cu  one 16-CTA cluster over distributed shared memory
        import os
        want = os.environ.get("VIAI_WAVENET_KERNEL", "auto")
        if want not in ("auto", "ws", "folded", "grid", "cluster"):
            raise RuntimeError("VIAI_WAVENET_KERNEL must be one of auto, ws, folded, grid, cluster (got %r)" % want)
        if want == "auto":
            want = ("ws" if Lh.viai_wavenet3_num_ctas(L, R, G, S, C, K, O, B) > 0 else
                    "folded" if Lh.viai_wavenet2_num_ctas(L, R, G, S, C, K, O, B) > 0 else "grid")
        cluster = want == "cluster" and Lh.viai_wavenet_cluster_supported(R, G, S, C, K, O, B) > 0
        if want == "cluster" and not cluster:
            raise RuntimeError("the cluster synthesis kernel does not support this configuration")
        ws = want == "ws" and Lh.viai_wavenet3_num_ctas(L, R, G, S, C, K, O, B) > 0
        folded = ws or (want == "folded" and Lh.viai_wavenet2_num_ctas(L, R, G, S, C, K, O, B) > 0)
        if want in ("folded", "ws") and not folded:
            raise RuntimeError("the %s synthesis kernel does not support this configuration" % want)
        self.last_synthesis_kernel = want
        nC = 16 if cluster else (Lh.viai_wavenet3_num_ctas(L, R, G, S, C, K, O, B) if ws else
                                 Lh.viai_wavenet2_num_ctas(L, R, G, S, C, K, O, B) if folded else
                                 Lh.viai_wavenet_num_ctas(R, G, S, C, K, O, B))
        if nC <= 0:
            raise RuntimeError("unsupported WaveNet configuration for the synthesis kernel (R=%d G=%d S=%d C=%d K=%d O=%d B=%d)"
                               % (R, G, S, C, K, O, B))
        # Re-linearised on every call, like the reference (clear_buffer() on entry, wavenet.py:263 -> conv.py:48-62): the weights
        # may have been changed by an optimizer step, load_state_dict or make_generation_fast_ since the last synthesis, and
        # packing 99 MB costs ~1 ms against seconds of synthesis.
        pk = self._packed = self.pack_for_synthesis_folded(nC) if folded else self.pack_for_synthesis(nC)
        nm = O // 3
        if uniforms is None:
            uniforms = torch.empty((T, B, nm + 1), device=dev).uniform_(1e-5, 1.0 - 1e-5)
        uniforms = uniforms.to(dev).float().contiguous()
        assert tuple(uniforms.shape) == (T, B, nm + 1)
        ring_len = [(K - 1) * 2 ** (l % self.layers_per_stack) + 1 for l in range(L)]
        offs, tot = [], 0
        for rl in ring_len:
            offs.append(tot)
            tot += rl * B * R
        ring = torch.zeros(tot, device=dev)
        ring_off = torch.tensor(offs, device=dev, dtype=torch.int64)
        # exchange buffers of 64-bit {value, stage tag} words, zero = "never written"
        nrep = Lh.viai_wavenet3_replicas() if ws else 1        # copies of every exchanged vector (spreads the polling over L2 slices)
        nbuf = 3 if folded else 1                                # h / x vectors in flight
        gbuf, sbuf, hbuf = (torch.zeros(2 * nbuf * nrep * B * (G // 2), device=dev), torch.zeros(2 * nrep * B * S, device=dev),
                            torch.zeros(2 * nrep * B * S, device=dev))
        bar = torch.zeros(2 * nbuf * nrep * B * R, device=dev, dtype=torch.int32)
        out = torch.empty((B, T), device=dev)
        logits = torch.empty((B, T, O), device=dev) if return_logits else None
        ti = None
        if test_inputs is not None:
            ti = test_inputs.to(dev).float().reshape(B, -1).contiguous()
        if cluster:
            _lib.check(Lh.viai_wavenet_synth_cluster(L, self.layers_per_stack, R, G, S, C, K, O, B, T, _p(pk["layers"]), _p(pk["first"]),
                                                     _p(pk["head1"]), _p(pk["head2"]), _p(cond), _p(uniforms), _p(ti),
                                                     0 if ti is None else ti.size(1), float(log_scale_min), _p(ring), _p(ring_off),
                                                     _p(out), _p(logits),
                                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "wavenet_synth_cluster")
            out = out.view(B, 1, T)
            return (out, logits) if return_logits else out
        if folded:
            fn = Lh.viai_wavenet_synth3 if ws else Lh.viai_wavenet_synth2
            _lib.check(fn(L, self.layers_per_stack, R, G, S, C, K, O, B, T, nC, _p(pk["layers"]), _p(pk["last"]), _p(pk["first"]),
                          _p(pk["head1"]), _p(pk["head2"]), _p(cond), _p(uniforms), _p(ti), 0 if ti is None else ti.size(1),
                          float(log_scale_min), _p(ring), _p(ring_off), _p(gbuf), _p(sbuf), _p(hbuf), _p(bar), _p(out), _p(logits),
                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "wavenet_synth3" if ws else "wavenet_synth2")
            out = out.view(B, 1, T)
            return (out, logits) if return_logits else out
        _lib.check(Lh.viai_wavenet_synth(L, self.layers_per_stack, R, G, S, C, K, O, B, T, nC, _p(pk["layers"]), _p(pk["first"]),
                                         _p(pk["head1"]), _p(pk["head2"]), _p(cond), _p(uniforms), _p(ti),
                                         0 if ti is None else ti.size(1), float(log_scale_min), _p(ring), _p(ring_off), _p(gbuf),
                                         _p(sbuf), _p(hbuf), _p(bar), _p(out), _p(logits),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "wavenet_synth")
        out = out.view(B, 1, T)
        return (out, logits) if return_logits else out

    def forward(self, x, c=None, g=None, softmax=False):
        """Teacher-forced, T-parallel outputs (B, out_channels, T) for x (B, 1, T) (reference :177-235), differentiable:
        every layer is one tensor-core GEMM over the B*T rows for the dilated convolution + conditioning, a gate kernel and
        two 1x1 GEMMs (``ResidualConv1dGLU.forward_rows``).  This is the training path."""
        if softmax:
            raise TypeError("softmax() got an unexpected keyword argument 'dim'")   # reference :233 fails the same way
        if g is not None:
            raise NotImplementedError("global conditioning is outside the VIAI hot path")
        B, _, T = x.size()                     # CPU tensors are refused by the first op (there is no CPU path)
        cond = None
        if c is not None:
            cond = self._upsample(c.float())
            assert cond.size(1) == T, "upsampled conditioning covers %d steps, x has %d" % (cond.size(1), T)
        h = rows_linear(x.float().reshape(B, T, 1), effective_weight(self.first_conv).reshape(-1, 1), self.first_conv.bias)
        skips = None
        for f in self.conv_layers:
            h, s = f.forward_rows(h, cond)
            skips = s if skips is None else ops.axpby(skips, math.sqrt(0.5), s, math.sqrt(0.5))      # :219-223
        y = skips
        for f in self.last_conv_layers:
            if isinstance(f, nn.ReLU):
                y = ops.norm_act(y.unsqueeze(0), None, "none", ops.ACT_RELU).squeeze(0)
            else:
                y = rows_linear(y, effective_weight(f).reshape(f.out_channels, -1), f.bias)
        return y.transpose(1, 2)

    def forward_incremental_kernel(self, x, c=None):
        """The same teacher-forced outputs evaluated by the synthesis kernel (incremental == batch); no
        autograd.  Kept as a cross-check of the two paths."""
        _, logits = self.incremental_forward(c=c, T=x.size(-1), test_inputs=x, return_logits=True)
        return logits.transpose(1, 2).contiguous()
