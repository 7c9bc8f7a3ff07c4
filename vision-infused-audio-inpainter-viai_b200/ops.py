"""Autograd operators of the VIAI hot path.  Every op launches kernels of libviai_b200.so through the C ABI;
PyTorch only owns memory, streams and the autograd tape.  Activations are NHWC fp32 CUDA tensors of shape
(N, H, W, C)."""
import ctypes

import torch

from . import _lib
from ._lib import ConvGeom

import os

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3

# Arithmetic of the convolutions (wherever the geometry is tensor-core shaped; CUDA cores otherwise):
#   "fp16x3" (default): tcgen05 tensor cores; FORWARD convolutions split both operands into fp16 pairs of the scaled value
#            (x * 8 = hi + lo, w * 2^k = hi + lo: 22 significand bits) and accumulate hi*hi + hi*lo + lo*hi in fp32: product
#            error ~2^-21 -- fp32-class activations (spectrogram ~1e-5 from the fp32 reference), which matters beyond the
#            spectrogram bound because every ReLU / LeakyReLU / L1-sign decision of the backward pass is taken on them: a
#            forward that is 30x less accurate (bf16 pairs) flips ~30x more of those decisions and the gradients land
#            ~sqrt(30)x further from the exact ones (measured).  Same MMA rate as bf16 pairs.  Data
#            gradients: bf16 pairs (below); weight gradients: one tf32 product on operands rounded in shared memory;
#   "bf16x3": tcgen05 tensor cores; forward convolutions AND data gradients split both operands into bf16 pairs
#            (x = hi + lo) and accumulate hi*hi + hi*lo + lo*hi in fp32 (product error ~2^-17: keeps spectrograms within
#            the 1e-3 parity bound, and keeps the gradient that is chained through ~30 layers at the reference's own fp32
#            envelope) at the bf16 MMA rate; weight gradients (one hop off the chain, nothing propagates) use one tf32
#            product.  A single-tf32 data gradient truncates both operands, which shrinks every layer's gradient by
#            ~2^-11 and compounds to 1.4e-2 at the encoder (scratch experiment reproduced on the CPU);
#            VIAI_DGRAD=tf32 / set_dgrad_x3(False) restores it for comparison;
#   "tf32x3": the same 3-term scheme on tf32 pairs (~2^-21 per product, half the MMA rate of bf16x3);
#   "tf32":  one tf32 product everywhere (what cuDNN does by default on Ampere+; ~1e-2 end to end on this network);
#   "fp32":  CUDA-core fp32 everywhere (the exact path, also the on-device validator of the other two).
_PRECISION = os.environ.get("VIAI_PRECISION", "fp16x3")
_FWD_MODE = {"fp16x3": (3, 8 | 32), "bf16x3": (2, 8), "tf32x3": (1, 4), "tf32": (0, 0)}   # precision -> (weight packing, viai_conv2d_tc flags)
# data gradients: the operand is a gradient tensor whose magnitude is not known in advance (1e-8 .. 1e-2), which rules out the
# fp16 pairs' fixed scale; bf16 pairs have fp32's exponent range, and the backward pass is linear in dy (no activation decisions
# depend on it), so their 2^-17 per product is ample
_DGRAD_MODE = {"fp16x3": (2, 8), "bf16x3": (2, 8), "tf32x3": (1, 4), "tf32": (0, 0)}
if os.environ.get("VIAI_DGRAD") == "tf32x3":          # experiment knob: 2^-21 data gradients at half the MMA rate
    _DGRAD_MODE = {k: (1, 4) for k in _DGRAD_MODE}
_WGRAD_FP32 = os.environ.get("VIAI_WGRAD") == "fp32"  # experiment knob: CUDA-core fp32 weight gradients
_WS = {}
# Fuse the first pass of a layer's norm backward (sum g, sum g*xhat) into the epilogue of the data-gradient convolution that
# produces dz (tensor-core path, BatchNorm batch statistics): viai_conv2d_tc_bwd_reduce.  Parity-tested, OFF by default: round 1
# measured it on every layer (15.5 ms vs 14.8 ms: on wide layers the epilogue warps are busy), round 2 on thin layers only, where
# the epilogue warps idle (15.49 ms vs 15.27 ms: the fused epilogue still has to read y, so only the dz read is saved, and the
# data gradient itself gets slower).  VIAI_FUSE_BWD_REDUCE = 0 (default): never, 1: every unit-stride layer, thin: <= 32 channels.
_FUSE_MODE = os.environ.get("VIAI_FUSE_BWD_REDUCE", "0")
_FUSE_BWD_REDUCE = _FUSE_MODE != "0"
_FUSE_MAX_C = 32 if _FUSE_MODE == "thin" else 1 << 30
_DGRAD_X3 = os.environ.get("VIAI_DGRAD", "x3") != "tf32"


# Fast paths of the ResNet stem (see include/viai_b200.h): the 7x7 / Cin <= 4 weight gradient as im2col + the tensor-core 1x1 weight
# gradient, and the max-pool backward that reads the forward output.  Validated on B200 (tests/test_fast_stem_gpu.py, round 2) and
# ON by default; VIAI_FAST_STEM=0 restores the CUDA-core stem weight gradient / gather-form max-pool backward.
_FAST_STEM = os.environ.get("VIAI_FAST_STEM", "1") == "1"


def set_precision(p):
    global _PRECISION
    if p not in ("fp16x3", "bf16x3", "tf32x3", "tf32", "fp32"):
        raise ValueError("precision must be 'fp16x3', 'bf16x3', 'tf32x3', 'tf32' or 'fp32'")
    prev, _PRECISION = _PRECISION, p
    return prev


def get_precision():
    return _PRECISION


def f16_overflow(reset=True):
    """Number of split-warp threads that saturated an activation (|x| >= 8188) in an fp16-pair forward convolution since the last
    reset.  Synchronises with the device: call it where the step's loss is read anyway.  Non-zero means the affected outputs are
    inaccurate (finite, never inf) -- use VIAI_PRECISION=bf16x3 / tf32x3 for such a model."""
    n = ctypes.c_uint(0)
    _lib.check(_lib.lib().viai_tc_f16_overflow(int(reset), ctypes.byref(n)), "tc_f16_overflow")
    return int(n.value)


def check_f16_overflow():
    n = f16_overflow(True)
    if n:
        raise RuntimeError("fp16x3 convolutions saturated activations beyond +-8188 in %d thread(s): results are inaccurate; "
                           "run with VIAI_PRECISION=bf16x3 (or tf32x3)" % n)


# Measurement hook of bench.py: while set to a list, every convolution / normalisation / resampling operator brackets its launches
# with CUDA events on the launching stream and appends (family, geometry key, algorithmic flops, algorithmic bytes, start, end).
# None (the default) costs one comparison per op.
_OP_LOG = None


class time_ops(object):
    """``with ops.time_ops() as log: step()`` -> after a device synchronise, ``summarize_ops(log)`` gives per-(family, geometry)
    launch counts, total device time, algorithmic FLOPs and bytes of one eager step."""

    def __enter__(self):
        global _OP_LOG
        _OP_LOG = []
        return _OP_LOG

    def __exit__(self, *exc):
        global _OP_LOG
        _OP_LOG = None
        return False


class _op_timer(object):
    __slots__ = ("fam", "key", "flops", "nbytes", "e0")

    def __init__(self, fam, key, flops, nbytes):
        self.fam, self.key, self.flops, self.nbytes = fam, key, flops, nbytes

    def __enter__(self):
        if _OP_LOG is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _OP_LOG is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _OP_LOG.append((self.fam, self.key, self.flops, self.nbytes, self.e0, e1))
        return False


def summarize_ops(log):
    """{(family, key): dict(launches, ms, flops, bytes)} from a ``time_ops`` log (call after torch.cuda.synchronize())."""
    out = {}
    for fam, key, flops, nbytes, e0, e1 in log:
        d = out.setdefault((fam, key), dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        d["launches"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += flops
        d["bytes"] += nbytes
    return out


def _conv_key(N, H, W, C, Ho, Wo, Cout, R, S, stride, transposed):
    return "%s%d->%d %dx%d s%d,%d in %dx%dx%d out %dx%d" % ("convT " if transposed else "conv ", C, Cout, R, S, stride[0], stride[1],
                                                           N, H, W, Ho, Wo)


# Diagnostic hook of the parity tests: while set to a list, every ReLU / LeakyReLU site of _NormActFn appends the bool tensor
# (output > 0) == (pre-activation > 0): the decisions the backward pass will take.  None (the default) costs nothing.
_ACT_TRACE = None


class trace_activation_decisions(object):
    def __enter__(self):
        global _ACT_TRACE
        _ACT_TRACE = []
        return _ACT_TRACE

    def __exit__(self, *exc):
        global _ACT_TRACE
        _ACT_TRACE = None
        return False


def set_dgrad_x3(flag):
    """True (default): data gradients use the forward's 3-term product; False: one tf32 product (round-1 behaviour)."""
    global _DGRAD_X3
    prev, _DGRAD_X3 = _DGRAD_X3, bool(flag)
    return prev


def _workspace(device, n):
    """Per-device fp32 scratch for the tensor-core weight gradient (grown on demand, reused by every layer)."""
    w = _WS.get(device)
    if w is None or w.numel() < n:
        w = torch.empty(max(n, 1 << 22), device=device, dtype=torch.float32)
        _WS[device] = w
    return w


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _ptr(t):
    return None if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError("VIAI ops need float32 CUDA tensors (got %s on %s); there is no CPU path" % (t.dtype, t.device))


def grad_target(param):
    """Gradient-bucket view installed by viai_b200.optim.GradBucket (fused wgrad accumulation), else None."""
    return None if param is None else getattr(param, "_viai_grad", None)


def _geom(N, Hin, Win, Cin, Hout, Wout, Cout, R, S, stride, pad, mode):
    return ConvGeom(N, Hin, Win, Cin, Hout, Wout, Cout, R, S, stride[0], stride[1], pad[0], pad[1], mode)


def _pack(weight, O_dim, I_dim, flip=False):
    """[O][R][S][I] operand from a (d0, d1, kh, kw) parameter; O_dim/I_dim say which of d0/d1 is O/I."""
    if _PACK_STATE["on"]:
        key = _pack_key(weight, O_dim, I_dim, -1 - int(flip))
        if _PACK_REC is not None:
            _PACK_REC.append((weight, O_dim, I_dim, -1 - int(flip)))
        hit = _PACK_CACHE.get(key)
        if hit is not None:
            return hit
        out = _PACK_CACHE[key] = _pack_uncached(weight, O_dim, I_dim, flip)
        return out
    return _pack_uncached(weight, O_dim, I_dim, flip)


def _pack_uncached(weight, O_dim, I_dim, flip=False):
    L = _lib.lib()
    O, I = weight.size(O_dim), weight.size(I_dim)
    R, S = weight.size(2), weight.size(3)
    out = torch.empty((O, R, S, I), device=weight.device, dtype=torch.float32)
    _lib.check(L.viai_pack_weight(_p(weight), _p(out), O, I, R, S, weight.stride(O_dim), weight.stride(I_dim),
                                  weight.stride(2), weight.stride(3), int(flip), _stream()), "pack_weight")
    return out


def _conv_out_size(H, k, s, p, transposed):
    return (H - 1) * s - 2 * p + k if transposed else (H + 2 * p - k) // s + 1


# Packed operands can be cached for the duration of ONE training step (GanTrainer brackets its step with pack_cache_begin /
# pack_cache_end): the discriminator runs its forward twice and its data gradient twice between two updates of its weights.
# FusedAdam rewrites the flat parameter buffer from a kernel (no autograd version bump), so it advances the epoch explicitly
# (weights_updated()).  Outside a bracketed step nothing is cached: a stale operand can never be picked up.
_PACK_CACHE = {}
_PACK_STATE = {"on": False, "epoch": 0}


def pack_cache_begin():
    _PACK_CACHE.clear()
    _PACK_STATE["on"] = True


def pack_cache_end():
    _PACK_CACHE.clear()
    _PACK_STATE["on"] = False


def weights_updated():
    _PACK_STATE["epoch"] += 1
    _PACK_CACHE.clear()
    if _PACK_REC is not None:
        _PACK_REC.append(None)           # segment boundary: what follows needs the updated weights


def _pack_key(weight, O_dim, I_dim, split):
    return (weight.data_ptr(), weight._version, _PACK_STATE["epoch"], O_dim, I_dim, split, tuple(weight.shape), tuple(weight.stride()))


def _pack_tc(weight, O_dim, I_dim, split):
    if not _PACK_STATE["on"]:
        return _pack_tc_uncached(weight, O_dim, I_dim, split)
    if _PACK_REC is not None:
        _PACK_REC.append((weight, O_dim, I_dim, split))
    key = _pack_key(weight, O_dim, I_dim, split)
    hit = _PACK_CACHE.get(key)
    if hit is None:
        hit = _PACK_CACHE[key] = _pack_tc_uncached(weight, O_dim, I_dim, split)
    return hit


# ---- batched packing: a captured training step re-lays every weight a segment needs (between two optimizer updates) in one launch
_PACK_REC = None


def pack_record_begin():
    """Starts recording the (weight, layout) requests of a bracketed step; weights_updated() marks the segment boundaries."""
    global _PACK_REC
    _PACK_REC = []


def pack_record_end():
    """-> list of segments, each a list of unique (parameter, O_dim, I_dim, split) requests (split < 0: the fp32 thin layout,
    -1 plain / -2 flipped).  Only nn.Parameters are kept: their storage is stable across replays of a captured step; derived
    tensors (weight-normed weights, the padded stem matrix) are packed where they are used, as before."""
    global _PACK_REC
    rec, _PACK_REC = _PACK_REC, None
    segs, seen = [[]], set()
    for r in rec or []:
        if r is None:
            segs.append([])
            seen = set()
            continue
        w, od, idim, split = r
        k = (id(w), od, idim, split)
        if isinstance(w, torch.nn.Parameter) and k not in seen:
            seen.add(k)
            segs[-1].append(r)
    return segs


def prepack(requests):
    """Packs every request of one segment with ONE viai_pack_weights_batched launch (per 24 tensors) and puts the results into the
    step's pack cache, where the convolutions find them."""
    if not requests or not _PACK_STATE["on"]:
        return
    L = _lib.lib()
    descs = (_lib.PackDesc * len(requests))()
    outs = []
    for d, (w, od, idim, split) in zip(descs, requests):
        O, I, R, S = w.size(od), w.size(idim), w.size(2), w.size(3)
        if split < 0:
            out = torch.empty((O, R, S, I), device=w.device, dtype=torch.float32)
            kind, flip = 0, int(split == -2)
        else:
            out = torch.empty(L.viai_tc_packed_size(O, I, R, S, split), device=w.device, dtype=torch.float32)
            kind, flip = 1 + split, 0
        d.src, d.dst, d.O, d.I, d.R, d.S = w.data_ptr(), out.data_ptr(), O, I, R, S
        d.so, d.si, d.sr, d.ss, d.flip, d.kind = w.stride(od), w.stride(idim), w.stride(2), w.stride(3), flip, kind
        outs.append(out)
    _lib.check(L.viai_pack_weights_batched(descs, len(requests), _stream()), "pack_weights_batched")
    for (w, od, idim, split), out in zip(requests, outs):
        _PACK_CACHE[_pack_key(w, od, idim, split)] = out


def _pack_tc_uncached(weight, O_dim, I_dim, split):
    L = _lib.lib()
    O, I = weight.size(O_dim), weight.size(I_dim)
    R, S = weight.size(2), weight.size(3)
    out = torch.empty(L.viai_tc_packed_size(O, I, R, S, split), device=weight.device, dtype=torch.float32)
    _lib.check(L.viai_pack_weight_tc(_p(weight), _p(out), O, I, R, S, weight.stride(O_dim), weight.stride(I_dim),
                                     weight.stride(2), weight.stride(3), 0, split, _stream()), "pack_weight_tc")
    return out


def _run_conv(g, x, weight, O_dim, I_dim, bias, y, stats=None, groups=1, forward=False, norm_ctx=None):
    """One gather convolution (forward or data gradient) on the tensor cores when possible, else on the CUDA cores.
    Returns True when ``stats`` (2 x groups*C doubles) was filled by the convolution's epilogue.  With ``norm_ctx`` (data
    gradients only) the epilogue reduction is the producing layer's norm-backward first pass instead of the statistics."""
    L = _lib.lib()
    if norm_ctx is not None:
        if _PRECISION == "fp32" or bias is not None or not L.viai_conv2d_tc_supported(ctypes.byref(g)) or \
                (L.viai_conv2d_thin_supported(ctypes.byref(g)) and x.data_ptr() % 16 == 0):
            norm_ctx, stats = None, None        # not fusable here: plain data gradient, the caller keeps the standalone pass
    if norm_ctx is not None:
        split, flags = _DGRAD_MODE[_PRECISION] if _DGRAD_X3 else (0, 0)
        wp = _pack_tc(weight, O_dim, I_dim, split)
        nb = _lib.NormBwdCtx(norm_ctx["y"].data_ptr(), _ptr(norm_ctx["mean"]), _ptr(norm_ctx["invstd"]), _ptr(norm_ctx["gamma"]),
                             _ptr(norm_ctx["beta"]), norm_ctx["act"], norm_ctx["slope"])
        _lib.check(L.viai_conv2d_tc_bwd_reduce(ctypes.byref(g), _p(x), _p(wp), _p(y), ctypes.byref(nb), _p(stats[0]), _p(stats[1]),
                                               flags, _stream()), "conv2d_tc_bwd_reduce")
        return True
    if L.viai_conv2d_thin_supported(ctypes.byref(g)) and x.data_ptr() % 16 == 0:
        wp = _pack(weight, O_dim, I_dim)
        if stats is not None and groups == 1 and L.viai_conv2d_thin_stats_supported(ctypes.byref(g)):
            # Cin == 1 layer in front of a BatchNorm: the statistics come out of the convolution kernel itself
            _lib.check(L.viai_conv2d_thin_stats(ctypes.byref(g), _p(x), _p(wp), _p(bias), _p(y), _p(stats[0]), _p(stats[1]),
                                                _stream()), "conv2d_thin_stats")
            return True
        _lib.check(L.viai_conv2d_thin(ctypes.byref(g), _p(x), _p(wp), _p(bias), _p(y), _stream()), "conv2d_thin")
        return False
    if _PRECISION != "fp32" and L.viai_conv2d_tc_supported(ctypes.byref(g)):
        split, flags = _FWD_MODE[_PRECISION] if forward else _DGRAD_MODE[_PRECISION] if _DGRAD_X3 else (0, 0)
        wp = _pack_tc(weight, O_dim, I_dim, split)
        _lib.check(L.viai_conv2d_tc(ctypes.byref(g), _p(x), _p(wp), _p(bias), _p(y), _p(stats[0]) if stats is not None else None,
                                    _p(stats[1]) if stats is not None else None, groups, flags, _stream()), "conv2d_tc")
        return stats is not None
    wp = _pack(weight, O_dim, I_dim)
    _lib.check(L.viai_conv2d_simt(ctypes.byref(g), _p(x), _p(wp), _p(bias), _p(y), _stream()), "conv2d")
    return False


class _ConvFn(torch.autograd.Function):
    """nn.Conv2d / nn.ConvTranspose2d forward + convolution_backward (see include/viai_b200.h for the site list).
    ``stat_groups`` > 0 asks for the per-(group, channel) sum / sum of squares of the output (what the following norm
    layer needs): the tensor-core kernel produces them in its epilogue, the CUDA-core path with viai_channel_stats."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, transposed, stat_groups):
        _require_cuda(x, weight, bias)
        x = x.contiguous()
        L = _lib.lib()
        N, H, W, C = x.shape
        R, S = weight.size(2), weight.size(3)
        if not transposed:
            Cout = weight.size(0)
            assert weight.size(1) == C, "conv: weight expects %d input channels, got %d" % (weight.size(1), C)
            od, idim, mode = 0, 1, 0
        else:
            Cout = weight.size(1)
            assert weight.size(0) == C, "convT: weight expects %d input channels, got %d" % (weight.size(0), C)
            od, idim, mode = 1, 0, 1
        Ho = _conv_out_size(H, R, stride[0], padding[0], transposed)
        Wo = _conv_out_size(W, S, stride[1], padding[1], transposed)
        if Ho <= 0 or Wo <= 0:
            raise RuntimeError("conv output size is non-positive: (%d, %d)" % (Ho, Wo))
        y = torch.empty((N, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
        g = _geom(N, H, W, C, Ho, Wo, Cout, R, S, stride, padding, mode)
        stats = None
        if stat_groups > 0:
            stats = torch.empty((2, stat_groups * Cout), device=x.device, dtype=torch.float64)
        macs = (N * Ho * Wo * Cout * C * R * S) if not transposed else (N * H * W * C * Cout * R * S)
        key = _conv_key(N, H, W, C, Ho, Wo, Cout, R, S, stride, transposed)
        # Many-tap few-channel stem (the ResNet 7x7, Cin 3 / 2): too few channels for the implicit-GEMM kernel, 49 taps too many for
        # CUDA cores (11 % of the C3 step).  Forward = im2col (channel-major K = Cin*kh*kw padded to a multiple of 32) + a 1x1
        # tensor-core convolution over the patch rows, statistics epilogue included; the weight gradient re-forms the patches
        # (_stem_wgrad) instead of keeping gigabytes of them alive.  Only when the input needs no gradient (it is the video).
        stem = (_FAST_STEM and not transposed and C <= 4 and R * S >= 25 and Cout % 4 == 0 and Cout >= 16 and _PRECISION != "fp32"
                and not ctx.needs_input_grad[0])
        with _op_timer("conv_fwd", key, 2.0 * macs, 4.0 * (x.numel() + y.numel() + weight.numel())):
            if stem:
                K = C * R * S
                Kpad = (K + 31) // 32 * 32
                cols = im2col(g, x, Kpad)
                w2 = torch.zeros((Cout, Kpad, 1, 1), device=x.device, dtype=torch.float32)
                w2.view(Cout, Kpad)[:, :K] = weight.detach().reshape(Cout, K)
                g1 = _geom(N, Ho, Wo, Kpad, Ho, Wo, Cout, 1, 1, (1, 1), (0, 0), 0)
                fused = _run_conv(g1, cols, w2, 0, 1, bias, y, stats, stat_groups, forward=True)
                del cols
            else:
                fused = _run_conv(g, x, weight, od, idim, bias, y, stats, stat_groups, forward=True)
            if not fused and stats is not None:
                _lib.check(L.viai_channel_stats(_p(y), (N * Ho * Wo) // stat_groups, stat_groups, Cout, _p(stats[0]), _p(stats[1]),
                                                _stream()), "channel_stats")
        ctx.op = (key, 2.0 * macs)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding, transposed, bias is not None)
        ctx.targets = (grad_target(weight), grad_target(bias))
        ctx.after_grad = getattr(weight, "_viai_after_grad", None)     # data-parallel overlap trigger (step.GanTrainer)
        ctx.in_norm = getattr(x, "_viai_norm_ctx", None) if _FUSE_BWD_REDUCE else None
        if stats is None:
            stats = torch.empty(0, device=x.device, dtype=torch.float64)
        ctx.mark_non_differentiable(stats)
        ctx.set_materialize_grads(False)      # no zero-filled float64 "gradient" of the statistics output
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        if dy is None:
            return None, None, None, None, None, None, None
        x, weight = ctx.saved_tensors
        stride, padding, transposed, has_bias = ctx.cfg
        L = _lib.lib()
        dy = dy.contiguous()
        N, H, W, C = x.shape
        _, Ho, Wo, Cout = dy.shape
        R, S = weight.size(2), weight.size(3)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            nc = ctx.in_norm
            # fused only for unit-stride layers: a strided convolution's data gradient is four low-K parity-class launches
            # whose short main loops cannot hide the epilogue's extra work
            if nc is not None and (nc["y"].shape != x.shape or not nc["y"].is_contiguous() or tuple(stride) != (1, 1) or C > _FUSE_MAX_C):
                nc = None
            bst = torch.empty((2, C), device=x.device, dtype=torch.float64) if nc is not None else None
            with _op_timer("conv_dgrad", ctx.op[0], ctx.op[1], 4.0 * (x.numel() + dy.numel() + weight.numel())):
                if not transposed:       # dgrad of Conv2d: transposed gather with wp[o=ci][r][s][i=co]
                    fused = _run_conv(_geom(N, Ho, Wo, Cout, H, W, C, R, S, stride, padding, 1), dy, weight, 1, 0, None, dx, bst,
                                      norm_ctx=nc)
                else:                    # dgrad of ConvTranspose2d: forward gather with wp[o=ci][r][s][i=co]
                    fused = _run_conv(_geom(N, Ho, Wo, Cout, H, W, C, R, S, stride, padding, 0), dy, weight, 0, 1, None, dx, bst,
                                      norm_ctx=nc)
            if nc is not None and fused:
                dx._viai_bwd_stats = (bst, nc["token"])      # consumed by the producing layer's _NormActFn.backward
        wt, bt = ctx.targets
        if ctx.needs_input_grad[1]:
            dw = wt if wt is not None else torch.empty_like(weight, memory_format=torch.contiguous_format)
            if not transposed:       # U = dOut (A = Cout), G = x (B = Cin)
                g = _geom(N, H, W, C, Ho, Wo, Cout, R, S, stride, padding, 0)
                U, G = dy, x
            else:                    # U = x (A = Cin_t), G = dOut (B = Cout_t)
                g = _geom(N, Ho, Wo, Cout, H, W, C, R, S, stride, padding, 0)
                U, G = x, dy
            with _op_timer("conv_wgrad", ctx.op[0], ctx.op[1], 4.0 * (x.numel() + dy.numel() + weight.numel())):
                # (the im2col route is for many-tap few-channel stems -- 7x7, Cin 2 / 3; the 1-channel 3x3 / 1x4 first layers of the
                # encoder / discriminator stay on the wide-tensor-stationary thin kernel, which is 3x faster there: measured)
                if _FAST_STEM and not transposed and C <= 4 and _PRECISION != "fp32" and Cout % 4 == 0 and Cout >= 16 and R * S >= 25:
                    _stem_wgrad(g, x, dy, dw, wt is not None)
                elif L.viai_conv2d_wgrad_thin_supported(ctypes.byref(g)) and U.data_ptr() % 16 == 0 and G.data_ptr() % 16 == 0:
                    ws = _workspace(dy.device, L.viai_wgrad_thin_workspace(ctypes.byref(g)))
                    _lib.check(L.viai_conv2d_wgrad_thin(ctypes.byref(g), _p(U), _p(G), _p(dw), dw.stride(0), dw.stride(1), dw.stride(2),
                                                        dw.stride(3), int(wt is not None), _p(ws), _stream()), "conv2d wgrad_thin")
                elif _PRECISION != "fp32" and not _WGRAD_FP32 and L.viai_conv2d_wgrad_tc_supported(ctypes.byref(g)):
                    ws = _workspace(dy.device, L.viai_wgrad_tc_workspace(ctypes.byref(g)))
                    _lib.check(L.viai_conv2d_wgrad_tc(ctypes.byref(g), _p(U), _p(G), _p(dw), dw.stride(0), dw.stride(1), dw.stride(2),
                                                      dw.stride(3), int(wt is not None), _p(ws), _stream()), "conv2d wgrad_tc")
                else:
                    _lib.check(L.viai_conv2d_wgrad_simt(ctypes.byref(g), _p(U), _p(G), _p(dw), dw.stride(0), dw.stride(1),
                                                        dw.stride(2), dw.stride(3), int(wt is not None), _stream()), "conv2d wgrad")
            if wt is not None:
                dw = None            # accumulated straight into the gradient bucket
        if has_bias and ctx.needs_input_grad[2]:
            rows = N * Ho * Wo
            acc = torch.empty(Cout, device=dy.device, dtype=torch.float64)
            db = bt if bt is not None else torch.empty(Cout, device=dy.device, dtype=torch.float32)
            _lib.check(L.viai_channel_stats(_p(dy), rows, 1, Cout, _p(acc), None, _stream()), "bias grad")
            _lib.check(L.viai_fold_groups(_p(acc), 1, Cout, _p(db), int(bt is not None), _stream()), "bias grad fold")
            if bt is not None:
                db = None
        if ctx.after_grad is not None and ctx.needs_input_grad[1]:
            ctx.after_grad()         # weight AND bias gradient of the trigger layer are launched: its bucket range is complete
        return dx, dw, db, None, None, None, None


def im2col(g, x, Kpad):
    """(N*Hout*Wout, Kpad) channel-major patches of the forward convolution ``g`` reads from x (viai_im2col)."""
    out = torch.empty((g.N * g.Hout * g.Wout, Kpad), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().viai_im2col(ctypes.byref(g), _p(x), Kpad, _p(out), _stream()), "im2col")
    return out


def pointwise_wgrad(U, G):
    """dW (A, B) = U^T G for row tensors U (P, A), G (P, B) on the tensor-core weight-gradient kernel (1x1 geometry)."""
    L = _lib.lib()
    P, A = U.shape
    B = G.size(1)
    W = next(w for w in (128, 64, 32, 16, 8, 4, 2, 1) if P % w == 0)
    g1 = _geom(1, P // W, W, B, P // W, W, A, 1, 1, (1, 1), (0, 0), 0)
    if not L.viai_conv2d_wgrad_tc_supported(ctypes.byref(g1)):
        raise RuntimeError("pointwise_wgrad: unsupported sizes (A %d, B %d)" % (A, B))
    dw = torch.empty((A, B), device=U.device, dtype=torch.float32)
    ws = _workspace(U.device, L.viai_wgrad_tc_workspace(ctypes.byref(g1)))
    _lib.check(L.viai_conv2d_wgrad_tc(ctypes.byref(g1), _p(U), _p(G), _p(dw), B, 1, 0, 0, 0, _p(ws), _stream()), "pointwise wgrad")
    return dw


def _stem_wgrad(g, x, dy, dw, accumulate):
    """Weight gradient of a few-channel Conv2d (the ResNet stem) as im2col + one 1x1 weight gradient; ``dw`` is the
    (Cout, Cin, kh, kw) target (a gradient-bucket view when ``accumulate``)."""
    K = g.Cin * g.R * g.S
    Kpad = (K + 31) // 32 * 32
    cols = im2col(g, x, Kpad)
    part = pointwise_wgrad(dy.reshape(-1, g.Cout), cols)[:, :K].reshape(dw.shape)
    if accumulate:
        dw.add_(part)
    else:
        dw.copy_(part)


def conv2d(x, weight, bias=None, stride=(1, 1), padding=(0, 0), transposed=False):
    return _ConvFn.apply(x, weight, bias, tuple(stride), tuple(padding), transposed, 0)[0]


def conv2d_stats(x, weight, bias, stride, padding, transposed, stat_groups):
    """Convolution that also returns the (2, groups*C) double statistics of its output for the following norm layer."""
    return _ConvFn.apply(x, weight, bias, tuple(stride), tuple(padding), transposed, int(stat_groups))


class _NormActFn(torch.autograd.Function):
    """norm in {'bn','in','none'} followed by an activation.  Replaces native_batch_norm / instance_norm +
    leaky_relu / relu / sigmoid (networks/Inpainting_Networks.py:72-76, networks/New_Inpainting_Networks.py:31-33,
    networks/Discriminator_Networks.py:38-49)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, nbt, norm, training, act, slope, eps, momentum, pre_stats=None,
                side=None):
        _require_cuda(y, gamma, beta)
        L = _lib.lib()
        y = y.contiguous()
        N, H, W, C = y.shape
        dev = y.device
        mean = invstd = fused_stats = None
        groups, rpg = 1, N * H * W
        if norm == "in":
            groups, rpg = N, H * W
        if norm in ("bn", "in"):
            use_batch = training or norm == "in" or running_mean is None
            if use_batch:
                mean = torch.empty(groups * C, device=dev, dtype=torch.float32)
                invstd = torch.empty(groups * C, device=dev, dtype=torch.float32)
                if pre_stats is not None and pre_stats.numel() == 2 * groups * C:
                    s = pre_stats          # produced by the convolution's epilogue
                else:
                    s = torch.empty((2, groups * C), device=dev, dtype=torch.float64)
                    _lib.check(L.viai_channel_stats(_p(y), rpg, groups, C, _p(s[0]), _p(s[1]), _stream()), "channel_stats")
                upd = norm == "bn" and training and running_mean is not None
                fused_stats = (s, upd)     # statistics are finalised inside the apply kernel (one launch instead of two)
            else:
                mean = running_mean
                invstd = torch.empty(C, device=dev, dtype=torch.float32)
                _lib.check(L.viai_rsqrt_eps(_p(running_var), C, eps, _p(invstd), _stream()), "rsqrt_eps")
        out = torch.empty_like(y)
        with _op_timer("norm_act_fwd", "%dx%dx%dx%d" % (N, H, W, C), 0.0, 8.0 * y.numel()):
            if fused_stats is not None:
                s, upd = fused_stats
                _lib.check(L.viai_norm_finalize_act_fwd(_p(y), rpg, groups, C, _p(s[0]), _p(s[1]), eps, _p(gamma), _p(beta), act, slope,
                                                        _p(out), _p(mean), _p(invstd), _p(running_mean) if upd else None,
                                                        _p(running_var) if upd else None, momentum, _p(nbt) if upd else None,
                                                        _stream()), "norm_finalize_act_fwd")
            else:
                _lib.check(L.viai_norm_act_fwd(_p(y), rpg, groups, C, _p(mean), _p(invstd), _p(gamma), _p(beta), act, slope,
                                               _p(out), _stream()), "norm_act_fwd")
        if _ACT_TRACE is not None and act in (ACT_RELU, ACT_LRELU):
            _ACT_TRACE.append(out > 0)
        ctx.save_for_backward(y, mean, invstd, gamma, beta)
        ctx.cfg = (norm, groups, rpg, act, slope, training or norm != "bn")
        ctx.targets = (grad_target(gamma), grad_target(beta))
        ctx.token = None
        if side is not None and groups == 1 and mean is not None and (training or norm != "bn") and act != ACT_SIGMOID:
            ctx.token = object()
            side.update(y=y, mean=mean, invstd=invstd, gamma=gamma, beta=beta, act=act, slope=float(slope), token=ctx.token)
        return out

    @staticmethod
    def backward(ctx, dz):
        y, mean, invstd, gamma, beta = ctx.saved_tensors
        norm, groups, rpg, act, slope, batch_stats = ctx.cfg
        if norm == "bn" and not batch_stats:
            raise RuntimeError("backward through eval-mode BatchNorm is not part of the VIAI hot path")
        L = _lib.lib()
        dz = dz.contiguous()
        C = y.size(3)
        dy = torch.empty_like(y)
        gt, bt = ctx.targets
        s = None
        pre = getattr(dz, "_viai_bwd_stats", None)
        with _op_timer("norm_act_bwd", "x".join(str(d) for d in y.shape), 0.0, 12.0 * y.numel()):
            if pre is not None and ctx.token is not None and pre[1] is ctx.token:
                s = pre[0]               # (sum g, sum g*xhat) came out of the epilogue of the convolution that produced dz
            elif mean is not None:
                s = torch.empty((2, groups * C), device=y.device, dtype=torch.float64)
                _lib.check(L.viai_norm_act_bwd_reduce(_p(dz), _p(y), rpg, groups, C, _p(mean), _p(invstd), _p(gamma), _p(beta),
                                                      act, slope, _p(s[0]), _p(s[1]), _stream()), "norm_act_bwd_reduce")
            want_g = s is not None and gamma is not None and ctx.needs_input_grad[1]
            want_b = s is not None and beta is not None and ctx.needs_input_grad[2]
            dgamma = (gt if gt is not None else torch.empty_like(gamma)) if want_g else None
            dbeta = (bt if bt is not None else torch.empty_like(beta)) if want_b else None
            # dgamma / dbeta = the group sums of s2 / s1, written (or accumulated into the gradient bucket) by the apply kernel's
            # first block when both go the same way; otherwise by viai_fold_groups
            in_kernel = mean is not None and (want_g or want_b) and (not (want_g and want_b) or (gt is None) == (bt is None))
            if in_kernel:
                acc = int((gt is not None) if want_g else (bt is not None))
                _lib.check(L.viai_norm_act_bwd_apply_fold(_p(dz), _p(y), rpg, groups, C, _p(mean), _p(invstd), _p(gamma), _p(beta), act,
                                                          slope, _p(s[0]), _p(s[1]), _p(dy), _p(dgamma), _p(dbeta), acc, _stream()),
                           "norm_act_bwd_apply_fold")
            else:
                _lib.check(L.viai_norm_act_bwd_apply(_p(dz), _p(y), rpg, groups, C, _p(mean), _p(invstd), _p(gamma), _p(beta), act,
                                                     slope, _p(s[0]) if s is not None else None, _p(s[1]) if s is not None else None,
                                                     _p(dy), None, None, _stream()), "norm_act_bwd_apply")
                if want_g:
                    _lib.check(L.viai_fold_groups(_p(s[1]), groups, C, _p(dgamma), int(gt is not None), _stream()), "dgamma")
                if want_b:
                    _lib.check(L.viai_fold_groups(_p(s[0]), groups, C, _p(dbeta), int(bt is not None), _stream()), "dbeta")
        if gt is not None:
            dgamma = None            # accumulated straight into the gradient bucket
        if bt is not None:
            dbeta = None
        return dy, dgamma, dbeta, None, None, None, None, None, None, None, None, None, None, None


def stat_groups_for(norm_module, norm, batch):
    """How many statistic groups the norm layer will want from the producing convolution (0: none)."""
    if norm_module is None or norm == "none":
        return 0
    if norm == "in":
        return batch
    return 1 if (norm_module.training or getattr(norm_module, "running_mean", None) is None) else 0


def norm_act(y, norm_module, norm, act, slope=0.0, pre_stats=None):
    """``norm_module`` is the nn.BatchNorm2d / nn.InstanceNorm2d parameter container (or None)."""
    if norm == "none" or norm_module is None:
        return _NormActFn.apply(y, None, None, None, None, None, "none", False, act, slope, 0.0, 0.0, None, None)
    gamma = getattr(norm_module, "weight", None)
    beta = getattr(norm_module, "bias", None)
    rm = getattr(norm_module, "running_mean", None)
    rv = getattr(norm_module, "running_var", None)
    nbt = getattr(norm_module, "num_batches_tracked", None)
    mom = norm_module.momentum if norm_module.momentum is not None else 0.1
    side = {}
    z = _NormActFn.apply(y, gamma, beta, rm, rv, nbt, norm, norm_module.training, act, slope, norm_module.eps, mom, pre_stats, side)
    if side and torch.is_grad_enabled() and z.requires_grad:
        z._viai_norm_ctx = side          # lets the consuming convolution's data gradient fuse this layer's backward reduction
    return z


class _BilinearCatFn(torch.autograd.Function):
    """F.interpolate(bilinear, align_corners=True) [+ torch.cat((out, skip), 1)] written straight into the
    concatenated buffer (networks/New_Inpainting_Networks.py:78-83)."""

    @staticmethod
    def forward(ctx, x, skip, Hout, Wout):
        _require_cuda(x, skip)
        L = _lib.lib()
        x = x.contiguous()
        N, H, W, C = x.shape
        Cs = 0 if skip is None else skip.size(3)
        out = torch.empty((N, Hout, Wout, C + Cs), device=x.device, dtype=torch.float32)
        _lib.check(L.viai_bilinear_fwd(_p(x), N, H, W, C, _p(out), Hout, Wout, C + Cs, 0, _stream()), "bilinear_fwd")
        if skip is not None:
            skip = skip.contiguous()
            assert skip.shape[:3] == (N, Hout, Wout), "cat: spatial sizes differ"
            _lib.check(L.viai_copy_channels(_p(skip), N * Hout * Wout, Cs, 0, _p(out), C + Cs, C, Cs, _stream()), "cat")
        ctx.shape = (N, H, W, C, Hout, Wout, Cs)
        return out

    @staticmethod
    def backward(ctx, dout):
        N, H, W, C, Hout, Wout, Cs = ctx.shape
        L = _lib.lib()
        dout = dout.contiguous()
        dx = dskip = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((N, H, W, C), device=dout.device, dtype=torch.float32)
            _lib.check(L.viai_bilinear_bwd(_p(dout), N, H, W, C, _p(dx), Hout, Wout, C + Cs, 0, _stream()), "bilinear_bwd")
        if Cs and ctx.needs_input_grad[1]:
            dskip = torch.empty((N, Hout, Wout, Cs), device=dout.device, dtype=torch.float32)
            _lib.check(L.viai_copy_channels(_p(dout), N * Hout * Wout, C + Cs, C, _p(dskip), Cs, 0, Cs, _stream()), "cat bwd")
        return dx, dskip, None, None


def bilinear_cat(x, size, skip=None):
    return _BilinearCatFn.apply(x, skip, int(size[0]), int(size[1]))


class _CatFn(torch.autograd.Function):
    """torch.cat along channels for NHWC tensors (networks/New_Inpainting_Networks.py:120)."""

    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, b)
        L = _lib.lib()
        a, b = a.contiguous(), b.contiguous()
        N, H, W, Ca = a.shape
        Cb = b.size(3)
        out = torch.empty((N, H, W, Ca + Cb), device=a.device, dtype=torch.float32)
        rows = N * H * W
        _lib.check(L.viai_copy_channels(_p(a), rows, Ca, 0, _p(out), Ca + Cb, 0, Ca, _stream()), "cat a")
        _lib.check(L.viai_copy_channels(_p(b), rows, Cb, 0, _p(out), Ca + Cb, Ca, Cb, _stream()), "cat b")
        ctx.shape = (N, H, W, Ca, Cb)
        return out

    @staticmethod
    def backward(ctx, dout):
        N, H, W, Ca, Cb = ctx.shape
        L = _lib.lib()
        dout = dout.contiguous()
        rows = N * H * W
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty((N, H, W, Ca), device=dout.device, dtype=torch.float32)
            _lib.check(L.viai_copy_channels(_p(dout), rows, Ca + Cb, 0, _p(da), Ca, 0, Ca, _stream()), "cat bwd a")
        if ctx.needs_input_grad[1]:
            db = torch.empty((N, H, W, Cb), device=dout.device, dtype=torch.float32)
            _lib.check(L.viai_copy_channels(_p(dout), rows, Ca + Cb, Ca, _p(db), Cb, 0, Cb, _stream()), "cat bwd b")
        return da, db


def cat_channels(a, b):
    return _CatFn.apply(a, b)


class _AvgPoolHFn(torch.autograd.Function):
    """nn.AvgPool2d((kh, 1)) (networks/Inpainting_Networks.py:65,77)."""

    @staticmethod
    def forward(ctx, x, kh):
        _require_cuda(x)
        L = _lib.lib()
        x = x.contiguous()
        N, H, W, C = x.shape
        if H < kh:
            raise RuntimeError("Given input size: (%dx%dx%d). Calculated output size: (%dx0x%d). Output size is too small"
                               % (C, H, W, C, W))
        out = torch.empty((N, H // kh, W, C), device=x.device, dtype=torch.float32)
        _lib.check(L.viai_avgpool_h_fwd(_p(x), N, H, W, C, kh, _p(out), _stream()), "avgpool_h_fwd")
        ctx.shape = (N, H, W, C, kh)
        return out

    @staticmethod
    def backward(ctx, dout):
        N, H, W, C, kh = ctx.shape
        L = _lib.lib()
        dout = dout.contiguous()
        dx = torch.empty((N, H, W, C), device=dout.device, dtype=torch.float32)
        _lib.check(L.viai_avgpool_h_bwd(_p(dout), N, H, W, C, kh, _p(dx), _stream()), "avgpool_h_bwd")
        return dx, None


def avgpool_h(x, kh):
    return _AvgPoolHFn.apply(x, kh)


class _MaxPoolFn(torch.autograd.Function):
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (networks/Image_Embedding.py:21).  With C % 4 == 0 the forward pass saves a
    one-byte argmax per output element and the backward pass gathers through it (viai_maxpool3s2_*_idx): neither the input nor the
    output of the pooling stays alive for the backward pass."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        L = _lib.lib()
        x = x.contiguous()
        N, H, W, C = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        out = torch.empty((N, Ho, Wo, C), device=x.device, dtype=torch.float32)
        ctx.shape = (N, H, W, C)
        if _FAST_STEM and C % 4 == 0:
            idx = torch.empty((N, Ho, Wo, C), device=x.device, dtype=torch.uint8)
            _lib.check(L.viai_maxpool3s2_fwd_idx(_p(x), N, H, W, C, _p(out), _p(idx), Ho, Wo, _stream()), "maxpool fwd (argmax)")
            ctx.save_for_backward(idx)
            ctx.mark_non_differentiable(idx)
        else:
            _lib.check(L.viai_maxpool3s2_fwd(_p(x), N, H, W, C, _p(out), Ho, Wo, _stream()), "maxpool fwd")
            ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors[0]
        L = _lib.lib()
        dout = dout.contiguous()
        N, H, W, C = ctx.shape
        dx = torch.empty((N, H, W, C), device=dout.device, dtype=torch.float32)
        if saved.dtype == torch.uint8:
            _lib.check(L.viai_maxpool3s2_bwd_idx(_p(saved), _p(dout), N, H, W, C, _p(dx), dout.size(1), dout.size(2), _stream()),
                       "maxpool bwd (argmax)")
        else:
            _lib.check(L.viai_maxpool3s2_bwd(_p(saved), _p(dout), N, H, W, C, _p(dx), dout.size(1), dout.size(2), _stream()), "maxpool bwd")
        return dx


def maxpool3s2(x):
    return _MaxPoolFn.apply(x)


class _AddActFn(torch.autograd.Function):
    """``out += residual; relu(out)`` (networks/ResNet.py:52-53)."""

    @staticmethod
    def forward(ctx, a, b, act):
        _require_cuda(a, b)
        L = _lib.lib()
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        _lib.check(L.viai_add_act(_p(a), _p(b), _p(out), a.numel(), act, _stream()), "add_act")
        ctx.save_for_backward(out)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        L = _lib.lib()
        dout = dout.contiguous()
        d = torch.empty_like(out)
        _lib.check(L.viai_add_act_bwd(_p(out), _p(dout), _p(d), out.numel(), ctx.act, _stream()), "add_act_bwd")
        return d, d, None


def add_act(a, b, act=ACT_RELU):
    return _AddActFn.apply(a, b, act)


class _MulFn(torch.autograd.Function):
    """mel * mask (bit exact: one IEEE multiply per element)."""

    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, b)
        L = _lib.lib()
        a, b = a.contiguous(), b.contiguous()
        assert a.shape == b.shape
        out = torch.empty_like(a)
        _lib.check(L.viai_mul(_p(a), _p(b), _p(out), a.numel(), _stream()), "mul")
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        L = _lib.lib()
        dout = dout.contiguous()
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(a)
            _lib.check(L.viai_mul(_p(dout), _p(b), _p(da), a.numel(), _stream()), "mul bwd")
        if ctx.needs_input_grad[1]:
            db = torch.empty_like(b)
            _lib.check(L.viai_mul(_p(dout), _p(a), _p(db), a.numel(), _stream()), "mul bwd")
        return da, db


def mul(a, b):
    return _MulFn.apply(a, b)


class _LossFn(torch.autograd.Function):
    """kind 0: MSE vs scalar, 1: BCE vs scalar (loss_functions.py:86-104), 2: L1 between two tensors."""

    @staticmethod
    def forward(ctx, p, q, kind, target):
        _require_cuda(p, q)
        L = _lib.lib()
        p = p.contiguous()
        q = q.contiguous() if q is not None else None
        acc = torch.empty(1, device=p.device, dtype=torch.float64)
        out = torch.empty((), device=p.device, dtype=torch.float32)
        _lib.check(L.viai_loss_fwd(kind, _p(p), _p(q), float(target), p.numel(), _p(acc), _p(out), _stream()), "loss_fwd")
        ctx.save_for_backward(p, q)
        ctx.cfg = (kind, float(target))
        return out

    @staticmethod
    def backward(ctx, gout):
        p, q = ctx.saved_tensors
        kind, target = ctx.cfg
        L = _lib.lib()
        gout = gout.contiguous().float()
        dp = torch.empty_like(p)
        _lib.check(L.viai_loss_bwd(kind, _p(p), _p(q), target, p.numel(), _p(gout), _p(dp), _stream()), "loss_bwd")
        dq = None
        if q is not None and ctx.needs_input_grad[1]:
            dq = -dp
        return dp, dq, None, None


def mse_scalar(p, target):
    return _LossFn.apply(p, None, 0, target)


def bce_scalar(p, target):
    return _LossFn.apply(p, None, 1, target)


def l1_loss(p, q):
    return _LossFn.apply(p, q, 2, 0.0)


class _LinComb2Fn(torch.autograd.Function):
    """wa * a + wb * b for 0-dim device scalars (loss weighting without leaving the library)."""

    @staticmethod
    def forward(ctx, a, b, wa, wb):
        L = _lib.lib()
        out = torch.empty((), device=a.device, dtype=torch.float32)
        _lib.check(L.viai_lincomb2(_p(a), wa, _p(b), wb, _p(out), _stream()), "lincomb2")
        ctx.w = (wa, wb, b is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        wa, wb, has_b = ctx.w
        L = _lib.lib()
        g = g.contiguous()
        da = torch.empty_like(g)
        _lib.check(L.viai_lincomb2(_p(g), wa, None, 0.0, _p(da), _stream()), "lincomb2 bwd")
        db = None
        if has_b:
            db = torch.empty_like(g)
            _lib.check(L.viai_lincomb2(_p(g), wb, None, 0.0, _p(db), _stream()), "lincomb2 bwd")
        return da, db, None, None


def lincomb2(a, wa, b=None, wb=0.0):
    return _LinComb2Fn.apply(a, b, float(wa), float(wb))


def to_nhwc(t):
    """(N,C,H,W) logical tensor -> contiguous (N,H,W,C).  Free when the tensor is already channels-last or C == 1."""
    return t.permute(0, 2, 3, 1).contiguous()


def to_nchw(t):
    """(N,H,W,C) -> (N,C,H,W) view (channels-last strides; no copy)."""
    return t.permute(0, 3, 1, 2)


# ---- WaveNet teacher-forced training path ------------------------------------------------------------------
class _ShiftCatFn(torch.autograd.Function):
    """Operand of a dilated causal Conv1d joined with the conditioning features (viai_shiftcat_fwd / _bwd); an optional
    {0,1} dropout mask and its 1/keep scale are applied to x on the fly."""

    @staticmethod
    def forward(ctx, x, c, K, dilation, Kpad, mask, scale):
        _require_cuda(x, c, mask)
        L = _lib.lib()
        x = x.contiguous()
        B, T, R = x.shape
        Cc = 0
        if c is not None:
            c = c.contiguous()
            Cc = c.size(2)
            assert tuple(c.shape[:2]) == (B, T), "conditioning features must cover the same (B, T) as the input"
        if mask is not None:
            mask = mask.contiguous()
            assert mask.shape == x.shape
        out = torch.empty((B, T, Kpad), device=x.device, dtype=torch.float32)
        _lib.check(L.viai_shiftcat_fwd(_p(x), _p(c), _p(mask), scale, B, T, R, Cc, K, dilation, Kpad, _p(out), _stream()), "shiftcat_fwd")
        ctx.save_for_backward(mask)
        ctx.cfg = (B, T, R, Cc, K, dilation, Kpad, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        B, T, R, Cc, K, dilation, Kpad, scale = ctx.cfg
        L = _lib.lib()
        dout = dout.contiguous()
        dx = torch.empty((B, T, R), device=dout.device, dtype=torch.float32)
        dc = torch.empty((B, T, Cc), device=dout.device, dtype=torch.float32) if (Cc > 0 and ctx.needs_input_grad[1]) else None
        _lib.check(L.viai_shiftcat_bwd(_p(dout), _p(mask), scale, B, T, R, Cc, K, dilation, Kpad, _p(dx), _p(dc), _stream()),
                   "shiftcat_bwd")
        return dx, dc, None, None, None, None, None


def shiftcat(x, c, K, dilation, Kpad, mask=None, scale=1.0):
    """x (B,T,R), c (B,T,Cc) or None -> (B,T,Kpad) with [x(t-(K-1)d) | ... | x(t) | c(t) | 0]; x is read as x * mask * scale
    when a dropout mask is given."""
    return _ShiftCatFn.apply(x, c, int(K), int(dilation), int(Kpad), mask, float(scale))


class _WeightNormFn(torch.autograd.Function):
    """w = v * g / ||v|| per slice of dim 0 (nn.utils.weight_norm, dim 0) in one kernel each way."""

    @staticmethod
    def forward(ctx, v, g):
        _require_cuda(v, g)
        L = _lib.lib()
        v, g = v.contiguous(), g.contiguous()
        rows = v.size(0)
        cols = v.numel() // rows
        w = torch.empty_like(v)
        nrm = torch.empty(rows, device=v.device, dtype=torch.float32)
        _lib.check(L.viai_weight_norm_fwd(_p(v), _p(g), rows, cols, _p(w), _p(nrm), _stream()), "weight_norm_fwd")
        ctx.save_for_backward(v, g, nrm)
        return w

    @staticmethod
    def backward(ctx, dw):
        v, g, nrm = ctx.saved_tensors
        L = _lib.lib()
        dw = dw.contiguous()
        rows = v.size(0)
        dv, dg = torch.empty_like(v), torch.empty_like(g)
        _lib.check(L.viai_weight_norm_bwd(_p(v), _p(g), _p(nrm), _p(dw), rows, v.numel() // rows, _p(dv), _p(dg), _stream()),
                   "weight_norm_bwd")
        return dv, dg


def weight_norm(v, g):
    return _WeightNormFn.apply(v, g)


class _GluFn(torch.autograd.Function):
    """tanh(a) * sigmoid(b) on the two channel halves (wavenet_vocoder/modules.py:180,196)."""

    @staticmethod
    def forward(ctx, y):
        _require_cuda(y)
        L = _lib.lib()
        y = y.contiguous()
        G = y.size(-1)
        rows = y.numel() // G
        out = torch.empty(y.shape[:-1] + (G // 2,), device=y.device, dtype=torch.float32)
        _lib.check(L.viai_glu_fwd(_p(y), rows, G, _p(out), _stream()), "glu_fwd")
        ctx.save_for_backward(y)
        return out

    @staticmethod
    def backward(ctx, dout):
        (y,) = ctx.saved_tensors
        L = _lib.lib()
        dout = dout.contiguous()
        G = y.size(-1)
        dy = torch.empty_like(y)
        _lib.check(L.viai_glu_bwd(_p(y), _p(dout), y.numel() // G, G, _p(dy), _stream()), "glu_bwd")
        return dy


def glu_tanh_sigmoid(y):
    return _GluFn.apply(y)


class _AxpbyFn(torch.autograd.Function):
    """alpha * a + beta * b (same shapes)."""

    @staticmethod
    def forward(ctx, a, b, alpha, beta):
        _require_cuda(a, b)
        L = _lib.lib()
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
        assert b is None or b.shape == a.shape
        out = torch.empty_like(a)
        _lib.check(L.viai_axpby(_p(a), alpha, _p(b), beta, _p(out), a.numel(), _stream()), "axpby")
        ctx.cfg = (alpha, beta, b is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        alpha, beta, has_b = ctx.cfg
        L = _lib.lib()
        g = g.contiguous()
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(g)
            _lib.check(L.viai_axpby(_p(g), alpha, None, 0.0, _p(da), g.numel(), _stream()), "axpby bwd")
        if has_b and ctx.needs_input_grad[1]:
            if da is not None and alpha == beta:
                db = da                  # (a + b) * s: both inputs receive the same gradient tensor
            else:
                db = torch.empty_like(g)
                _lib.check(L.viai_axpby(_p(g), beta, None, 0.0, _p(db), g.numel(), _stream()), "axpby bwd")
        return da, db, None, None


def axpby(a, alpha, b=None, beta=0.0):
    return _AxpbyFn.apply(a, b, float(alpha), float(beta))


def axpby_(a, alpha, b, beta):
    """In place on ``a`` (no autograd): the EMA update of loss_functions.py:70-73."""
    _require_cuda(a, b)
    assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape
    _lib.check(_lib.lib().viai_axpby(_p(a), float(alpha), _p(b), float(beta), _p(a), a.numel(), _stream()), "axpby_")
    return a


class _DmolNllFn(torch.autograd.Function):
    """discretized_mix_logistic_loss(reduce=False) on rows of 3*nr_mix parameters (wavenet_vocoder/mixture.py:25-105)."""

    @staticmethod
    def forward(ctx, y_hat, target, num_classes, log_scale_min):
        _require_cuda(y_hat, target)
        L = _lib.lib()
        y_hat = y_hat.contiguous()
        target = target.contiguous()
        C = y_hat.size(-1)
        assert C % 3 == 0
        rows = y_hat.numel() // C
        assert target.numel() == rows, "one target per row of y_hat"
        nll = torch.empty(y_hat.shape[:-1], device=y_hat.device, dtype=torch.float32)
        _lib.check(L.viai_dmol_nll(_p(y_hat), _p(target), rows, C // 3, num_classes, log_scale_min, _p(nll), None, None, _stream()),
                   "dmol_nll")
        ctx.save_for_backward(y_hat, target)
        ctx.cfg = (rows, C // 3, num_classes, log_scale_min)
        return nll

    @staticmethod
    def backward(ctx, dnll):
        y_hat, target = ctx.saved_tensors
        rows, nm, num_classes, lsm = ctx.cfg
        L = _lib.lib()
        dnll = dnll.contiguous()
        dy = torch.empty_like(y_hat)
        _lib.check(L.viai_dmol_nll(_p(y_hat), _p(target), rows, nm, num_classes, lsm, None, _p(dnll), _p(dy), _stream()), "dmol_nll bwd")
        return dy, None, None, None


def dmol_nll(y_hat_rows, target, num_classes=256, log_scale_min=-7.0):
    """y_hat_rows (..., 3*nr_mix), target (...) -> per-sample negative log-likelihood (...)."""
    return _DmolNllFn.apply(y_hat_rows, target, int(num_classes), float(log_scale_min))


def dmol_sample(y_hat_rows, uniforms, log_scale_min=-7.0):
    """y_hat_rows (..., 3*nr_mix), uniforms (..., nr_mix + 1) -> samples (...) in [-1, 1] (no gradient, as in the reference)."""
    _require_cuda(y_hat_rows, uniforms)
    y, u = y_hat_rows.detach().contiguous(), uniforms.contiguous()
    nm = y.size(-1) // 3
    assert y.size(-1) == 3 * nm and u.size(-1) == nm + 1 and u.numel() // (nm + 1) == y.numel() // (3 * nm)
    out = torch.empty(y.shape[:-1], device=y.device, dtype=torch.float32)
    _lib.check(_lib.lib().viai_dmol_sample(_p(y), _p(u), out.numel(), nm, float(log_scale_min), _p(out), _stream()), "dmol_sample")
    return out


class _MaskedSumFn(torch.autograd.Function):
    """(v * mask).sum() / mask.sum() (mean=True) or (v * mask).sum(); mask None = ones."""

    @staticmethod
    def forward(ctx, v, mask, mean):
        _require_cuda(v, mask)
        L = _lib.lib()
        v = v.contiguous()
        mask = mask.contiguous() if mask is not None else None
        assert mask is None or mask.numel() == v.numel()
        acc = torch.empty(2, device=v.device, dtype=torch.float64)
        out = torch.empty((), device=v.device, dtype=torch.float32)
        _lib.check(L.viai_masked_sum_fwd(_p(v), _p(mask), v.numel(), int(mean), _p(acc), _p(out), _stream()), "masked_sum_fwd")
        ctx.save_for_backward(mask, acc)
        ctx.cfg = (tuple(v.shape), int(mean))
        return out

    @staticmethod
    def backward(ctx, gout):
        mask, acc = ctx.saved_tensors
        shape, mean = ctx.cfg
        L = _lib.lib()
        gout = gout.contiguous().float()
        dv = torch.empty(shape, device=gout.device, dtype=torch.float32)
        _lib.check(L.viai_masked_sum_bwd(_p(mask), dv.numel(), mean, _p(acc), _p(gout), _p(dv), _stream()), "masked_sum_bwd")
        return dv, None, None


def masked_sum(v, mask=None, mean=True):
    return _MaskedSumFn.apply(v, mask, bool(mean))


def sequence_mask(lengths, max_len):
    """loss_functions.py:11-21 on the device: (B, max_len) float mask, 1 where t < lengths[b]."""
    if not lengths.is_cuda:
        raise RuntimeError("VIAI ops need CUDA tensors; there is no CPU path")
    lengths = lengths.to(torch.int64).contiguous()
    B = lengths.numel()
    out = torch.empty((B, int(max_len)), device=lengths.device, dtype=torch.float32)
    _lib.check(_lib.lib().viai_sequence_mask(_p(lengths), B, int(max_len), _p(out), _stream()), "sequence_mask")
    return out


# ---- audio-visual synchronisation heads ----------------------------------------------------------------------
class _L2NormFn(torch.autograd.Function):
    """F.normalize(x, p=2, dim=1) for (rows, cols) (utils/util.py:94-96)."""

    @staticmethod
    def forward(ctx, x, eps):
        _require_cuda(x)
        L = _lib.lib()
        x = x.contiguous()
        rows, cols = x.shape
        y = torch.empty_like(x)
        nrm = torch.empty(rows, device=x.device, dtype=torch.float32)
        _lib.check(L.viai_l2norm_fwd(_p(x), rows, cols, eps, _p(y), _p(nrm), _stream()), "l2norm_fwd")
        ctx.save_for_backward(y, nrm)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        y, nrm = ctx.saved_tensors
        L = _lib.lib()
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        _lib.check(L.viai_l2norm_bwd(_p(y), _p(nrm), _p(dy), y.size(0), y.size(1), ctx.eps, _p(dx), _stream()), "l2norm_bwd")
        return dx, None


def l2_normalize(x, eps=1e-12):
    return _L2NormFn.apply(x, float(eps))


class _PairDistFn(torch.autograd.Function):
    """scores[a][b] = ||f1[a] - f2[b]||_2 (loss_functions.py:106-108)."""

    @staticmethod
    def forward(ctx, f1, f2):
        _require_cuda(f1, f2)
        L = _lib.lib()
        f1, f2 = f1.contiguous(), f2.contiguous()
        assert f1.dim() == 2 and f2.dim() == 2 and f1.size(1) == f2.size(1)
        s = torch.empty((f1.size(0), f2.size(0)), device=f1.device, dtype=torch.float32)
        _lib.check(L.viai_pairdist_fwd(_p(f1), _p(f2), f1.size(0), f2.size(0), f1.size(1), _p(s), _stream()), "pairdist_fwd")
        ctx.save_for_backward(f1, f2, s)
        return s

    @staticmethod
    def backward(ctx, ds):
        f1, f2, s = ctx.saved_tensors
        L = _lib.lib()
        ds = ds.contiguous()
        df1 = torch.empty_like(f1) if ctx.needs_input_grad[0] else None
        df2 = torch.empty_like(f2) if ctx.needs_input_grad[1] else None
        if df1 is None and df2 is None:
            return None, None
        _lib.check(L.viai_pairdist_bwd(_p(f1), _p(f2), _p(s), _p(ds), f1.size(0), f2.size(0), f1.size(1), _p(df1), _p(df2), _stream()),
                   "pairdist_bwd")
        return df1, df2


def pairdist(f1, f2):
    return _PairDistFn.apply(f1, f2)


def retrieval_ranks(dist):
    """(ranks, top1) int64 (n1,) of a (n1 captions x n2 clips) distance matrix: see viai_retrieval_ranks."""
    _require_cuda(dist)
    d = dist.detach().contiguous()
    n1, n2 = d.shape
    ranks = torch.empty(n1, device=d.device, dtype=torch.int64)
    top1 = torch.empty(n1, device=d.device, dtype=torch.int64)
    _lib.check(_lib.lib().viai_retrieval_ranks(_p(d), n1, n2, _p(ranks), _p(top1), _stream()), "retrieval_ranks")
    return ranks, top1


class _L2ContrastiveFn(torch.autograd.Function):
    """The hinge / diagonal reduction of L2ContrastiveLoss on a (B, B) score matrix (loss_functions.py:127-148)."""

    @staticmethod
    def forward(ctx, scores, margin, max_violation):
        _require_cuda(scores)
        L = _lib.lib()
        scores = scores.contiguous()
        B = scores.size(0)
        assert scores.dim() == 2 and scores.size(1) == B
        out = torch.empty((), device=scores.device, dtype=torch.float32)
        _lib.check(L.viai_l2_contrastive(_p(scores), B, margin, int(max_violation), _p(out), None, None, _stream()), "l2_contrastive")
        ctx.save_for_backward(scores)
        ctx.cfg = (margin, int(max_violation))
        return out

    @staticmethod
    def backward(ctx, gout):
        (scores,) = ctx.saved_tensors
        margin, mv = ctx.cfg
        L = _lib.lib()
        gout = gout.contiguous().float()
        ds = torch.empty_like(scores)
        _lib.check(L.viai_l2_contrastive(_p(scores), scores.size(0), margin, mv, None, _p(gout), _p(ds), _stream()), "l2_contrastive bwd")
        return ds, None, None


def l2_contrastive(scores, margin=0.0, max_violation=False):
    return _L2ContrastiveFn.apply(scores, float(margin), bool(max_violation))


# ---- loader: frame preprocessing -------------------------------------------------------------------------------
def frames_preprocess(frames_u8, out, c_off, resize_hw, flip, crop_rc, swap_rb):
    """frames_u8: (n, h, w[, c]) uint8 CUDA tensor of decoded frames; writes channels [c_off, c_off + c) of the float NHWC block
    ``out`` (n, out_h, out_w, out_c) with cv2.resize -> BGR->RGB -> fliplr -> (v - 127) / 128 -> crop (bit exact)."""
    if not (frames_u8.is_cuda and out.is_cuda) or frames_u8.dtype != torch.uint8 or out.dtype != torch.float32:
        raise RuntimeError("frames_preprocess needs a uint8 CUDA source and a float32 CUDA block; there is no CPU path")
    src = frames_u8.contiguous()
    n, h, w = src.shape[:3]
    c = src.size(3) if src.dim() == 4 else 1
    assert out.is_contiguous() and out.size(0) == n
    _lib.check(_lib.lib().viai_frames_preprocess(_p(src), n, h, w, c, int(swap_rb), int(resize_hw[0]), int(resize_hw[1]), int(flip),
                                                 int(crop_rc[0]), int(crop_rc[1]), out.size(1), out.size(2), out.size(3), int(c_off),
                                                 _p(out), _stream()), "frames_preprocess")
    return out
