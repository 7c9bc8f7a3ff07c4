#!/usr/bin/env python
"""bench.py -- GAN-step spectrogram-frames/sec of the VIAI hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one D update + one G update on a batch of synthetic mel spectrograms.
Workload of the headline line at every N: BASELINE config C2 per GPU (B=32, 256x256 mel, 50% centre time-band mask, BatchNorm,
LSGAN + 100*L1, Adam) -- weak scaling, frames = B * W per step per GPU.

  value     : whole-job frames/s with the inputs resident in HBM (CUDA-graph replay of the step), CUDA events, max over ranks
  e2e       : the same step through the public API with HOST (pinned) inputs: H2D copy of mel+mask and D2H read of the loss
              inside the timed region
  sustained : the same replay loop kept up for >= 3 s (clocks sampled over that window)
  roofline  : the by-time DOMINANT operator of the step, found live: one eager step with every convolution / normalisation
              operator bracketed by CUDA events on the launching stream (ops.time_ops); achieved = its algorithmic FLOPs (or
              bytes) / its device time, against the measured tensor (or HBM) peak; ``traffic`` = DRAM bytes of that launch from
              the committed ncu capture (profiles/r02_traffic.json) or null; ``generator_stack`` = 3*F_G / (time in the
              generator's forward + backward kernels x sustained bf16 peak) -- the north star's quantity; ``top_ops`` = the table
  strong    : (N > 1) the strong-scaling point: global batch 32 split over the N GPUs
  c3/c4/c5  : the other BASELINE configs: vision-infused step (ResNet-18 x 2 fused at the bottleneck), WaveNet synthesis of 10 s
              of 16 kHz audio (T = 160 000; N = 1 only) and the free-form-mask 128 / 256 / 512 sweep
  cpu_baseline : the reference's OWN nn.Modules (vendored copy oracle/_ref, else the oracle port) on the host cores, N = 1 only
`--impl reference` times that CPU implementation alone (all host threads) and prints the same JSON line.
"""
import argparse
import gc
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GAN-step spectrogram-frames/sec"
UNIT = "frames/s"
B, HMEL, WFR = 32, 256, 256            # config C2
F_G, F_D = 183.91e9, 233.07e9          # forward conv FLOPs (2*MAC) of generator / discriminator at C2
STEP_FLOPS = 3 * F_G + 8 * F_D


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU arm ------------------------------------------------------------------------------------------------------------------
def cpu_gan(steps, warmup, batch, budget_s):
    """The reference's GAN step on all host threads: (frames/s, cores, ms/step, timed steps, warm-up steps, kind)."""
    from oracle import build_ref
    if build_ref.available() or os.path.isdir("/root/reference/networks"):
        try:
            from oracle import ref_step
            v, cores, ms, n, w = ref_step.time_gan_steps(batch, HMEL, WFR, steps, warmup, budget_s)
            return v, cores, ms, n, w, "reference"
        except Exception as e:                     # a broken vendored copy must not cost the baseline: fall back to the port
            sys.stderr.write("reference modules unusable (%r); timing the oracle port instead\n" % (e,))
    import torch
    from oracle import viai_oracle as O            # fallback: the oracle port
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import viai_test_helpers as H
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    esd, gsd, dsd = H.filled(H.encoder_sd("bn")), H.filled(H.decoder_sd("bn")), H.filled(H.discriminator_sd("bn"))
    mel = torch.rand(batch, 1, HMEL, WFR)
    mask = O.time_band_mask(mel.shape, WFR // 4, WFR // 2)
    opt = {"G": {}, "D": {}}
    t_begin, times, w = time.perf_counter(), [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = O.gan_step(esd, gsd, dsd, mel, mask, HMEL, opt_state=opt)
        esd, gsd, dsd = r["enc"], r["dec"], r["dis"]
        if i < warmup:
            w += 1
        else:
            times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 2:
            break
    ms = 1e3 * sum(times) / len(times)
    return batch * WFR / (ms / 1e3), cores, ms, len(times), w, "port"


def run_reference(args, rank):
    if rank != 0:
        return
    v, cores, ms, n, w, kind = cpu_gan(args.steps, args.warmup, B, budget_s=150.0)
    sample = ("full C2 batch (B=%d, 256x256) per step, %d timed + %d warm-up steps (of --steps %d --warmup %d, bounded to ~150 s), %s, "
              "torch CPU fp32, %d threads" % (B, n, w, args.steps, args.warmup,
                                              "the reference's own nn.Modules (oracle/_ref) + step glue of SURVEY 3.1" if kind == "reference"
                                              else "oracle port of the reference", cores))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": w,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: audio-only GAN train step (1 D + 1 G update), B=32, 256x256 mel, 50% centre time-band mask, "
                                   "BatchNorm, LSGAN+100*L1, Adam (CPU arm: one replica on the host cores)",
                       "global_batch": B, "mel_bins": HMEL, "frames": WFR},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- GPU arm helpers ------------------------------------------------------------------------------------------------------------
class Ctx(object):
    def __init__(self, torch, dist, rank, world):
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """ms per call of ``fn`` over ``steps`` calls: barrier + synchronize on both sides, CUDA events, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1) / steps
        if self.world > 1:
            t = torch.tensor([ms], device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def free(self):
        gc.collect()
        self.torch.cuda.empty_cache()


def band_mask(torch, shape):
    m = torch.ones(shape)
    W = shape[-1]
    m[..., W // 4:W // 4 + W // 2] = 0.0            # 50 % centre time band, all mel bins
    return m


def make_trainer(ctx, mel_bins, decoder="MelDecoder", video_encoder=None):
    from viai_b200 import Options_inpainting
    from viai_b200.step import GanTrainer
    hp = Options_inpainting.Inpainting_Config(cin_channels=mel_bins)
    return GanTrainer(hp, "cuda", decoder=decoder, video_encoder=video_encoder, world_size=ctx.world)     # broadcasts rank 0's weights


def dominant_op_and_generator_stack(ctx, tr, mel_d, mask_d, pk):
    """One eager step with per-operator CUDA events: the by-time table, the dominant operator's roofline and the generator-stack
    fraction.  Rank 0, N = 1."""
    torch = ctx.torch
    from viai_b200 import ops
    for _ in range(2):
        tr.train_step(mel_d, mask_d)
    torch.cuda.synchronize()
    tr.segment_events = {}
    with ops.time_ops() as log:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.train_step(mel_d, mask_d)
        e1.record()
    seg = tr.segment_ms()
    tr.segment_events = None
    step_ms = e0.elapsed_time(e1)
    table = ops.summarize_ops(log)
    rows = sorted(table.items(), key=lambda kv: -kv[1]["ms"])
    ridge = pk["tc_sustained"] * 1e12 / (pk["hbm"] * 1e9)
    top = []
    for (fam, key), d in rows[:8]:
        top.append({"op": fam, "geometry": key, "launches": d["launches"], "ms": round(d["ms"], 4), "share_of_step": round(d["ms"] / step_ms, 4),
                    "tflops": round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1) if d["flops"] else None,
                    "gbs": round(d["bytes"] / (d["ms"] * 1e-3) / 1e9, 1)})
    (fam, key), d = rows[0]
    per = d["ms"] / d["launches"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("%s|%s" % (fam, key))
    if d["flops"] and d["flops"] / d["bytes"] >= ridge:
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tc_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tc_sustained"],
                "peak_source": pk["src"] + " bf16 sustained (kernel timed inside the step)"}
    else:
        ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "peak_source": pk["src"] + " copy bandwidth"}
    roof.update({"traffic": traffic, "kernel": "%s %s" % (fam, key), "launches_per_step": d["launches"], "ms_per_launch": per,
                 "share_of_step": d["ms"] / step_ms, "algorithmic_flops_per_launch": d["flops"] / d["launches"],
                 "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
                 "how": "dominant (operator, geometry) by device time in one eager C2 step, every operator bracketed by CUDA events "
                        "on the launching stream (in-situ cache state; each launch streams more than the 126 MB L2)"})
    g_ms = seg.get("g_fwd", 0.0) + seg.get("g_bwd", 0.0)
    conv_ms = sum(d["ms"] for (fam, _), d in table.items() if fam.startswith("conv"))
    gen = {"g_fwd_ms": seg.get("g_fwd"), "g_bwd_ms": seg.get("g_bwd"), "flops": 3 * F_G,
           "tflops": 3 * F_G / (g_ms * 1e-3) / 1e12 if g_ms else None,
           "frac_of_sustained_bf16_peak": 3 * F_G / (g_ms * 1e-3) / 1e12 / pk["tc_sustained"] if g_ms else None,
           "definition": "3*F_G / (device time of the generator's forward + backward kernels, conv + norm + resample) / sustained bf16 peak"}
    whole = {"flops": STEP_FLOPS, "eager_step_ms": step_ms, "conv_ops_ms": conv_ms,
             "frac_of_sustained_bf16_peak_eager": STEP_FLOPS / (step_ms * 1e-3) / 1e12 / pk["tc_sustained"]}
    return roof, gen, whole, top


def run_c3(ctx, pk):
    """Vision-infused step: ResNet-18 x 2 ImageEmbedding fused at the bottleneck through MelDecoderImage; 256x256 mel, T = 128
    video frames of 224x224 per sample.  Per-GPU batch: the largest of 16 / 8 / 4 that fits (B = 32 needs ~170 GB of saved
    ResNet activations for its 4096 frame pairs)."""
    torch = ctx.torch
    from viai_b200 import Options_inpainting
    from viai_b200.networks.Image_Embedding import ImageEmbedding
    T = (HMEL // 32 // 3) * WFR // 4                                   # H5 * W / 4 = 128
    err = None
    for bs in (16, 8, 4):
        tr = ve = None
        try:
            hp = Options_inpainting.Inpainting_Config(cin_channels=HMEL)
            torch.manual_seed(100 + ctx.rank)
            ve = ImageEmbedding(hp).cuda()
            tr = make_trainer(ctx, HMEL, "MelDecoderImage", ve)
            g = torch.Generator(device="cuda").manual_seed(7 + ctx.rank)
            mel = torch.rand(bs, 1, HMEL, WFR, device="cuda", generator=g)
            mask = band_mask(torch, mel.shape).cuda()
            video = torch.randn(bs, T, 3, 224, 224, device="cuda", generator=g).clamp_(-1, 1)
            flow = torch.randn(bs, T, 2, 224, 224, device="cuda", generator=g).clamp_(-1, 1)
            step = lambda: tr.train_step(mel, mask, video, flow)
            for _ in range(2):
                step()
            ms = ctx.timed(step, 3)
            tr.segment_events = {}                 # one more step on EVERY rank (it contains collectives), with segment events
            step()
            seg = tr.segment_ms()
            tr.segment_events = None
            peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
            v_ms = (seg.get("v_fwd") or 0.0) + (seg.get("v_bwd") or 0.0)
            v_flops = 3 * 2 * 3.63e9 * bs * T                           # fwd + dgrad + wgrad, two streams, 3.63 GFLOP per 224^2 frame
            out = {"workload": "C3: vision-infused GAN step, B=%d per GPU (largest of 16/8/4 that fits), 256x256 mel, T=%d frames of "
                               "224x224 RGB + flow per sample, MelDecoderImage + ImageEmbedding (ResNet-18 x 2), eager step" % (bs, T),
                   "batch_per_gpu": bs, "n_gpus": ctx.world, "ms_per_step": ms, "value": bs * WFR * ctx.world / (ms * 1e-3), "unit": UNIT,
                   "video_frames_per_s": bs * T * ctx.world / (ms * 1e-3), "peak_mem_gib": round(peak_mem, 1),
                   "visual_encoder_ms": v_ms or None, "generator_discriminator_ms": (ms - v_ms) if v_ms else None,
                   "visual_encoder_tflops": v_flops / (v_ms * 1e-3) / 1e12 if v_ms else None,
                   "visual_encoder_frac_of_sustained_bf16_peak": v_flops / (v_ms * 1e-3) / 1e12 / pk["tc_sustained"] if v_ms else None,
                   "launches_per_step": int(tr.launches_per_step)}
            del tr, ve, mel, mask, video, flow
            ctx.free()
            return out
        except torch.cuda.OutOfMemoryError as e:
            err = "out of memory at B=%d" % bs
            del tr, ve
            ctx.free()
            if ctx.world > 1:                      # every rank must take the same branch: report instead of retrying out of step
                break
    return {"error": err}


def run_c4(ctx, pk, T=160000):
    """WaveNet synthesis of 10 s of 16 kHz audio (T = 160 000 sequential samples), N = 1; the reference's incremental_forward on
    the host cores beside it (bounded sample, extrapolation stated)."""
    torch = ctx.torch
    from viai_b200.wavenet_vocoder import WaveNet
    torch.manual_seed(0)
    m = WaveNet().cuda().eval()
    m.make_generation_fast_()
    c = torch.rand(1, 80, T // 160).cuda()
    with torch.no_grad():
        m.incremental_forward(c=c[:, :, :5], T=800)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = m.incremental_forward(c=c, T=T)
        e1.record()
        e1.synchronize()
    ms = e0.elapsed_time(e1)
    rate = T / ms * 1e3
    kern = getattr(m, "last_synthesis_kernel", "grid")
    wbytes = float(m._packed["layers"].numel() + m._packed["head1"].numel() + m._packed["head2"].numel()) * 4
    # four utterances at once (the reference's incremental_forward takes a batch; the dependent chain is shared)
    T4 = 16000
    c4 = torch.rand(4, 80, T4 // 160).cuda()
    with torch.no_grad():
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        out4 = m.incremental_forward(c=c4, T=T4)
        e3.record()
        e3.synchronize()
    ms4 = e2.elapsed_time(e3)
    res = {"metric": "WaveNet samples/sec", "value": rate, "unit": "samples/s", "T": T, "seconds_of_audio": T / 16000.0, "ms": ms,
           "real_time_factor": (T / 16000.0) / (ms * 1e-3), "finite": bool(torch.isfinite(out).all()),
           "in_range": bool((out.abs() <= 1).all()), "kernel": kern,
           "config": "C4: 24 layers / 4 stacks, 512/512/256 channels, 80-bin local conditioning, B=1, scalar (DMoL) output, 16 kHz",
           "batch4": {"value": 4 * T4 / ms4 * 1e3, "unit": "samples/s", "B": 4, "T": T4, "ms": ms4,
                      "finite": bool(torch.isfinite(out4).all())},
           "roofline": {"bound": "hbm", "achieved": wbytes * rate / 1e9, "peak": pk["hbm"], "unit": "GB/s", "frac": wbytes * rate / 1e9 / pk["hbm"],
                        "traffic": 119.3e6 if kern == "ws" else None,
                        "note": "algorithmic bytes = the %.1f MB of packed fp32 parameters every sample touches; the per-sample sweep is cyclic "
                                "and larger than what the 126 MB L2 keeps, so they stream from HBM every sample (ncu: 119 MB of DRAM reads per "
                                "sample, profiles/r02_wavenet_folded_ncu_summary.txt); the loop is bound by the %d dependent cross-CTA "
                                "exchanges per sample (latency), not by HBM" % (wbytes / 1e6, 26 if kern in ("ws", "folded") else 50)}}
    del m
    ctx.free()
    try:
        from oracle import build_ref, ref_step
        if build_ref.available() or os.path.isdir("/root/reference/networks"):
            fast, cores, dt = ref_step.time_wavenet_synthesis(1600, True)
            slow, _, dt2 = ref_step.time_wavenet_synthesis(800, False)
            res["cpu_baseline"] = {"value": fast, "unit": "samples/s", "cores": cores, "kind": "reference",
                                   "sample": "the reference's WaveNet.incremental_forward (oracle/_ref): 1600 samples after make_generation_fast_() "
                                             "in %.1f s (%.0f samples/s; without it 800 samples at %.0f samples/s); 160 000 samples would take "
                                             "%.0f s -- extrapolated, not run" % (dt, fast, slow, T / fast)}
    except Exception as e:
        res["cpu_baseline"] = {"error": repr(e)}
    return res


def time_wavenet_train(torch, B=4, T=8000, steps=5, warmup=3, graph=True):
    """WaveNet teacher-forced training step (full C4 network) through ``WaveNetTrainer``: forward over B*T samples
    in parallel, masked DMoL loss on the shifted targets, backward, fused Adam + EMA, captured as one CUDA graph.  The timed
    region includes the host->device copy of each step's audio / conditioning from pinned memory and the read-back of the
    loss.  samples/s = B*T / step time; algorithmic FLOPs 3 x 49.3 MFLOP/sample."""
    from viai_b200.wavenet_step import WaveNetTrainer
    from viai_b200.wavenet_vocoder import WaveNet
    torch.manual_seed(0)
    m = WaveNet().cuda().train()
    tr = WaveNetTrainer(m)
    x_h = (torch.rand(B, 1, T) * 2 - 1).pin_memory()
    c_h = torch.rand(B, 80, T // 160).pin_memory()
    mask = torch.ones(B, T, 1).cuda()
    x, c = x_h.cuda(), c_h.cuda()
    y = x.transpose(1, 2).contiguous()
    if graph:
        tr.capture(x, y, c, mask, warmup=2)
        step = lambda: tr.replay(x_h, x_h, c_h)          # y is the same signal as x for raw audio: (B,1,T) and (B,T,1) share memory
    else:
        def step():
            xd = x_h.cuda(non_blocking=True)
            return tr.train_step(xd, xd.transpose(1, 2), c_h.cuda(non_blocking=True), mask)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = float(step())
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"metric": "WaveNet teacher-forced training samples/sec", "value": B * T / ms * 1e3, "unit": "samples/s", "B": B, "T": T,
            "ms_per_step": ms, "algorithmic_tflops": 3 * 49.30e6 * B * T / (ms * 1e-3) / 1e12, "loss": loss,
            "launches_per_step": int(tr.launches_per_step), "cuda_graph": bool(graph),
            "config": "24 layers / 4 stacks, 512/512/256 channels, 80-bin local conditioning, dropout 0.05, masked DMoL loss, "
                      "Adam + EMA; inputs from pinned host memory and loss read back every step"}


def run_stft(ctx, pk, seconds=3600):
    """STFT -> mel front end (utils/audio.py:70-75): one launch over an hour of 16 kHz audio (T = 57.6 M samples, 230 MB: larger than
    L2), frames/s and algorithmic HBM bytes (hop * 4 bytes of new samples read + n_mels * 4 bytes written per frame)."""
    torch = ctx.torch
    from viai_b200.utils import audio
    T = 16000 * seconds
    y = torch.rand(T, device="cuda") * 2 - 1
    for _ in range(2):
        mel = audio.melspectrogram_cuda(y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        mel = audio.melspectrogram_cuda(y)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 5
    M = mel.size(1)
    nbytes = 4.0 * (T + mel.numel())
    return {"metric": "STFT->mel frames/sec", "value": M / (ms * 1e-3), "unit": "frames/s", "frames": M, "ms": ms,
            "config": "fft 1024 / hop 160 / 80 mels, one waveform of %d s at 16 kHz, fused frame + window + rFFT + |.| + mel + dB + normalise" % seconds,
            "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": nbytes / (ms * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                         "algorithmic_bytes_per_frame": nbytes / M, "flops_per_frame": 35e3,
                         "note": "35 kFLOP per ~1 kB frame: the kernel is fp32-issue / shared-memory bound (one warp per frame, three "
                                 "radix-8 passes), not HBM bound"}}


def run_c5(ctx, steps=8):
    """Free-form (seeded random-walk stroke) masks at 128 / 256 / 512 square mels, B=32 per GPU, one captured step per size."""
    torch = ctx.torch
    from viai_b200.utils.masks import freeform_mask
    sizes, per = (128, 256, 512), {}
    tot_frames = tot_ms = 0.0
    for size in sizes:
        torch.manual_seed(200 + ctx.rank)
        tr = make_trainer(ctx, size)
        mel = torch.rand(B, 1, size, size).cuda()
        mask = freeform_mask(mel.shape, seed=size + 1000 * ctx.rank).cuda()
        tr.capture(mel, mask, warmup=2)
        for _ in range(2):
            tr.replay()
        ms = ctx.timed(tr.replay, steps)
        per[str(size)] = {"ms_per_step": ms, "value": B * size * ctx.world / (ms * 1e-3), "masked_fraction": round(float((mask == 0).float().mean()), 4)}
        tot_frames += B * size * ctx.world
        tot_ms += ms
        del tr, mel, mask
        ctx.free()
    return {"workload": "C5: free-form irregular masks (viai_b200.utils.masks.freeform_mask), square mels 128/256/512, B=32 per GPU, "
                        "one size per step cycling 128 -> 256 -> 512 (CUDA-graph replay per size)",
            "n_gpus": ctx.world, "unit": UNIT, "per_size": per, "value": tot_frames / (tot_ms * 1e-3),
            "blended": "frames of one 128+256+512 cycle / time of the cycle"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="viai_b200", choices=["viai_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-wavenet", action="store_true", help="skip C4 (WaveNet synthesis, ~25 s)")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C5 / strong-scaling / sustained / per-operator sub-records")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the VIAI hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # (a short collective timeout: a rank that falls out of step must abort the run in minutes, not hang the box)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=240))
    ctx = Ctx(torch, dist, rank, world)
    from viai_b200 import ops
    pk = peaks()

    torch.manual_seed(rank)
    tr = make_trainer(ctx, HMEL)
    g = torch.Generator().manual_seed(1000 + rank)
    mel_h = torch.rand(B, 1, HMEL, WFR, generator=g).pin_memory()
    mask_h = band_mask(torch, mel_h.shape).pin_memory()
    mel_d, mask_d = mel_h.cuda(), mask_h.cuda()

    use_graph = not args.no_graph
    if use_graph:
        tr.capture(mel_d, mask_d, warmup=2)
        step_dev = tr.replay
    else:
        step_dev = lambda: tr.train_step(mel_d, mask_d)
    for _ in range(args.warmup):
        step_dev()
    launches = tr.launches_per_step
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = ctx.timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: pinned host inputs, H2D inside, loss read back every step
    sink = [0.0]
    if use_graph:
        def step_e2e():
            out = tr.replay()                           # consumes the inputs put in flight by the previous prefetch()
            tr.prefetch(mel_h, mask_h)                  # next step's H2D on a side stream, overlapping this step
            sink[0] += float(out["loss_L1"])            # D2H read of the step's result
        tr.prefetch(mel_h, mask_h)
    else:
        def step_e2e():
            out = tr.train_step(mel_h.cuda(non_blocking=True), mask_h.cuda(non_blocking=True))
            sink[0] += float(out["loss_L1"])
    for _ in range(3):
        step_e2e()
    ms_e2e = ctx.timed(step_e2e, args.steps)
    overflow = ops.f16_overflow()

    extra = not args.no_extra
    sustained = strong = None
    if extra:
        n_sus = max(args.steps, int(math.ceil(3200.0 / ms_dev)))
        s2 = ClockSampler(local_rank)
        if rank == 0:
            s2.start()
        ms_sus = ctx.timed(step_dev, n_sus)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"steps": n_sus, "seconds": n_sus * ms_sus * 1e-3, "ms_per_step": ms_sus, "value": B * WFR * world / (ms_sus * 1e-3),
                     "unit": UNIT, "clocks": c2}

    roof = gen = whole = top = cpu = None
    if rank == 0 and world == 1 and extra:
        roof, gen, whole, top = dominant_op_and_generator_stack(ctx, tr, mel_d, mask_d, pk)
        whole["graph_step_ms"] = ms_dev
        whole["frac_of_sustained_bf16_peak"] = STEP_FLOPS / (ms_dev * 1e-3) / 1e12 / pk["tc_sustained"]
    del tr
    ctx.free()

    if extra and world > 1 and B % world == 0:
        bs = B // world
        torch.manual_seed(rank)
        tr2 = make_trainer(ctx, HMEL)
        tr2.capture(mel_d[:bs].contiguous(), mask_d[:bs].contiguous(), warmup=2)
        for _ in range(3):
            tr2.replay()
        ms_s = ctx.timed(tr2.replay, args.steps)
        strong = {"scaling": "strong", "global_batch": B, "batch_per_gpu": bs, "ms_per_step": ms_s, "value": B * WFR / (ms_s * 1e-3), "unit": UNIT}
        del tr2
        ctx.free()

    c3 = c4 = c5 = None
    if extra:
        try:
            c5 = run_c5(ctx)
        except Exception as e:                       # the headline line must survive a sub-record failure
            c5 = {"error": repr(e)}
        try:
            c3 = run_c3(ctx, pk)
        except Exception as e:
            c3 = {"error": repr(e)}
    stft = wn_train = None
    if rank == 0 and world == 1 and extra:
        try:
            wn_train = time_wavenet_train(torch)
        except Exception as e:
            wn_train = {"error": repr(e)}
        ctx.free()
    if rank == 0 and world == 1 and extra:
        try:
            stft = run_stft(ctx, pk)
        except Exception as e:
            stft = {"error": repr(e)}
        ctx.free()
    if rank == 0 and world == 1 and not args.no_wavenet:
        try:
            c4 = run_c4(ctx, pk)
        except Exception as e:
            c4 = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, cores, cms, n, w, kind = cpu_gan(3, 1, B, budget_s=45.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "full C2 batch (B=32, 256x256) per step, %d timed steps after %d warm-up, %.0f ms/step, %s, torch CPU fp32" % (
                       n, w, cms, "the reference's own nn.Modules (oracle/_ref) + the step glue of SURVEY 3.1" if kind == "reference"
                       else "oracle port of the reference")}
        except Exception as e:                       # the GPU line must survive a failure of the CPU arm
            cpu = {"error": repr(e)}
    if rank == 0:
        frames = B * WFR * world
        line = {"metric": METRIC, "value": frames / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C2: audio-only GAN train step (1 D + 1 G update), B=32 per GPU, 256x256 mel, 50% centre "
                                       "time-band mask, BatchNorm, LSGAN+100*L1, Adam",
                           "global_batch": B * world, "mel_bins": HMEL, "frames": WFR, "parallelism": "dp%d" % world,
                           "cuda_graph": use_graph,
                           "precision": "fp32 tensors in HBM; tcgen05 convolutions: forward = 3-term fp16-pair products of power-of-two "
                                        "scaled operands (fp32 accumulate, ~2^-21 per product), data gradient = 3-term bf16-pair "
                                        "products (~2^-17), weight gradient = one tf32 product on operands rounded in shared memory "
                                        "(precision '%s')" % ops.get_precision(),
                           "l2": "no flush: the step streams >5 GB of activations per replay (>> 126 MB L2)"},
                "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": 2 * mel_h.numel() * 4, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches) * args.steps, "launches_per_step": int(launches), "f16_saturations": overflow,
                "clocks": clocks, "sustained": sustained, "roofline": roof, "generator_stack": gen, "whole_step": whole, "top_ops": top,
                "strong": strong, "cpu_baseline": cpu, "c3": c3, "c4": c4, "c5": c5, "stft": stft, "wavenet": c4, "wavenet_train": wn_train}
        print(json.dumps(line), flush=True)
    if world > 1:
        # No barrier / destroy_process_group here: tearing down the NCCL communicator after its all-reduces were captured in CUDA
        # graphs was observed to hang at exit (2 x B200, round 2).  The last collective of every rank (the MAX over ranks inside
        # Ctx.timed) has completed on the host by now, so leaving directly is safe; torchrun sees exit code 0 from every rank.
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
