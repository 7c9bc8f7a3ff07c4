#!/usr/bin/env python
"""bench.py -- GAN-step spectrogram-frames/sec of the VIAI hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one D update + one G update on a batch of synthetic mel spectrograms.
Workload at every N: BASELINE config C2 per GPU (B=32, 256x256 mel, 50% centre time-band mask, BatchNorm, LSGAN +
100*L1, Adam) -- weak scaling, frames = B * W per step per GPU.

  value : whole-job frames/s with the inputs resident in HBM (CUDA-graph replay of the step), CUDA events, max over ranks
  e2e   : the same step driven through the public API with HOST (pinned) inputs: H2D copy of mel+mask and D2H read
          of the loss inside the timed region
  roofline     : the dominant kernel (the discriminator's 256->512 3x3 convolution, 154.6 GFLOP ALGORITHMIC per launch at
                 C2; the bf16x3 forward executes 3 MMAs per MAC) timed alone with CUDA events, L2 flushed between
                 launches, against the measured bf16 tensor peak; traffic = DRAM bytes of that launch from the committed
                 ncu capture (profiles/)
  wavenet      : (N=1 only) the second metric BASELINE.json names: WaveNet-vocoder synthesis samples/s (24 layers, 4 stacks,
                 512/512/256, 16 kHz; a bounded T of the C4 workload, the loop is strictly sequential so the rate is
                 T-independent)
  cpu_baseline : the oracle (CPU restatement of the reference, oracle/viai_oracle.py) timed on the host cores on a
                 bounded sample of the same workload
`--impl reference` times that CPU implementation alone (all host threads) and prints the same JSON line.
"""
import argparse
import json
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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


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


def cpu_reference_run(steps, warmup, batch, budget_s=None):
    """Times oracle.gan_step (the CPU port of the reference path) on all host threads.  Returns (frames/s, cores, ms/step, n)."""
    import torch
    from oracle import viai_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import viai_test_helpers as H
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    esd, gsd, dsd = H.filled(H.encoder_sd("bn")), H.filled(H.decoder_sd("bn")), H.filled(H.discriminator_sd("bn"))
    mel = torch.rand(batch, 1, HMEL, WFR)
    mask = O.time_band_mask(mel.shape, WFR // 4, WFR // 2)
    opt = {"G": {}, "D": {}}
    for _ in range(warmup):
        r = O.gan_step(esd, gsd, dsd, mel, mask, HMEL, opt_state=opt)
        esd, gsd, dsd = r["enc"], r["dec"], r["dis"]
    times = []
    t_begin = time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        r = O.gan_step(esd, gsd, dsd, mel, mask, HMEL, opt_state=opt)
        esd, gsd, dsd = r["enc"], r["dec"], r["dis"]
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s and len(times) >= 2:
            break
    ms = 1e3 * sum(times) / len(times)
    return batch * WFR / (ms / 1e3), cores, ms, len(times)


def run_reference(args, rank):
    if rank != 0:
        return
    batch = 4
    v, cores, ms, n = cpu_reference_run(args.steps, min(args.warmup, 1), batch)
    sample = "B=%d of the B=32 256x256 C2 batch per step, %d timed steps, oracle port of the reference (torch CPU fp32)" % (batch, n)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: audio-only GAN train step, 256x256 mel, 50%% centre time-band mask (CPU arm: B=%d sample)" % batch,
                       "global_batch": batch, "mel_bins": HMEL, "frames": WFR},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def time_dominant_kernel(torch, iters=10):
    """D.conv3 (256->512, 3x3, s1) forward at C2: (32,64,32,256) -> (32,64,32,512), 154.6 GFLOP per launch."""
    from viai_b200 import ops
    x = torch.randn(B, 64, 32, 256, device="cuda")
    w = torch.randn(512, 256, 3, 3, device="cuda") * 0.02
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            ops.conv2d(x, w, None, (1, 1), (1, 1), False)
        tot = 0.0
        for _ in range(iters):
            flush.fill_(0.0)                      # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv2d(x, w, None, (1, 1), (1, 1), False)
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
    ms = tot / iters                               # includes the (tiny) weight re-layout launch
    flops = 2.0 * B * 64 * 32 * 512 * 256 * 9
    return flops, ms


# dram__bytes_read.sum + dram__bytes_write.sum of one D.conv3 forward launch (ncu --set full, profiles/r01_conv_tc_d_conv3_fwd_ncu.csv)
DOMINANT_KERNEL_DRAM_BYTES = 72.07e6 + 80.0e6


def time_wavenet(torch, T=8000):
    """WaveNet synthesis (BASELINE config C4 shapes, bounded T): samples/s of the persistent synthesis kernel."""
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
    return {"metric": "WaveNet samples/sec", "value": T / ms * 1e3, "unit": "samples/s", "T": T, "ms": ms,
            "config": "24 layers / 4 stacks, 512/512/256 channels, 80-bin local conditioning, B=1, scalar (DMoL) output",
            "finite": bool(torch.isfinite(out).all())}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="viai_b200", choices=["viai_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-wavenet", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the VIAI hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from viai_b200 import Options_inpainting, _lib
    from viai_b200.step import GanTrainer

    torch.manual_seed(rank)
    hp = Options_inpainting.Inpainting_Config(cin_channels=HMEL)
    tr = GanTrainer(hp, "cuda", world_size=world)
    if world > 1:                                  # identical initial weights on every rank
        for opt in (tr.optimizer_G, tr.optimizer_D):
            dist.broadcast(opt.flat_param, 0)
    g = torch.Generator().manual_seed(1000 + rank)
    mel_h = torch.rand(B, 1, HMEL, WFR, generator=g).pin_memory()
    mask_h = torch.ones_like(mel_h)
    mask_h[..., WFR // 4:WFR // 4 + WFR // 2] = 0.0            # 50 % centre time band (columns 64:192), all mel bins
    mask_h = mask_h.pin_memory()
    mel_d, mask_d = mel_h.cuda(), mask_h.cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.no_graph
    if use_graph:
        tr.capture(mel_d, mask_d, warmup=2)
        step_dev = lambda: tr.replay()

        def step_e2e(last=False):
            """One step through the public API from pinned HOST inputs: this step's inputs were put in flight by prefetch()
            (side-stream H2D, overlapping the previous step); the next step's H2D is started right after this step is launched."""
            out = tr.replay()
            if not last:
                tr.prefetch(mel_h, mask_h)
            return out
    else:
        step_dev = lambda: tr.train_step(mel_d, mask_d)
        step_e2e = lambda last=False: tr.train_step(mel_h.cuda(non_blocking=True), mask_h.cuda(non_blocking=True))

    def timed(fn, steps, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sink = 0.0
        if read_loss and use_graph:
            tr.prefetch(mel_h, mask_h)                  # H2D of the first timed step's inputs (inside the timed region)
        for i in range(steps):
            out = fn(i == steps - 1) if read_loss else fn()
            if read_loss:
                sink += float(out["loss_L1"])           # D2H read of the step's result
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for _ in range(args.warmup):
        step_dev()
    launches = tr.launches_per_step
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_dev, args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    if use_graph:
        tr.prefetch(mel_h, mask_h)
    for _ in range(2):
        step_e2e()
    step_e2e(True)                                      # drains the warm-up prefetch: every timed step copies its own inputs
    ms_e2e = timed(step_e2e, args.steps, True)

    roof = cpu = None
    if rank == 0:
        pk = peaks()
        flops, kms = time_dominant_kernel(torch)
        ach = flops / (kms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tc_burst"], "unit": "TFLOP/s", "frac": ach / pk["tc_burst"],
                "traffic": DOMINANT_KERNEL_DRAM_BYTES, "algorithmic_flops": flops, "executed_flops": 3 * flops,
                "kernel": "conv2d fwd 256->512 3x3 (D.conv3) B=32 64x32, bf16x3 (3 MMAs per MAC)",
                "peak_source": pk["src"] + " bf16 burst", "ms_per_launch": kms}
        if not args.no_cpu_baseline and world == 1:
            v, cores, cms, n = cpu_reference_run(6, 1, 4, budget_s=20.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "B=4 slice of the C2 batch, %d timed G+D steps of the oracle (torch CPU fp32), %.0f ms/step" % (n, cms)}
    wn = None
    if rank == 0 and world == 1 and not args.no_wavenet:
        try:
            wn = time_wavenet(torch)
        except Exception as e:                      # the GAN line must survive a WaveNet failure
            wn = {"error": repr(e)}
        try:
            wn["train"] = time_wavenet_train(torch)
        except Exception as e:
            wn["train"] = {"error": repr(e)}
    if rank == 0:
        frames = B * WFR * world
        line = {"metric": METRIC, "value": frames / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C2: audio-only GAN train step (1 D + 1 G update), B=32 per GPU, 256x256 mel, 50% centre "
                                       "time-band mask, BatchNorm, LSGAN+100*L1, Adam",
                           "global_batch": B * world, "mel_bins": HMEL, "frames": WFR, "parallelism": "dp%d" % world,
                           "cuda_graph": use_graph,
                           "precision": "fp32 tensors; forward convolutions as 3-term bf16-pair tensor-core products (fp32 accumulate, "
                                        "~2^-17 per product), data/weight gradients one tf32 product",
                           "l2": "no flush: the step streams >5 GB of activations per replay (>> 126 MB L2)"},
                "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": 2 * mel_h.numel() * 4, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches) * args.steps, "launches_per_step": int(launches),
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "wavenet": wn}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
