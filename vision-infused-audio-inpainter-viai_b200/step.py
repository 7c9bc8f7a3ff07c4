"""The GAN training step of VIAI (one D update + one G update) on the CUDA path.

The reference's ``Models/Whole_Sync_inpainting_modify.py`` (class ``AudioModel``) is missing from its tree; the step
follows the call contract of /root/reference/train_whole_sync.py:49-112 and the pix2pix update order restated in
SURVEY.md 3.1.  ``GanTrainer`` owns the networks, the two optimizers and (optionally) a CUDA graph of the step."""
import torch
import torch.nn as nn

from . import _lib, ops
from .loss_functions import GANLoss, L1Loss
from .networks.Discriminator_Networks import MelDiscriminator
from .networks.Inpainting_Networks import MelEncoder
from .networks import New_Inpainting_Networks as NIN
from .optim import FusedAdam


def set_requires_grad(net, flag):
    for p in net.parameters():
        p.requires_grad_(flag)


class GanTrainer(object):
    def __init__(self, hparams, device="cuda", norm_layer_d=nn.BatchNorm2d, norm_layer_e=nn.BatchNorm2d,
                 decoder="MelDecoder", video_encoder=None, world_size=1, process_group=None):
        self.hparams = hparams
        self.device = torch.device(device)
        self.world_size = world_size
        self.Mel_Encoder = MelEncoder(hparams, norm_layer=norm_layer_e).to(self.device)
        self.Mel_Decoder = getattr(NIN, decoder)(hparams, norm_layer=hparams.normlayer).to(self.device)
        self.netD = MelDiscriminator(norm_layer=norm_layer_d).to(self.device)
        self.VideoEncoder = video_encoder
        self.uses_video = decoder in ("MelDecoderImage", "MelDecoderImage2")
        self.criterionGAN = GANLoss(use_lsgan=getattr(hparams, "use_lsgan", True), device=self.device).to(self.device)
        self.criterionL1 = L1Loss()
        lr = getattr(hparams, "lr", 2e-4)
        b1 = getattr(hparams, "beta1", 0.5)
        g_params = list(self.Mel_Encoder.parameters()) + list(self.Mel_Decoder.parameters())
        if self.VideoEncoder is not None:
            g_params += list(self.VideoEncoder.parameters())
        self.optimizer_G = FusedAdam(g_params, lr=lr, betas=(b1, 0.999), world_size=world_size, process_group=process_group)
        self.optimizer_D = FusedAdam(self.netD.parameters(), lr=lr, betas=(b1, 0.999), world_size=world_size,
                                     process_group=process_group)
        self.lambda_L1 = float(getattr(hparams, "lambda_L1", 100.0))
        self.process_group = process_group
        self.sync_replicas()
        import os
        self.overlap = world_size > 1 and os.environ.get("VIAI_DDP_OVERLAP", "1") != "0"
        self._triggers = []
        if self.overlap:
            self._plan_overlap()
        self._graphs = None
        self._static = None
        self.launches_per_step = None
        # Measurement aid (bench.py): set to {} before an EAGER step and the step fills it with CUDA events bracketing the
        # generator's kernels -- "g_fwd" (encoder [+ visual encoder] + decoder forward) and "g_bwd" (from the moment d loss / d fake
        # is complete to the end of the generator's backward), "v_fwd" / "v_bwd" for the visual encoder alone.
        self.segment_events = None
        self._pack_plan = None                   # set by capture(): batched weight packing per step segment

    def segment_ms(self):
        """{segment: milliseconds} of the last eager step run with ``segment_events = {}`` (synchronises)."""
        torch.cuda.synchronize()
        ev = self.segment_events or {}
        return {k[:-2]: ev[k].elapsed_time(ev[k[:-2] + "_1"]) for k in ev if k.endswith("_0") and (k[:-2] + "_1") in ev}

    def _mark(self, name):
        if self.segment_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.segment_events[name] = e

    # ---- data-parallel overlap -----------------------------------------------------------------------------------------
    def _plan_overlap(self):
        """Bucket ranges whose gradients are complete before the backward pass ends, with the parameter whose weight gradient is the
        LAST one written into the range (backward runs in reverse forward order, the bucket is in registration order):
          D: [conv3.weight, end)  after conv3's weight gradient of the second discriminator pass of the D phase (76 % of D's bytes;
             the three large-image layers conv2_2 / conv2_1 / conv1 are still to come);
          G: [convblock2, decoder end) after convblock2's first layer; [decoder start, convblock2) after the decoder's first layer
             (the encoder's backward is still to come).  What is left (D: 24 %, G: the encoder [+ the visual encoder]) is reduced at
             the end of the backward pass; ranges that are zero on every rank are part of a range anyway or skipped."""
        bD, bG = self.optimizer_D.bucket, self.optimizer_G.bucket
        D, dec = self.netD, self.Mel_Decoder
        dec_params = list(dec.parameters())
        dec_start = bG.offset_of(dec_params[0])
        dec_end = bG.offset_of(dec_params[-1]) + dec_params[-1].numel()
        blk2 = bG.offset_of(next(dec.convblock2.parameters()))
        first = dec.deconv1_1_1 if self.uses_video else dec.deconv1_1
        self._add_trigger(D.conv3.weight, 2, bD, bD.offset_of(D.conv3.weight), bD.numel)
        self._add_trigger(next(dec.convblock2.parameters()), 1, bG, blk2, dec_end)
        self._add_trigger(first.weight, 1, bG, dec_start, blk2)

    def _add_trigger(self, param, count, bucket, start, end):
        st = {"n": 0}

        def fire():
            st["n"] += 1
            if st["n"] == count:
                bucket.reduce_range_async(start, end)
        param._viai_after_grad = fire
        self._triggers.append(st)

    def _reduce_begin(self, opt):
        for st in self._triggers:
            st["n"] = 0
        opt.bucket.begin_overlap()

    def _reduce_finish(self, opt):
        if self.overlap:
            opt.bucket.finish_overlap()
        else:
            opt.all_reduce_grads()

    def sync_replicas(self, src=0):
        """Identical weights AND buffers (BatchNorm running statistics) on every rank, from rank ``src`` -- what
        nn.DataParallel's replicate does every step in the reference (utils/model_util.py:137).  No-op for one rank."""
        if self.world_size <= 1:
            return
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("GanTrainer(world_size=%d) needs an initialised torch.distributed process group" % self.world_size)
        for opt in (self.optimizer_G, self.optimizer_D):
            opt.bucket.broadcast_params(src)
        mods = [self.Mel_Encoder, self.Mel_Decoder, self.netD] + ([self.VideoEncoder] if self.VideoEncoder is not None else [])
        for m in mods:
            for b in m.buffers():
                dist.broadcast(b, src, group=self.process_group)

    # ---- the three segments of a step; NCCL all-reduces sit between them -------------------------------------
    def _seg_forward_and_d_backward(self, mel, mask, video=None, flow=None):
        B = mel.size(0)
        H = self.hparams.cin_channels
        real = mel.reshape(B, 1, H, -1)
        self.real = real
        if self._pack_plan:
            ops.prepack(self._pack_plan[0])          # every weight re-layout up to the discriminator's update: one launch
        self._mark("g_fwd_0")
        masked = ops.mul(real.reshape(B, H, -1, 1), mask.reshape(B, H, -1, 1)).reshape(real.shape)
        feats = self.Mel_Encoder(masked)
        if self.uses_video:
            self._mark("v_fwd_0")
            vnet = self.VideoEncoder(video, flow)
            self._mark("v_fwd_1")
            if self.segment_events is not None and vnet.requires_grad:
                vnet.register_hook(lambda g: self._mark("v_bwd_0"))
            self.fake = self.Mel_Decoder(feats, real.shape, vnet)
        else:
            self.fake = self.Mel_Decoder(feats, real.shape)
        self._mark("g_fwd_1")
        set_requires_grad(self.netD, True)
        self.optimizer_D.zero_grad()
        pred_fake = self.netD(self.fake.detach())
        pred_real = self.netD(real)
        self.loss_D_fake = self.criterionGAN(pred_fake, False)
        self.loss_D_real = self.criterionGAN(pred_real, True)
        self.loss_D = ops.lincomb2(self.loss_D_fake, 0.5, self.loss_D_real, 0.5)
        self.loss_D.backward()

    def _seg_d_update_and_g_backward(self):
        self.optimizer_D.step()
        if self._pack_plan and len(self._pack_plan) > 1:
            ops.prepack(self._pack_plan[1])          # ... and from there (the updated discriminator) to the generator's update
        set_requires_grad(self.netD, False)
        self.optimizer_G.zero_grad()
        pred_fake = self.netD(self.fake)
        self.loss_G_GAN = self.criterionGAN(pred_fake, True)
        self.loss_L1 = self.criterionL1(self.fake, self.real)
        self.loss_G = ops.lincomb2(self.loss_G_GAN, 1.0, self.loss_L1, self.lambda_L1)
        if self.segment_events is not None:
            self.fake.register_hook(lambda g: self._mark("g_bwd_0"))     # fires once d loss / d fake is complete (D's part is done)
        self.loss_G.backward()
        self._mark("g_bwd_1")
        if self.segment_events is not None and "v_bwd_0" in self.segment_events:
            self.segment_events["v_bwd_1"] = self.segment_events["g_bwd_1"]
        set_requires_grad(self.netD, True)

    def _seg_g_update(self):
        self.optimizer_G.step()

    def _drop_step_state(self):
        for k in ("real", "fake", "loss_D_fake", "loss_D_real", "loss_D", "loss_G_GAN", "loss_L1", "loss_G"):
            setattr(self, k, None)
        import gc
        gc.collect()

    def _outputs(self):
        return dict(fake=self.fake.detach(), loss_D=self.loss_D.detach(), loss_G_GAN=self.loss_G_GAN.detach(),
                    loss_L1=self.loss_L1.detach(), loss_G=self.loss_G.detach())

    def _step_body(self, mel, mask, video=None, flow=None):
        """The whole step with its collectives (world_size > 1): the bucket ranges of _plan_overlap are all-reduced while the rest
        of the backward pass runs, the remainder right after it."""
        if self.world_size > 1:
            self._reduce_begin(self.optimizer_D)
        self._seg_forward_and_d_backward(mel, mask, video, flow)
        if self.world_size > 1:
            self._reduce_finish(self.optimizer_D)
            self._reduce_begin(self.optimizer_G)
        self._seg_d_update_and_g_backward()
        if self.world_size > 1:
            self._reduce_finish(self.optimizer_G)
        self._seg_g_update()

    # ---- eager step -------------------------------------------------------------------------------------------
    def train_step(self, mel, mask, video=None, flow=None):
        """mel (B,1,H,W) or (B,H,W) fp32 in [0,1]; mask same shape, {0,1}.  Returns dict of device tensors."""
        n0 = _lib.launch_count()
        ops.pack_cache_begin()
        try:
            self._step_body(mel, mask, video, flow)
        finally:
            ops.pack_cache_end()
        self._pack_plan = None                   # the plan is baked into the graph; eager steps pack on demand
        self.launches_per_step = _lib.launch_count() - n0
        return self._outputs()

    # ---- CUDA-graph step --------------------------------------------------------------------------------------
    def _modules(self):
        return [self.Mel_Encoder, self.Mel_Decoder, self.netD] + ([self.VideoEncoder] if self.VideoEncoder is not None else [])

    def _snapshot(self):
        opts = [(o.flat_param.clone(), o.flat_m.clone(), o.flat_v.clone(), o.step_dev.clone()) for o in (self.optimizer_G, self.optimizer_D)]
        return opts, [b.clone() for m in self._modules() for b in m.buffers()]

    def _restore(self, snap):
        opts, bufs = snap
        for o, (p, m, v, s) in zip((self.optimizer_G, self.optimizer_D), opts):
            o.flat_param.copy_(p); o.flat_m.copy_(m); o.flat_v.copy_(v); o.step_dev.copy_(s)
        for b, saved in zip([b for m in self._modules() for b in m.buffers()], bufs):
            b.copy_(saved)
        ops.weights_updated()

    def capture(self, mel, mask, video=None, flow=None, warmup=2, preserve_state=False):
        """Captures the step into CUDA graphs (one graph per segment so that the NCCL all-reduces stay eager when
        world_size > 1; a single graph otherwise).  ``mel``/``mask`` become the static input buffers.  The warm-up steps are
        real optimisation steps; ``preserve_state`` puts weights, Adam moments / step counters and BatchNorm buffers back
        afterwards, so that the first ``replay()`` is the first step (what ``AudioModel.optimize_parameters`` needs)."""
        self._graphs = None                  # release a previous capture's memory pool before allocating the new one
        self._static = dict(mel=mel.clone(), mask=mask.clone(),
                            video=None if video is None else video.clone(), flow=None if flow is None else flow.clone())
        st = self._static
        snap = self._snapshot() if preserve_state else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._pack_plan = None
            for i in range(warmup):
                if i == warmup - 1:
                    ops.pack_record_begin()          # learn which operand layouts each segment of the step asks for
                self.train_step(st["mel"], st["mask"], st["video"], st["flow"])
            plan = ops.pack_record_end() if warmup > 0 else None
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # Drop every autograd graph built during warm-up: the AccumulateGrad nodes they keep alive are bound to the
        # warm-up stream, and a backward inside the capture would then wait on that (uncaptured) stream.
        self._drop_step_state()
        if snap is not None:
            self._restore(snap)
            torch.cuda.synchronize()
        import os
        # Batched weight packing (ops.prepack: one launch per step segment instead of ~60 small ones).  Measured on B200 at C2
        # (B = 32): 325 instead of 383 launches per step but 15.46 ms instead of 15.31 ms -- packing a layer's operand right before
        # the convolution that reads it leaves it hot in L2, packing everything up front does not.  Off by default;
        # VIAI_BATCHED_PACK=1 enables it (launch-bound small-batch regimes).
        self._pack_plan = plan if (plan and os.environ.get("VIAI_BATCHED_PACK", "0") == "1") else None
        n0 = _lib.launch_count()
        ops.pack_cache_begin()
        if self.world_size == 1 or self.overlap:
            # ONE graph for the whole step; with world_size > 1 the NCCL all-reduces are captured inside it (on NCCL's own stream,
            # forked from / joined to the capture stream by events), overlapping the backward pass
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_body(st["mel"], st["mask"], st["video"], st["flow"])
                self._static_out = self._outputs()
            self._graphs = [g]
            self._segmented = False
        else:
            self._segmented = True
            pool = torch.cuda.graph_pool_handle()
            g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1, pool=pool):
                self._seg_forward_and_d_backward(st["mel"], st["mask"], st["video"], st["flow"])
            self.optimizer_D.all_reduce_grads()
            with torch.cuda.graph(g2, pool=pool):
                self._seg_d_update_and_g_backward()
            self.optimizer_G.all_reduce_grads()
            with torch.cuda.graph(g3, pool=pool):
                self._seg_g_update()
                self._static_out = self._outputs()
            self._graphs = [g1, g2, g3]
        ops.pack_cache_end()
        self.launches_per_step = _lib.launch_count() - n0
        return self

    def prefetch(self, mel, mask, video=None, flow=None):
        """Starts the host->device copy of the NEXT step's inputs on a side stream while the current step runs (the input
        half of a double-buffered loader: the reference's DataLoader uses pinned memory for the same purpose,
        Data_loaders/audio_loader.py:573).  The following ``replay()`` without arguments consumes them."""
        if self._static is None:
            raise RuntimeError("prefetch() needs a captured step (call capture() first)")
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._staging = {k: (None if v is None else torch.empty_like(v)) for k, v in self._static.items()}
            self._staging_ready = torch.cuda.Event()
            self._staging_consumed = torch.cuda.Event()
            self._staging_consumed.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._staging_consumed)       # the previous device-side hand-over has finished
            for k, src in (("mel", mel), ("mask", mask), ("video", video), ("flow", flow)):
                if src is not None:
                    self._staging[k].copy_(src.reshape(self._staging[k].shape), non_blocking=True)
            self._staging_ready.record(self._copy_stream)
        self._prefetched = [k for k, src in (("mel", mel), ("mask", mask), ("video", video), ("flow", flow)) if src is not None]

    def replay(self, mel=None, mask=None, video=None, flow=None):
        st = self._static
        if mel is None and mask is None and getattr(self, "_prefetched", None):
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staging_ready)
            for k in self._prefetched:
                st[k].copy_(self._staging[k], non_blocking=True)      # device-to-device hand-over (microseconds)
            self._staging_consumed.record(cur)
            self._prefetched = None
        if mel is not None:
            st["mel"].copy_(mel, non_blocking=True)
        if mask is not None:
            st["mask"].copy_(mask, non_blocking=True)
        if video is not None:
            st["video"].copy_(video, non_blocking=True)
        if flow is not None:
            st["flow"].copy_(flow, non_blocking=True)
        self.optimizer_D.sync_lr()          # a changed param_groups[0]['lr'] reaches the captured Adam kernels
        self.optimizer_G.sync_lr()
        if not getattr(self, "_segmented", False):
            self._graphs[0].replay()
        else:
            self._graphs[0].replay()
            self.optimizer_D.all_reduce_grads()
            self._graphs[1].replay()
            self.optimizer_G.all_reduce_grads()
            self._graphs[2].replay()
        return self._static_out
