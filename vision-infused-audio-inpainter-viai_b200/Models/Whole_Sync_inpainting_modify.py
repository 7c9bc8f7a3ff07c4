"""``AudioModel`` -- the model-glue class that /root/reference/train_whole_sync.py drives (imported there as
``Models.Whole_Sync_inpainting_modify``, line 10) and that is MISSING from the reference tree.

Only its call contract exists upstream: the methods and attributes used at train_whole_sync.py:49-112,159-167 and the
attributes read by utils/util.py:146-173 (``Mel_Encoder``, ``Mel_Decoder``, ``netD``, ``optimizer_G``, ``optimizer_D``).  This
class implements that contract on top of :class:`viai_b200.step.GanTrainer`; everything the contract leaves open (mask
schedule, loss weights, optimizer settings) is OURS and read from ``hparams`` with the defaults of SURVEY.md 3.1:

    blank_length   : frames of the contiguous time band that is blanked (``hparams.blank_length``, default W // 2), optionally
                     grown linearly from ``blank_length_start`` over ``blank_warmup_steps`` steps
    input tuple    : the 8-tuple of Data_loaders/audio_loader.py:532
                     (video_batch, flow_batch, c_batch (B, n_mel, W), x_batch, y_batch, g_batch, input_lengths, path_batch)
    update_wavenet : (``hparams.update_wavenet``, default False) also runs one teacher-forced WaveNet step per batch on the
                     waveform (x_batch, y_batch, input_lengths) conditioned on the inpainted mel (detached) and reports its masked
                     DMoL loss as ``reconstruct_loss_item`` (train_whole_sync.py:105-107); ``hparams.wavenet_kwargs`` overrides the
                     WaveNet constructor defaults
    cuda_graph     : (``hparams.cuda_graph``, default True) ``optimize_parameters`` captures the whole D + G update as a CUDA
                     graph on the first batch of a given shape (weights, Adam state and BatchNorm buffers are restored after the
                     capture's warm-up, so the first replay IS the first step) and replays it afterwards; the mask is a graph
                     input, so the ``blank_length`` schedule does not force a re-capture; a new batch shape does
    EmbeddingL2    : ``test()`` reports the L2 contrastive loss between the l2-normalised audio bottleneck and visual embedding
                     (``EmbeddingL2_item``, train_whole_sync.py:109) when the two have the same width (native 80-bin mels)
"""
import contextlib
import os
from collections import OrderedDict

import numpy as np
import torch

from .. import ops
from ..loss_functions import L2ContrastiveLoss, sequence_mask
from ..networks.Image_Embedding import ImageEmbedding
from ..step import GanTrainer
from ..utils.util import l2_norm


@contextlib.contextmanager
def _frozen_norm_buffers(modules):
    """Evaluation forwards use batch statistics like the training forwards (the pix2pix convention the reference borrows,
    README.md:40) but must not move the BatchNorm running buffers: momentum 0 keeps running_mean / running_var bit-identical,
    num_batches_tracked is put back afterwards."""
    bns = [m for mod in modules if mod is not None for m in mod.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    saved = [(m.momentum, None if m.num_batches_tracked is None else m.num_batches_tracked.clone()) for m in bns]
    for m in bns:
        m.momentum = 0.0
    try:
        yield
    finally:
        for m, (mom, nbt) in zip(bns, saved):
            m.momentum = mom
            if nbt is not None:
                m.num_batches_tracked.copy_(nbt)


class AudioModel(object):
    def __init__(self, hparams, device=torch.device("cuda")):
        self.hparams = hparams
        self.device = torch.device(device)
        self.use_video = bool(getattr(hparams, "image", False) or getattr(hparams, "flow", False))
        ve = ImageEmbedding(hparams).to(self.device) if self.use_video else None
        self.trainer = GanTrainer(hparams, self.device, decoder="MelDecoderImage" if self.use_video else "MelDecoder",
                                  video_encoder=ve, world_size=int(getattr(hparams, "world_size", 1)))
        t = self.trainer
        self.Mel_Encoder, self.Mel_Decoder, self.netD, self.VideoEncoder = t.Mel_Encoder, t.Mel_Decoder, t.netD, ve
        self.optimizer_G, self.optimizer_D = t.optimizer_G, t.optimizer_D
        self.train = 1
        self.update_wavenet = bool(getattr(hparams, "update_wavenet", False))
        self.wavenet = self.wavenet_trainer = None
        if self.update_wavenet:
            from ..wavenet_step import WaveNetTrainer
            from ..wavenet_vocoder import WaveNet
            kw = dict(cin_channels=hparams.cin_channels)
            kw.update(getattr(hparams, "wavenet_kwargs", {}))
            self.wavenet = WaveNet(**kw).to(self.device)
            self.wavenet_trainer = WaveNetTrainer(self.wavenet, lr=float(getattr(hparams, "wavenet_lr", 1e-3)),
                                                  world_size=int(getattr(hparams, "world_size", 1)))
        self.criterionEmbedding = L2ContrastiveLoss(margin=float(getattr(hparams, "embedding_margin", 0.0)))
        self.audio = self.audio_target = self.input_lengths = None
        self.blank_length = 0
        self.reconstruct_loss_item = 0.0
        self.EmbeddingL2_item = 0.0
        self.loss_mel_L1_item = 0.0
        self.loss_D_item = self.loss_G_GAN_item = 0.0
        self.mel_net_norm = self.video_net_norm = None
        self.current_lr = self.optimizer_G.param_groups[0]["lr"]
        self._out = None
        self.mel = self.mask = self.video = self.flow = None
        self.use_graph = bool(getattr(hparams, "cuda_graph", True))
        self._graph_key = None

    # ---- inputs ------------------------------------------------------------------------------------------------------
    def get_blank_space_length(self, global_step):
        hp = self.hparams
        W = int(getattr(hp, "max_mel_lengths", 256))
        full = int(getattr(hp, "blank_length", W // 2))
        warm = int(getattr(hp, "blank_warmup_steps", 0))
        if warm > 0:
            start = int(getattr(hp, "blank_length_start", max(1, full // 4)))
            full = start + (full - start) * min(global_step, warm) // warm
        self.blank_length = full
        return full

    def set_inputs(self, data):
        video, flow, c_batch = data[0], data[1], data[2]
        mel = c_batch.to(self.device, non_blocking=True).float()              # (B, n_mel, W) in [0, 1]
        B, Hm, W = mel.shape
        self.mel = mel.reshape(B, 1, Hm, W)
        bl = min(self.blank_length if self.blank_length > 0 else W // 2, W)
        t0 = (W - bl) // 2
        mask = torch.ones_like(self.mel)
        mask[..., t0:t0 + bl] = 0.0
        self.mask = mask
        if self.use_video:
            self.video = video.to(self.device, non_blocking=True).float().reshape(B, -1, 3, video.size(-2), video.size(-1))
            self.flow = flow.to(self.device, non_blocking=True).float().reshape(B, -1, 2, flow.size(-2), flow.size(-1))
        if self.update_wavenet:
            self.audio = data[3].to(self.device, non_blocking=True).float()            # x_batch (B, 1, T)
            self.audio_target = data[4].to(self.device, non_blocking=True).float()     # y_batch (B, T, 1)
            self.input_lengths = data[6].to(self.device, non_blocking=True)

    # ---- one optimisation step / one evaluation forward ----------------------------------------------------------------
    def optimize_parameters(self, global_step):
        t = self.trainer
        if self.use_graph:
            key = tuple(None if x is None else tuple(x.shape) for x in (self.mel, self.video, self.flow))
            if key != self._graph_key:
                t.capture(self.mel, self.mask, self.video, self.flow, warmup=1, preserve_state=True)
                self._graph_key = key
            self._out = dict(t.replay(self.mel, self.mask, self.video, self.flow))
        else:
            self._out = t.train_step(self.mel, self.mask, self.video, self.flow)
        self.fake = self._out["fake"]
        if self.update_wavenet:
            T = self.audio.size(-1)
            mask = sequence_mask(self.input_lengths, T).unsqueeze(-1)
            self._out["reconstruct_loss"] = self.wavenet_trainer.train_step(self.audio, self.audio_target,
                                                                            self.fake.detach()[:, 0], mask)

    def test(self):
        t = self.trainer
        B, _, Hm, W = self.mel.shape
        masked = ops.mul(self.mel.reshape(B, Hm, W, 1), self.mask.reshape(B, Hm, W, 1)).reshape(self.mel.shape)
        with _frozen_norm_buffers((t.Mel_Encoder, t.Mel_Decoder, self.VideoEncoder)):
            feats = t.Mel_Encoder(masked)
            vnet = self.VideoEncoder(self.video, self.flow) if self.use_video else None
            self.fake = t.Mel_Decoder(feats, self.mel.shape, vnet) if self.use_video else t.Mel_Decoder(feats, self.mel.shape)
        l1 = t.criterionL1(self.fake, self.mel)
        self._out = dict(fake=self.fake, loss_L1=l1, loss_D=torch.zeros((), device=self.device),
                         loss_G_GAN=torch.zeros((), device=self.device), loss_G=l1 * t.lambda_L1)
        self.mel_net_norm = l2_norm(feats[-1].reshape(B, -1))
        self.video_net_norm = l2_norm(vnet.reshape(B, -1)) if vnet is not None else torch.zeros_like(self.mel_net_norm)
        if vnet is not None and self.video_net_norm.shape == self.mel_net_norm.shape:
            self._out["EmbeddingL2"] = self.criterionEmbedding(self.mel_net_norm, self.video_net_norm)

    def get_loss_items(self):
        o = self._out
        ops.check_f16_overflow()          # the step's only host synchronisation point: also the place to learn about saturation
        self.loss_mel_L1_item = float(o["loss_L1"])
        self.loss_D_item = float(o["loss_D"])
        self.loss_G_GAN_item = float(o["loss_G_GAN"])
        self.EmbeddingL2_item = float(o["EmbeddingL2"]) if "EmbeddingL2" in o else 0.0
        self.reconstruct_loss_item = float(o["reconstruct_loss"]) if "reconstruct_loss" in o else 0.0

    def get_current_errors(self):
        return OrderedDict([("loss_D", self.loss_D_item), ("loss_G_GAN", self.loss_G_GAN_item), ("loss_mel_L1", self.loss_mel_L1_item)])

    def get_current_visuals(self):
        img = lambda t: (t[0, 0].detach().float().clamp(0, 1) * 255).byte().cpu().numpy()[:, :, None].repeat(3, axis=2)
        return OrderedDict([("real_mel", img(self.mel)), ("masked_mel", img(self.mel * self.mask)), ("fake_mel", img(self.fake))])

    def TF_writer(self, writer, step):
        name = getattr(self.hparams, "name", "viai")
        for k, v in self.get_current_errors().items():
            writer.add_scalar("%s_%s" % (name, k), v, step)

    def del_no_need(self):
        self._out = None
        self.fake = None

    def eval_model_test(self, global_step, eval_dir):
        """Writes the current inpainted mel next to the ground truth (vocoding it is WaveNet.incremental_forward's job)."""
        with torch.no_grad():
            self.test()
        np.save(os.path.join(eval_dir, "step%09d_fake_mel.npy" % global_step), self.fake[0, 0].cpu().numpy())
        np.save(os.path.join(eval_dir, "step%09d_real_mel.npy" % global_step), self.mel[0, 0].cpu().numpy())

    # ---- checkpoints: the format of utils/util.py:146-162 ----------------------------------------------------------------
    def save_inpainting_checkpoint(self, global_step, global_test_step, checkpoint_dir, epoch, hparams=None):
        hp = hparams if hparams is not None else self.hparams
        path = os.path.join(checkpoint_dir, getattr(hp, "name", "viai") + "_checkpoint_step{:09d}.pth.tar".format(global_step))
        save_opt = getattr(hp, "save_optimizer_state", True)
        ck = {"Mel_Encoder": self.Mel_Encoder.state_dict(), "Mel_Decoder": self.Mel_Decoder.state_dict(),
              "netD": self.netD.state_dict(),
              "optimizer_G": self.optimizer_G.state_dict() if save_opt else None,
              "optimizer_D": self.optimizer_D.state_dict() if save_opt else None,
              "global_step": global_step, "global_epoch": epoch, "global_test_step": global_test_step}
        # Beyond utils/util.py:146-162 (whose VideoEncoder line is commented out): optimizer_G's state covers the visual
        # encoder's parameters, so a resumed image/flow run needs its weights too; likewise the vocoder when it is trained here.
        if self.VideoEncoder is not None:
            ck["VideoEncoder"] = self.VideoEncoder.state_dict()
        if self.wavenet is not None:
            ck["wavenet"] = self.wavenet.state_dict()
            ck["wavenet_optimizer"] = self.wavenet_trainer.optimizer.state_dict() if save_opt else None
            if self.wavenet_trainer.ema_flat is not None:
                ck["wavenet_ema"] = self.wavenet_trainer.ema_state_dict()
        torch.save(ck, path)
        print("Saved checkpoint:", path)
        return path

    def load_inpainting_checkpoint(self, path, reset_optimizer=False):
        ck = torch.load(path, map_location=self.device, weights_only=False)
        self.Mel_Encoder.load_state_dict(ck["Mel_Encoder"])
        self.Mel_Decoder.load_state_dict(ck["Mel_Decoder"])
        self.netD.load_state_dict(ck["netD"])
        if self.VideoEncoder is not None and "VideoEncoder" in ck:      # absent from reference-format / older checkpoints
            self.VideoEncoder.load_state_dict(ck["VideoEncoder"])
        if self.wavenet is not None and "wavenet" in ck:
            self.wavenet.load_state_dict(ck["wavenet"])
            if self.wavenet_trainer.ema_flat is not None and "wavenet_ema" in ck:
                self.wavenet_trainer.load_ema_state_dict(ck["wavenet_ema"])
            if not reset_optimizer and ck.get("wavenet_optimizer") is not None:
                self.wavenet_trainer.optimizer.load_state_dict(ck["wavenet_optimizer"])
        if not reset_optimizer:
            if ck.get("optimizer_G") is not None:
                self.optimizer_G.load_state_dict(ck["optimizer_G"])
            if ck.get("optimizer_D") is not None:
                self.optimizer_D.load_state_dict(ck["optimizer_D"])
        ops.weights_updated()
        return ck["global_step"], ck["global_epoch"], ck["global_test_step"]

    def load_part_checkpoint(self, path=None):
        """utils/util.py:165-173: shape-tolerant copy of the generator weights only."""
        path = path if path is not None else getattr(self.hparams, "pretrain_path", None)
        ck = torch.load(path, map_location=self.device, weights_only=False)
        for key, mod in (("Mel_Encoder", self.Mel_Encoder), ("Mel_Decoder", self.Mel_Decoder)):
            tgt = mod.state_dict()
            for name, param in ck[key].items():
                if name in tgt and tuple(param.shape) == tuple(tgt[name].shape):
                    tgt[name].copy_(param)
                elif name in tgt:
                    print("mismatch:", name, tuple(param.shape), tuple(tgt[name].shape))
        ops.weights_updated()
        return self
