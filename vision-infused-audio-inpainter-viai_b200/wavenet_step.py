"""One teacher-forced WaveNet training step on the CUDA path.

The reference accumulates a WaveNet reconstruction loss when ``update_wavenet`` is set (train_whole_sync.py:105-107) but the
model file that would run it is missing; the step here is the upstream wavenet_vocoder recipe on the reference's own classes:
``y_hat = model(x, c)`` (wavenet_vocoder/wavenet.py:177-235), ``DiscretizedMixturelogisticLoss`` (loss_functions.py:43-59) of
``y_hat[:, :, :-1]`` against ``y[:, 1:, :]`` under ``mask[:, 1:, :]``, Adam, and an exponential moving average of the
parameters (loss_functions.py:62-76).  Like ``GanTrainer`` the step can be captured into CUDA graphs (one graph, or
forward+backward | all-reduce | update when world_size > 1)."""
import torch

from . import _lib, ops
from .optim import FusedAdam


class WaveNetTrainer(object):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, ema_decay=0.9999, num_classes=65536, log_scale_min=-32.23619130191664,
                 world_size=1, process_group=None):
        self.model = model
        self.optimizer = FusedAdam(model.parameters(), lr=lr, betas=betas, eps=eps, world_size=world_size, process_group=process_group)
        self.world_size = world_size
        self.ema_decay = ema_decay
        # one flat shadow buffer: the EMA of every parameter is a single axpby over the optimizer's flat parameter buffer
        self.ema_flat = self.optimizer.flat_param.detach().clone() if ema_decay else None
        self.num_classes, self.log_scale_min = int(num_classes), float(log_scale_min)
        self._graphs = None
        self._static = None
        self.loss = None
        self.launches_per_step = None

    # ---- pieces ------------------------------------------------------------------------------------------------------
    @staticmethod
    def shifted_targets(y, mask):
        """(target, weight) rows aligned with y_hat so that sum_t w[t] * nll(y_hat[t], target[t]) / sum_t w[t] equals the loss of
        y_hat[:, :, :-1] vs y[:, 1:, :] under mask[:, 1:, :] -- without slicing the (B, T, 30) network output."""
        B = y.size(0)
        y2, m2 = y.reshape(B, -1), mask.reshape(B, -1).float()
        z = y2.new_zeros(B, 1)
        return torch.cat((y2[:, 1:], z), 1), torch.cat((m2[:, 1:], z), 1)

    def _forward_backward(self, x, y, c, mask):
        self.optimizer.zero_grad()
        y_hat = self.model(x, c)                                        # (B, 30, T), a view of contiguous (B, T, 30) rows
        target, weight = self.shifted_targets(y, mask)
        nll = ops.dmol_nll(y_hat.transpose(1, 2), target, self.num_classes, self.log_scale_min)
        self.loss = ops.masked_sum(nll, weight, mean=True)
        self.loss.backward()

    def _update(self):
        self.optimizer.step()
        if self.ema_flat is not None:
            ops.axpby_(self.ema_flat, self.ema_decay, self.optimizer.flat_param, 1.0 - self.ema_decay)

    def ema_state_dict(self):
        """EMA weights under the model's parameter names (what the upstream recipe saves next to the raw checkpoint)."""
        out, off = {}, 0
        names = {id(p): n for n, p in self.model.named_parameters()}
        for p, o in zip(self.optimizer._plist, self.optimizer.bucket.offsets):
            out[names[id(p)]] = self.ema_flat[o:o + p.numel()].view(p.shape).clone()
        return out

    def load_ema_state_dict(self, sd):
        names = {id(p): n for n, p in self.model.named_parameters()}
        for p, o in zip(self.optimizer._plist, self.optimizer.bucket.offsets):
            n = names[id(p)]
            if n in sd:
                self.ema_flat[o:o + p.numel()].copy_(sd[n].reshape(-1))

    # ---- eager / captured step -----------------------------------------------------------------------------------------
    def train_step(self, x, y, c, mask):
        """x (B,1,T) input, y (B,T,1) target (the same signal for raw audio), c (B,cin,T/hop), mask (B,T,1) or (B,T) of {0,1}.
        Returns the loss (device scalar)."""
        n0 = _lib.launch_count()
        self._forward_backward(x, y, c, mask)
        self.optimizer.all_reduce_grads()
        self._update()
        self.launches_per_step = _lib.launch_count() - n0
        return self.loss.detach()

    def capture(self, x, y, c, mask, warmup=2):
        self._static = dict(x=x.clone(), y=y.clone(), c=c.clone(), mask=mask.clone().float())
        st = self._static
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.train_step(st["x"], st["y"], st["c"], st["mask"])
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.loss = None                     # drop the warm-up autograd graph (its AccumulateGrad nodes are bound to the side stream)
        import gc
        gc.collect()
        n0 = _lib.launch_count()
        if self.world_size == 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._forward_backward(st["x"], st["y"], st["c"], st["mask"])
                self._update()
                self._static_loss = self.loss.detach()
            self._graphs = [g]
        else:
            pool = torch.cuda.graph_pool_handle()
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1, pool=pool):
                self._forward_backward(st["x"], st["y"], st["c"], st["mask"])
                self._static_loss = self.loss.detach()
            self.optimizer.all_reduce_grads()
            with torch.cuda.graph(g2, pool=pool):
                self._update()
            self._graphs = [g1, g2]
        self.launches_per_step = _lib.launch_count() - n0
        return self

    def replay(self, x=None, y=None, c=None, mask=None):
        st = self._static
        for k, v in (("x", x), ("y", y), ("c", c), ("mask", mask)):
            if v is not None:
                st[k].copy_(v.reshape(st[k].shape), non_blocking=True)
        self.optimizer.sync_lr()
        self._graphs[0].replay()
        if len(self._graphs) > 1:
            self.optimizer.all_reduce_grads()
            self._graphs[1].replay()
        return self._static_loss
