"""Flat parameter / gradient buckets and the fused Adam that steps them with one kernel launch.

The bucket is the B200-native replacement for what ``torch.nn.DataParallel`` does in the reference
(/root/reference/utils/model_util.py:137): gradients of one optimizer live in ONE flat fp32 buffer that the
weight-gradient kernels accumulate into directly (``param._viai_grad`` views) and that is all-reduced once per
optimizer step over NCCL."""
import ctypes

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class GradBucket(object):
    """One flat fp32 parameter buffer + one flat gradient buffer for a list of parameters (any device).

    ``p.data`` and ``p.grad`` become views of the flat buffers, in list order; ``p._viai_grad`` is the view the
    weight-gradient kernels accumulate into.  Parameters that never receive a gradient (the reference's dead
    ``convblock1``, /root/reference/networks/New_Inpainting_Networks.py:57,82) simply keep a zero slice, so every rank
    all-reduces the same layout whatever its autograd graph touched.  Device-agnostic: the world_size-2 gloo tests
    exercise exactly this class on CPU."""

    def __init__(self, params, world_size=1, process_group=None):
        self.params = list(params)
        if not self.params:
            raise ValueError("GradBucket needs at least one parameter")
        self.world_size = int(world_size)
        self.process_group = process_group
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.numel = n
        self.flat_param = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.offsets = []
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("GradBucket: all parameters must be float32 on one device")
            k = p.numel()
            view = self.flat_param[off:off + k].view(p.shape)
            view.copy_(p.data)
            p.data = view
            gview = self.flat_grad[off:off + k].view(p.shape)
            p._viai_grad = gview          # weight-gradient kernels accumulate here (ops.grad_target)
            p.grad = gview
            self.offsets.append(off)
            off += k

    def rebind(self):
        """Keeps ``.grad`` pointing at the bucket (autograd may have replaced or dropped it)."""
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != p._viai_grad.data_ptr():
                p.grad = p._viai_grad

    def all_reduce(self):
        """One all-reduce (sum) over the whole bucket; the caller folds 1/world_size into its update."""
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.process_group)

    # ---- overlapped reduction: ranges of the bucket are all-reduced asynchronously (NCCL's own stream) as soon as the backward
    # pass has finished writing them; finish_overlap() reduces whatever is left and joins.  Capturable in a CUDA graph.
    def begin_overlap(self):
        self._works, self._reduced = [], []

    def reduce_range_async(self, start, end):
        if self.world_size > 1 and end > start:
            import torch.distributed as dist
            self._works.append(dist.all_reduce(self.flat_grad[start:end], op=dist.ReduceOp.SUM, group=self.process_group, async_op=True))
            self._reduced.append((start, end))

    def finish_overlap(self, skip=()):
        """All-reduces every range not yet reduced (and not in ``skip``: slices that are zero on every rank, e.g. the reference's
        never-called convblock1), then waits for the asynchronous ones."""
        if self.world_size <= 1:
            return
        import torch.distributed as dist
        covered = sorted(list(getattr(self, "_reduced", [])) + [tuple(r) for r in skip])
        pos = 0
        for s, e in covered + [(self.numel, self.numel)]:
            if s > pos:
                dist.all_reduce(self.flat_grad[pos:s], op=dist.ReduceOp.SUM, group=self.process_group)
            pos = max(pos, e)
        for w in getattr(self, "_works", []):
            w.wait()
        self._works, self._reduced = [], []

    def offset_of(self, param):
        for p, off in zip(self.params, self.offsets):
            if p is param:
                return off
        raise KeyError("parameter is not in this bucket")

    def broadcast_params(self, src=0):
        """Identical initial weights on every rank (what nn.DataParallel's replicate does each step in the reference)."""
        if self.world_size > 1:
            import torch.distributed as dist
            dist.broadcast(self.flat_param, src, group=self.process_group)


def shard_batch(n_items, rank, world_size):
    """[begin, end) of the samples rank ``rank`` owns when a global batch of ``n_items`` is split as evenly as possible
    in rank order (the scatter of nn.DataParallel, /root/reference/utils/model_util.py:137: chunks along dim 0)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    chunk = -(-n_items // world_size)          # torch.chunk semantics: ceil-sized leading chunks
    b = min(rank * chunk, n_items)
    e = min(b + chunk, n_items)
    if e <= b:
        # torch.chunk would hand this rank nothing (e.g. 5 samples over 4 ranks -> 2,2,1,0): the kernels need n > 0 and a rank
        # that skips its all-reduce would hang the others, so refuse loudly instead
        raise ValueError("shard_batch: a global batch of %d leaves rank %d of %d without samples" % (n_items, rank, world_size))
    return b, e


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad).  ``state_dict()`` has the usual per-parameter layout
    (``step``, ``exp_avg``, ``exp_avg_sq``) so reference-style checkpoints (/root/reference/utils/util.py:149-150)
    round-trip; internally the three are views of flat buffers."""

    def __init__(self, params, lr=2e-4, betas=(0.5, 0.999), eps=1e-8, world_size=1, process_group=None):
        params = list(params)
        super(FusedAdam, self).__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.world_size = world_size
        self.process_group = process_group
        plist = [p for g in self.param_groups for p in g["params"]]
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam supports a single param group")
        dev = plist[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam needs CUDA parameters; there is no CPU path")
        self.bucket = GradBucket(plist, world_size, process_group)
        n = self.numel = self.bucket.numel
        self.flat_param, self.flat_grad = self.bucket.flat_param, self.bucket.flat_grad
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        self.lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self._lr_host = float(lr)
        for p, off in zip(plist, self.bucket.offsets):
            k = p.numel()
            self.state[p] = dict(step=self.step_dev[0], exp_avg=self.flat_m[off:off + k].view(p.shape),
                                 exp_avg_sq=self.flat_v[off:off + k].view(p.shape))
        self._plist = plist

    def zero_grad(self, set_to_none=False):
        L = _lib.lib()
        _lib.check(L.viai_fill(_p(self.flat_grad), self.numel, 0.0, _stream()), "zero_grad")
        self.bucket.rebind()

    def all_reduce_grads(self):
        """One NCCL all-reduce (sum) over the whole bucket; the 1/world_size is folded into the Adam kernel."""
        self.bucket.all_reduce()

    def sync_lr(self):
        """Copies ``param_groups[0]['lr']`` into the device scalar the Adam kernel reads.  ``step()`` does it itself; a captured
        step never runs ``step()`` again, so ``GanTrainer.replay`` / ``WaveNetTrainer.replay`` call this before launching their
        graphs (LR schedules, ``load_state_dict`` after capture)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()                # LR schedules: refresh the device scalar (outside any captured graph)
        L = _lib.lib()
        b1, b2 = g["betas"]
        _lib.check(L.viai_adam_step(_p(self.flat_param), _p(self.flat_grad), _p(self.flat_m), _p(self.flat_v),
                                    self.numel, _p(self.lr_dev), float(b1), float(b2), float(g["eps"]),
                                    _p(self.step_dev), 1, 1.0 / self.world_size, _stream()), "adam_step")
        from . import ops
        ops.weights_updated()             # the kernel rewrote the parameters in place: cached packed operands are stale

    def load_state_dict(self, state_dict):
        sd = state_dict["state"]
        for i, p in enumerate(self._plist):
            if i in sd:
                st = sd[i]
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                self.step_dev.fill_(float(st["step"]))
        pg = state_dict["param_groups"][0]
        self.param_groups[0]["lr"] = pg["lr"]
        self.param_groups[0]["betas"] = tuple(pg["betas"])
        self.param_groups[0]["eps"] = pg["eps"]
