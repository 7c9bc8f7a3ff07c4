#!/usr/bin/env python
"""N-rank NCCL check of the OVERLAPPED data-parallel GAN step (torchrun --nproc-per-node N).

(1) eager step: bucket ranges all-reduced asynchronously while the backward pass runs (step.GanTrainer._plan_overlap) must
    leave the same post-step weights as the plain one-all-reduce-per-optimizer step (VIAI_DDP_OVERLAP=0 semantics, run here by
    calling the segments by hand), identical on every rank;
(2) the whole step, NCCL all-reduces included, captured as ONE CUDA graph and replayed: weights stay identical across ranks,
    losses finite, and the replay equals the eager overlapped step from the same state;
(3) per-step time of the captured step at B = 32 / GPU, 256 x 256 (weak) and at global B = 32 (strong)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def same_on_all_ranks(t, what, rank):
    t0 = t.clone()
    dist.broadcast(t0, 0)
    assert torch.equal(t, t0), "rank %d: %s differs from rank 0" % (rank, what)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])), timeout=datetime.timedelta(seconds=120))
    from viai_b200 import Options_inpainting as OI
    from viai_b200.step import GanTrainer
    hp = OI.Inpainting_Config(cin_channels=80)
    g = torch.Generator().manual_seed(7 + rank)
    mel = torch.rand(2, 1, 80, 64, generator=g).cuda()
    mask = torch.ones_like(mel); mask[..., 16:48] = 0

    def fresh(seed):
        torch.manual_seed(seed + rank)          # different weights per rank: the constructor's broadcast must fix that
        return GanTrainer(hp, "cuda", world_size=world)

    # (1) overlapped eager step vs the plain segments + whole-bucket all-reduces
    a = fresh(1)
    assert a.overlap
    for opt in (a.optimizer_G, a.optimizer_D):
        same_on_all_ranks(opt.flat_param, "initial weights", rank)
    ra = a.train_step(mel, mask)
    os.environ["VIAI_DDP_OVERLAP"] = "0"
    c = fresh(1)
    os.environ["VIAI_DDP_OVERLAP"] = "1"
    assert not c.overlap
    rc = c.train_step(mel, mask)
    torch.cuda.synchronize()
    for oa, oc, name in ((a.optimizer_G, c.optimizer_G, "G"), (a.optimizer_D, c.optimizer_D, "D")):
        same_on_all_ranks(oa.flat_param, "post-step %s weights (overlap)" % name, rank)
        err = float((oa.flat_grad - oc.flat_grad).norm() / oc.flat_grad.norm())
        assert err < 1e-3, "%s gradient bucket: overlapped vs plain reduction differ by %.3e" % (name, err)
    # (2) one graph with NCCL inside
    d = fresh(2)
    state = [(o.flat_param.clone(), o.flat_m.clone(), o.flat_v.clone(), o.step_dev.clone()) for o in (d.optimizer_G, d.optimizer_D)]
    d.capture(mel, mask, warmup=1, preserve_state=True)
    assert len(d._graphs) == 1
    out = d.replay(mel, mask)
    torch.cuda.synchronize()
    e = fresh(2)
    re = e.train_step(mel, mask)
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(out[k]).all()) for k in ("loss_D", "loss_G", "loss_L1"))
    for od, oe, name in ((d.optimizer_G, e.optimizer_G, "G"), (d.optimizer_D, e.optimizer_D, "D")):
        same_on_all_ranks(od.flat_param, "post-replay %s weights" % name, rank)
        err = float((od.flat_grad - oe.flat_grad).norm() / oe.flat_grad.norm())
        assert err < 1e-3, "%s gradient bucket: graph replay vs eager differ by %.3e" % (name, err)
    for _ in range(3):
        d.replay(mel, mask)
    torch.cuda.synchronize()
    same_on_all_ranks(d.optimizer_G.flat_param, "weights after 4 replays", rank)
    if rank == 0:
        print("overlap OK (world %d): eager overlapped == plain reduction, one-graph replay == eager, ranks identical" % world)
    # (3) timing at C2
    del a, c, d, e
    torch.cuda.empty_cache()
    hp2 = OI.Inpainting_Config(cin_channels=256)
    for tag, bs in (("weak B=32/GPU", 32), ("strong global B=32", 32 // world)):
        for mode in ("1", "0"):
            os.environ["VIAI_DDP_OVERLAP"] = mode
            torch.manual_seed(3)
            tr = GanTrainer(hp2, "cuda", world_size=world)
            m2 = torch.rand(bs, 1, 256, 256, device="cuda"); k2 = torch.ones_like(m2); k2[..., 64:192] = 0
            tr.capture(m2, k2, warmup=2)
            for _ in range(3):
                tr.replay()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                tr.replay()
            e1.record(); dist.barrier(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                print("%s, overlap=%s (%s): %.3f ms/step, %d graph(s)" % (tag, mode, "one graph, NCCL inside" if mode == "1" else "3 graphs + 2 eager all-reduces",
                                                                        float(ms), len(tr._graphs)))
            del tr
            torch.cuda.empty_cache()
    os.environ["VIAI_DDP_OVERLAP"] = "1"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
