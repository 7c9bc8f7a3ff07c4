#!/usr/bin/env python
"""2-rank NCCL parity of the data-parallel GAN step (run with torchrun --nproc-per-node 2).

Each rank takes its shard of a global batch (optim.shard_batch = the scatter of the reference's nn.DataParallel,
/root/reference/utils/model_util.py:137), runs GanTrainer.train_step; the gradient buckets are all-reduced by NCCL.  Rank 0
then checks the averaged bucket against the mean of the per-replica gradients of the CPU oracle (BatchNorm statistics are
per replica in both) and that both ranks hold identical weights after the step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from viai_b200 import Options_inpainting as OI
    from viai_b200.optim import shard_batch
    from viai_b200.step import GanTrainer
    from oracle import viai_oracle as O
    import viai_test_helpers as H
    hp = OI.Inpainting_Config(cin_channels=80)
    torch.manual_seed(100 + rank)                       # ranks start from DIFFERENT weights; the broadcast must fix that
    tr = GanTrainer(hp, "cuda", world_size=world)
    for opt in (tr.optimizer_G, tr.optimizer_D):
        opt.bucket.broadcast_params(0)
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    esd, gsd, dsd = cpu(tr.Mel_Encoder), cpu(tr.Mel_Decoder), cpu(tr.netD)
    g = torch.Generator().manual_seed(7)
    mel = torch.rand(4, 1, 80, 64, generator=g)
    mask = O.time_band_mask(mel.shape, 16, 32)
    b, e = shard_batch(mel.size(0), rank, world)
    # gradients only (no update) through the same segments the step uses
    tr._seg_forward_and_d_backward(mel[b:e].cuda(), mask[b:e].cuda())
    tr.optimizer_D.all_reduce_grads()
    gD = tr.optimizer_D.flat_grad.clone() / world
    tr._seg_d_update_and_g_backward()
    tr.optimizer_G.all_reduce_grads()
    gG = tr.optimizer_G.flat_grad.clone() / world
    tr._seg_g_update()
    torch.cuda.synchronize()
    # identical weights on both ranks after the step
    for name, opt in (("G", tr.optimizer_G), ("D", tr.optimizer_D)):
        w = opt.flat_param.clone()
        w0 = w.clone()
        dist.broadcast(w0, 0)
        assert torch.equal(w, w0), "rank %d: %s weights diverged" % (rank, name)
    if rank == 0:
        # oracle: the D phase of every replica from the same initial weights; expected bucket = mean over replicas
        parts = [O.gan_step(esd, gsd, dsd, mel[s:t], mask[s:t], 80, update=False) for (s, t) in (shard_batch(4, r, world) for r in range(world))]
        ps = dict(tr.netD.named_parameters())
        off = dict(zip([n for n, _ in tr.netD.named_parameters()], tr.optimizer_D.bucket.offsets))
        want = {k: sum(p_["grads_D"][k].double() for p_ in parts) / world for k in parts[0]["grads_D"]}
        got = {k: gD[off[k]:off[k] + ps[k].numel()].view(ps[k].shape) for k in want}
        l2, cos, worst = H.whole_net_metrics(got, want)
        print("D gradient bucket vs mean of per-replica oracle gradients: L2 %.3e cos %.6f worst %.3e" % (l2, cos, worst))
        assert l2 < 5e-2 and cos > 0.999
        print("ddp parity OK (world %d): identical post-step weights on all ranks; averaged D bucket matches the oracle" % world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
