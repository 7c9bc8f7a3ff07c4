"""Host-side logic of the data-parallel GAN step on CPU: two ``gloo`` ranks, one flat gradient bucket per optimizer.

The reference's only parallelism is ``torch.nn.DataParallel`` (/root/reference/utils/model_util.py:137): the batch is
chunked along dim 0, every replica runs the step on its chunk with per-replica BatchNorm statistics and the gradients are
summed.  The B200 path does the same with one process per GPU and one all-reduce per optimizer (optim.GradBucket).  These
tests drive that plumbing with the CPU oracle standing in for the kernels: what must hold is that the all-reduced bucket
equals the mean of the per-replica gradients, that parameters no rank touched (the dead ``convblock1``) stay zero in the same
slots on every rank, and that the rank-order sharding covers the batch exactly once."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import viai_oracle as O
import viai_test_helpers as H

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _modules():
    import torch.nn as nn
    from viai_b200 import Options_inpainting
    from viai_b200.networks.Discriminator_Networks import MelDiscriminator
    from viai_b200.networks.Inpainting_Networks import MelEncoder
    from viai_b200.networks.New_Inpainting_Networks import MelDecoder
    hp = Options_inpainting.Inpainting_Config(cin_channels=80)
    torch.manual_seed(1234)
    return MelEncoder(hp, nn.BatchNorm2d), MelDecoder(hp, nn.BatchNorm2d), MelDiscriminator(norm_layer=nn.BatchNorm2d)


def _batch():
    g = torch.Generator().manual_seed(7)
    mel = torch.rand(4, 1, 80, 32, generator=g)
    mask = O.time_band_mask(mel.shape, 8, 16)
    return mel, mask


def _oracle_grads(enc, dec, dis, mel, mask):
    sd = lambda m: {k: v.detach().clone() for k, v in m.state_dict().items()}
    r = O.gan_step(sd(enc), sd(dec), sd(dis), mel, mask, 80, "bn", "bn", update=False)
    return r["grads_E"], r["grads_Dec"], r["grads_D"]


def _fill(bucket_module_pairs, grads_by_module):
    """Writes oracle gradients where the kernels would accumulate them: into ``param._viai_grad``."""
    for (module, grads) in zip(bucket_module_pairs, grads_by_module):
        for name, p in module.named_parameters():
            if name in grads:
                p._viai_grad.copy_(grads[name])


def _worker(rank, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from viai_b200.optim import GradBucket, shard_batch
        torch.set_num_threads(2)
        enc, dec, dis = _modules()
        if rank != 0:                                 # ranks start from different weights; broadcast must fix that
            with torch.no_grad():
                for m in (enc, dec, dis):
                    for p in m.parameters():
                        p.add_(1.0)
        bG = GradBucket(list(enc.parameters()) + list(dec.parameters()), WORLD)
        bD = GradBucket(list(dis.parameters()), WORLD)
        bG.broadcast_params(0)
        bD.broadcast_params(0)
        mel, mask = _batch()
        b, e = shard_batch(mel.size(0), rank, WORLD)
        gE, gDec, gD = _oracle_grads(enc, dec, dis, mel[b:e], mask[b:e])
        _fill((enc, dec), (gE, gDec))
        _fill((dis,), (gD,))
        bG.all_reduce()
        bD.all_reduce()
        full_G = bG.flat_grad.clone()
        # the overlapped form (ranges reduced asynchronously as the backward pass completes them, the remainder at the end, the
        # all-zero convblock1 slice skipped) must leave exactly the same bucket
        _fill((enc, dec), (gE, gDec))
        dec_params = list(dec.parameters())
        dec_start = bG.offset_of(dec_params[0])
        blk1 = bG.offset_of(next(dec.convblock1.parameters()))
        blk2 = bG.offset_of(next(dec.convblock2.parameters()))
        bG.begin_overlap()
        bG.reduce_range_async(blk2, bG.numel)
        bG.reduce_range_async(dec_start, blk1)
        bG.finish_overlap(skip=[(blk1, blk2)])
        overlap_equal = bool(torch.equal(bG.flat_grad, full_G))
        torch.save(dict(G=full_G / WORLD, D=bD.flat_grad / WORLD, Gp=bG.flat_param.clone(), Dp=bD.flat_param.clone(),
                        shard=(b, e), overlap_equal=overlap_equal), os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_bucket_allreduce_matches_mean_of_replica_grads(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    # sharding: rank order, disjoint, covering
    assert r0["shard"] == (0, 2) and r1["shard"] == (2, 4)
    # broadcast made the replicas identical; all-reduce left identical buckets on both ranks
    for k in ("G", "D", "Gp", "Dp"):
        assert torch.equal(r0[k], r1[k]), k
    assert r0["overlap_equal"] and r1["overlap_equal"]        # GradBucket.reduce_range_async / finish_overlap == one all-reduce
    # expected: mean over replicas of the per-replica gradients (BatchNorm statistics are per replica, as in DataParallel)
    from viai_b200.optim import GradBucket
    enc, dec, dis = _modules()
    mel, mask = _batch()
    parts = [_oracle_grads(enc, dec, dis, mel[b:e], mask[b:e]) for (b, e) in ((0, 2), (2, 4))]
    bG = GradBucket(list(enc.parameters()) + list(dec.parameters()))
    bD = GradBucket(list(dis.parameters()))
    assert torch.equal(bG.flat_param, r0["Gp"]) and torch.equal(bD.flat_param, r0["Dp"])
    want_G, want_D = torch.zeros_like(bG.flat_grad), torch.zeros_like(bD.flat_grad)
    for gE, gDec, gD in parts:
        _fill((enc, dec), (gE, gDec))
        _fill((dis,), (gD,))
        want_G += bG.flat_grad / WORLD
        want_D += bD.flat_grad / WORLD
    assert H.relerr(r0["G"], want_G) < 1e-6
    assert H.relerr(r0["D"], want_D) < 1e-6
    # the reference's dead parameters (convblock1, never called: New_Inpainting_Networks.py:57,82) own all-zero slices
    names = [n for n, _ in enc.named_parameters()] + [n for n, _ in dec.named_parameters()]
    offs = dict(zip(names[len(list(enc.parameters())):], bG.offsets[len(list(enc.parameters())):]))
    dead = [(n, p) for n, p in dec.named_parameters() if n.startswith("convblock1.")]
    assert dead, "MelDecoder should carry the reference's unused convblock1"
    for n, p in dead:
        sl = r0["G"][offs[n]:offs[n] + p.numel()]
        assert torch.count_nonzero(sl) == 0, n
    assert torch.count_nonzero(r0["G"]) > 0.5 * (r0["G"].numel() - sum(p.numel() for _, p in dead))


def test_shard_batch_covers_every_sample_once():
    """Same chunks as torch.chunk (DataParallel's scatter); a split that would leave a rank empty is refused loudly (an empty
    shard fails the kernels' n > 0 checks or hangs the other ranks' all-reduce)."""
    from viai_b200.optim import shard_batch
    for n in (1, 5, 32, 33):
        for w in (1, 2, 3, 8):
            want = [c.numel() for c in torch.arange(n).chunk(w)]
            if len(want) < w:
                with pytest.raises(ValueError, match="without samples"):
                    [shard_batch(n, r, w) for r in range(w)]
                continue
            seen = []
            for r in range(w):
                b, e = shard_batch(n, r, w)
                assert 0 <= b < e <= n
                seen += list(range(b, e))
            assert seen == list(range(n)), (n, w)
            assert [shard_batch(n, r, w)[1] - shard_batch(n, r, w)[0] for r in range(w)] == want
    with pytest.raises(ValueError):
        shard_batch(4, 2, 2)
    with pytest.raises(ValueError, match="without samples"):
        shard_batch(5, 3, 4)


def test_grad_bucket_views_alias_flat_buffers():
    from viai_b200.optim import GradBucket
    _, _, dis = _modules()
    ref = {n: p.detach().clone() for n, p in dis.named_parameters()}
    b = GradBucket(list(dis.parameters()))
    assert b.numel == sum(p.numel() for p in dis.parameters()) == 1555072
    for n, p in dis.named_parameters():
        assert torch.equal(p.detach(), ref[n])                      # values preserved by the re-homing
        assert p.data_ptr() >= b.flat_param.data_ptr() and p.grad.data_ptr() >= b.flat_grad.data_ptr()
    b.flat_param.zero_()
    assert all(torch.count_nonzero(p) == 0 for p in dis.parameters())
    for p in dis.parameters():
        p.grad = None
    b.rebind()
    assert all(p.grad is not None and p.grad.data_ptr() == p._viai_grad.data_ptr() for p in dis.parameters())
