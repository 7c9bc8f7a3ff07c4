"""Helpers of /root/reference/utils/util.py that sit on the AV-sync path: ``l2_norm`` :94-96, ``L2retrieval`` :99-121,
``copy_state_dict`` :124-144.  (Checkpoint helpers :146-173 live on ``AudioModel``.)"""
import numpy as np
import torch.nn as nn

from .. import ops


def l2_norm(x):
    """F.normalize(x, p=2, dim=1) for (B, F) embeddings."""
    return ops.l2_normalize(x)


def L2retrieval(clips_embed, captions_embed, return_ranks=False):
    """Recall@{1,5,10,50} (percent), median and mean rank (1-based) of the matching clip of every caption (utils/util.py:99-121).
    Embeddings are CUDA tensors (or numpy arrays, moved to the device).  The (captions x clips) Euclidean distances come from
    viai_pairdist_fwd and the rank of caption i's own clip is COUNTED on the device (viai_retrieval_ranks: how many clips are
    closer) -- the reference's argsort of the whole matrix is not needed; only the n ranks travel to the host."""
    import torch
    to_dev = lambda a: a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()
    clips, caps = to_dev(clips_embed).float(), to_dev(captions_embed).float()
    with torch.no_grad():
        ranks_t, top1_t = ops.retrieval_ranks(ops.pairdist(caps, clips))
    ranks, top1 = ranks_t.cpu().numpy(), top1_t.cpu().numpy()
    recall = lambda k: 100.0 * float((ranks < k).mean())
    metrics = (recall(1), recall(5), recall(10), recall(50), float(np.floor(np.median(ranks)) + 1), float(ranks.mean() + 1))
    return (metrics, (ranks, top1)) if return_ranks else metrics


def copy_state_dict(state_dict, model, strip=None):
    tgt = model.state_dict()
    for name, param in state_dict.items():
        if strip is not None and name.startswith(strip):
            name = name[len(strip):]
        if name not in tgt or tgt[name].size() != param.size():
            continue
        tgt[name].copy_(param.data if isinstance(param, nn.Parameter) else param)
    return model
