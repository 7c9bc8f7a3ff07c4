"""Helpers of /root/reference/utils/util.py that sit on the AV-sync path: ``l2_norm`` :94-96, ``L2retrieval`` :99-121,
``copy_state_dict`` :124-144.  (Checkpoint helpers :146-173 live on ``AudioModel``.)"""
import numpy as np
import torch.nn as nn

from .. import ops


def l2_norm(x):
    """F.normalize(x, p=2, dim=1) for (B, F) embeddings."""
    return ops.l2_normalize(x)


def L2retrieval(clips_embed, captions_embed, return_ranks=False):
    """Recall@{1,5,10,50}, median and mean rank of the matching clip for every caption.  Embeddings are CUDA tensors (or numpy
    arrays, moved to the device); the (captions x clips) Euclidean distance matrix comes from viai_pairdist_fwd, the ranking
    itself is index work on the host, as in the reference."""
    import torch
    to_dev = lambda a: a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()
    clips, caps = to_dev(clips_embed).float(), to_dev(captions_embed).float()
    captions_num = caps.shape[0]
    with torch.no_grad():
        d = ops.pairdist(caps, clips).cpu().numpy()
    inds = np.argsort(d)
    num = np.arange(captions_num).reshape(captions_num, 1)
    ranks = np.where(inds == num)[1]
    top1 = inds[:, 0]
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    r50 = 100.0 * len(np.where(ranks < 50)[0]) / len(ranks)
    medr = np.floor(np.median(ranks)) + 1
    meanr = ranks.mean() + 1
    if return_ranks:
        return (r1, r5, r10, r50, medr, meanr), (ranks, top1)
    return (r1, r5, r10, r50, medr, meanr)


def copy_state_dict(state_dict, model, strip=None):
    tgt = model.state_dict()
    for name, param in state_dict.items():
        if strip is not None and name.startswith(strip):
            name = name[len(strip):]
        if name not in tgt or tgt[name].size() != param.size():
            continue
        tgt[name].copy_(param.data if isinstance(param, nn.Parameter) else param)
    return model
