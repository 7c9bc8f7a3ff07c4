"""TEST INFRASTRUCTURE ONLY: plain-torch stand-ins for ``viai_b200.ops`` so that the HOST logic of the product modules
(tensor layouts, weight re-packing, shape glue, state_dict contracts) can be exercised by the ``-m "not gpu"`` tests in a
container without a GPU.  Nothing in the product package imports this file; the product ops have no CPU path and raise on
CPU tensors.  ``with installed():`` swaps the functions in for the duration of one test and restores the CUDA-only originals."""
import contextlib

import torch
import torch.nn.functional as F


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def conv2d(x, weight, bias=None, stride=(1, 1), padding=(0, 0), transposed=False):
    f = F.conv_transpose2d if transposed else F.conv2d
    return _nhwc(f(_nchw(x), weight, bias, tuple(stride), tuple(padding)))


def conv2d_stats(x, weight, bias, stride, padding, transposed, stat_groups):
    return conv2d(x, weight, bias, stride, padding, transposed), torch.empty(0, dtype=torch.float64)


def _act(x, act, slope):
    if act == 1:
        return F.relu(x)
    if act == 2:
        return F.leaky_relu(x, slope)
    if act == 3:
        return torch.sigmoid(x)
    return x


def norm_act(y, norm_module, norm, act, slope=0.0, pre_stats=None):
    if norm == "none" or norm_module is None:
        return _act(y, act, slope)
    x = _nchw(y)
    m = norm_module
    if norm == "bn":
        if m.training and m.num_batches_tracked is not None:
            m.num_batches_tracked += 1
        out = F.batch_norm(x, m.running_mean, m.running_var, m.weight, m.bias, m.training or m.running_mean is None,
                           m.momentum if m.momentum is not None else 0.1, m.eps)
    else:
        out = F.instance_norm(x, None, None, m.weight, m.bias, True, 0.1, m.eps)
    return _act(_nhwc(out), act, slope)


def cat_channels(a, b):
    return torch.cat((a, b), 3)


def mul(a, b):
    return a * b


def avgpool_h(x, kh):
    return _nhwc(F.avg_pool2d(_nchw(x), (kh, 1)))


def maxpool3s2(x):
    return _nhwc(F.max_pool2d(_nchw(x), 3, 2, 1))


def add_act(a, b, act=1):
    return _act(a + b, act, 0.0)


def weight_norm(v, g):
    nrm = v.reshape(v.size(0), -1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / nrm)


def shiftcat(x, c, K, dilation, Kpad, mask=None, scale=1.0):
    B, T, R = x.shape
    if mask is not None:
        x = x * mask * scale
    cols = []
    for k in range(K):
        sh = (K - 1 - k) * dilation
        cols.append(F.pad(x, (0, 0, sh, 0))[:, :T])
    if c is not None:
        cols.append(c)
    out = torch.cat(cols, 2)
    return F.pad(out, (0, Kpad - out.size(2)))


def glu_tanh_sigmoid(y):
    a, b = y.split(y.size(-1) // 2, dim=-1)
    return torch.tanh(a) * torch.sigmoid(b)


def axpby(a, alpha, b=None, beta=0.0):
    return alpha * a if b is None else alpha * a + beta * b


def axpby_(a, alpha, b, beta):
    a.mul_(alpha).add_(b, alpha=beta)
    return a


def dmol_nll(rows, target, num_classes=256, log_scale_min=-7.0):
    from oracle import viai_oracle as O
    shp = rows.shape[:-1]
    r = rows.reshape(1, -1, rows.size(-1))
    return O.dmol_nll(r.transpose(1, 2), target.reshape(1, -1, 1), num_classes, log_scale_min).reshape(shp)


def dmol_sample(rows, uniforms, log_scale_min=-7.0):
    from oracle import viai_oracle as O
    nm = rows.size(-1) // 3
    r, u = rows.reshape(-1, 3 * nm), uniforms.reshape(-1, nm + 1)
    return O.sample_dmol(r, u[:, :nm], u[:, nm], log_scale_min).reshape(rows.shape[:-1])


def masked_sum(v, mask=None, mean=True):
    if mask is None:
        return v.mean() if mean else v.sum()
    s = (v * mask).sum()
    return s / mask.sum() if mean else s


def sequence_mask(lengths, max_len):
    return (torch.arange(0, int(max_len)).unsqueeze(0) < lengths.unsqueeze(1)).float()


def l2_normalize(x, eps=1e-12):
    return F.normalize(x, p=2, dim=1, eps=eps)


def pairdist(f1, f2):
    return torch.norm(f1.unsqueeze(1) - f2.unsqueeze(0), p=2, dim=2)


def retrieval_ranks(dist):
    n1 = dist.size(0)
    own = dist.diagonal()[:n1].unsqueeze(1)
    j = torch.arange(dist.size(1)).unsqueeze(0)
    i = torch.arange(n1).unsqueeze(1)
    ranks = ((dist < own) | ((dist == own) & (j < i))).sum(1)
    return ranks.to(torch.int64), dist.argmin(1)


def l2_contrastive(scores, margin=0.0, max_violation=False):
    cost = (margin - scores).clamp(min=0).masked_fill(torch.eye(scores.size(0)) > 0.5, 0)
    if max_violation:
        cost = cost.max(1)[0]
    return (torch.sum(cost ** 2) + torch.sum(scores.diag() ** 2)) / (2 * scores.size(0))


def frames_preprocess(frames_u8, out, c_off, resize_hw, flip, crop_rc, swap_rb):
    from oracle import loader_oracle as LO
    src = frames_u8.numpy()
    for i in range(src.shape[0]):
        img = LO.resize_linear_u8(src[i], int(resize_hw[1]), int(resize_hw[0]))
        img = img.reshape(img.shape[0], img.shape[1], -1)
        if swap_rb:
            img = img[:, :, ::-1]
        if flip:
            img = img[:, ::-1]
        img = img[crop_rc[0]:crop_rc[0] + out.size(1), crop_rc[1]:crop_rc[1] + out.size(2)]
        out[i, :, :, c_off:c_off + img.shape[2]] = torch.from_numpy(((img.astype("float64") - 127.) / 128.).astype("float32"))
    return out


def im2col(g, x, Kpad):
    cols = F.unfold(_nchw(x), (g.R, g.S), padding=(g.pad_h, g.pad_w), stride=(g.stride_h, g.stride_w))     # (N, C*R*S, L), channel-major
    cols = cols.transpose(1, 2).reshape(-1, cols.size(1))
    return F.pad(cols, (0, Kpad - cols.size(1)))


def pointwise_wgrad(U, G):
    return U.t() @ G


_NAMES = ["conv2d", "conv2d_stats", "norm_act", "cat_channels", "mul", "avgpool_h", "maxpool3s2", "add_act", "shiftcat", "weight_norm",
          "glu_tanh_sigmoid", "axpby", "axpby_", "dmol_nll", "dmol_sample", "masked_sum", "sequence_mask", "l2_normalize", "pairdist",
          "l2_contrastive", "frames_preprocess", "im2col", "pointwise_wgrad", "retrieval_ranks"]


@contextlib.contextmanager
def installed():
    from viai_b200 import ops
    saved = {n: getattr(ops, n) for n in _NAMES}
    try:
        for n in _NAMES:
            setattr(ops, n, globals()[n])
        yield ops
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
