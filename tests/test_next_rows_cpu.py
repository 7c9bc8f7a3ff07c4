"""SURVEY 8f "next" rows on the CPU: (1) the oracle restatements against golden vectors produced by the reference itself
(tests/golden/next_rows.pt, oracle/make_golden.py ``next_rows_fixture``): DMoL loss values + gradients, masked mean, EMA,
L2 contrastive loss, l2_norm / l2_sim / L2retrieval, Inpainting_Dis, DomainDis, ImageEmbedding_single/_finetune/2, one
teacher-forced WaveNet training step; (2) the HOST logic of the product modules (layouts, weight linearisation, operand
packing, autograd glue) dry-run with plain-torch stand-ins for the CUDA ops (tests/torch_ops_shim.py -- test infrastructure;
the product ops themselves have no CPU path) against the same golden vectors.  The CUDA kernels are checked in
tests/test_next_rows_gpu.py."""
import math

import pytest
import torch

import torch_ops_shim as shim
import viai_test_helpers as H
from oracle import fixtures as FX
from oracle import make_golden as MG
from oracle import viai_oracle as O


@pytest.fixture(scope="module")
def fx():
    return H.load_golden("next_rows.pt")


@pytest.fixture(scope="module")
def inp():
    return MG.next_rows_inputs()


def _filled(module, salt=0):
    sd = module.state_dict()
    FX.deterministic_fill(sd, salt)
    module.load_state_dict(sd)
    return module


def _sd(shapes):
    return FX.deterministic_fill({k: torch.zeros(s, dtype=torch.long if "num_batches" in k else torch.float32) for k, s in shapes.items()})


# ---- (1) oracle vs the reference's golden vectors --------------------------------------------------------------------
@pytest.mark.parametrize("tag,nc,lsm", [("256", 256, -7.0), ("65536", 65536, math.log(1e-14))])
def test_oracle_dmol_loss_and_gradient(fx, inp, tag, nc, lsm):
    g = fx["dmol_" + tag]
    yh = inp["dmol_yhat"].clone().requires_grad_(True)
    nll = O.dmol_nll(yh, inp["dmol_y"], nc, lsm)
    assert H.relerr(nll, g["nll"]) < 1e-6
    assert abs(float(nll.sum()) - g["total"]) / abs(g["total"]) < 1e-6
    grad, = torch.autograd.grad(nll.sum(), yh)
    assert H.relerr(grad, g["grad"]) < 1e-5
    # the closed form the CUDA kernel evaluates
    assert H.relerr(O.dmol_nll_grad(inp["dmol_yhat"], inp["dmol_y"], nc, lsm), g["grad"]) < 1e-5


def test_oracle_masked_loss_ema_mask(fx, inp):
    g = fx["dmol_masked"]
    yh = inp["dmol_yhat"].clone().requires_grad_(True)
    loss = O.masked_dmol_loss(yh, inp["dmol_y"], lengths=inp["dmol_len"])
    assert abs(float(loss) - g["loss"]) / abs(g["loss"]) < 1e-6
    grad, = torch.autograd.grad(loss, yh)
    assert H.relerr(grad, g["grad"]) < 1e-5
    assert torch.equal(O.sequence_mask(inp["dmol_len"]), g["mask"])
    assert H.relerr(O.ema_update(inp["f1"], inp["f2"], 0.9), fx["ema"]) < 1e-6
    u = inp["dmol_u"].reshape(-1, 11)
    got = O.sample_dmol(inp["dmol_yhat"].transpose(1, 2).reshape(-1, 30), u[:, :10], u[:, 10], -7.0).reshape(2, 40)
    assert H.relerr(got, fx["dmol_sample"]) < 1e-6


@pytest.mark.parametrize("tag,margin,mv", [("m0", 0.0, False), ("m8", 8.0, False), ("m8max", 8.0, True)])
def test_oracle_contrastive(fx, inp, tag, margin, mv):
    g = fx["ctr_" + tag]
    a, b = inp["f1"].clone().requires_grad_(True), inp["f2"].clone().requires_grad_(True)
    loss = O.l2_contrastive(a, b, margin, mv)
    assert abs(float(loss) - g["loss"]) / g["loss"] < 1e-6
    ga, gb = torch.autograd.grad(loss, (a, b))
    assert H.relerr(ga, g["g1"]) < 1e-5 and H.relerr(gb, g["g2"]) < 1e-5


def test_oracle_sim_norm_retrieval(fx, inp):
    assert H.relerr(O.pairdist(inp["f1"], inp["f2"]), fx["l2_sim"]) < 1e-6
    assert H.relerr(O.l2_normalize(inp["f1"]), fx["l2_norm"]) < 1e-6
    assert O.l2_retrieval(inp["f1"], inp["f2"]) == pytest.approx(fx["retrieval"])
    assert O.l2_retrieval(inp["f1"], inp["f3"]) == pytest.approx(fx["retrieval_noisy"])


def test_oracle_discriminator_heads(fx, inp):
    g = fx["inpainting_dis"]
    sd = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in _sd(g["shapes"]).items()}
    out = O.inpainting_dis_forward(sd, inp["dis_mel"], inp["dis_fea"], training=True)
    assert H.relerr(out, g["out"]) < 1e-5
    out.pow(2).sum().backward()
    for k, want in g["grads"].items():
        assert H.relerr(sd[k].grad, want) < 1e-4, k
    for k, want in g["rm"].items():
        assert H.relerr(sd[k], want) < 1e-5, k
    g = fx["domain_dis"]
    assert H.relerr(O.domain_dis_forward(_sd(g["shapes"]), inp["dom_x"]), g["out"]) < 1e-5


def test_oracle_visual_branches(fx, inp):
    g = fx["ie2"]
    out, fea = O.image_embedding2_forward(_sd(g["shapes"]), inp["video"], inp["flow"])
    assert H.relerr(out, g["out"]) < 1e-4 and H.relerr(fea, g["fea"]) < 1e-4
    g = fx["ie_single"]
    sd = _sd(g["shapes"])
    assert H.relerr(O.image_embedding_single_forward(sd, inp["video"]), g["out"]) < 1e-4
    assert H.relerr(sd["bn_1.running_mean"], g["bn_1_running_mean"]) < 1e-4
    g = fx["ie_finetune"]
    assert H.relerr(O.image_embedding_finetune_forward(_sd(g["shapes"]), inp["feat"]), g["out"]) < 1e-5


def _wavenet_train_inputs(fx):
    kw = MG.WAVENET_TRAIN_KW
    T = fx["wavenet_train"]["T"]
    x = torch.cat((FX.uniform("wav_xsmall", (1, 1, T), -1.0, 1.0), FX.uniform("wav_x2", (1, 1, T), -1.0, 1.0)), 0)
    c = torch.cat((FX.uniform("wav_csmall", (1, kw["cin_channels"], T // 8)), FX.uniform("wav_c2", (1, kw["cin_channels"], T // 8))), 0)
    return kw, T, x, c


def test_oracle_wavenet_training_step(fx):
    g = fx["wavenet_train"]
    kw, T, x, c = _wavenet_train_inputs(fx)
    sd = {k: v.requires_grad_(True) for k, v in H.filled(H.wavenet_sd(**kw)).items()}
    y_hat = O.wavenet_forward(sd, x, c, kw["layers"] // kw["stacks"], kw["upsample_scales"])
    assert H.relerr(y_hat, g["y_hat"]) < 1e-5
    loss = O.masked_dmol_loss(y_hat[:, :, :-1], x.transpose(1, 2)[:, 1:, :], lengths=g["lengths"])
    assert abs(float(loss) - g["loss"]) / g["loss"] < 1e-5
    loss.backward()
    for k, want in g["grads"].items():
        assert H.relerr(sd[k].grad, want) < 1e-3, k           # fp32 summation order (north star: 1e-3 rel)
    for k in g["dead"]:
        assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0


# ---- (2) product host logic, dry-run on the CPU with torch stand-ins for the CUDA ops -------------------------------------
@pytest.fixture()
def cpu_ops():
    with shim.installed() as ops:
        yield ops


def test_product_wavenet_forward_and_training_glue(fx, cpu_ops):
    from viai_b200.wavenet_vocoder import WaveNet
    from viai_b200 import loss_functions as LF
    g = fx["wavenet_train"]
    kw, T, x, c = _wavenet_train_inputs(fx)
    m = _filled(WaveNet(**kw)).eval()
    y_hat = m(x, c)                # the T-parallel forward is pure glue over ops.*: that glue is what is dry-run here
    assert tuple(y_hat.shape) == tuple(g["y_hat"].shape)
    assert H.relerr(y_hat, g["y_hat"]) < 1e-5
    loss = LF.DiscretizedMixturelogisticLoss()(y_hat[:, :, :-1], x.transpose(1, 2)[:, 1:, :], lengths=g["lengths"])
    assert abs(float(loss) - g["loss"]) / g["loss"] < 1e-5
    loss.backward()
    params = dict(m.named_parameters())
    for k, want in g["grads"].items():
        assert H.relerr(params[k].grad, want) < 1e-3, k
    for k, want in g["gnorm"].items():
        assert abs(float(params[k].grad.norm()) - want) <= 1e-3 * max(want, 1e-3), k
    for k in g["dead"]:
        assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0


def test_product_wavenet_forward_refuses_cpu_tensors():
    from viai_b200.wavenet_vocoder import WaveNet
    m = WaveNet(**MG.WAVENET_TRAIN_KW)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 1, 16), torch.zeros(1, 8, 2))
    from viai_b200 import ops
    for call in (lambda: ops.shiftcat(torch.zeros(1, 4, 4), None, 3, 1, 32), lambda: ops.glu_tanh_sigmoid(torch.zeros(2, 8)),
                 lambda: ops.dmol_nll(torch.zeros(2, 30), torch.zeros(2)), lambda: ops.pairdist(torch.zeros(2, 4), torch.zeros(2, 4)),
                 lambda: ops.l2_normalize(torch.zeros(2, 4)), lambda: ops.sequence_mask(torch.tensor([1, 2]), 3)):
        with pytest.raises(RuntimeError, match="no CPU path"):
            call()


def test_product_residual_layer_reference_layout(cpu_ops):
    """ResidualConv1dGLU.forward in the reference's (B, C, T) layout == the oracle's single layer."""
    from viai_b200.wavenet_vocoder.modules import ResidualConv1dGLU
    import torch.nn.functional as F
    torch.manual_seed(0)
    lay = ResidualConv1dGLU(16, 32, 3, skip_out_channels=8, cin_channels=4, dilation=4).eval()
    x, c = torch.randn(2, 16, 21), torch.randn(2, 4, 21)
    xo, s = lay(x, c)
    sd = O.fold_weight_norm({k: v.detach() for k, v in lay.state_dict().items()})
    z = F.conv1d(x, sd["conv.weight"], sd["conv.bias"], padding=8, dilation=4)[:, :, :21] + F.conv1d(c, sd["conv1x1c.weight"], sd["conv1x1c.bias"])
    a, b = z.split(16, dim=1)
    gte = torch.tanh(a) * torch.sigmoid(b)
    assert H.relerr(s, F.conv1d(gte, sd["conv1x1_skip.weight"], sd["conv1x1_skip.bias"])) < 1e-5
    assert H.relerr(xo, (F.conv1d(gte, sd["conv1x1_out.weight"], sd["conv1x1_out.bias"]) + x) * math.sqrt(0.5)) < 1e-5
    lay.train()
    torch.manual_seed(1)
    xo2, _ = lay(x, c)                      # dropout active: different from the eval output, same shape
    assert xo2.shape == xo.shape and not torch.allclose(xo2, xo)


def test_product_losses_and_metrics(fx, inp, cpu_ops):
    from viai_b200 import loss_functions as LF
    from viai_b200.utils import util
    from viai_b200.wavenet_vocoder.mixture import discretized_mix_logistic_loss
    g = fx["dmol_256"]
    assert H.relerr(discretized_mix_logistic_loss(inp["dmol_yhat"], inp["dmol_y"], 256, -7.0, reduce=False), g["nll"]) < 1e-6
    assert abs(float(discretized_mix_logistic_loss(inp["dmol_yhat"], inp["dmol_y"], 256, -7.0)) - g["total"]) / abs(g["total"]) < 1e-6
    assert abs(float(LF.DiscretizedMixturelogisticLoss()(inp["dmol_yhat"], inp["dmol_y"], lengths=inp["dmol_len"])) -
               fx["dmol_masked"]["loss"]) / fx["dmol_masked"]["loss"] < 1e-6
    from viai_b200.wavenet_vocoder.mixture import sample_from_discretized_mix_logistic
    smp = sample_from_discretized_mix_logistic(inp["dmol_yhat"], -7.0, uniforms=inp["dmol_u"])
    assert tuple(smp.shape) == (2, 40) and H.relerr(smp, fx["dmol_sample"]) < 1e-6
    with pytest.raises(RuntimeError, match="either lengths or mask"):
        LF.DiscretizedMixturelogisticLoss()(inp["dmol_yhat"], inp["dmol_y"])
    assert torch.equal(LF.sequence_mask(inp["dmol_len"]), fx["dmol_masked"]["mask"])
    ema = LF.ExponentialMovingAverage(0.9)
    ema.register("w", inp["f1"])
    ema.update("w", inp["f2"])
    assert H.relerr(ema.shadow["w"], fx["ema"]) < 1e-6
    for tag, margin, mv in (("m0", 0, False), ("m8", 8.0, False), ("m8max", 8.0, True)):
        assert abs(float(LF.L2ContrastiveLoss(margin=margin, max_violation=mv)(inp["f1"], inp["f2"])) - fx["ctr_" + tag]["loss"]) < 1e-5
    assert H.relerr(LF.l2_sim(inp["f1"], inp["f2"]), fx["l2_sim"]) < 1e-6
    assert H.relerr(util.l2_norm(inp["f1"]), fx["l2_norm"]) < 1e-6
    assert util.L2retrieval(inp["f1"], inp["f3"]) == pytest.approx(fx["retrieval_noisy"])
    m, (ranks, top1) = util.L2retrieval(inp["f1"], inp["f2"], return_ranks=True)
    assert m == pytest.approx(fx["retrieval"]) and ranks.shape == (6,) and top1.shape == (6,)


def test_product_discriminator_heads(fx, inp, cpu_ops):
    from viai_b200.networks.Discriminator_Networks import DomainDis, Inpainting_Dis
    g = fx["inpainting_dis"]
    D = _filled(Inpainting_Dis())
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == g["shapes"]
    out = D(inp["dis_mel"], inp["dis_fea"])
    assert tuple(out.shape) == tuple(g["out"].shape) and H.relerr(out, g["out"]) < 1e-5
    out.pow(2).sum().backward()
    params = dict(D.named_parameters())
    for k, want in g["grads"].items():
        assert H.relerr(params[k].grad, want) < 1e-4, k
    for k, want in g["rm"].items():
        assert H.relerr(D.state_dict()[k], want) < 1e-5, k
    with pytest.raises(RuntimeError, match="80-bin"):
        D(torch.rand(1, 1, 96, 64), torch.randn(1, 512, 16))
    g = fx["domain_dis"]
    D = _filled(DomainDis())
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == g["shapes"]
    out = D(inp["dom_x"])
    assert tuple(out.shape) == (3, 1) and H.relerr(out, g["out"]) < 1e-5
    out.sum().backward()
    for k, want in g["gnorm"].items():
        assert abs(float(dict(D.named_parameters())[k].grad.norm()) - want) <= 1e-4 * max(want, 1e-3), k


def test_product_visual_branches(fx, inp, cpu_ops):
    from viai_b200.networks import Image_Embedding as IE
    g0 = H.load_golden("image_embedding.pt")
    M = _filled(IE.ImageEmbedding())
    out = M(inp["video"], inp["flow"])
    assert tuple(out.shape) == tuple(g0["out"].shape) and H.relerr(out, g0["out"]) < 1e-4
    assert H.relerr(M.state_dict()["bn_1.running_mean"], g0["bn_1_running_mean"]) < 1e-4
    assert H.relerr(M.state_dict()["image_single_model.bn1.running_mean"], g0["img_bn1_running_mean"]) < 1e-4
    g = fx["ie2"]
    M = _filled(IE.ImageEmbedding2())
    assert {k: tuple(v.shape) for k, v in M.state_dict().items()} == g["shapes"]
    out, fea = M(inp["video"], inp["flow"])
    assert tuple(out.shape) == tuple(g["out"].shape) and tuple(fea.shape) == tuple(g["fea"].shape)
    assert H.relerr(out, g["out"]) < 1e-4 and H.relerr(fea, g["fea"]) < 1e-4
    g = fx["ie_single"]
    M = _filled(IE.ImageEmbedding_single(image=1))
    assert {k: tuple(v.shape) for k, v in M.state_dict().items()} == g["shapes"]
    out = M(inp["video"])
    assert tuple(out.shape) == tuple(g["out"].shape) and H.relerr(out, g["out"]) < 1e-4
    assert H.relerr(M.state_dict()["bn_1.running_mean"], g["bn_1_running_mean"]) < 1e-4
    g = fx["ie_finetune"]
    M = _filled(IE.ImageEmbedding_finetune())
    assert {k: tuple(v.shape) for k, v in M.state_dict().items()} == g["shapes"]
    out = M(inp["feat"])
    assert tuple(out.shape) == tuple(g["out"].shape) and H.relerr(out, g["out"]) < 1e-5


def test_trainer_shifted_targets_equal_the_sliced_loss(fx, inp):
    """WaveNetTrainer aligns targets / weights with y_hat instead of slicing y_hat[:, :, :-1]: same loss (pure torch glue)."""
    from viai_b200.wavenet_step import WaveNetTrainer
    yh, y = inp["dmol_yhat"], inp["dmol_y"]
    mask = O.sequence_mask(inp["dmol_len"], 40).unsqueeze(-1)
    want = O.masked_dmol_loss(yh[:, :, :-1], y[:, 1:, :], mask=mask[:, 1:, :], num_classes=256, log_scale_min=-7.0)
    target, weight = WaveNetTrainer.shifted_targets(y, mask)
    nll = O.dmol_nll(yh, target.unsqueeze(-1), 256, -7.0).squeeze(-1)
    got = (nll * weight).sum() / weight.sum()
    assert abs(float(got) - float(want)) / float(want) < 1e-6
