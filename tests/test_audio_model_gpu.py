"""The train_whole_sync.py step contract (/root/reference/train_whole_sync.py:49-112,159-167) driven on synthetic 8-tuples of
the loader's layout (Data_loaders/audio_loader.py:532): every method / attribute the script touches exists and behaves."""
import pytest
import torch

import viai_test_helpers as H

pytestmark = pytest.mark.gpu

CONTRACT_METHODS = ["get_blank_space_length", "set_inputs", "eval_model_test", "optimize_parameters", "test", "get_loss_items",
                    "get_current_visuals", "get_current_errors", "save_inpainting_checkpoint", "load_inpainting_checkpoint",
                    "load_part_checkpoint", "TF_writer", "del_no_need"]
CONTRACT_ATTRS = ["train", "blank_length", "update_wavenet", "reconstruct_loss_item", "EmbeddingL2_item", "loss_mel_L1_item",
                  "mel_net_norm", "video_net_norm", "current_lr", "Mel_Encoder", "Mel_Decoder", "netD", "optimizer_G", "optimizer_D"]


class _Writer(object):
    def __init__(self):
        self.rows = []

    def add_scalar(self, name, value, step):
        self.rows.append((name, float(value), step))


def _batch(B=2, W=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(B, 80, W, generator=g)
    return (torch.zeros(B, 1), torch.zeros(B, 1), c, torch.zeros(B, 1, W * 160), torch.zeros(B, W * 160, 1), None,
            torch.full((B,), W * 160, dtype=torch.long), ["clip%d" % i for i in range(B)])


def test_audio_model_runs_the_reference_step_contract(tmp_path):
    from viai_b200 import Options_inpainting as OI
    from viai_b200.Models.Whole_Sync_inpainting_modify import AudioModel
    hp = OI.Inpainting_Config(cin_channels=80)
    hp.max_mel_lengths, hp.name, hp.save_optimizer_state = 64, "viai_test", True
    torch.manual_seed(3)
    model = AudioModel(hp, device=torch.device("cuda"))
    for m in CONTRACT_METHODS:
        assert callable(getattr(model, m)), m
    for a in CONTRACT_ATTRS:
        assert hasattr(model, a), a
    writer = _Writer()
    global_step, l1 = 0, []
    for step in range(3):                                      # train phase (train_whole_sync.py:74-77)
        model.get_blank_space_length(global_step)
        model.set_inputs(_batch(seed=0))
        model.train = 1
        model.optimize_parameters(global_step)
        global_step += 1
        model.get_loss_items()
        errs = model.get_current_errors()
        assert set(errs) == {"loss_D", "loss_G_GAN", "loss_mel_L1"} and all(v == v for v in errs.values())
        model.TF_writer(writer, step=global_step)
        l1.append(model.loss_mel_L1_item)
        vis = model.get_current_visuals()
        assert vis["fake_mel"].shape == (80, 64, 3)
        model.del_no_need()
    assert model.blank_length == 32 and l1[-1] < l1[0]        # same batch three times: the L1 term goes down
    assert len(writer.rows) == 9
    # test phase (train_whole_sync.py:78-84)
    model.set_inputs(_batch(seed=1))
    model.train = 0
    with torch.no_grad():
        model.test()
    model.get_loss_items()
    assert tuple(model.mel_net_norm.shape) == (2, 256 * 1 * 4) and tuple(model.video_net_norm.shape) == (2, 1024)
    assert torch.allclose(model.mel_net_norm.norm(dim=1), torch.ones(2, device="cuda"), atol=1e-4)
    # checkpoints in the reference's format (utils/util.py:146-162) and resume
    path = model.save_inpainting_checkpoint(global_step, 1, str(tmp_path), 0, hparams=hp)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"Mel_Encoder", "Mel_Decoder", "netD", "optimizer_G", "optimizer_D", "global_step", "global_epoch", "global_test_step"}
    model2 = AudioModel(hp, device=torch.device("cuda"))
    assert model2.load_inpainting_checkpoint(path, reset_optimizer=False) == (3, 0, 1)
    model.eval_model_test(global_step, str(tmp_path))
    model2.set_inputs(_batch(seed=1)); model2.blank_length = model.blank_length; model2.set_inputs(_batch(seed=1))
    with torch.no_grad():
        model2.test()
    assert torch.allclose(model2.fake, model.fake, atol=1e-5)


def test_audio_model_update_wavenet_reports_the_reconstruction_loss():
    """hparams.update_wavenet: one teacher-forced WaveNet step per batch on (x_batch, y_batch, input_lengths), conditioned on the
    inpainted mel; train_whole_sync.py:105-107 accumulates ``reconstruct_loss_item``."""
    from viai_b200 import Options_inpainting as OI
    from viai_b200.Models.Whole_Sync_inpainting_modify import AudioModel
    hp = OI.Inpainting_Config(cin_channels=80)
    hp.max_mel_lengths, hp.update_wavenet = 64, True
    hp.wavenet_kwargs = dict(layers=4, stacks=2, residual_channels=32, gate_channels=32, skip_out_channels=16, dropout=0.0)
    torch.manual_seed(5)
    model = AudioModel(hp, device=torch.device("cuda"))
    assert model.update_wavenet and model.wavenet is not None
    B, W = 2, 64
    g = torch.Generator().manual_seed(1)
    wav = (torch.rand(B, W * 160, generator=g) * 2 - 1) * 0.5
    batch = (torch.zeros(B, 1), torch.zeros(B, 1), torch.rand(B, 80, W, generator=g), wav.unsqueeze(1), wav.unsqueeze(2), None,
             torch.tensor([W * 160, W * 160 - 999]), ["a", "b"])
    losses = []
    for step in range(4):
        model.get_blank_space_length(step)
        model.set_inputs(batch)
        model.optimize_parameters(step)
        model.get_loss_items()
        losses.append(model.reconstruct_loss_item)
        model.del_no_need()
    assert all(l == l and l > 0 for l in losses) and min(losses[1:]) < losses[0]      # Adam at 1e-3 on the same batch: it goes down
    assert model.loss_mel_L1_item > 0


@pytest.mark.bf16x3
def test_audio_model_embedding_l2_matches_oracle_contrastive_loss():
    """Vision-infused model, eval forward: EmbeddingL2_item == L2ContrastiveLoss(l2_norm(audio bottleneck), l2_norm(visual
    embedding)) (train_whole_sync.py:109), checked with the oracle on the embeddings the model exposes."""
    from oracle import viai_oracle as O
    from viai_b200 import Options_inpainting as OI
    from viai_b200.Models.Whole_Sync_inpainting_modify import AudioModel
    hp = OI.Inpainting_Config(cin_channels=80)
    hp.max_mel_lengths, hp.image, hp.embedding_margin = 64, True, 0.5
    torch.manual_seed(6)
    model = AudioModel(hp, device=torch.device("cuda"))
    B, W = 2, 64
    g = torch.Generator().manual_seed(2)
    batch = (torch.randn(B, W // 4, 3, 224, 224, generator=g).clamp(-1, 1), torch.randn(B, W // 4, 2, 224, 224, generator=g).clamp(-1, 1),
             torch.rand(B, 80, W, generator=g), torch.zeros(B, 1, 8), torch.zeros(B, 8, 1), None, torch.tensor([8, 8]), ["a", "b"])
    model.set_inputs(batch)
    with torch.no_grad():
        model.test()
    model.get_loss_items()
    assert model.mel_net_norm.shape == model.video_net_norm.shape == (B, 1024)
    want = float(O.l2_contrastive(model.mel_net_norm.cpu(), model.video_net_norm.cpu(), 0.5))
    assert want > 0 and abs(model.EmbeddingL2_item - want) / want < 1e-4


def _hp(W=64, **kw):
    from viai_b200 import Options_inpainting as OI
    hp = OI.Inpainting_Config(cin_channels=80)
    hp.max_mel_lengths = W
    for k, v in kw.items():
        setattr(hp, k, v)
    return hp


def _drive(model, steps, seeds):
    out = []
    for step, seed in zip(range(steps), seeds):
        model.get_blank_space_length(step)
        model.set_inputs(_batch(seed=seed))
        model.optimize_parameters(step)
        model.get_loss_items()
        out.append((model.loss_D_item, model.loss_G_GAN_item, model.loss_mel_L1_item))
        model.del_no_need()
    return out


def test_optimize_parameters_replays_a_cuda_graph_and_equals_the_eager_step():
    """The drop-in entry point gets the benched step: ``optimize_parameters`` captures the D + G update on the first batch
    (restoring every piece of state after the capture's warm-up) and replays it; three steps on three different batches give the
    losses and weights of the eager step, and a blank_length change does not re-capture."""
    from viai_b200.Models.Whole_Sync_inpainting_modify import AudioModel
    torch.manual_seed(11)
    a = AudioModel(_hp(cuda_graph=True, blank_length=32, blank_warmup_steps=2, blank_length_start=8), device=torch.device("cuda"))
    torch.manual_seed(11)
    b = AudioModel(_hp(cuda_graph=False, blank_length=32, blank_warmup_steps=2, blank_length_start=8), device=torch.device("cuda"))
    la, lb = _drive(a, 3, (0, 1, 2)), _drive(b, 3, (0, 1, 2))
    assert a._graph_key is not None and b._graph_key is None
    graphs = a.trainer._graphs
    for x, y in zip(la, lb):
        for u, v in zip(x, y):
            assert abs(u - v) <= 2e-3 * max(abs(v), 1e-6), (la, lb)          # atomics order differs run to run; Adam amplifies
    assert a.trainer._graphs is graphs                                     # the mask schedule moved, the capture did not
    for oa, ob in ((a.optimizer_G, b.optimizer_G), (a.optimizer_D, b.optimizer_D)):
        assert float(oa.step_dev) == float(ob.step_dev) == 3.0
        assert H.relerr_l2(oa.flat_param, ob.flat_param) < 1e-3
    assert int(a.netD.bn1.num_batches_tracked) == int(b.netD.bn1.num_batches_tracked) == 9
    # a new batch shape re-captures
    a.set_inputs((torch.zeros(3, 1), torch.zeros(3, 1), torch.rand(3, 80, 64), torch.zeros(3, 1, 8), torch.zeros(3, 8, 1), None,
                  torch.full((3,), 8, dtype=torch.long), ["x"] * 3))
    a.optimize_parameters(3)
    assert a.trainer._graphs is not graphs and float(a.optimizer_G.step_dev) == 4.0


def test_optimize_parameters_is_as_fast_as_the_captured_trainer():
    """20 steps through the train_whole_sync.py:49-112 call order (set_inputs from host tensors, optimize_parameters,
    get_loss_items with its device->host read) within 1.25x of bare GanTrainer.replay() on the same shapes."""
    import time
    from viai_b200.Models.Whole_Sync_inpainting_modify import AudioModel
    from viai_b200.step import GanTrainer
    B, W = 8, 256
    hp = _hp(W)
    model = AudioModel(hp, device=torch.device("cuda"))
    g = torch.Generator().manual_seed(0)
    c = torch.rand(B, 80, W, generator=g).pin_memory()
    batch = (torch.zeros(B, 1), torch.zeros(B, 1), c, torch.zeros(B, 1, 8), torch.zeros(B, 8, 1), None, torch.full((B,), 8, dtype=torch.long), ["x"] * B)

    def loop(n):
        for step in range(n):
            model.get_blank_space_length(step)
            model.set_inputs(batch)
            model.optimize_parameters(step)
            model.get_loss_items()
            model.del_no_need()
    loop(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(20); torch.cuda.synchronize(); t_model = (time.perf_counter() - t0) / 20
    tr = GanTrainer(hp, "cuda")
    tr.capture(model.mel, model.mask, warmup=1)
    for _ in range(3):
        tr.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        tr.replay()
    torch.cuda.synchronize()
    t_bare = (time.perf_counter() - t0) / 20
    print("AudioModel loop %.3f ms/step, bare replay %.3f ms/step" % (t_model * 1e3, t_bare * 1e3))
    assert t_model <= 1.25 * t_bare + 2e-4


def test_lr_change_reaches_the_captured_step():
    """FusedAdam's learning rate lives in a device scalar read by the captured Adam kernel; replay() refreshes it from
    param_groups (LR schedules, load_state_dict after capture)."""
    from viai_b200.Models.Whole_Sync_inpainting_modify import AudioModel
    torch.manual_seed(5)
    m = AudioModel(_hp(), device=torch.device("cuda"))
    _drive(m, 1, (0,))
    w0 = m.optimizer_D.flat_param.clone()
    for opt in (m.optimizer_G, m.optimizer_D):
        opt.param_groups[0]["lr"] = 0.0
    _drive(m, 1, (1,))
    assert torch.equal(m.optimizer_D.flat_param, w0)                        # lr 0: the update is a no-op
    for opt in (m.optimizer_G, m.optimizer_D):
        opt.param_groups[0]["lr"] = 2e-4
    _drive(m, 1, (2,))
    assert not torch.equal(m.optimizer_D.flat_param, w0)


def test_evaluation_forward_leaves_the_batchnorm_buffers_alone():
    from viai_b200.Models.Whole_Sync_inpainting_modify import AudioModel
    torch.manual_seed(6)
    m = AudioModel(_hp(), device=torch.device("cuda"))
    _drive(m, 1, (0,))
    before = {k: v.clone() for mod in (m.Mel_Encoder, m.Mel_Decoder) for k, v in mod.state_dict().items() if "running" in k or "tracked" in k}
    m.set_inputs(_batch(seed=3))
    with torch.no_grad():
        m.test()
    after = {k: v for mod in (m.Mel_Encoder, m.Mel_Decoder) for k, v in mod.state_dict().items() if "running" in k or "tracked" in k}
    assert before and all(torch.equal(before[k], after[k]) for k in before)
