"""The train_whole_sync.py step contract (/root/reference/train_whole_sync.py:49-112,159-167) driven on synthetic 8-tuples of
the loader's layout (Data_loaders/audio_loader.py:532): every method / attribute the script touches exists and behaves."""
import pytest
import torch

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
