"""CPU-side checks: the C-ABI library builds, loads and exports every symbol declared in include/viai_b200.h; the
product modules expose the reference's constructor signatures and state_dict layouts; host logic that needs no GPU."""
import inspect
import os
import re

import pytest
import torch
import torch.nn as nn

import viai_test_helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from viai_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "viai_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(viai_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libviai_b200.so does not export %s" % name
    bound = set(lib.SIGNATURES) | set(lib.VALUE_FUNCS) | {"viai_last_error", "viai_version", "viai_launch_count"}
    assert declared == bound, declared ^ bound
    assert L.viai_version() >= 100


def test_host_side_selectors_need_no_gpu(lib):
    """Geometry / tuning selectors of the C ABI are plain host functions: which Cin = 1 layers get their BatchNorm statistics from
    the convolution kernel, and the size threshold of the alternating normalisation sweeps (set, query, restore)."""
    import ctypes
    from viai_b200 import ops
    L = lib.lib()
    sup = lambda *a: L.viai_conv2d_thin_stats_supported(ctypes.byref(ops._geom(*a)))
    if os.environ.get("VIAI_CIN1_STATS", "1") != "0":
        assert sup(32, 256, 256, 1, 256, 128, 64, 1, 4, (1, 2), (0, 1), 0) == 1      # MelDiscriminator.conv1 at C2
        assert sup(32, 256, 256, 1, 128, 128, 32, 3, 3, (2, 2), (1, 1), 0) == 1      # MelEncoder.conv1 at C2
    assert sup(2, 16, 16, 1, 16, 16, 128, 3, 3, (1, 1), (1, 1), 0) == 0              # more than 64 output channels
    assert sup(2, 16, 16, 1, 16, 16, 32, 5, 5, (1, 1), (2, 2), 0) == 0               # other filter shapes: generic kernel
    assert sup(2, 16, 16, 32, 16, 16, 1, 3, 3, (1, 1), (1, 1), 0) == 0               # Cout == 1 is the other thin kernel
    assert sup(2, 16, 16, 32, 16, 16, 32, 3, 3, (1, 1), (1, 1), 0) == 0              # tensor-core layer
    prev = L.viai_norm_walk_mb(-1)
    assert prev >= 0
    try:
        assert L.viai_norm_walk_mb(0) == prev and L.viai_norm_walk_mb(-1) == 0
        assert L.viai_norm_walk_mb(512) == 0 and L.viai_norm_walk_mb(-5) == 512
    finally:
        L.viai_norm_walk_mb(prev)
    assert L.viai_norm_walk_mb(-1) == prev


def test_no_cpu_fallback(lib):
    from viai_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.conv2d(torch.zeros(1, 4, 4, 2), torch.zeros(3, 2, 3, 3))
    from viai_b200.networks.Inpainting_Networks import MelEncoder
    with pytest.raises(RuntimeError):
        MelEncoder()(torch.rand(1, 80, 64))


@pytest.mark.parametrize("norm", ["bn", "in"])
def test_state_dict_layouts_match_reference(lib, norm):
    from viai_b200.networks import Discriminator_Networks as DN, Inpainting_Networks as IN, New_Inpainting_Networks as NN
    nl = nn.BatchNorm2d if norm == "bn" else nn.InstanceNorm2d
    pairs = [(IN.MelEncoder(norm_layer=nl), H.encoder_sd(norm)), (DN.MelDiscriminator(norm_layer=nl), H.discriminator_sd(norm))]
    for v in ("MelDecoder", "MelDecoderImage", "MelDecoderImage2", "MelDecoder_old"):
        pairs.append((getattr(NN, v)(norm_layer=nl), H.decoder_sd(norm, v)))
    for m, want in pairs:
        got = {k: tuple(t.shape) for k, t in m.state_dict().items()}
        assert got == {k: tuple(t.shape) for k, t in want.items()}, type(m).__name__
    n = lambda m: sum(p.numel() for p in m.parameters())
    if norm == "bn":     # SURVEY 8c KAT (7)
        assert (n(pairs[0][0]), n(pairs[2][0]), n(pairs[1][0])) == (978656, 3202497, 1555072)
    else:
        assert (n(pairs[0][0]), n(pairs[2][0]), n(pairs[1][0])) == (979392, 3200097, 1554113)


def test_constructor_signatures(lib):
    from viai_b200.networks import Discriminator_Networks as DN, Inpainting_Networks as IN, New_Inpainting_Networks as NN
    from viai_b200.loss_functions import GANLoss
    assert list(inspect.signature(IN.MelEncoder.__init__).parameters) == ["self", "hparams", "norm_layer"]
    assert list(inspect.signature(NN.MelDecoder.__init__).parameters) == ["self", "hparams", "norm_layer"]
    assert list(inspect.signature(NN.MelDecoderImage.forward).parameters) == ["self", "net", "x_size", "video_net"]
    assert list(inspect.signature(NN.TransConvBlock.__init__).parameters) == [
        "self", "inplanes", "outplanes", "name", "nums", "kernel_size", "padding", "stride", "norm_layer"]
    assert list(inspect.signature(DN.MelDiscriminator.__init__).parameters) == [
        "self", "input_nc", "ndf", "n_layers", "norm_layer", "use_sigmoid"]
    assert list(inspect.signature(GANLoss.__init__).parameters) == [
        "self", "use_lsgan", "device", "target_real_label", "target_fake_label"]
    with pytest.raises(Exception, match="name should be str"):
        NN.TransConvBlock(4, 4, 1)
    g = GANLoss()
    assert set(dict(g.named_buffers())) == {"real_label", "fake_label"}


def test_audio_model_exposes_the_train_script_contract():
    """Every AudioModel method train_whole_sync.py calls (/root/reference/train_whole_sync.py:49-112,159-167) is defined
    (constructing the model needs a GPU; the signature check does not)."""
    import inspect
    from viai_b200.Models.Whole_Sync_inpainting_modify import AudioModel
    want = {"get_blank_space_length": ["global_step"], "set_inputs": ["data"], "eval_model_test": ["global_step", "eval_dir"],
            "optimize_parameters": ["global_step"], "test": [], "get_loss_items": [], "get_current_visuals": [],
            "get_current_errors": [], "TF_writer": ["writer", "step"], "del_no_need": [],
            "save_inpainting_checkpoint": ["global_step", "global_test_step", "checkpoint_dir", "epoch", "hparams"],
            "load_inpainting_checkpoint": ["path", "reset_optimizer"], "load_part_checkpoint": ["path"]}
    for name, params in want.items():
        fn = getattr(AudioModel, name)
        assert list(inspect.signature(fn).parameters)[1:] == params, name
    assert list(inspect.signature(AudioModel.__init__).parameters)[1:] == ["hparams", "device"]


def test_lr_schedule_known_answers():
    """SURVEY 8c KAT (4) on the product module (no reference needed)."""
    import math
    from viai_b200.utils import lrschedule as P
    assert math.isclose(P.noam_learning_rate_decay(1e-3, 0), 5e-7, rel_tol=1e-3)
    assert math.isclose(P.noam_learning_rate_decay(1e-3, 1999), 1e-3, rel_tol=1e-9)
    assert math.isclose(P.step_learning_rate_decay(1e-3, 100000), 9.604e-4, rel_tol=1e-9)
    assert math.isclose(P.cyclic_cosine_annealing(1e-3, 1, 1000, 5), 1e-3, rel_tol=1e-12)


def test_wavenet_synthesis_kernel_selection_is_host_logic(lib):
    """viai_wavenet{,2,3}_num_ctas are pure host functions (shared-memory budget, one warp per output unit): the C4 network gets 128
    cooperating CTAs for every batch the kernels accept, a batch of 5 or a 2-layer network is refused (the host side then falls
    back to the next kernel), the small test network of tests/test_wavenet_gpu.py gets 16."""
    lib = lib.lib()
    for B in (1, 2, 3, 4):
        assert lib.viai_wavenet_num_ctas(512, 512, 256, 80, 3, 30, B) == 128
        assert lib.viai_wavenet2_num_ctas(24, 512, 512, 256, 80, 3, 30, B) == 128
        assert lib.viai_wavenet3_num_ctas(24, 512, 512, 256, 80, 3, 30, B) == 128
    assert lib.viai_wavenet3_num_ctas(24, 512, 512, 256, 80, 3, 30, 5) == 0 and lib.viai_wavenet2_num_ctas(24, 512, 512, 256, 80, 3, 30, 5) == 0
    assert lib.viai_wavenet3_num_ctas(2, 512, 512, 256, 80, 3, 30, 1) == 0           # the folded schedule needs >= 3 layers
    assert lib.viai_wavenet3_num_ctas(8, 32, 32, 32, 80, 3, 30, 3) == 16
    assert lib.viai_wavenet3_replicas() == 8
