"""``Options_inpainting.Inpainting_Config`` -- the reference imports this module in eight files
(e.g. /root/reference/networks/Inpainting_Networks.py:3,10) but does not ship it.  This is the product-side
configuration object with the attributes the hot path reads; users may pass their own
object with the same attributes as ``hparams=`` exactly as with the reference."""
import torch.nn as nn


class Inpainting_Config(object):
    cin_channels = 80             # mel bins; forced by 80->40->20->10->5->3 -> AvgPool2d((3,1)) -> 1
    max_mel_lengths = 256
    normlayer = nn.BatchNorm2d    # or nn.InstanceNorm2d
    length_feature = 256
    image_size = 224
    resnet_pretrain = False
    resnet_pretrain_path = ""
    sample_rate = 16000
    hop_size = 160
    batch_size = 32
    name = "viai_b200"
    save_optimizer_state = True
    # step glue (the reference's AudioModel file is missing; pix2pix defaults)
    lr = 2e-4
    beta1 = 0.5
    lambda_L1 = 100.0
    use_lsgan = True
    blank_ratio = 0.5             # fraction of time frames blanked by the time-band mask
    checkpoint_dir = "checkpoints"
    # loader (Data_loaders/audio_loader.py; SURVEY.md Appendix A "Loader" row -- free parameters, upstream-style defaults)
    new_split_name = "_new_split.txt"
    load_num = 1
    image_hope_size = 1           # video frames per step of the window (25 fps: 0.04 s each)
    image_rescal_size = 256       # training frames are resized to this, then randomly cropped to image_size
    max_time_steps = 40960        # samples per clip window (= 64 frames * 4 mel frames * hop 160)
    image = False
    flow = False

    def __init__(self, **overrides):
        for k, v in overrides.items():
            setattr(self, k, v)
