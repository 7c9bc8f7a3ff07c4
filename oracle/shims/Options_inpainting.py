"""Stand-in for the reference's missing ``Options_inpainting`` module.

TEST INFRASTRUCTURE ONLY.  The reference imports ``Options_inpainting`` at module import time in
eight files (e.g. /root/reference/networks/Inpainting_Networks.py:10) but does not ship it.  The
attribute list and the values forced by the reference's own shape arithmetic are catalogued in
SURVEY.md Appendix A.
"""
import torch.nn as nn


class Inpainting_Config(object):
    cin_channels = 80            # forced: 80->40->20->10->5->3 -> AvgPool2d((3,1)) -> 1
    max_mel_lengths = 256
    normlayer = nn.BatchNorm2d   # or nn.InstanceNorm2d
    length_feature = 256         # forced by deconv1_1_1(512, ...)
    image_size = 224             # forced by AvgPool2d(7) -> fc(512)
    resnet_pretrain = False
    resnet_pretrain_path = ""
    sample_rate = 16000
    name = "viai_oracle"
    save_optimizer_state = True
    batch_size = 32
