"""``Config.Config`` -- audio / WaveNet hyper-parameters read by /root/reference/utils/audio.py:16-144 and
/root/reference/wavenet_vocoder/wavenet.py:99-108; missing from the reference, values per SURVEY.md Appendix A."""
import math


class Config(object):
    sample_rate = 16000
    fft_size = 1024
    hop_size = 160
    frame_shift_ms = None
    num_mels = 80
    fmin = 125
    fmax = 7600
    min_level_db = -100
    ref_level_db = 20
    allow_clipping_in_normalization = True
    silence_threshold = 2
    rescaling = True
    rescaling_max = 0.999
    out_channels = 10 * 3
    decode_layers = 24
    decode_stacks = 4
    residual_channels = 512
    gate_channels = 512
    skip_out_channels = 256
    kernel_size = 3
    dropout = 1 - 0.95
    cin_channels = 80
    gin_channels = -1
    n_speakers = None
    weight_normalization = True
    upsample_conditional_features = True
    upsample_scales = [4, 4, 10]
    freq_axis_kernel_size = 3
    quantize_channels = 65536
    log_scale_min = float(math.log(1e-14))
    input_type = "raw"

    def __init__(self, **overrides):
        for k, v in overrides.items():
            setattr(self, k, v)
