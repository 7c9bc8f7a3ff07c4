"""B200-native WaveNet vocoder synthesis.  Mirrors the public surface of /root/reference/wavenet_vocoder (r9y9
0.0.5+2092a64): ``WaveNet``, ``receptive_field_size``, ``ResidualConv1dGLU``, ``builder.wavenet``."""
from .version import __version__  # noqa: F401
from .wavenet import WaveNet, receptive_field_size  # noqa: F401
from .modules import ResidualConv1dGLU  # noqa: F401
