"""viai_b200 -- B200-native hot path of the Vision-Infused Audio Inpainter (VIAI).

The directory is named ``vision-infused-audio-inpainter-viai_b200``; import it as ``viai_b200`` (alias package at
the repository root)."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
