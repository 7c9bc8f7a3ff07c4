"""Import alias: ``viai_b200`` -> ``vision-infused-audio-inpainter-viai_b200/`` (a hyphenated directory name cannot be
imported directly)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "vision-infused-audio-inpainter-viai_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
