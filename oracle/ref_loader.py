"""Import the UNMODIFIED reference modules: from /root/reference (this container) or, where that does not exist (the GPU
box), from the byte-identical copy ``oracle/build_ref.py`` vendored into ``oracle/_ref/``.

TEST / BASELINE INFRASTRUCTURE ONLY -- used by ``oracle/make_golden.py``, by the ``not gpu`` tests that pin the oracle
restatement against the real reference classes, and by the CPU-baseline legs of ``bench.py`` (``oracle/ref_step.py``).  Nothing
under the product package imports it; ``-m gpu`` tests, ``smoke()`` and bench.py's GPU arm never read /root/reference.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("VIAI_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "networks")) and os.path.isdir(os.path.join(_HERE, "_ref", "networks")):
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")
_SHIMS = os.path.join(_HERE, "shims")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks"))


def load(normlayer=None, cin_channels=None):
    """Returns a namespace of reference modules.  ``normlayer``/``cin_channels`` patch the shim config
    *before* import because the reference evaluates them as default arguments at import time
    (/root/reference/networks/New_Inpainting_Networks.py:49)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the reference imports two loader modules it never uses on this path (``from Data_loaders import mel_loader``): stand-ins
    # for them and for their parent package (the vendored copy has no Data_loaders directory at all)
    pkg = sys.modules.get("Data_loaders")
    if pkg is None or not hasattr(pkg, "__path__"):
        pkg = types.ModuleType("Data_loaders")
        pkg.__path__ = []
        sys.modules["Data_loaders"] = pkg
    for short in ("mel_loader", "AV_loader"):
        name = "Data_loaders." + short
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
        setattr(pkg, short, sys.modules[name])
    import Options_inpainting
    if normlayer is not None:
        Options_inpainting.Inpainting_Config.normlayer = normlayer
    if cin_channels is not None:
        Options_inpainting.Inpainting_Config.cin_channels = cin_channels
    ns = types.SimpleNamespace()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ns.Inpainting_Networks = importlib.import_module("networks.Inpainting_Networks")
        ns.New_Inpainting_Networks = importlib.import_module("networks.New_Inpainting_Networks")
        ns.Discriminator_Networks = importlib.import_module("networks.Discriminator_Networks")
        ns.loss_functions = importlib.import_module("loss_functions")
        ns.wavenet = importlib.import_module("wavenet_vocoder.wavenet")
        ns.mixture = importlib.import_module("wavenet_vocoder.mixture")
        ns.lrschedule = importlib.import_module("utils.lrschedule")
        try:
            ns.Image_Embedding = importlib.import_module("networks.Image_Embedding")
        except Exception as e:  # cv2 missing etc.
            ns.Image_Embedding = None
            ns.Image_Embedding_error = e
    ns.Options_inpainting = Options_inpainting
    return ns
