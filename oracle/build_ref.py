"""Recipe that vendors the UNMODIFIED reference modules of the hot path into ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored:
it travels to the GPU box like a built .so, and never enters the repository's history).

TEST / BASELINE INFRASTRUCTURE ONLY.  ``bench.py --impl reference`` and the ``cpu_baseline`` leg time these files (the
reference's own ``nn.Module``s on the host cores, ``cpu_baseline.kind = "reference"``); when ``oracle/_ref`` is absent they fall
back to the oracle port (``kind = "port"``).  The reference is pure Python without a build system, so "building" it is a copy of
the files the path imports, byte for byte (checked), from where they lie under /root/reference:

    python -m oracle.build_ref            (also run by __graft_entry__.build() when /root/reference is present)
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("VIAI_REFERENCE_ROOT", "/root/reference")
FILES = [
    "networks/Inpainting_Networks.py", "networks/New_Inpainting_Networks.py", "networks/Discriminator_Networks.py",
    "networks/Image_Embedding.py", "networks/ResNet.py", "loss_functions.py",
    "wavenet_vocoder/__init__.py", "wavenet_vocoder/builder.py", "wavenet_vocoder/conv.py", "wavenet_vocoder/mixture.py",
    "wavenet_vocoder/modules.py", "wavenet_vocoder/util.py", "wavenet_vocoder/version.py", "wavenet_vocoder/wavenet.py",
    "wavenet_vocoder/wavenet_encoder.py", "utils/__init__.py", "utils/util.py", "utils/lrschedule.py",
]


def build(verbose=False):
    """Copies FILES from the reference tree; returns DEST, or None when the reference tree is not present."""
    if not os.path.isdir(os.path.join(SRC, "networks")) or os.path.realpath(SRC) == os.path.realpath(DEST):
        return None
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print("vendored", rel)
    return DEST


def available():
    return all(os.path.exists(os.path.join(DEST, rel)) for rel in FILES)


if __name__ == "__main__":
    print(build(verbose=True) or "reference tree not present; oracle/_ref left as is (%s)" % ("complete" if available() else "absent"))
