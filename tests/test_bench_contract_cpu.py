"""bench.py's reference arm (`--impl reference`: the reference's own nn.Modules from the vendored copy oracle/_ref -- or the oracle
port where that copy is absent -- on the host cores) runs without a GPU and prints ONE JSON line with the keys the measurement
contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GAN-step spectrogram-frames/sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    from oracle import build_ref
    want_kind = "reference" if (build_ref.available() or os.path.isdir("/root/reference/networks")) else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["config"]["global_batch"] == 32                  # the same C2 batch as the GPU arm
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_vendoring_recipe_copies_the_reference_byte_for_byte():
    """oracle/build_ref.py: the files under oracle/_ref are the reference's, unmodified (build container only)."""
    import filecmp
    import pytest
    from oracle import build_ref
    if not os.path.isdir("/root/reference/networks"):
        pytest.skip("/root/reference not present")
    assert build_ref.build() == build_ref.DEST and build_ref.available()
    for rel in build_ref.FILES:
        assert filecmp.cmp(os.path.join("/root/reference", rel), os.path.join(build_ref.DEST, rel), shallow=False), rel
