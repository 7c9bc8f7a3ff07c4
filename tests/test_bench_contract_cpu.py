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


def test_launch_list_summariser_cuts_out_the_last_step(tmp_path):
    """scripts/summarize_launches.py --last-step KERNEL N: of an ncu launch list that holds several eager steps, only the launches
    after the previous step's last KERNEL launch up to this step's N-th one are aggregated (profiles/r02_final_launch_shares.csv)."""
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC",' \
          '"Section Name","Metric Name","Metric Unit","Metric Value"\n'
    one = [("conv(int)", 100.0), ("adam_kernel(float*)", 10.0), ("conv(int)", 200.0), ("apply(float)", 50.0), ("adam_kernel(float*)", 10.0)]
    rows = [("fill_init()", 7.0)] * 3 + one + one + one
    p = tmp_path / "l.csv"
    with open(p, "w") as f:
        f.write("==PROF== noise line\n" + hdr)
        for i, (k, us) in enumerate(rows):
            f.write('"%d","1","python","box","%s","1","7","(256,1,1)","(148,1,1)","0","10.0","Command line profiler metrics",'
                    '"gpu__time_duration.sum","us","%.1f"\n' % (i, k, us))
    run = lambda *a: subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), str(p)] + list(a),
                                    capture_output=True, text=True, timeout=60)
    allr = run()
    assert allr.returncode == 0 and allr.stdout.splitlines()[-1] == "TOTAL,18,1131.0,1.0"
    last = run("--last-step", "adam_kernel", "2")
    assert last.returncode == 0, last.stderr
    lines = last.stdout.splitlines()
    assert lines[0] == "kernel,launches,total_us,share" and lines[-1] == "TOTAL,5,370.0,1.0"
    assert lines[1].startswith("conv,2,300.0,")
