"""bench.py's `--impl reference` arm runs on the host cores, so its JSON contract can be pinned on a
box without a GPU (tiny vocabulary / batch so the test takes seconds)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--vocab",
                          "20011", "--cpu-batch", "256", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "samples/s" and j["higher_is_better"] is True
    assert j["steps"] == 2 and j["warmup"] == 1 and j["n_gpus"] == 1 and j["scaling"] == "weak"
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["vs_baseline"] is None
    assert j["dtype"] == "f32" and j["data"] == "synthetic" and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    e = j["e2e"]
    assert e["value"] == j["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_bench_never_reads_the_reference_checkout_at_run_time():
    src = open(os.path.join(ROOT, "bench.py")).read() + open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert "/root/reference" not in src
