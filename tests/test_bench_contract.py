"""CPU test of the part of bench.py's contract that runs without a GPU: the `--impl reference` arm (the CPU restatement of
the reference's clustering step timed on a bounded, density-preserving sample) prints one JSON line with the keys the
driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--config", "C2", "--cpu-sample-reads", "20000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "reads_clustered_per_s" and j["unit"] == "reads/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1 and j["n_gpus"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] == 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "sample" in j["config"]


def test_density_preserving_sample_keeps_the_pair_tests_per_read():
    """The CPU sample shortens the genome with the read count (bench.cpu_port_run): pair tests per read stay those of the
    full table, where thinning the reads alone would lose the chance overlaps."""
    sys.path.insert(0, ROOT)
    import bench
    _, full, n_full = bench.cpu_port_run("C2", 100_000)
    _, samp, n_samp = bench.cpu_port_run("C2", 25_000)
    a, b = full["pair_tests"] / n_full, samp["pair_tests"] / n_samp
    assert abs(a - b) / a < 0.1, (a, b)
