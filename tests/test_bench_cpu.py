"""The parts of bench.py that run without a GPU: the reference arm's JSON line (the contract the driver
parses) and the size cap of the line of record."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("batched 1D C2C forward N=4096")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_is_rank_0_only_under_torchrun():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_the_line_of_record_stays_small():
    """BENCH_r01 lost its headline because the JSON line outgrew what the driver captures: the line is capped."""
    sys.path.insert(0, ROOT)
    import bench
    line = {"metric": "m", "value": 1.23456789012, "also": {f"w{i}": {"ms": 1.0, "note": "x" * 300} for i in range(40)}}
    text = bench.format_line(line)
    assert len(text) <= bench.MAX_LINE_BYTES
    d = json.loads(text)
    assert d["value"] == 1.23457 and d["metric"] == "m" and d.get("also_truncated") is True
    small = bench.format_line({"metric": "m", "value": 2.0, "also": {"a": {"ms": 1.0}}})
    assert json.loads(small)["also"] == {"a": {"ms": 1.0}}
    assert "run_optin_probes" not in dir(bench)  # measurement side jobs live in tools/, never in the run of record
