"""The parts of bench.py that run without a GPU: the reference arm's JSON line (the contract the driver
parses) and the probe runner's handling of children that fail or never end."""
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


def test_probe_runner_survives_failing_and_hanging_children(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    real = subprocess.Popen

    def fake(cmd, **kw):
        what = cmd[-1]
        if "fastbench.py" in cmd[1]:
            what = "variant"
        code = {"latency": "import json; print(json.dumps({'what': 'single_call_latency', 'us_per_call': 7.5, 'ok': True, 'junk': 1}))",
                "early": "raise SystemExit(3)",
                "variant": "print('{\"workload\": \"c2c4096\", \"ms\": 2.7, \"ok\": true, \"label\": \"x\"}')",
                "real2d": "import time; print('{\"what\": \"r2c2d\", \"ms\": 1.0, \"ok\": true}', flush=True); time.sleep(60)"}.get(what, "print('{}')")
        return real([sys.executable, "-c", code], **{k: v for k, v in kw.items() if k != "env"})

    monkeypatch.setattr(bench.subprocess, "Popen", fake)
    out = bench.run_optin_probes(seconds_each=2)
    assert out[0] == {"what": "single_call_latency", "us_per_call": 7.5, "ok": True}   # unknown keys dropped
    assert out[1]["what"] == "early" and out[1]["error"].startswith("exit 3")           # a clean failure: carry on
    variants = [d for d in out if d.get("workload") == "c2c4096"]                       # base + every prebuilt variant (if any travelled)
    assert len(variants) in (0, 2, 3, 4, 5) and all(d["label"] == "x" for d in variants)
    assert out[-1]["what"] == "real2d" and "killed" in out[-1]["error"] and out[-1]["lines"] == [{"what": "r2c2d", "ms": 1.0, "ok": True}]
    assert not any(d.get("what") == "cluster" for d in out)                             # nothing is started after a kill
