"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints exactly ONE JSON line with the
keys the driver reads; the CUDA arm's workload table names every BASELINE config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 3 and d["n_gpus"] == 1      # never fewer than 3 warm-up steps
    assert d["spread"]["steps"] == 2 and d["spread"]["min"] <= d["value"] <= d["spread"]["max"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_the_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_workload_table_names_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.WORKLOADS) == {"c2", "c3", "c4", "c5"}
    assert bench.WORKLOADS["c5"]["envs"] == 1 << 21 and bench.WORKLOADS["c5"]["positions"] == [-3, -2, -1, 0, 1, 2, 3]
    assert bench.WORKLOADS["c3"]["envs"] == 65_536 and bench.WORKLOADS["c3"]["windows"] == 64
    assert bench.WORKLOADS["c2"]["envs"] == 4096 and bench.WORKLOADS["c2"]["windows"] is None
    assert bench.WORKLOADS["c4"]["n_datasets"] == 32 and bench.WORKLOADS["c4"]["rows"] == 1_000_000
    assert bench.algorithmic_bytes(64) == (106, 3064) and bench.algorithmic_bytes(None) == (106, 40)
