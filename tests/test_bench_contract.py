"""bench.py's reference arm runs on the host cores, so its JSON line - and with it the key contract the driver reads
from both arms - can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "Gtexel-lights/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_baseline_metric_and_configs_are_the_ones_benched():
    sys.path.insert(0, ROOT)
    import bench

    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.CONFIGS["c2"]["workload"] == base["configs"][1]
    assert bench.CONFIGS["c3"]["workload"] == base["configs"][2]
    assert bench.C4["workload"] == base["configs"][3]
    assert (bench.CONFIGS["c2"]["B"], bench.CONFIGS["c2"]["H"], bench.CONFIGS["c2"]["L"]) == (64, 1024, 1)
    assert (bench.CONFIGS["c3"]["B"], bench.CONFIGS["c3"]["H"], bench.CONFIGS["c3"]["L"]) == (16, 2048, 16)
    assert (bench.CONFIGS["c5"]["B"], bench.CONFIGS["c5"]["H"], bench.CONFIGS["c5"]["L"]) == (512, 512, 8)
