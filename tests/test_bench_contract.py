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
    # the other shading configs of BASELINE.json ride along as bounded samples on the same host cores
    assert set(d["configs"]) == {"c1", "c3", "c5"}
    for k, c in d["configs"].items():
        assert c["value"] > 0 and c["kind"] == "port" and c["cores"] >= 1 and c["sample"] and c["workload"], k
    assert d["configs"]["c1"]["fwd_us"] > 0 and d["configs"]["c1"]["fwd_bwd_us"] >= d["configs"]["c1"]["fwd_us"] * 0.5


def test_reference_arm_times_the_conversion_and_blend_pipeline():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c4", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][0])
    assert d["impl"] == "reference" and d["unit"] == "Gtexel/s" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert "2048x2048" in d["config"]["sample"]


def test_traffic_capture_is_dropped_when_the_kernel_sources_changed(tmp_path, monkeypatch):
    """roofline.traffic comes from a committed ncu capture; it is only reported while profiles/traffic.json carries the hash
    of the very kernel sources the library is built from."""
    sys.path.insert(0, ROOT)
    import bench

    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "pypbr_b200" / "csrc")
    os.makedirs(tmp_path / "include")
    (tmp_path / "pypbr_b200" / "csrc" / "k.cu").write_text("__global__ void k() {}")
    (tmp_path / "include" / "pbrcuda.h").write_text("/* abi */")
    h = bench.csrc_sha16()
    (prof / "traffic.json").write_text(json.dumps({"_csrc_sha16": h, "c2:backward": {"bytes": 5.0e9}}))
    assert bench.measured_traffic("c2:backward") == {"bytes": 5.0e9}
    (tmp_path / "pypbr_b200" / "csrc" / "k.cu").write_text("__global__ void k() { }")
    assert bench.measured_traffic("c2:backward") is None


def test_baseline_metric_and_configs_are_the_ones_benched():
    sys.path.insert(0, ROOT)
    import bench

    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.C1["workload"] == base["configs"][0]
    assert bench.CONFIGS["c2"]["workload"] == base["configs"][1]
    assert bench.CONFIGS["c3"]["workload"] == base["configs"][2]
    assert bench.C4["workload"] == base["configs"][3]
    assert (bench.CONFIGS["c2"]["B"], bench.CONFIGS["c2"]["H"], bench.CONFIGS["c2"]["L"]) == (64, 1024, 1)
    assert (bench.CONFIGS["c3"]["B"], bench.CONFIGS["c3"]["H"], bench.CONFIGS["c3"]["L"]) == (16, 2048, 16)
    assert (bench.CONFIGS["c5"]["B"], bench.CONFIGS["c5"]["H"], bench.CONFIGS["c5"]["L"]) == (512, 512, 8)
