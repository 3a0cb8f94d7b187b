"""The bench.py JSON contract (CPU): the reference arm is run here on a tiny sample; the B200 arm's committed line
(profiles/r01_bench_n1.json, written by a real run on the GPU box) is checked for the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches")


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: the reference's algorithm (oracle port) on the host cores, same metric / unit /
    workload name as the B200 arm, `impl`, `cpu_baseline` describing this run, `e2e` equal to the line's value."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--K", "1024", "--ref-clips", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in BASE_KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "contrastive_step_clips_per_sec" and d["unit"] == "clips/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["value"] - 2 / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "clips/step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_committed_b200_line_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")) as f:
        d = json.load(f)
    for k in BASE_KEYS + ("clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "contrastive_step_clips_per_sec" and d["unit"] == "clips/s" and d["n_gpus"] == 1
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 1e8 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.01
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)
    assert not any(x in c["reasons"] for x in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"))
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == d["unit"] and cb["sample"]
