"""The JSON line `bench.py` prints is a contract with the driver (metric / config / roofline / cpu_baseline / e2e keys).  The
committed 1-GPU line of the round (profiles/r2_bench_C2_1gpu.json, produced on a B200) and a live run of the reference arm
(CPU, bounded sample) are checked against it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_committed_bench_line_carries_the_contract():
    d = _line(os.path.join(ROOT, "profiles", "r2_bench_C2_1gpu.json"))
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == "pose_point_loss_evals_per_sec" and "pose" in base["metric"] and d["unit"] == "pose*point/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["config"]["workload"].startswith("C2") and d["config"]["n_points"] == 1_000_000 and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is not None
    for key in ("roofline_score", "roofline_refine"):
        assert 0.0 < d[key]["frac"] < 2.0 and d[key]["issue_frac"] and d[key]["l1tex_frac"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 40_000_000 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
    assert abs(d["value"] - d["n_gpus"] * (1800 + 600) * 1e6 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_reference_arm_prints_the_same_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "pose_point_loss_evals_per_sec" and d["unit"] == "pose*point/s"
    assert d["config"]["workload"].startswith("C2") and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
