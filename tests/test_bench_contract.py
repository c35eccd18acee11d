"""The bench contract, checked without a GPU: the reference arm runs live on the host cores (it is the CPU
implementation), and the committed B200 lines under profiles/ must carry every key the driver reads."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"]


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), key=lambda p: int(p.rsplit("_v", 1)[1].split(".")[0]))
    assert files, pattern
    return files[-1]


def test_reference_arm_prints_one_json_line(native_libs):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line, got %d" % len(lines)
    d = json.loads(lines[0])
    for k in BASE_KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["unit"] == "samples/s" and d["higher_is_better"] is True
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_committed_b200_line_has_the_contract_keys():
    d = json.load(open(_latest("r01_bench_c2cap_v*.json")))
    for k in BASE_KEYS + ["gpu_launches", "clocks", "roofline", "cpu_baseline"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] == d["steps"] > 0
    assert d["vs_baseline"] is None          # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "l2" in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    # the headline: end to end through the plugins against the reference's own CPU implementation on the same box
    assert e["value"] / cb["value"] > 50


def test_committed_scaling_lines_are_weak_and_linear():
    one = json.load(open(_latest("r01_bench_c2cap_v*.json")))
    for name, n in (("r01_bench_c2cap_n2_v*.json", 2), ("r01_bench_c2cap_n8_v*.json", 8)):
        d = json.load(open(_latest(name)))
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        assert d["config"]["streams_total"] == n * d["config"]["streams_per_gpu"]
        assert 0.95 < d["value"] / (n * one["value"]) < 1.05
