"""The bench line contract (task statement: `bench.py` JSON keys) checked on the committed end-of-round evidence in
profiles/ -- guards the format of what bench.py prints without needing a GPU."""
import glob
import json
import os

import pytest

from conftest import ROOT


def _last_line(path):
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def _latest(pattern):
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert paths, pattern
    return paths[-1]


def test_own_arm_line_has_every_contract_key():
    j = _last_line(_latest("r01_v*_bench_B4096.json"))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in j, k
    assert "48kHz mono" in j["metric"] and "frames/sec" in j["metric"] and j["unit"] == "frames/s"
    assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["data"].startswith("synthetic")
    assert "workload" in j["config"] and j["n_gpus"] == 1 and j["warmup"] >= 3
    assert j["gpu_launches"] > 0
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < j["value"]
    c = j["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not (set(j["clocks"]["reasons"]) & bad) and j["clocks"]["sm_mhz"] > 0.9 * j["clocks"]["sm_max_mhz"]


def test_reference_arm_line():
    j = _last_line(_latest("r01_v*_bench_reference_arm.json"))
    assert j["impl"] == "reference" and j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["e2e"]["value"] == j["value"] == j["cpu_baseline"]["value"]
    own = _last_line(_latest("r01_v*_bench_B4096.json"))
    assert j["metric"] == own["metric"] and j["unit"] == own["unit"]
    assert j["config"]["streams_per_gpu"] == own["config"]["streams_per_gpu"] and j["config"]["frames_per_step"] == own["config"]["frames_per_step"]


@pytest.mark.parametrize("n", [2, 8])
def test_multi_gpu_lines_scale(n):
    j = _last_line(_latest("r01_v*_bench_%dgpu.json" % n))
    one = _last_line(_latest("r01_v*_bench_B4096.json"))
    assert j["n_gpus"] == n and j["scaling"] == "weak"
    assert j["value"] > 0.9 * n * one["value"]          # independent stream shards: near-linear


def test_bench_labels_follow_baseline_json():
    """bench.py's metric label is the throughput clause of BASELINE.json's metric; both arms share metric and workload."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert base["metric"].startswith(b.METRIC)
    assert "batch=4096 independent mono streams" in b.workload_name(4096, 100) and "batch=4096 independent mono streams" in base["configs"][1]
