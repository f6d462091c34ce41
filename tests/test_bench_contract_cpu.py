"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm prints exactly one JSON line with the keys the
driver reads, and the B200 arm refuses to run (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--batch", "1", "--clip", "32", "64", "64")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True and d["value"] > 0
    # the unmodified reference when tools/install_reference.py has placed it under baseline/_ref, else the oracle port
    kind = "reference" if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "models")) else "port"
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "CSN-152" in d["metric"] and "TubeR_CSN152_AVA21" in d["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_committed_bench_line_keeps_the_contract():
    """profiles/r2_bench_final.json (the line bench.py printed on the B200 at the round's last build) carries every key of the bench
    contract, and its derived numbers are consistent with each other."""
    import json
    with open(os.path.join(ROOT, "profiles", "r2_bench_final.json")) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "stages", "stage_ms", "kernels"):
        assert k in d, k
    assert "CSN-152" in d["metric"] and "CSN152_AVA21" in d["config"]["workload"] and d["vs_baseline"] is None and d["warmup"] >= 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert abs(d["value"] - d["config"]["global_batch"] / d["ms_per_step"] * 1e3) < 1e-6 * d["value"]
    assert d["gpu_launches"] == d["launches_per_step"] * d["steps"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == d["config"]["per_gpu_batch"] * 3 * 32 * 256 * 256 * 4 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= 1.05 * d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1 and r["traffic"] > 0
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    names = [s["stage"] for s in d["stages"]]
    assert names == list(d["stage_ms"].keys()) and len(names) == 10
    for s in d["stages"]:
        assert s["bound"] in ("hbm", "tensor") and 0 <= s["hbm_frac"] < 1 and 0 <= s["tensor_frac_issued"] < 1
        assert abs(s["tensor_frac_issued"] - 3 * s["tensor_frac"]) < 2e-4
