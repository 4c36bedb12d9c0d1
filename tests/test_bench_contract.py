"""bench.py's output contract (one JSON line on stdout with the driver's keys), checked on the CPU: the reference arm
runs for real on a tiny sample; the GPU arm's line is checked on the line committed with the round's profiles."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config"}


def test_reference_arm_prints_one_valid_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--triangles", "20000", "--ommatidia", "500", "--samples", "64", "--ref-rays-per-step", "20000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one line on stdout"
    j = json.loads(lines[0])
    assert BASE_KEYS <= set(j) and j["impl"] == "reference" and j["unit"] == "rays/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


def test_committed_gpu_bench_line_carries_every_contract_key():
    path = os.path.join(ROOT, "profiles", "r01f_bench_1gpu.json")
    j = json.load(open(path))
    assert BASE_KEYS | {"clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"} <= set(j)
    assert j["metric"] == "rays_per_sec" and j["n_gpus"] == 1 and j["scaling"] == "weak" and j["vs_baseline"] is None and j["dtype"] == "f32"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"]) and not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(j["e2e"]) and 0 < j["e2e"]["value"] < j["value"]
    assert j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0 and j["gpu_launches"] > 0
    r = j["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] and r["traffic"] > 0
    c = j["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port" and c["value"] > 0
    assert "workload" in j["config"] and "l2_policy" in j["config"]


def test_issue_roofline_from_the_committed_ncu_capture():
    """roofline.issue: warp instructions per ray from the ncu capture x live rays/s against SMs x 4 schedulers x clock.
    With the committed line's 20.8 Grays/s at 1965 MHz it lands a few percent under ncu's own issue-active figure
    (the step also contains the entry-frontier and ordered-sum launches)."""
    code = ("import sys, json; sys.argv=['bench.py']; import bench; "
            "tj=json.load(open('profiles/k1_traffic.json')); "
            "r=bench.issue_roofline(tj, 34, 10240000, 20.803e9, 1965.0); "
            "bad=[bench.issue_roofline({}, 34, 10240000, 1e9, 1965.0), bench.issue_roofline(tj, 8, 10240000, 1e9, 1965.0), "
            "bench.issue_roofline(tj, 34, 10240000, 1e9, None)]; "
            "sys.stderr.write('RESULT ' + json.dumps([r, bad]))")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    r, bad = json.loads(p.stderr.split("RESULT ", 1)[1])
    assert bad == [None, None, None]
    assert r["bound"] == "issue" and abs(r["peak"] - 148 * 4 * 1.965) < 1e-6 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 30 < r["warp_inst_per_ray"] < 40 and 0.5 < r["frac"] < r["ncu_issue_active_pct"] / 100
