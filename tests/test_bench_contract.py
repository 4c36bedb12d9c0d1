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


def _check_gpu_line(j):
    assert BASE_KEYS | {"clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"} <= set(j)
    assert j["metric"] == "rays_per_sec" and j["n_gpus"] == 1 and j["scaling"] == "weak" and j["vs_baseline"] is None and j["dtype"] == "f32"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"]) and not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(j["e2e"]) and 0 < j["e2e"]["value"] < j["value"]
    assert j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0 and j["gpu_launches"] > 0
    c = j["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port" and c["value"] > 0
    assert "workload" in j["config"] and "l2_policy" in j["config"]


def test_committed_round1_gpu_bench_line():
    j = json.load(open(os.path.join(ROOT, "profiles", "r01f_bench_1gpu.json")))
    _check_gpu_line(j)
    r = j["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] and r["traffic"] > 0


def test_committed_round2_gpu_bench_line_carries_its_own_evidence():
    """whatever --steps the driver picks, the line must carry non-null traffic / dram_frac / issue;
    the binding roof (warp-instruction issue) comes first, the SURVEY 8(d) HBM line -- strictly its formula -- second;
    the value is the median of repeated batches; the CPU baseline has a single-thread and an all-core figure."""
    path = os.path.join(ROOT, "profiles", "r02_bench_1gpu.json")
    j = json.load(open(path))
    _check_gpu_line(j)
    r = j["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "hbm", "issue"} <= set(r) and r["bound"] == "issue"
    assert r["traffic"] and r["traffic"] > 0 and 0.3 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    h = r["hbm"]
    assert h["bound"] == "hbm" and h["unit"] == "GB/s" and abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-9
    F, S = r["frames_per_launch"], j["config"]["samples_per_ommatidium"]
    assert abs(h["bytes_per_ray"] - (64.0 * r["nodes_per_ray"] + 48.0 * r["tris_per_ray"] + 64.0 / F + 48.0 / S)) < 1e-6
    assert h["dram_frac"] is not None and 0 < h["dram_frac"] < 0.2                   # the kernel is not HBM-bound
    assert j["repeats"] >= 5 and len(j["batch_ms"]) == j["repeats"]
    med = sorted(j["batch_ms"])[len(j["batch_ms"]) // 2]
    assert abs(j["ms_per_step"] - med / j["steps"]) < 1e-9
    c = j["cpu_baseline"]
    assert c["single_thread"]["cores"] == 1 and 0 < c["single_thread"]["value"] < c["value"] and c["repetitions"] >= 3
    assert c["same_bits_as_checker_build"] is True and c["cores"] >= 2
    assert set(j["modes"]) >= {"ordered", "fused", "fused_fast"}


def test_ncu_model_reconstructs_any_frames_per_launch():
    """bench.ncu_model: DRAM bytes and warp instructions per launch are affine in the frames per launch; the committed
    captures (profiles/k1_traffic.json) give them back exactly at the captured F and a line in between / beyond."""
    code = ("import sys, json; sys.argv=['bench.py']; import bench; "
            "tj=json.load(open('profiles/k1_traffic.json')); caps=tj['modes']['fused']['captures']; "
            "out=[bench.ncu_model('fused', c['frames_per_launch']) for c in caps] + [bench.ncu_model('fused', 20), bench.ncu_model('fused', 57), bench.ncu_model('nope', 20)]; "
            "sys.stderr.write('RESULT ' + json.dumps([caps, out]))")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    caps, out = json.loads(p.stderr.split("RESULT ", 1)[1])
    assert len(caps) >= 2 and out[-1] is None
    for c, m in zip(caps, out):
        assert abs(m["dram_bytes"] - c["dram_bytes"]) <= 0.02 * c["dram_bytes"] and abs(m["inst_executed"] - c["inst_executed"]) <= 0.02 * c["inst_executed"]
    m20, m57 = out[len(caps)], out[len(caps) + 1]
    rays = 10.24e6
    for m, F in ((m20, 20), (m57, 57)):
        assert 25 < m["inst_executed"] / (rays * F) < 45          # warp instructions per ray
        assert 0.5 < m["dram_bytes"] / (rays * F) < 40            # DRAM bytes per ray
    assert m20["dram_bytes"] / 20 > m57["dram_bytes"] / 57         # the per-launch RNG-state term is amortised over more frames
