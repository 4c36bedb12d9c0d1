#!/usr/bin/env python
"""bench.py -- headline benchmark of the compound-eye render path (BASELINE.json metric: rays/s and
ommatidia-frames/s at S samples/ommatidium on 1/2/4/8 B200).

Workload (config.workload): BASELINE.json configs[3] headline -- synthetic speed-test terrain of
~1M triangles (per-vertex u16 colours, simple_sky) seen by a 10 000-ommatidia Fibonacci eye,
S samples per ommatidium, projection single_dimension_fast.  A "step" is one frame = N*S rays.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--triangles T] [--ommatidia N]
  python bench.py --impl reference ...     CPU oracle port on the host cores (the reference has no
                                           CPU renderer and its OptiX build cannot be built here)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "compound-ray_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

BENCH_DIR = os.environ.get("CR_BENCH_DIR", "/tmp/crb200_bench")

# stdout carries exactly ONE JSON line: native libraries (NCCL's version banner, loader chatter) write to
# fd 1 directly, so fd 1 is pointed at stderr for the whole run and the result goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())
L2_BYTES = 126e6


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_workload(triangles, ommatidia, rank=0):
    """Writes the synthetic scene + eye once per (T, N) into BENCH_DIR (rank-private file names)."""
    from tools import synth
    os.makedirs(BENCH_DIR, exist_ok=True)
    tag = f"terrain_T{triangles}_N{ommatidia}_r{rank}"
    gltf = os.path.join(BENCH_DIR, tag + ".gltf")
    eye = os.path.join(BENCH_DIR, tag + ".eye")
    if not (os.path.exists(gltf) and os.path.exists(gltf + ".bin") and os.path.exists(eye)):
        synth.write_eye(eye, synth.fibonacci_eye(ommatidia))
        info = synth.write_terrain_gltf(gltf, triangles=triangles, eye_file=os.path.basename(eye))
        log(f"[bench] wrote {gltf}: {info}")
    return gltf, eye


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms from warm-up to the end of the e2e loop."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def poses_for(cam_pos, axes, count, first_index):
    """Deterministic pose sequence: a small drift around the scene's compound camera."""
    k = np.arange(first_index, first_index + count, dtype=np.float64)
    pos = np.stack([cam_pos[0] + 3.0 * np.sin(0.37 * k), cam_pos[1] + 0.5 * np.sin(0.11 * k) + 0.5,
                    cam_pos[2] + 3.0 * np.cos(0.23 * k)], axis=1).astype(np.float32)
    import eye_renderer as er
    return er.make_poses(pos, x=axes[0:3], y=axes[3:6], z=axes[6:9])


def traversal_counters(lib, er, S_probe=32):
    """Per-ray BVH nodes fetched / triangles tested, counted ON THE DEVICE by the dump variant of the trace
    kernel (same code path, entry frontier included) on a probe frame of S_probe samples, plus the same
    figures for a plain root-to-leaf traversal counted by the CPU oracle's instrumented walk of the IDENTICAL
    device BVH on the product's own rays (which also re-checks the hit ids)."""
    from oracle import oracle as O
    N = lib.getCurrentEyeOmmatidialCount()
    S_keep = lib.getCurrentEyeSamplesPerOmmatidium()
    lib.setCurrentEyeSamplesPerOmmatidium(S_probe)
    lib.crDebugSetEntryFrontier(1, -1, 0)            # the probe frame is small: keep the frontier pass of the timed frames
    lib.crDebugSetRayDump(True)
    lib.renderFrame()
    lib.crDebugSetEntryFrontier(1, -1, 3 << 18)
    n = N * S_probe
    o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); h = np.zeros((n, 4), np.int32)
    lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data)
    cnt_dev = np.zeros((n, 2), np.int32)
    assert lib.crDebugCopyLastRayCounts(cnt_dev.ctypes.data) == n
    lib.crDebugSetRayDump(False)
    T = lib.crDebugGetTriangleCount(); nn = lib.crDebugGetBvhNodeCount()
    nodes = np.zeros((nn, 16), np.float32); tris = np.zeros((T, 12), np.float32)
    lib.crDebugCopyBvh(nodes.ctypes.data, tris.ctypes.data)
    tm = np.zeros(n, np.float32)
    hits, cnt = O.trace_device_bvh(nodes, tris, o, d, tm)
    assert np.array_equal(hits["prim"], h[:, 0]), "oracle traversal of the device BVH disagrees with the kernel"
    lib.setCurrentEyeSamplesPerOmmatidium(S_keep)
    return {"nodes": float(cnt_dev[:, 0].mean()), "tris": float(cnt_dev[:, 1].mean()), "hit_fraction": float((h[:, 0] >= 0).mean()),
            "nodes_from_root": cnt[0] / n, "tris_from_root": cnt[1] / n}


def issue_roofline(tj, frames_per_launch, rays_per_step, rays_per_sec_per_gpu, sm_mhz, n_sms=148):
    """Second roofline of K1, the one that binds it: warp-instruction issue slots.  The instructions one launch
    executes were counted by ncu (profiles/k1_traffic.json, same kernel, same workload, same frames per launch);
    achieved = that count per ray x the LIVE rays/s; peak = SMs x 4 schedulers x 1 warp instruction per clock at the
    SM clock sampled during the timed region."""
    try:
        inst = float(tj["inst_executed_per_launch"])
        if int(tj["frames_per_launch"]) != int(frames_per_launch) or not sm_mhz:
            return None
        per_ray = inst / (float(rays_per_step) * frames_per_launch)
        peak = n_sms * 4 * float(sm_mhz) * 1e6
        achieved = rays_per_sec_per_gpu * per_ray
        return {"bound": "issue", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": achieved / peak,
                "warp_inst_per_ray": per_ray, "active_lanes_per_inst": tj.get("thread_inst_per_warp_inst"),
                "ncu_issue_active_pct": tj.get("issue_active_pct"),
                "note": "K1 is bound by instruction issue and load latency, not by HBM; the lanes idle inside the traversal loop "
                        "(active_lanes_per_inst of 32) are the remaining headroom.  achieved uses the whole step's rays/s (entry "
                        "frontier and ordered sum included), so K1 alone sits a few percent higher (ncu_issue_active_pct)"}
    except Exception:
        return None


def cpu_baseline(gltf, S, target_seconds=12.0, threads=None):
    """Times the CPU oracle port (raygen + BVH traversal + shading, all host threads) on a bounded
    sample of the same workload: the first n ommatidia of the eye at S samples each."""
    from oracle import gltf_loader, oracle as O
    sc = gltf_loader.load_scene(gltf)
    sh = O.SceneHandle(sc)
    cam = [c for c in sc.cameras if c.kind == "compound"][0]
    if threads:
        O.lib().cro_set_num_threads(int(threads))
    cores = O.lib().cro_num_threads()
    t0 = time.perf_counter(); sh.bvh(); build_s = time.perf_counter() - t0
    pose = O.pose_from_camera(cam)

    def run(n_omm):
        sub = cam.ommatidia[::max(1, len(cam.ommatidia) // n_omm)][:n_omm]     # evenly strided: unbiased in direction
        eye = O.CompoundEyeOracle(sh, sub, pose, "single_dimension_fast", samples=S)
        eye.set_render_size(n_omm, 1)
        eye.render_frame(method="bvh", project=False)        # frame 0 pays the stream initialisation
        t = time.perf_counter()
        eye.render_frame(method="bvh", project=False)
        return n_omm * S / (time.perf_counter() - t)

    n0 = max(1, min(len(cam.ommatidia), 65536 // S or 1))
    rate = run(n0)
    n1 = int(max(n0, min(len(cam.ommatidia), rate * target_seconds / S)))
    rate = run(n1)
    return {"value": rate, "unit": "rays/s", "cores": int(cores), "kind": "port",
            "sample": f"{n1} evenly strided of {len(cam.ommatidia)} ommatidia x {S} samples = {n1 * S} rays, 1 frame after RNG init "
                      f"(oracle BVH build {build_s:.2f} s excluded)"}


def run_reference(args):
    """--impl reference: the reference has no CPU implementation of this path and its OptiX build
    cannot be compiled here (no OptiX SDK), so this arm times the CPU oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gltf, _ = make_workload(args.triangles, args.ommatidia)
    from oracle import gltf_loader, oracle as O
    sc = gltf_loader.load_scene(gltf)
    sh = O.SceneHandle(sc)
    cam = [c for c in sc.cameras if c.kind == "compound"][0]
    sh.bvh()
    cores = O.lib().cro_num_threads()
    n_omm = max(1, min(len(cam.ommatidia), int(args.ref_rays_per_step // args.samples) or 1))
    sub = cam.ommatidia[::max(1, len(cam.ommatidia) // n_omm)][:n_omm]         # evenly strided: unbiased in direction
    eye = O.CompoundEyeOracle(sh, sub, O.pose_from_camera(cam), "single_dimension_fast", samples=args.samples)
    eye.set_render_size(n_omm, 1)
    for _ in range(max(args.warmup, 1)):
        eye.render_frame(method="bvh")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eye.render_frame(method="bvh")
    dt = time.perf_counter() - t0
    rays = n_omm * args.samples * args.steps
    val = rays / dt
    sample = f"{n_omm} evenly strided of {len(cam.ommatidia)} ommatidia x {args.samples} samples per step"
    out = {"impl": "reference", "metric": "rays_per_sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args),
           "cpu_baseline": {"value": val, "unit": "rays/s", "cores": int(cores), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "ommatidia_frames_per_sec": val / args.samples}
    emit(out)


def workload_config(args):
    return {"workload": f"speed-test terrain {args.triangles} triangles (u16 vertex colours, simple_sky) + "
                        f"{args.ommatidia}-ommatidia Fibonacci eye, S={args.samples}, single_dimension_fast",
            "triangles": args.triangles, "ommatidia": args.ommatidia, "samples_per_ommatidium": args.samples,
            "rays_per_step": args.ommatidia * args.samples,
            "l2_policy": "inputs larger than L2: RNG state %.0f MB + samples %.0f MB streamed per step, BVH %.0f MB "
                         "(8 octant variants of 64 B nodes at leaf size 2 + 48 B triangles), vs 126 MB L2" % (
                32e-6 * args.ommatidia * args.samples, 12e-6 * args.ommatidia * args.samples, (256 + 48) * 1e-6 * args.triangles)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--triangles", type=int, default=1_000_000)
    ap.add_argument("--ommatidia", type=int, default=10_000)
    ap.add_argument("--samples", type=int, default=1024)
    ap.add_argument("--ref-rays-per-step", type=float, default=2e6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="also report rays/s for S in {1,32,64,256}")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import eye_renderer as er
    gltf, _ = make_workload(args.triangles, args.ommatidia, rank=local)
    lib = er.load_library(device=local)
    lib.setVerbosity(False)
    lib.loadGlTFscene(gltf.encode())
    if not lib.gotoCameraByName(b"compound-cam"):
        raise SystemExit("compound-cam not found")
    N, S, K, W = args.ommatidia, args.samples, args.steps, args.warmup
    assert lib.getCurrentEyeOmmatidialCount() == N
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    pose0 = np.zeros(12, np.float32)
    lib.crDebugCopyCameraPose(pose0.ctypes.data)
    cam_pos, axes = pose0[:3].copy(), pose0[3:].copy()

    # rank r renders global frames [r*(W+K), (r+1)*(W+K)): its streams start at that frame
    first = rank * (W + K)
    lib.crSetFirstFrame(first)
    launches0 = lib.crGetLaunchCount()

    # ------------------------------------------------------------------ device-resident value
    out_dev = None
    if world > 1:
        import torch
        send = torch.empty((K, N, 4), dtype=torch.uint8, device="cuda")
        gathered = torch.empty((world * K, N, 4), dtype=torch.uint8, device="cuda")
        out_dev = send.data_ptr()
    sampler = ClockSampler(local)
    sampler.start()                                     # sampled from warm-up to the end of the e2e loop
    warm = poses_for(cam_pos, axes, W, first)
    er.renderPoseBatch(lib, warm)                       # W untimed warm-up frames (incl. RNG init)
    timed = poses_for(cam_pos, axes, K, first + W)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    launches1 = lib.crGetLaunchCount()
    if world > 1:
        _, host_ms = er.renderPoseBatch(lib, timed, out_device_ptr=out_dev)
        dev_ms = lib.crGetLastTraceMs()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather_into_tensor(gathered.view(-1), send.view(-1))
        e1.record()
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1)
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
        dist.barrier()
    else:
        out_host, host_ms = er.renderPoseBatch(lib, timed)      # D2H of K*N*4 bytes happens AFTER the event pair
        dev_ms = lib.crGetLastTraceMs()
    launches_timed = lib.crGetLaunchCount() - launches1
    lib.crGetLastBatchFrames.restype = C.c_int
    frames_per_launch = lib.crGetLastBatchFrames()
    rays_per_step = N * S
    value = world * K * rays_per_step / (dev_ms * 1e-3)

    # ------------------------------------------------------------------ e2e through the reference-facing C ABI
    # every rank drives its own GPU through the per-frame API at the same time: host pose in, host frame out
    Ke = K
    for k in range(2):
        lib.setCameraPosition(float(cam_pos[0]), float(cam_pos[1] + 0.5), float(cam_pos[2]))
        lib.renderFrame(); lib.getFramePointer()
    pe = poses_for(cam_pos, axes, Ke, first + W + K)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    checksum = 0
    for k in range(Ke):
        lib.setCameraPosition(float(pe[k, 0]), float(pe[k, 1]), float(pe[k, 2]))       # host pose in
        lib.renderFrame()
        fr = lib.getFramePointer()                                                       # D2H of the frame
        checksum += int(fr[0, 0, 0])
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop()

    out = None
    if rank == 0:
        e2e = {"value": world * Ke * rays_per_step / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": 4 * N,
               "ms_per_step": 1e3 * e2e_s / Ke, "api": "setCameraPosition + renderFrame + getFramePointer (ctypes), every rank "
               "concurrently on its own GPU; max over ranks"}

        # -------------------------------------------------------------- roofline of the dominant kernel (K1)
        tc = traversal_counters(lib, er)
        n_node, n_tri, hit_frac = tc["nodes"], tc["tris"], tc["hit_fraction"]
        lib.crGetLastBatchFrames.restype = C.c_int
        F = max(1, int(frames_per_launch))
        # per ray: node + triangle fetches, RNG state read+write once per F-frame launch, 12 B sample write +
        # 12 B read by the ordered sum, ommatidium row + result amortised over S
        bytes_per_ray = 64.0 * n_node + 48.0 * n_tri + 64.0 / F + 24.0 + 48.0 / S
        peak, peak_src = measured_peak_gbs()
        achieved = (value / world) * bytes_per_ray / 1e9
        traffic = None
        tj = {}
        tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch") if tj.get("frames_per_launch") == F else None
            except Exception:
                traffic, tj = None, {}
        dram_gbs = (traffic / (dev_ms / K * F * 1e-3) / 1e9) if traffic else None # measured DRAM bytes (ncu) over the live launch time
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "dram_achieved_gbs": dram_gbs, "dram_frac": (dram_gbs / peak) if dram_gbs else None,
                    "kernel": "k_traceCompound<false,true>", "frames_per_launch": F,
                    "algorithmic_bytes_per_launch": bytes_per_ray * rays_per_step * F, "bytes_per_ray": bytes_per_ray, "nodes_per_ray": n_node, "tris_per_ray": n_tri,
                    "nodes_per_ray_from_root": tc["nodes_from_root"], "tris_per_ray_from_root": tc["tris_from_root"],
                    "hit_fraction": hit_frac, "peak_source": peak_src,
                    "note": "achieved = SURVEY 8(d) algorithmic bytes (64*nodes + 48*tris + 64/F RNG r+w + 24 sample w+r + 48/S per ray) / time, "
                            "with nodes/tris per ray counted on the device by the dump variant of the kernel (entry frontier active; "
                            "*_from_root = what a root-to-leaf walk of the same tree fetches; the per-ommatidium frontier pass adds < 0.1 B/ray). "
                            "The BVH bytes are L1/L2 hits (one viewpoint per frame): the kernel is issue/latency-bound, not HBM-bound. "
                            "dram_* = ncu-measured DRAM bytes per launch (RNG state + samples) / live launch time: the true HBM "
                            "utilisation -- see profiles/ for the ncu summaries"}
        roofline["issue"] = issue_roofline(tj, F, rays_per_step, value / world, clocks.get("sm_mhz"))
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(gltf, S)
        out = {"metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": workload_config(args), "clocks": clocks, "e2e": e2e,
               "gpu_launches": int(launches_timed), "roofline": roofline, "cpu_baseline": cpu,
               "ommatidia_frames_per_sec": value / S, "bvh_build_ms": lib.crGetBvhBuildMs(),
               "timing": "CUDA events around the K fused trace+pack launches on the library stream"
                         + (", plus torch events around ncclAllGather; max over ranks" if world > 1 else "")}
        if args.sweep:
            sweep = {}
            for s in (1, 32, 64, 256):
                lib.setCurrentEyeSamplesPerOmmatidium(s)
                er.renderPoseBatch(lib, warm)
                er.renderPoseBatch(lib, timed)
                sweep[str(s)] = K * N * s / (lib.crGetLastTraceMs() * 1e-3)
            out["sweep_rays_per_sec_by_S"] = sweep
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


if __name__ == "__main__":
    main()
