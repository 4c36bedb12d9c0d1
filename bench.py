#!/usr/bin/env python
"""bench.py -- headline benchmark of the compound-eye render path (BASELINE.json metric: rays/s and
ommatidia-frames/s at S samples/ommatidium on 1/2/4/8 B200).

Workload (config.workload): BASELINE.json configs[3] headline -- synthetic speed-test terrain of
~1M triangles (per-vertex u16 colours, simple_sky) seen by a 10 000-ommatidia Fibonacci eye,
S samples per ommatidium, projection single_dimension_fast.  A "step" is one frame = N*S rays.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--repeats R] [--mode fused|ordered|fused_fast|ordered_fast]
                  [--samples S] [--triangles T] [--ommatidia N]
  python bench.py --impl reference ...     CPU oracle port on the host cores (the reference has no
                                           CPU renderer and its OptiX build cannot be built here)

Timing: W untimed warm-up frames, then R batches of EXACTLY K frames; every batch is bracketed by a barrier and a
device synchronisation and timed with CUDA events on the library's stream (plus torch events around the NCCL
all-gather for N > 1), maximum over ranks per batch; `value` and `ms_per_step` come from the MEDIAN batch, all R
batch times are listed in `batch_ms`.

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
import zlib

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

# crSetRenderMode(fused, fast) per named mode.  "ordered" = the reference's sequential sum and cr_math: every bit equal
# to the checker.  "fused" (headline) = same rays/hits/colours, in-kernel reduction in a fixed order the checker restates.
MODES = {"ordered": (0, 0), "fused": (1, 0), "fused_fast": (1, 1), "ordered_fast": (0, 1)}
KERNEL_OF_MODE = {"ordered": "k_traceCompound<false,true,false,false>", "fused": "k_traceCompound<false,true,true,false>",
                  "fused_fast": "k_traceCompound<false,true,true,true>", "ordered_fast": "k_traceCompound<false,true,false,true>"}


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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
        # nvidia-smi numbers the PHYSICAL GPUs; the CUDA ordinal is an index into CUDA_VISIBLE_DEVICES when that is set
        # (sampling GPU 0 while the work runs on another one reads an idle clock).  Prefer the UUID torch reports.
        self.device = str(device)
        uuid = None
        try:
            import torch
            uuid = getattr(torch.cuda.get_device_properties(int(device)), "uuid", None)
        except Exception:
            uuid = None
        if uuid is not None:
            u = str(uuid)
            self.device = u if u.startswith("GPU-") else "GPU-" + u
        else:
            ids = [x.strip() for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
            if ids and int(device) < len(ids):
                self.device = ids[int(device)]
        if self.device != str(device):                  # make sure nvidia-smi accepts that name, else the plain ordinal
            try:
                ok = subprocess.run(["nvidia-smi", "-i", self.device, "--query-gpu=index", "--format=csv,noheader"],
                                    capture_output=True, text=True, timeout=10).returncode == 0
            except Exception:
                ok = False
            if not ok:
                self.device = str(device)
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
    kernel (same code path, entry frontier included) on a probe frame of S_probe samples, plus the same figures for a plain root-to-leaf walk counted by
    the CPU oracle's instrumented traversal of the IDENTICAL device BVH on the product's own rays (which also
    re-checks the hit ids)."""
    from oracle import oracle as O
    N = lib.getCurrentEyeOmmatidialCount()
    S_keep = lib.getCurrentEyeSamplesPerOmmatidium()
    lib.setCurrentEyeSamplesPerOmmatidium(S_probe)
    lib.crDebugSetEntryFrontier(1, -1, 0)            # the probe frame is small and single: keep the frontier pass of the timed frames
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
    hit_ids_equal = bool(np.array_equal(hits["prim"], h[:, 0]))
    lib.setCurrentEyeSamplesPerOmmatidium(S_keep)
    hit = h[:, 0] >= 0
    trav = cnt_dev[:, 0] > 0
    return {"nodes": float(cnt_dev[:, 0].mean()), "tris": float(cnt_dev[:, 1].mean()), "hit_fraction": float(hit.mean()),
            "traversing_fraction": float(trav.mean()),
            "nodes_per_traversing_ray": float(cnt_dev[trav, 0].mean()) if trav.any() else 0.0,
            "tris_per_traversing_ray": float(cnt_dev[trav, 1].mean()) if trav.any() else 0.0,
            "nodes_from_root": cnt[0] / n, "tris_from_root": cnt[1] / n, "hit_ids_equal_oracle_walk": hit_ids_equal}


def ncu_model(mode, F):
    """Per-launch DRAM bytes and warp instructions of the trace kernel at F frames per launch, from the committed ncu
    captures (profiles/k1_traffic.json: `--set full --clock-control none` launches of this kernel on this workload at
    two or more F).  Both quantities are affine in F (a per-launch term -- the RNG state read and written once -- plus
    a per-frame term), so a line through the captures reconstructs any F the driver picks."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        caps = sorted(tj["modes"][mode]["captures"], key=lambda c: c["frames_per_launch"])
    except Exception:
        return None
    if not caps:
        return None

    def at(key):
        xs = np.array([c["frames_per_launch"] for c in caps], float)
        ys = np.array([c[key] for c in caps], float)
        if len(caps) == 1 or np.ptp(xs) == 0:
            return float(ys[0] * F / xs[0])
        b, a = np.polyfit(xs, ys, 1)
        return float(a + b * F)

    near = min(caps, key=lambda c: abs(c["frames_per_launch"] - F))
    return {"dram_bytes": at("dram_bytes"), "inst_executed": at("inst_executed"),
            "issue_active_pct": near.get("issue_active_pct"), "thread_inst_per_warp_inst": near.get("thread_inst_per_warp_inst"),
            "captures_at_frames_per_launch": [c["frames_per_launch"] for c in caps], "source": tj["modes"][mode].get("source"),
            "model": "exact capture" if any(c["frames_per_launch"] == F for c in caps) else
                     ("line through the captures" if len(caps) > 1 else "proportional to the single capture")}


def sample_checksum(O, sh, cam, S):
    """CRC of the float RGB of a fixed small frame (64 strided ommatidia, frame 0): equal for every build / thread count."""
    sub = cam.ommatidia[::max(1, len(cam.ommatidia) // 64)][:64]
    eye = O.CompoundEyeOracle(sh, sub, O.pose_from_camera(cam), "single_dimension_fast", samples=S)
    eye.set_render_size(len(sub), 1)
    eye.render_frame(method="bvh", project=False)
    return zlib.crc32(eye.last["summed"].tobytes())


def run_reference(args):
    """--impl reference: the reference has no CPU implementation of this path and its OptiX build cannot be compiled
    here (no OptiX SDK), so this arm times the CPU oracle port: the SAME source as the checker, compiled on
    the machine it is timed on with `gcc -O3 -march=native` (no contraction, no fast-math: same bits), on all host
    threads (OMP_NUM_THREADS is overridden: torchrun sets it to 1).  A step = one frame of a bounded sample of the
    workload; the value is the median step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gltf, _ = make_workload(args.triangles, args.ommatidia)
    build = "portable build (gcc -O2 -march=x86-64-v3)"
    if not args.ref_portable:
        from oracle import oracle as O0
        native = O0.build_native(BENCH_DIR)
        if native:
            os.environ["CR_ORACLE_SO"] = native
            build = "gcc -O3 -march=native -ffp-contract=off, compiled on this host"
    from oracle import gltf_loader, oracle as O
    threads = int(args.ref_threads) if args.ref_threads else host_threads()
    O.lib().cro_set_num_threads(threads)
    cores = O.lib().cro_num_threads()
    sc = gltf_loader.load_scene(gltf)
    sh = O.SceneHandle(sc)
    cam = [c for c in sc.cameras if c.kind == "compound"][0]
    sh.bvh()
    crc = sample_checksum(O, sh, cam, min(args.samples, 64))
    n_omm = max(1, min(len(cam.ommatidia), int(args.ref_rays_per_step // args.samples) or 1))
    sub = cam.ommatidia[::max(1, len(cam.ommatidia) // n_omm)][:n_omm]         # evenly strided: unbiased in direction
    eye = O.CompoundEyeOracle(sh, sub, O.pose_from_camera(cam), "single_dimension_fast", samples=args.samples)
    eye.set_render_size(n_omm, 1)
    for _ in range(max(args.warmup, 1)):
        eye.render_frame(method="bvh")
    step_s = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        eye.render_frame(method="bvh")
        step_s.append(time.perf_counter() - t0)
    med = float(np.median(step_s))
    rays = n_omm * args.samples
    val = rays / med
    sample = (f"{n_omm} evenly strided of {len(cam.ommatidia)} ommatidia x {args.samples} samples = {rays} rays per step; "
              f"median of {args.steps} steps after {max(args.warmup, 1)} warm-up; {build}")
    out = {"impl": "reference", "metric": "rays_per_sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * med, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args),
           "cpu_baseline": {"value": val, "unit": "rays/s", "cores": int(cores), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "ommatidia_frames_per_sec": val / args.samples, "step_ms": [1e3 * s for s in step_s],
           "sample_rgb_crc32": crc, "host_threads_available": host_threads()}
    emit(out)


def cpu_baseline(args):
    """The reference arm above, run twice as a subprocess (so the timing build and thread count cannot leak into this
    process): single-threaded and on all host threads, median of 3 steps each."""
    def run(threads, rays):
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
               "--triangles", str(args.triangles), "--ommatidia", str(args.ommatidia), "--samples", str(args.samples),
               "--ref-threads", str(threads), "--ref-rays-per-step", str(rays)]
        env = dict(os.environ)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS"):
            env.pop(k, None)
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
        return json.loads(r.stdout.strip().splitlines()[-1])
    try:
        one = run(1, 1.0e6)
        n = host_threads()
        allc = run(n, min(float(args.ommatidia * args.samples), max(2.0e6, one["value"] * n * 2.0)))
        from oracle import gltf_loader, oracle as O          # the portable checker build, in THIS process: same bits?
        gltf, _ = make_workload(args.triangles, args.ommatidia)
        sc = gltf_loader.load_scene(gltf)
        crc = sample_checksum(O, O.SceneHandle(sc), [c for c in sc.cameras if c.kind == "compound"][0], min(args.samples, 64))
        cb = dict(allc["cpu_baseline"])
        cb["single_thread"] = {"value": one["value"], "cores": one["cpu_baseline"]["cores"], "sample": one["cpu_baseline"]["sample"],
                               "step_ms": one["step_ms"]}
        cb["step_ms"] = allc["step_ms"]
        cb["repetitions"] = 3
        cb["same_bits_as_checker_build"] = bool(crc == allc["sample_rgb_crc32"] == one["sample_rgb_crc32"])
        return cb
    except Exception as e:          # the baseline is a reported figure; it must not take the GPU numbers down with it
        return {"value": None, "unit": "rays/s", "cores": None, "kind": "port", "sample": f"failed: {e!r}"}


def workload_config(args):
    return {"workload": f"speed-test terrain {args.triangles} triangles (u16 vertex colours, simple_sky) + "
                        f"{args.ommatidia}-ommatidia Fibonacci eye, S={args.samples}, single_dimension_fast",
            "triangles": args.triangles, "ommatidia": args.ommatidia, "samples_per_ommatidium": args.samples,
            "rays_per_step": args.ommatidia * args.samples,
            "l2_policy": "inputs larger than L2: RNG state %.0f MB streamed per launch, BVH %.0f MB (8 octant variants of 64 B "
                         "nodes at leaf size 2 + 48 B triangles), vs 126 MB L2; a new camera pose every frame" % (
                32e-6 * args.ommatidia * args.samples, (256 + 48) * 1e-6 * args.triangles)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--repeats", type=int, default=5, help="timed batches of --steps frames; the median is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fused", choices=sorted(MODES))
    ap.add_argument("--triangles", type=int, default=1_000_000)
    ap.add_argument("--ommatidia", type=int, default=10_000)
    ap.add_argument("--samples", type=int, default=1024)
    ap.add_argument("--ref-rays-per-step", type=float, default=4e6)
    ap.add_argument("--ref-threads", type=int, default=0, help="reference arm: host threads (0 = all)")
    ap.add_argument("--ref-portable", action="store_true", help="reference arm: time the shipped -O2 checker build")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the rays/s of the other render modes")
    ap.add_argument("--sweep", action="store_true", help="also report rays/s for S in {1,32,64,256}")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = torch = None
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
    N, S, K, W, R = args.ommatidia, args.samples, args.steps, args.warmup, max(1, args.repeats)
    assert lib.getCurrentEyeOmmatidialCount() == N
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, N, 1)
    lib.crSetRenderMode(*MODES[args.mode])
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    pose0 = np.zeros(12, np.float32)
    lib.crDebugCopyCameraPose(pose0.ctypes.data)
    cam_pos, axes = pose0[:3].copy(), pose0[3:].copy()
    rays_per_step = N * S

    # rank r renders global frames [r*span, (r+1)*span): its streams start at that frame
    span = W + (R + 1) * K
    first = rank * span
    lib.crSetFirstFrame(first)

    send = gathered = None
    out_dev = None
    if world > 1:
        send = torch.empty((K, N, 4), dtype=torch.uint8, device="cuda")
        gathered = torch.empty((world * K, N, 4), dtype=torch.uint8, device="cuda")
        out_dev = send.data_ptr()
        torch.cuda.synchronize()                          # the renderer writes `send` from its own (non-blocking) stream
    sampler = ClockSampler(local)
    sampler.start()                                       # sampled from warm-up to the end of the e2e loop

    # Every timed batch, on every rank, renders the SAME K camera poses (a new pose every frame within the batch): the
    # per-rank work of this weak-scaling run is then identical by construction, not just statistically, and the scaling
    # figure measures the machine, not the luck of a rank's pose draw (pose sets differ by +-3 % in cost).  The sample
    # streams do advance from batch to batch and differ between ranks (crSetFirstFrame), so no two batches trace the same rays.
    timed_poses = poses_for(cam_pos, axes, K, W)
    rank_ms = []                                          # per batch: (this rank's trace ms, this rank's gather ms)

    def timed_batches(repeats, with_gather=True):
        """`repeats` batches of K frames; per batch: barrier + sync, CUDA events around the fused launches on the library
        stream (+ torch events around ncclAllGather), max over ranks."""
        out = []
        for b in range(repeats):
            timed = timed_poses
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
                er.renderPoseBatch(lib, timed, out_device_ptr=out_dev)       # returns after the stream is idle
                dev_ms = lib.crGetLastTraceMs()
                g_ms = 0.0
                if with_gather:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    dist.all_gather_into_tensor(gathered.view(-1), send.view(-1))
                    e1.record()
                    torch.cuda.synchronize()
                    g_ms = e0.elapsed_time(e1)
                rank_ms.append((dev_ms, g_ms))
                dev_ms += g_ms
                t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dev_ms = float(t.item())
            else:
                er.renderPoseBatch(lib, timed)                                # D2H of K*N*4 bytes happens AFTER the event pair
                dev_ms = lib.crGetLastTraceMs()
            out.append(dev_ms)
        return out

    er.renderPoseBatch(lib, poses_for(cam_pos, axes, W, first), out_device_ptr=None)   # W untimed warm-up frames (incl. RNG init)
    if world > 1:                                         # ... and the collective with its real shape (first call builds channels)
        er.renderPoseBatch(lib, timed_poses, out_device_ptr=out_dev)
        for _ in range(2):
            dist.all_gather_into_tensor(gathered.view(-1), send.view(-1))
        torch.cuda.synchronize()
    else:
        er.renderPoseBatch(lib, timed_poses)                                  # same shape as the timed batches (allocations)
    launches1 = lib.crGetLaunchCount()
    batch_ms = timed_batches(R)
    launches_timed = lib.crGetLaunchCount() - launches1
    frames_per_launch = int(lib.crGetLastBatchFrames())
    dev_ms = float(np.median(batch_ms))
    value = world * K * rays_per_step / (dev_ms * 1e-3)

    # ------------------------------------------------------------------ e2e through the reference-facing C ABI
    # every rank drives its own GPU through the per-frame API at the same time: host pose in, host frame out
    def e2e_run(frames, first_frame):
        for _ in range(2):
            lib.setCameraPosition(float(cam_pos[0]), float(cam_pos[1] + 0.5), float(cam_pos[2]))
            lib.renderFrame(); lib.getFramePointer()
        pe = poses_for(cam_pos, axes, frames, first_frame)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        checksum = 0
        for k in range(frames):
            lib.setCameraPosition(float(pe[k, 0]), float(pe[k, 1]), float(pe[k, 2]))       # host pose in
            lib.renderFrame()
            fr = lib.getFramePointer()                                                       # D2H of the frame
            checksum += int(fr[0, 0, 0])
        s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s = float(t.item())
        return s
    Ke = K
    e2e_s = e2e_run(Ke, first + W + (R + 1) * K)

    # ------------------------------------------------------------------ the other render modes (one GPU only)
    modes = None
    if world == 1 and not args.no_modes:
        modes = {args.mode: {"rays_per_sec": value, "e2e_rays_per_sec": Ke * rays_per_step / e2e_s}}
        for name in ("ordered", "fused", "fused_fast"):
            if name in modes:
                continue
            lib.crSetRenderMode(*MODES[name])
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            er.renderPoseBatch(lib, poses_for(cam_pos, axes, K, 0))
            ms = float(np.median(timed_batches(3)))
            es = e2e_run(Ke, 4 * K)
            modes[name] = {"rays_per_sec": K * rays_per_step / (ms * 1e-3), "e2e_rays_per_sec": Ke * rays_per_step / es,
                           "frames_per_launch": int(lib.crGetLastBatchFrames())}
        modes["note"] = ("ordered = reference's sequential per-sample sum + cr_math (bit-exact vs the checker); fused = in-kernel "
                         "reduction in a fixed order the checker restates (same rays, hits, colours); fused_fast = fused + "
                         "hardware sin/cos/log/pow as the reference's --use_fast_math build (tolerance 1/255)")
        lib.crSetRenderMode(*MODES[args.mode])
        lib.setCurrentEyeSamplesPerOmmatidium(S)
    clocks = sampler.stop()
    per_rank = None
    if world > 1:                                         # diagnostics: every rank's median trace time and gather time (incl. waiting)
        mine = torch.tensor([float(np.median([a for a, _ in rank_ms[:R]])), float(np.median([g for _, g in rank_ms[:R]]))],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"trace_ms_per_batch": [float(t[0]) for t in allr], "gather_ms_per_batch_incl_wait": [float(t[1]) for t in allr]}

    out = None
    if rank == 0:
        e2e = {"value": world * Ke * rays_per_step / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": 4 * N,
               "ms_per_step": 1e3 * e2e_s / Ke, "api": "setCameraPosition + renderFrame + getFramePointer (ctypes), every rank "
               "concurrently on its own GPU; max over ranks", "mode": args.mode}

        # -------------------------------------------------------------- rooflines of the dominant kernel (K1)
        tc = traversal_counters(lib, er)
        n_node, n_tri = tc["nodes"], tc["tris"]
        F = max(1, frames_per_launch)
        # SURVEY 8(d), strictly: per ray 64 B per BVH node fetched + 48 B per triangle tested + RNG state (32 B read + 32 B
        # written once per F-frame launch) + ommatidium row (32 B) and result (16 B) amortised over S
        bytes_per_ray = 64.0 * n_node + 48.0 * n_tri + 64.0 / F + 48.0 / S
        peak, peak_src = measured_peak_gbs()
        rate = value / world                                            # rays/s per GPU
        launch_s = dev_ms * 1e-3 * F / K                                # average duration of one F-frame launch group, live
        nm = ncu_model(args.mode, F)
        traffic = nm["dram_bytes"] if nm else None
        hbm = {"bound": "hbm", "achieved": rate * bytes_per_ray / 1e9, "peak": peak, "unit": "GB/s",
               "frac": rate * bytes_per_ray / 1e9 / peak, "bytes_per_ray": bytes_per_ray,
               "algorithmic_bytes_per_launch": bytes_per_ray * rays_per_step * F,
               "dram_achieved_gbs": (traffic / launch_s / 1e9) if traffic else None,
               "dram_frac": (traffic / launch_s / 1e9 / peak) if traffic else None, "peak_source": peak_src,
               "note": "SURVEY 8(d) formula: 64*nodes + 48*tris + 64/F (RNG state r+w per launch) + 48/S per ray, nodes/tris counted on "
                       "the device.  These bytes are L1/L2 hits (one viewpoint per frame): dram_* = ncu-measured DRAM bytes of the "
                       "launch over its live duration is the true HBM utilisation, and it is small -- the kernel is not HBM-bound"}
        issue = None
        if nm and clocks.get("sm_mhz"):
            per_ray = nm["inst_executed"] / (float(rays_per_step) * F)
            ipeak = 148 * 4 * float(clocks["sm_mhz"]) * 1e6
            issue = {"achieved": rate * per_ray / 1e9, "peak": ipeak / 1e9, "unit": "G warp-inst/s", "frac": rate * per_ray / ipeak,
                     "warp_inst_per_ray": per_ray, "active_lanes_per_inst": nm["thread_inst_per_warp_inst"],
                     "ncu_issue_active_pct": nm["issue_active_pct"]}
        roofline = {"bound": "issue", "kernel": KERNEL_OF_MODE[args.mode], "frames_per_launch": F,
                    "achieved": issue["achieved"] if issue else None, "peak": issue["peak"] if issue else None,
                    "unit": "G warp-inst/s", "frac": issue["frac"] if issue else None, "traffic": traffic,
                    "issue": issue, "hbm": hbm, "ncu": nm,
                    "nodes_per_ray": n_node, "tris_per_ray": n_tri, "nodes_per_ray_from_root": tc["nodes_from_root"],
                    "tris_per_ray_from_root": tc["tris_from_root"], "hit_fraction": tc["hit_fraction"],
                    "traversing_fraction": tc["traversing_fraction"], "nodes_per_traversing_ray": tc["nodes_per_traversing_ray"],
                    "tris_per_traversing_ray": tc["tris_per_traversing_ray"],
                    "hit_only_rays_per_sec_equivalent": rate * tc["hit_fraction"],
                    "hit_ids_equal_oracle_walk": tc["hit_ids_equal_oracle_walk"],
                    "note": "K1 is bound by warp-instruction issue (and the load latency behind it), not by HBM: `achieved` = warp "
                            "instructions per launch counted by ncu (profiles/k1_traffic.json, same kernel and workload, reconstructed "
                            "at this run's frames per launch) x live launches/s; `peak` = 148 SMs x 4 schedulers x the SM clock sampled "
                            "during the run.  `traffic` = ncu DRAM bytes per launch.  `hbm` is the SURVEY 8(d) line.  About half of the "
                            "eye looks at the sky (hit_fraction): those rays never enter the BVH"}
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(args)
        out = {"metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": dict(workload_config(args), mode=args.mode), "clocks": clocks, "e2e": e2e,
               "gpu_launches": int(launches_timed), "gpu_launches_per_batch": int(launches_timed) // R,
               "repeats": R, "batch_ms": batch_ms, "per_rank": per_rank, "roofline": roofline, "cpu_baseline": cpu, "modes": modes,
               "ommatidia_frames_per_sec": value / S, "bvh_build_ms": lib.crGetBvhBuildMs(),
               "timing": f"median of {R} batches of {K} frames; CUDA events around each batch's fused trace+reduce launches on the "
                         "library stream" + (", plus torch events around ncclAllGather of the batch's rows; max over ranks per batch"
                                             if world > 1 else "")}
        if args.sweep:
            sweep = {}
            for s in (1, 32, 64, 256):
                lib.setCurrentEyeSamplesPerOmmatidium(s)
                er.renderPoseBatch(lib, poses_for(cam_pos, axes, W, 0))
                er.renderPoseBatch(lib, poses_for(cam_pos, axes, K, W))
                sweep[str(s)] = K * N * s / (lib.crGetLastTraceMs() * 1e-3)
            out["sweep_rays_per_sec_by_S"] = sweep
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


if __name__ == "__main__":
    main()
