#!/usr/bin/env python
"""benchmarks/wavefront_sweep.py -- the headline workload (bench.py's) with the wavefront queue off / on at several refill
thresholds and queue sizes: device-resident rays/s (crRenderPoseBatch, CUDA events), rays through the queue, and the
per-frame ABI (wall clock and the library's own event pair around the frame's kernels, i.e. what the host side adds).

  python benchmarks/wavefront_sweep.py [--frames 20] [--repeats 3] [--out gpurun_out/wavefront.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--samples", type=int, default=1024)
    ap.add_argument("--triangles", type=int, default=1_000_000)
    ap.add_argument("--ommatidia", type=int, default=10_000)
    ap.add_argument("--refills", default="16,24")
    ap.add_argument("--node-lanes", default="8,12,16", help="phase-switch thresholds tried with the queue on")
    ap.add_argument("--inline-lanes", default="16", help="thresholds tried with the queue off (matter only to a CR_INLINE_PHASED build)")
    ap.add_argument("--modes", default="1:0,1:1", help="fused:fast pairs")
    ap.add_argument("--e2e-frames", type=int, default=40)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import eye_renderer as er
    gltf, _ = bench.make_workload(a.triangles, a.ommatidia)
    lib = er.load_library(device=0)
    lib.setVerbosity(False)
    lib.loadGlTFscene(gltf.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    N, S = a.ommatidia, a.samples
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, N, 1)
    pose0 = np.zeros(12, np.float32)
    lib.crDebugCopyCameraPose(pose0.ctypes.data)
    poses = bench.poses_for(pose0[:3].copy(), pose0[3:].copy(), a.frames, 0)
    rows = []
    ref = {}

    def batch(tag, fused, fast, lists, wf, refill, frac, nl=16):
        lib.crDebugSetNodeLanes(nl)
        lib.crDebugSetCandidateLists(lists)
        lib.crDebugSetWavefront(wf, refill, frac)
        lib.crSetRenderMode(fused, fast)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        er.renderPoseBatch(lib, poses[:4])
        ms = []
        for _ in range(a.repeats):
            er.renderPoseBatch(lib, poses)
            ms.append(lib.crGetLastTraceMs())
        queued = int(lib.crDebugLastQueuedRays()) if wf else 0
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        chk, _ = er.renderPoseBatch(lib, poses[:4])
        key = (fused, fast)
        if key not in ref:
            ref[key] = chk.copy()
        row = {"what": tag, "fused": fused, "fast": fast, "lists": lists, "wavefront": wf, "refill_below": refill, "queue_fraction": frac, "node_lanes": nl,
               "grays_device": a.frames * N * S / (np.median(ms) * 1e-3) / 1e9, "ms_per_frame": float(np.median(ms)) / a.frames,
               "ms_all": [round(m, 3) for m in ms], "queued_fraction": queued / (a.frames * N * S),
               "rows_equal_first_variant": bool(np.array_equal(chk, ref[key]))}
        rows.append(row)
        print(json.dumps(row), flush=True)

    def per_frame(tag, fused, fast, lists, wf, refill, nl=16):
        lib.crDebugSetNodeLanes(nl)
        lib.crDebugSetCandidateLists(lists)
        lib.crDebugSetWavefront(wf, refill, 0.35)
        lib.crSetRenderMode(fused, fast)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        lib.crGetLastTraceMs()
        for k in range(3):
            lib.renderFrame(); lib.getFramePointer()
        ev = []
        t0 = time.perf_counter()
        for k in range(a.e2e_frames):
            p = poses[k % len(poses)]
            lib.setCameraPosition(float(p[0]), float(p[1]), float(p[2]))
            lib.renderFrame(); lib.getFramePointer()
            ev.append(lib.crGetLastTraceMs())
        wall = (time.perf_counter() - t0) / a.e2e_frames * 1e3
        lib.crDebugSetFrameProfile(1)                      # device-side breakdown, outside the timed loop
        bd = []
        for k in range(8):
            p = poses[k % len(poses)]
            lib.setCameraPosition(float(p[0]), float(p[1]), float(p[2]))
            lib.renderFrame(); lib.getFramePointer()
            b3 = np.zeros(3, np.float32); lib.crDebugFrameBreakdown(b3.ctypes.data); bd.append(b3)
        lib.crDebugSetFrameProfile(0)
        bd = np.median(np.asarray(bd), axis=0)
        lib.setCameraPosition(float(pose0[0]), float(pose0[1]), float(pose0[2]))
        row = {"what": tag, "fused": fused, "fast": fast, "lists": lists, "wavefront": wf, "refill_below": refill, "node_lanes": nl,
               "per_frame_wall_ms": wall, "per_frame_kernels_ms": float(np.median(ev)), "host_side_ms": wall - float(np.median(ev)),
               "frontier_ms": float(bd[0]), "trace_ms": float(bd[1]), "reduce_ms": float(bd[2]),
               "grays_per_frame_abi": N * S / (wall * 1e-3) / 1e9}
        rows.append(row)
        print(json.dumps(row), flush=True)

    inline_lanes = [int(v) for v in a.inline_lanes.split(",")]
    for spec in a.modes.split(","):
        fused, fast = (int(v) for v in spec.split(":"))
        lanes = [int(v) for v in a.node_lanes.split(",")]
        for nl in inline_lanes:
            batch("batch, queue off", fused, fast, 1, 0, 24, 0.35, nl)
        for nl in lanes:
            for refill in [int(v) for v in a.refills.split(",")]:
                batch("batch, queue on", fused, fast, 1, 1, refill, 0.35, nl)
        batch("batch, no lists, no queue", fused, fast, 0, 0, 24, 0.35, inline_lanes[0])
        for nl in inline_lanes:
            per_frame("per frame, no lists", fused, fast, 1, 0, 24, nl)
        per_frame("per frame, lists", fused, fast, 2, 0, 24, inline_lanes[0])
        for nl in lanes:
            per_frame("per frame, lists + queue", fused, fast, 2, 1, int(a.refills.split(",")[0]), nl)
    lib.crDebugSetNodeLanes(16)
    lib.crDebugSetCandidateLists(1)
    lib.crDebugSetWavefront(1, 24, 0.35)
    lib.crSetRenderMode(0, 0)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump({"workload": f"terrain {a.triangles} tris, {N} ommatidia, S={S}, {a.frames} frames per batch", "rows": rows},
                  open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
