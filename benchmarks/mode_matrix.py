#!/usr/bin/env python
"""benchmarks/mode_matrix.py -- rays/s of the headline workload (bench.py's: 1M-triangle terrain, 10k-ommatidia eye,
S = 1024) for every combination of the trace kernel's switches: per-ommatidium candidate lists on/off, ordered vs fused
reduction, cr_math vs hardware elementary functions; device-resident (crRenderPoseBatch, CUDA events) and through
the per-frame ABI (setCameraPosition + renderFrame + getFramePointer, wall clock).

  python benchmarks/mode_matrix.py [--frames 34] [--repeats 3] [--samples 1024] [--out gpurun_out/mode_matrix.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload generator and pose sequence)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=34)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--samples", type=int, default=1024)
    ap.add_argument("--triangles", type=int, default=1_000_000)
    ap.add_argument("--ommatidia", type=int, default=10_000)
    ap.add_argument("--lists", default="0,1", help="crDebugSetCandidateLists values: 0 never, 1 batches only (default), 2 always")
    ap.add_argument("--e2e-frames", type=int, default=20)
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
    ref_rows = None
    # candidate-list statistics of one frame
    lib.crDebugSetCandidateLists(2)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    lib.renderFrame()
    rec = np.zeros((N, 16), np.int32)
    got = lib.crDebugCopyCandidateLists(rec.ctypes.data, N)
    hdr = rec[:got, 0]
    stats = {"records": int(got), "fallback_fraction": float((hdr < 0).mean()) if got else None,
             "empty_fraction": float((hdr == 0).mean()) if got else None,
             "mean_elements_of_listed": float(hdr[hdr > 0].mean()) if (hdr > 0).any() else None,
             "histogram": np.bincount(hdr[hdr >= 0], minlength=16).tolist() if got else None}
    print(json.dumps({"candidate_list_stats": stats}), flush=True)
    for cone in [int(c) for c in a.lists.split(",")]:
        for fused in (0, 1):
            for fast in (0, 1):
                lib.crDebugSetCandidateLists(cone)
                lib.crSetRenderMode(fused, fast)
                lib.setCurrentEyeSamplesPerOmmatidium(S)
                er.renderPoseBatch(lib, poses[:3])                       # warm-up (stream init, allocations)
                ms = []
                for _ in range(a.repeats):
                    out, _ = er.renderPoseBatch(lib, poses)
                    ms.append(lib.crGetLastTraceMs())
                dev = a.frames * N * S / (np.median(ms) * 1e-3)
                # per-frame ABI
                for k in range(2):
                    lib.renderFrame(); lib.getFramePointer()
                t0 = time.perf_counter()
                for k in range(a.e2e_frames):
                    lib.setCameraPosition(float(poses[k % len(poses), 0]), float(poses[k % len(poses), 1]), float(poses[k % len(poses), 2]))
                    lib.renderFrame(); lib.getFramePointer()
                e2e = a.e2e_frames * N * S / (time.perf_counter() - t0)
                lib.setCameraPosition(float(pose0[0]), float(pose0[1]), float(pose0[2]))
                # same streams, same poses: exact-math rows must not depend on the traversal switch
                lib.setCurrentEyeSamplesPerOmmatidium(S)
                chk, _ = er.renderPoseBatch(lib, poses[:2])
                if fast == 0 and fused == 0:
                    if ref_rows is None:
                        ref_rows = chk.copy()
                    same = bool(np.array_equal(chk, ref_rows))
                else:
                    d = np.abs(chk.astype(int) - ref_rows.astype(int))
                    same = f"max {int(d.max())} step, {float((d > 0).mean()):.2e} of bytes"
                row = {"candidate_lists": cone, "fused": fused, "fast_math": fast, "frames_per_launch": int(lib.crGetLastBatchFrames()),
                       "grays_device": dev / 1e9, "ms_per_frame": float(np.median(ms)) / a.frames, "ms_all": ms,
                       "grays_per_frame_abi": e2e / 1e9, "rows_vs_ordered_exact": same}
                rows.append(row)
                print(json.dumps(row), flush=True)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump({"workload": f"terrain {a.triangles} tris, {N} ommatidia, S={S}", "candidate_list_stats": stats, "rows": rows},
                  open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
