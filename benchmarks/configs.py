#!/usr/bin/env python
"""benchmarks/configs.py -- one line per BASELINE.json configuration other than the headline (bench.py covers configs[3]
at 1M triangles), at the sizes SURVEY 8(d) fixes, on one GPU.  Writes JSON (stdout and --out).

  cfg2  data/natural-standin-sky.gltf + AM_60185 geometry (6 374 ommatidia) at S = 64: rays/s and ommatidia-frames/s,
        per-frame ABI and batched
  cfg3  synthetic ofstad arena (JPEG-textured cylinder) + icosahedral 12-ommatidia eye and 1000-equidistant eye:
        for S in {1, 2, 4, ..., 1024}, `--frames` consecutive frames through the per-frame ABI exactly as
        data/tools/minimumSampleRateFinder.py:36-47,271-282 takes them (the first "frame" of each S is the stale frame
        of the previous S, as in the script) -> max per-ommatidium frame-to-frame SD, rays/s
  cfg4  speed-test terrain at 10^4 / 10^5 / 10^6 / 10^7 triangles, 10 000-ommatidia eye, S = 1024 (batched, fused mode)
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
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import speed_test  # noqa: E402


def variance_image(frames):
    allImages = np.vstack(frames).astype(np.float64)
    diff = allImages - allImages.mean(axis=0)
    mag = np.linalg.norm(diff, axis=2)
    return (mag * mag).sum(axis=0) / (len(frames) - 1)


def cfg2(lib, er, data, mode):
    lib.loadGlTFscene(os.path.join(data, "data", "natural-standin-sky.gltf").encode())
    assert lib.gotoCameraByName(b"insect-eye-spherical-projector")
    omm = np.asarray([[*o.position, *o.direction, o.acceptanceAngle, o.focalpointOffset]
                      for o in er.readEyeFile(os.path.join(data, "sim-environment", "eyes", "AM_60185-real.eye"))], np.float32)
    omm[:, 0:3] *= np.float32(0.1); omm[:, 7] *= np.float32(0.1)
    er.setOmmatidiaFromArray(lib, omm)
    N, S = len(omm), 64
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, N, 1)
    lib.crSetRenderMode(*bench.MODES[mode])
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    for _ in range(20):
        lib.renderFrame(); lib.getFramePointer()
    K = 500
    t0 = time.perf_counter()
    for _ in range(K):
        lib.renderFrame(); lib.getFramePointer()
    abi = K / (time.perf_counter() - t0)                   # standing camera (the read-ahead renders several frames per launch)
    pose = np.zeros(12, np.float32); lib.crDebugCopyCameraPose(pose.ctypes.data)
    t0 = time.perf_counter()
    for k in range(K):                                     # moving camera: a new pose every frame
        lib.setCameraPosition(float(pose[0] + 0.02 * (k % 50)), float(pose[1]), float(pose[2]))
        lib.renderFrame(); lib.getFramePointer()
    abi_moving = K / (time.perf_counter() - t0)
    lib.setCameraPosition(float(pose[0]), float(pose[1]), float(pose[2]))
    poses = np.tile(pose, (2048, 1)); poses[:, 0] += np.linspace(-20, 20, 2048, dtype=np.float32)
    er.renderPoseBatch(lib, poses[:64])
    ms = []
    for _ in range(3):
        er.renderPoseBatch(lib, poses); ms.append(lib.crGetLastTraceMs())
    fps = len(poses) / (np.median(ms) * 1e-3)
    return {"config": "cfg2: natural-standin-sky.gltf (24 200 triangles, 1024^2 texture, simple_sky) + AM_60185 geometry, 6 374 ommatidia, S=64",
            "mode": mode, "ommatidia": N, "samples": S,
            "per_frame_abi": {"frames_per_sec": abi, "rays_per_sec": abi * N * S, "ommatidia_frames_per_sec": abi * N,
                              "camera": "standing (renderFrame's read-ahead applies)"},
            "per_frame_abi_moving_camera": {"frames_per_sec": abi_moving, "rays_per_sec": abi_moving * N * S, "ommatidia_frames_per_sec": abi_moving * N},
            "batched": {"frames_per_sec": fps, "rays_per_sec": fps * N * S, "ommatidia_frames_per_sec": fps * N}}


def cfg3(lib, er, data, mode, frames):
    from tools import synth
    os.makedirs(bench.BENCH_DIR, exist_ok=True)
    gltf = os.path.join(bench.BENCH_DIR, "arena.gltf")
    synth.write_eye(os.path.join(bench.BENCH_DIR, "ico.eye"), synth.ico_eye())
    info = synth.write_arena_gltf(gltf, eye_file="ico.eye")
    lib.loadGlTFscene(gltf.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    lib.crSetRenderMode(*bench.MODES[mode])
    out = {"config": f"cfg3: synthetic ofstad arena ({info['triangles']} triangles, 1024^2 JPEG pattern, default background), "
                     f"minimumSampleRateFinder statistic over {frames} consecutive frames per S", "mode": mode, "eyes": {}}
    eyes = {"icosahedral-12 (1 sr each)": er.getIcoOmmatidia(),
            "1000-equidistant.eye": er.readEyeFile(os.path.join(data, "data", "eyes", "1000-equidistant.eye"))}
    lib.setCameraPose(3.0, 2.5, -4.0, 0.3, -0.8, 0.1)
    for name, omms in eyes.items():
        er.setOmmatidiaFromOmmatidiumList(lib, omms)
        N = len(omms)
        lib.setCurrentEyeShaderName(b"single_dimension_fast")
        er.setRenderSize(lib, N, 1)
        lib.setCurrentEyeSamplesPerOmmatidium(1)
        lib.renderFrame()
        rows = []
        for S in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            imgs = [np.copy(lib.getFramePointer()[:, :, :3])]            # the stale frame, as the script does (:37-38)
            t0 = time.perf_counter()
            for _ in range(frames - 1):
                lib.renderFrame()
                imgs.append(np.copy(lib.getFramePointer()[:, :, :3]))
            dt = time.perf_counter() - t0
            rows.append({"S": S, "max_sd": float(np.sqrt(variance_image(imgs).max())), "frames_per_sec": (frames - 1) / dt,
                         "rays_per_sec": (frames - 1) * N * S / dt})
        out["eyes"][name] = {"ommatidia": N, "sweep": rows,
                             "sd_limit_1pct": 0.01 * float(np.linalg.norm([255.0] * 3))}
    return out


def cfg4(lib, er, mode, sizes):
    rows = []
    for T in sizes:
        gltf, _ = bench.make_workload(T, 10000)
        lib.loadGlTFscene(gltf.encode())
        assert lib.gotoCameraByName(b"compound-cam")
        lib.setCurrentEyeShaderName(b"single_dimension_fast")
        er.setRenderSize(lib, 10000, 1)
        lib.crSetRenderMode(*bench.MODES[mode])
        lib.setCurrentEyeSamplesPerOmmatidium(1024)
        pose = np.zeros(12, np.float32); lib.crDebugCopyCameraPose(pose.ctypes.data)
        poses = bench.poses_for(pose[:3].copy(), pose[3:].copy(), 20, 0)
        er.renderPoseBatch(lib, poses[:3]); er.renderPoseBatch(lib, poses)
        ms = []
        for _ in range(5):
            er.renderPoseBatch(lib, poses); ms.append(lib.crGetLastTraceMs())
        for _ in range(3):
            lib.renderFrame(); lib.getFramePointer()
        t0 = time.perf_counter()
        for k in range(20):
            lib.setCameraPosition(float(poses[k, 0]), float(poses[k, 1]), float(poses[k, 2]))
            lib.renderFrame(); lib.getFramePointer()
        abi = 20 * 10000 * 1024 / (time.perf_counter() - t0)
        rows.append({"triangles": int(lib.crDebugGetTriangleCount()), "bvh_build_ms": lib.crGetBvhBuildMs(),
                     "rays_per_sec_batched": 20 * 10000 * 1024 / (np.median(ms) * 1e-3), "rays_per_sec_per_frame_abi": abi})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    return {"config": "cfg4: speed-test terrain sweep, 10 000-ommatidia eye, S=1024", "mode": mode, "sweep": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="fused", choices=sorted(bench.MODES))
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--sizes", default="10000,100000,1000000,10000000")
    ap.add_argument("--only", default="cfg2,cfg3,cfg4")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import eye_renderer as er
    data = speed_test.fixtures()
    lib = er.load_library(device=0)
    lib.setVerbosity(False)
    res = {}
    if "cfg2" in a.only:
        res["cfg2"] = cfg2(lib, er, data, a.mode)
    if "cfg3" in a.only:
        res["cfg3"] = cfg3(lib, er, data, a.mode, a.frames)
    if "cfg4" in a.only:
        res["cfg4"] = cfg4(lib, er, a.mode, [int(x) for x in a.sizes.split(",")])
    bench.emit(res)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
