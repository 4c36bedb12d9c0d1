#!/usr/bin/env python
"""benchmarks/small_batch_ab.py -- pose batches of SMALL frames (BASELINE config 5's eye: 6 374 ommatidia x 64 samples on
env_2.gltf) with the frame groups and the candidate lists on and off, in one process, repeated: wall clock and the
library's own event time per call of crRenderPoseBatch."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "compound-ray_b200"), os.path.join(ROOT, "benchmarks")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import eye_renderer as er
    import speed_test
    from tools import synth
    poses_n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    data = speed_test.fixtures()
    lib = er.load_library(device=0)
    lib.setVerbosity(False)
    lib.crSetRenderMode(1, 0)
    lib.loadGlTFscene(os.path.join(data, "sim-environment", "env_2.gltf").encode())
    lib.gotoCameraByName(b"compound-cam")
    base = np.array([[*o.position, *o.direction, o.acceptanceAngle, o.focalpointOffset] for o in
                     er.readEyeFile(os.path.join(data, "sim-environment", "eyes", "AM_60185-real.eye"))], np.float32)
    er.setOmmatidiaFromArray(lib, synth.heterogeneous_eye(base))
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N, S = lib.getCurrentEyeOmmatidialCount(), 64
    er.setRenderSize(lib, N, 1)
    pose = np.zeros(12, np.float32)
    lib.crDebugCopyCameraPose(pose.ctypes.data)
    pos = np.random.default_rng(0).uniform(-25, 25, (poses_n, 3)).astype(np.float32)
    poses = er.make_poses(pos, x=pose[3:6], y=pose[6:9], z=pose[9:12])
    import torch
    out = torch.empty((poses_n, N, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    rows = []
    ref = None
    import sharding
    lib.crDebugSetFrameGroups(1)
    lib.crDebugSetCandidateLists(0)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    for chunk in (0, 2048, 512):                        # the library's sharded entry point on one GPU (no communicator): per-chunk calls
        for rep in range(2):
            lib.crSetFirstFrame(0)
            t0 = time.perf_counter()
            _, ms = sharding.render_pose_batch_sharded(lib, poses, chunk=chunk, out_device_ptr=out.data_ptr())
            wall = time.perf_counter() - t0
            row = {"sharded_entry_point": True, "chunk": chunk, "rep": rep, "wall_s": wall, "returned_ms": ms, "events_s": lib.crGetLastTraceMs() * 1e-3,
                   "grays_wall": poses_n * N * S / wall / 1e9}
            print(json.dumps(row), flush=True)
    for rep in range(2):
        for groups, lists in ((0, 1), (1, 1), (1, 0), (0, 0)):
            lib.crDebugSetFrameGroups(groups)
            lib.crDebugSetCandidateLists(lists)
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            er.renderPoseBatch(lib, poses[:16], out_device_ptr=out.data_ptr())
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            t0 = time.perf_counter()
            er.renderPoseBatch(lib, poses, out_device_ptr=out.data_ptr())
            wall = time.perf_counter() - t0
            ev = lib.crGetLastTraceMs() * 1e-3
            chk = int(out.to(torch.int64).sum().item())
            ref = chk if ref is None else ref
            row = {"rep": rep, "frame_groups": groups, "candidate_lists": lists, "wall_s": wall, "events_s": ev,
                   "grays_wall": poses_n * N * S / wall / 1e9, "grays_events": poses_n * N * S / ev / 1e9, "same_checksum": chk == ref}
            rows.append(row)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
