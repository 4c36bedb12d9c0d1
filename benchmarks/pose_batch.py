#!/usr/bin/env python
"""benchmarks/pose_batch.py -- BASELINE config 5: pose-batched rendering of the position-estimation
toy experiment (env_2.gltf, AM_60185-real eye geometry with heterogeneous acceptance angles,
positions uniform in the 50 mm cube, numpy default_rng(0)), 1..8 GPUs, NCCL allgather of the rows.

  python benchmarks/pose_batch.py [--poses 512] [--samples 64] [--chunk C]
  torchrun --nproc-per-node N benchmarks/pose_batch.py ...

--chunk C (> 0): render C poses at a time and issue each chunk's allgather asynchronously while the next chunk is
traced (sharding.ChunkedPoseGather); the result and its checksum are the same as with one gather.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "compound-ray_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "benchmarks")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=512)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--native", action="store_true", help="the library's own NCCL data plane (crRenderPoseBatchSharded, csrc/cr_comm.cpp): "
                    "every rank passes all poses; the C++ host shards, positions the streams and gathers")
    ap.add_argument("--id-file", default=None, help="--native without torch.distributed: file that carries the NCCL unique id")
    ap.add_argument("--mode", default="ordered", choices=["ordered", "fused", "fused_fast"])
    ap.add_argument("--reps", type=int, default=3, help="--native: repetitions of the whole job; `seconds` is their median")
    args = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import eye_renderer as er
    import sharding
    import speed_test
    from tools import synth
    data = speed_test.fixtures()
    use_torch = world > 1 and not (args.native and args.id_file)
    if use_torch:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = er.load_library(device=local)
    lib.setVerbosity(False)
    lib.crSetRenderMode({"ordered": 0, "fused": 1, "fused_fast": 1}[args.mode], 1 if args.mode == "fused_fast" else 0)
    lib.loadGlTFscene(os.path.join(data, "sim-environment", "env_2.gltf").encode())
    lib.gotoCameraByName(b"compound-cam")
    base = np.array([[*o.position, *o.direction, o.acceptanceAngle, o.focalpointOffset] for o in
                     er.readEyeFile(os.path.join(data, "sim-environment", "eyes", "AM_60185-real.eye"))], np.float32)
    omm = synth.heterogeneous_eye(base)
    er.setOmmatidiaFromArray(lib, omm)
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N, S, P = len(omm), args.samples, args.poses
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    pose = np.zeros(12, np.float32)
    lib.crDebugCopyCameraPose(pose.ctypes.data)
    pos = np.random.default_rng(0).uniform(-25, 25, (P, 3)).astype(np.float32)
    lo, hi = sharding.pose_block(rank, world, P)
    all_poses = er.make_poses(pos, x=pose[3:6], y=pose[6:9], z=pose[9:12])
    poses = all_poses[lo:hi]
    lib.crSetFirstFrame(lo)
    # warm-up: allocations, module load and -- a fresh process on an idle GPU needs several hundred ms of load before a batch
    # runs at its steady rate (the same 100 000 poses take 1.55 s in a process that has been rendering, 2.1-2.6 s right
    # after start-up: profiles/r03h_small_batch_ab_100k.json) -- clocks; then rewind the streams
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 1.5:
        er.renderPoseBatch(lib, poses[:min(4096, len(poses))])
    lib.crSetFirstFrame(lo)
    if args.native:
        sharding.init_library_comm(lib, rank, world, dist if use_torch else None, args.id_file)
        sharding.render_pose_batch_sharded(lib, all_poses[:min(1024 * world, P)], chunk=args.chunk)   # warm-up incl. the collective
        dev_out = None
        try:                                        # result stays in device memory on every rank, as in the torch legs below
            import torch
            torch.cuda.set_device(local)
            dev_out = torch.empty((P, N, 4), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
        except Exception:
            pass
        reps = []
        for rep in range(args.reps):                 # the same logical job, repeated (streams repositioned at frame 0 by the call itself)
            if use_torch:
                dist.barrier()
            t0 = time.perf_counter()
            if dev_out is not None:
                sharding.render_pose_batch_sharded(lib, all_poses, chunk=args.chunk, out_device_ptr=dev_out.data_ptr())
            else:
                rows, _ = sharding.render_pose_batch_sharded(lib, all_poses, chunk=args.chunk)   # all P rows on every rank, host side
            dt = time.perf_counter() - t0
            if use_torch:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            reps.append(dt)
        dt = float(np.median(reps))
        all_reps = reps
        checksum = int(dev_out.to(torch.int64).sum().item()) if dev_out is not None else int(rows.astype(np.int64).sum())
        lib.crCommDestroy()
    elif args.chunk > 0:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        cg = sharding.ChunkedPoseGather(rank, world, P, (N, 4), torch.uint8, "cuda", args.chunk, dist)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in range(len(cg.chunks)):
            a, b = cg.local_poses(c)
            if b > a:                                                       # rows land in the chunk's send slot
                er.renderPoseBatch(lib, poses[a:b], out_device_ptr=cg.send_rows(c).data_ptr())
            cg.issue(c)                                                     # async: overlaps the next chunk's trace
        rows = cg.finish()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        checksum = int(rows.to(torch.int64).sum().item())
    elif world > 1:
        blk = sharding.padded_block_size(world, P)
        send = torch.zeros((blk, N, 4), dtype=torch.uint8, device="cuda")
        gathered = torch.empty((world * blk, N, 4), dtype=torch.uint8, device="cuda")
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        er.renderPoseBatch(lib, poses, out_device_ptr=send.data_ptr())
        dist.all_gather_into_tensor(gathered.view(-1), send.view(-1))
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        checksum = int(sharding.unpad(gathered, world, P).to(torch.int64).sum().item())
    else:
        try:                                        # as the multi-GPU legs: the rows stay in device memory
            import torch
            torch.cuda.set_device(local)
            dev_out = torch.empty((P, N, 4), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            er.renderPoseBatch(lib, poses, out_device_ptr=dev_out.data_ptr())
            dt = time.perf_counter() - t0
            checksum = int(dev_out.to(torch.int64).sum().item())
        except ImportError:
            t0 = time.perf_counter()
            rows, _ = er.renderPoseBatch(lib, poses)
            dt = time.perf_counter() - t0
            checksum = int(rows.astype(np.int64).sum())
    if rank == 0:
        out = {"benchmark": "pose batch (BASELINE config 5)", "n_gpus": world, "poses": P, "ommatidia": N, "samples": S,
               "chunk": args.chunk, "mode": args.mode,
               "data_plane": "library (crRenderPoseBatchSharded: ncclBroadcast groups per chunk, C++)" if args.native else "torch.distributed", "seconds": dt, "poses_per_sec": P / dt, "rays_per_sec": P * N * S / dt, "ommatidia_frames_per_sec": P * N / dt,
               "seconds_all_repetitions": locals().get("all_reps"), "checksum": checksum, "timing": "host wall clock incl. pose upload, render and gather; every rank ends with all rows in DEVICE memory; max over ranks"}
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if use_torch:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
