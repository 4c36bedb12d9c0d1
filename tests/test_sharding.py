"""Multi-GPU host logic on CPU: pose sharding + allgather layout with torch.distributed (gloo,
world size 2).  Each rank renders its pose block with the CPU oracle (the product needs a GPU),
positions its RNG streams at its first frame, and the gathered result must equal the sequential
single-process run -- the property bench.py's N>1 path and crSetFirstFrame rely on."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, PKG


def test_pose_blocks_partition():
    import sharding
    for world in (1, 2, 3, 8):
        for P in (0, 1, 7, 8, 100):
            blocks = [sharding.pose_block(r, world, P) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == P
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) <= sharding.padded_block_size(world, P) if P else True
    assert [sharding.draws_before_frame(k) for k in range(6)] == [0, 3, 4, 7, 8, 11]


def _worker(rank, world, port, out_dir, data_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
    import torch
    import torch.distributed as dist
    import sharding
    from oracle import gltf_loader, oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    path = os.path.join(data_dir, "data", "test-scene", "test-scene.gltf")
    sc = gltf_loader.load_scene(path)
    cam = [c for c in sc.cameras if c.name == "insect-cam-2"][0]
    sh = O.SceneHandle(sc)
    P, S, N = 7, 6, len(cam.ommatidia)
    rng = np.random.default_rng(0)
    positions = rng.uniform(-1, 1, (P, 3)).astype(np.float32) + cam.position
    lo, hi = sharding.pose_block(rank, world, P)
    blk = sharding.padded_block_size(world, P)
    eye = O.CompoundEyeOracle(sh, cam.ommatidia, O.pose_from_camera(cam), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    eye.set_first_frame(lo)                                    # == crSetFirstFrame(lo) in the product
    send = torch.zeros((blk, N, 4), dtype=torch.uint8)
    for i, p in enumerate(range(lo, hi)):
        eye.pose = O.make_pose(positions[p], cam.x_axis, cam.y_axis, cam.z_axis)
        send[i] = torch.from_numpy(eye.render_frame(method="brute")[0].copy())
    gathered = torch.zeros((world * blk, N, 4), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered.view(-1), send.view(-1))
    result = sharding.unpad(gathered.numpy(), world, P)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), result)
    # the same run with the gather issued in chunks of 2 poses while the next chunk renders
    eye.set_first_frame(lo)
    cg = sharding.ChunkedPoseGather(rank, world, P, (N, 4), torch.uint8, "cpu", 2, dist)
    assert cg.chunks == [(0, 2), (2, 4)] and cg.blk == blk
    for c in range(len(cg.chunks)):
        a, b = cg.local_poses(c)
        rows = cg.send_rows(c)
        for i in range(a, b):
            eye.pose = O.make_pose(positions[lo + i], cam.x_axis, cam.y_axis, cam.z_axis)
            rows[i - cg.chunks[c][0]] = torch.from_numpy(eye.render_frame(method="brute")[0].copy())
        cg.issue(c)
    np.save(os.path.join(out_dir, f"chunked_rank{rank}.npy"), cg.finish().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_equals_sequential_run(ref_data, oracle, loader, tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path), ref_data), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy"); r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1), "every rank holds the full result"
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"chunked_rank{r}.npy"), r0), "chunked, overlapped gather = single gather"
    # sequential reference: one process, frames 0..P-1
    path = os.path.join(ref_data, "data", "test-scene", "test-scene.gltf")
    sc = loader.load_scene(path)
    cam = [c for c in sc.cameras if c.name == "insect-cam-2"][0]
    sh = oracle.SceneHandle(sc)
    P, S, N = 7, 6, len(cam.ommatidia)
    positions = np.random.default_rng(0).uniform(-1, 1, (P, 3)).astype(np.float32) + cam.position
    eye = oracle.CompoundEyeOracle(sh, cam.ommatidia, oracle.pose_from_camera(cam), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    for p in range(P):
        eye.pose = oracle.make_pose(positions[p], cam.x_axis, cam.y_axis, cam.z_axis)
        assert np.array_equal(r0[p], eye.render_frame(method="brute")[0]), f"pose {p}"


def _omm_worker(rank, world, port, out_dir, data_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
    import torch
    import torch.distributed as dist
    import sharding
    from oracle import gltf_loader, oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    path = os.path.join(data_dir, "data", "test-scene", "test-scene.gltf")
    sc = gltf_loader.load_scene(path)
    cam = [c for c in sc.cameras if c.name == "insect-cam-2"][0]
    sh = O.SceneHandle(sc)
    omm = np.asarray(cam.ommatidia, np.float32).reshape(-1, 8)
    N, S = len(omm), 5
    lo, hi = sharding.ommatidia_block(rank, world, N)
    eye = O.CompoundEyeOracle(sh, omm[lo:hi], O.pose_from_camera(cam), "single_dimension_fast", samples=S)
    eye.set_render_size(hi - lo, 1)
    eye.set_shard(N, lo)                                       # == crSetOmmatidialShard(N, lo) in the product
    out = []
    for frame in range(2):
        eye.render_frame(method="brute", project=False)
        out.append(sharding.allgather_rows(torch.from_numpy(eye.last["summed"].copy()), world, N, dist).numpy())
    np.save(os.path.join(out_dir, f"omm_rank{rank}.npy"), np.stack(out))
    dist.barrier()
    dist.destroy_process_group()


def test_ommatidium_range_shards_equal_the_whole_eye(ref_data, oracle, loader, tmp_path):
    """SURVEY 8e secondary partition: ranks own ommatidium ranges of ONE pose; stream ids stay global, so the
    allgathered per-ommatidium float RGB equals the unsharded frame bit for bit (frames 0 and 1)."""
    import sharding
    import torch.multiprocessing as mp
    for world in (2, 3, 7):
        blocks = [sharding.ommatidia_block(r, world, 100) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 100 and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    world, port = 2, 31500 + (os.getpid() % 2000)
    mp.spawn(_omm_worker, args=(world, port, str(tmp_path), ref_data), nprocs=world, join=True)
    r0 = np.load(tmp_path / "omm_rank0.npy"); r1 = np.load(tmp_path / "omm_rank1.npy")
    assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32)), "every rank holds the full result"
    path = os.path.join(ref_data, "data", "test-scene", "test-scene.gltf")
    sc = loader.load_scene(path)
    cam = [c for c in sc.cameras if c.name == "insect-cam-2"][0]
    eye = oracle.CompoundEyeOracle(oracle.SceneHandle(sc), cam.ommatidia, oracle.pose_from_camera(cam), "single_dimension_fast", samples=5)
    for frame in range(2):
        eye.render_frame(method="brute", project=False)
        assert np.array_equal(r0[frame].view(np.uint32), eye.last["summed"].view(np.uint32)), f"frame {frame}"
