"""GPU parity at the BASELINE.json configuration sizes (cfg3/cfg4/cfg5), through the C ABI.

Full-size frames cannot be brute-forced on the CPU, so these use (i) the oracle's own SAH BVH
(independent of the product's LBVH) on the full scene and (ii) size-independent properties:
batch == per-frame API, sharded start == sequential run, variance ~ 1/S."""
import os

import numpy as np
import pytest

from conftest import load_oracle_scene

pytestmark = pytest.mark.gpu

HIT4 = np.dtype([("prim", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])


@pytest.fixture(scope="module")
def terrain(synth_dir):
    from tools import synth
    gltf = os.path.join(synth_dir, "terrain1m.gltf")
    synth.write_eye(os.path.join(synth_dir, "eye10k.eye"), synth.fibonacci_eye(10000))
    info = synth.write_terrain_gltf(gltf, triangles=1_000_000, eye_file="eye10k.eye")
    assert info["triangles"] >= 1_000_000
    return gltf


def test_cfg4_million_triangle_terrain_10k_eye(lib, er, loader, oracle, terrain):
    """Headline workload geometry: 1M-triangle terrain, 10k-ommatidia eye.  One S=8 frame (80k rays):
    rays, hit ids, (t,u,v), per-ommatidium RGB and the frame are bit-exact vs the oracle, whose hits
    come from its own binned-SAH BVH over the same million triangles."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.crDebugGetTriangleCount() >= 1_000_000
    assert lib.gotoCameraByName(b"compound-cam")
    N, S = 10000, 8
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    sc, sh, ocam = load_oracle_scene(loader, oracle, terrain, "compound-cam")
    T = lib.crDebugGetTriangleCount()
    tris = np.zeros((T, 9), np.float32)
    lib.crDebugCopyTriangles(tris.ctypes.data)
    assert np.array_equal(tris.view(np.uint32), sc.tris.view(np.uint32))
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    lib.crDebugSetRayDump(True)
    for frame in range(2):
        lib.renderFrame()
        eye.render_frame(method="bvh")
        o = np.zeros((N * S, 3), np.float32); d = np.zeros((N * S, 3), np.float32); h = np.zeros(N * S, HIT4)
        assert lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data) == N * S
        assert np.array_equal(d.view(np.uint32), eye.last["dirs"].view(np.uint32))
        assert np.array_equal(h["prim"], eye.last["hits"]["prim"])
        hit = h["prim"] >= 0
        assert 0.3 < hit.mean() < 0.7
        assert np.array_equal(h["t"][hit].view(np.uint32), eye.last["hits"]["t"][hit].view(np.uint32))
        assert np.array_equal(er.getOmmatidialData(lib).view(np.uint32), eye.last["summed"].view(np.uint32))
        assert np.array_equal(er.getFrame(lib, N, 1), eye.frame)
    lib.crDebugSetRayDump(False)
    # the device BVH traversed by the oracle's instrumented loop gives the same hits (roofline counters)
    nn = lib.crDebugGetBvhNodeCount()
    nodes = np.zeros((nn, 16), np.float32); dtris = np.zeros((T, 12), np.float32)
    lib.crDebugCopyBvh(nodes.ctypes.data, dtris.ctypes.data)
    hits2, cnt = oracle.trace_device_bvh(nodes, dtris, o, d, np.zeros(N * S, np.float32))
    assert np.array_equal(hits2["prim"], h["prim"]) and cnt[0] / (N * S) < 40


def _stress_eye(n, seed):
    """Ommatidia that exercise every branch of the entry-frontier pass: zero / narrow / wide / very wide
    acceptance angles, focal offsets, positions off the eye centre, exactly vertical axes, axes that are
    not unit length (frontier must fall back to the root) and the perp-sum==0 quirk (az == ax)."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    omm = np.zeros((n, 8), np.float32)
    omm[:, 0:3] = d * 0.05 + rng.normal(scale=0.01, size=(n, 3))
    omm[:, 3:6] = d
    omm[:, 6] = np.radians(rng.choice([0.0, 0.05, 1.0, 2.3, 8.0, 20.0, 35.0, 60.0, 120.0], size=n))
    omm[:, 7] = rng.choice([0.0, 0.0, 0.01, 0.5, -0.2, -8.0], size=n)   # negative: tmin < 0, hits behind the origin count
    omm[0, 3:6] = (0, 1, 0); omm[1, 3:6] = (0, -1, 0)
    omm[2, 3:6] = (0.6, 0.52915026, 0.6)                 # az == ax: perp falls back to (0,0,1), not perpendicular
    omm[3:40, 3:6] *= rng.uniform(0.5, 2.0, size=(37, 1)).astype(np.float32)   # non-unit axes
    omm[40:60, 3:6] *= np.float32(1.00002)               # inside the unit-length tolerance
    return omm


def test_entry_frontier_equals_root_traversal(lib, er, terrain):
    """The per-ommatidium entry frontier (k_buildEntries) only removes work: with it switched on and off,
    the same RNG streams give bit-identical hits (prim, t, u, v), ommatidial RGB and batch rows, for a stress eye
    at poses above, inside, beside and far from the terrain, axis-aligned and rotated."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    N, S = 2048, 32
    omm = _stress_eye(N, 7)
    er.setOmmatidiaFromArray(lib, omm)
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, N, 1)
    rng = np.random.default_rng(11)
    poses = []
    for pos in ([0, 12, 0], [3.3, 0.7, -41.0], [49.9, 2.0, 49.9], [0, -5, 0], [400, 50, 0], [0.001, 3.0, 0.001], [-20, 0.05, 7]):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        poses.append(np.concatenate([np.float32(pos), q[:, 0], q[:, 1], q[:, 2]]))
        poses.append(np.concatenate([np.float32(pos), [1, 0, 0, 0, 1, 0, 0, 0, 1]]))
    poses = np.asarray(poses, np.float32)
    res = {}
    for mode in (0, 1):
        lib.crDebugSetEntryFrontier(mode, 8, 0)
        lib.setCurrentEyeSamplesPerOmmatidium(S)          # resets the streams: both modes draw the same rays
        lib.crDebugSetRayDump(True)
        per_pose = []
        for p in poses:
            lib.setCameraPosition(float(p[0]), float(p[1]), float(p[2]))
            er.setCameraLocalSpace(lib, p[3:12].reshape(3, 3).T)
            lib.renderFrame()
            o = np.zeros((N * S, 3), np.float32); d = np.zeros((N * S, 3), np.float32); h = np.zeros(N * S, HIT4)
            assert lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data) == N * S
            cnt = np.zeros((N * S, 2), np.int32)
            assert lib.crDebugCopyLastRayCounts(cnt.ctypes.data) == N * S
            per_pose.append((d.copy(), h.copy(), er.getOmmatidialData(lib).copy(), cnt.sum(axis=0)))
        lib.crDebugSetRayDump(False)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        rows, _ = er.renderPoseBatch(lib, poses)
        res[mode] = (per_pose, rows)
    lib.crDebugSetEntryFrontier(1, 2, 0)
    n_hits = 0
    visits = np.zeros((2, 2), np.int64)
    for (d0, h0, c0, n0), (d1, h1, c1, n1) in zip(res[0][0], res[1][0]):
        visits[0] += n0; visits[1] += n1
        assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
        assert np.array_equal(h0["prim"], h1["prim"])
        hit = h0["prim"] >= 0
        n_hits += int(hit.sum())
        for k in ("t", "u", "v"):
            assert np.array_equal(h0[k][hit].view(np.uint32), h1[k][hit].view(np.uint32)), k
        assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32))
    assert n_hits > 100000
    assert np.array_equal(res[0][1], res[1][1])
    assert visits[1, 0] < visits[0, 0], "the frontier must only remove node fetches"


def test_entry_frontier_fuzz(lib):
    """A slice of compound-ray_b200/tools/frontier_fuzz.py (random soups x eyes x poses, frontier off vs on, all bits
    equal).  The full campaign -- 4 800 scenes, 14 400 poses, 118 M rays -- ran clean when the pass was added."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compound-ray_b200", "tools",
                                                     "frontier_fuzz.py"), "--configs", "120", "--seed", "9"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["mismatching_frames"] == 0 and res["hits"] > 50000


def test_bvh_fuzz_against_bruteforce(lib):
    """A slice of compound-ray_b200/tools/bvh_fuzz.py: device BVH vs the oracle's brute-force loop on random soups and
    adversarial rays, bit-identical hits; only numerically degenerate ray/triangle pairs (|det| < 1e-5 |d||e1||e2|,
    where binary32 Moller-Trumbore is itself off by per cent) may differ and are counted."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compound-ray_b200", "tools",
                                                     "bvh_fuzz.py"), "--configs", "80", "--seed", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["mismatching_configs"] == 0 and res["hits"] > 50000 and res["degenerate_differences"] <= 4


def test_ommatidium_range_shards_equal_the_whole_eye(lib, er, loader, oracle, terrain):
    """crSetOmmatidialShard: three uneven ommatidium ranges rendered one after another reproduce the per-ommatidium
    float RGB and 8-bit rows of the unsharded eye bit for bit (frames 0 and 1), and shard 1 equals the oracle."""
    import sharding
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    sc, sh, ocam = load_oracle_scene(loader, oracle, terrain, "compound-cam")
    omm = np.asarray(ocam.ommatidia, np.float32).reshape(-1, 8)[:1001].copy()
    N, S = len(omm), 16
    er.setOmmatidiaFromArray(lib, omm)
    lib.crSetOmmatidialShard(0, 0)
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    whole = []
    for frame in range(2):
        lib.renderFrame()
        whole.append((er.getOmmatidialData(lib).copy(), er.getFrame(lib, N, 1).copy()))
    world = 3
    parts = [[], []]
    for rank in range(world):
        lo, hi = sharding.configure_ommatidia_shard(lib, er, omm, rank, world)
        er.setRenderSize(lib, hi - lo, 1)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        if rank == 1:
            eye = oracle.CompoundEyeOracle(sh, omm[lo:hi], oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
            eye.set_shard(N, lo)
        for frame in range(2):
            lib.renderFrame()
            rgb = er.getOmmatidialData(lib).copy()
            parts[frame].append((rgb, er.getFrame(lib, hi - lo, 1).copy()))
            if rank == 1:
                eye.render_frame(method="bvh", project=False)
                assert np.array_equal(rgb.view(np.uint32), eye.last["summed"].view(np.uint32))
    lib.crSetOmmatidialShard(0, 0)
    for frame in range(2):
        rgb = np.concatenate([p[0] for p in parts[frame]]); row = np.concatenate([p[1] for p in parts[frame]], axis=1)
        assert np.array_equal(rgb.view(np.uint32), whole[frame][0].view(np.uint32)), f"frame {frame} RGB"
        assert np.array_equal(row, whole[frame][1]), f"frame {frame} row"


def test_cfg4_full_sample_count_properties(lib, er, terrain):
    """S=1024 on the headline workload (10.24M rays/frame): the batch path, the per-frame ABI and a
    restarted (sharded) run produce byte-identical rows."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N = 10000
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(1024)
    pose = np.zeros(12, np.float32)
    lib.crDebugCopyCameraPose(pose.ctypes.data)
    pos = pose[:3] + np.array([[0, 0, 0], [1, 0.5, -2], [-3, 1, 0.5], [0.2, 2, 4]], np.float32)
    poses = er.make_poses(pos, x=pose[3:6], y=pose[6:9], z=pose[9:12])
    rows, _ = er.renderPoseBatch(lib, poses)
    lib.setCurrentEyeSamplesPerOmmatidium(1024)                      # reset streams, then frame by frame
    for p in range(4):
        lib.setCameraPosition(*[float(v) for v in pos[p]])
        lib.renderFrame()
        assert np.array_equal(er.getFrame(lib, N, 1)[0], rows[p]), f"frame {p}"
    lib.crSetFirstFrame(2)
    rows2, _ = er.renderPoseBatch(lib, poses[2:])
    lib.crSetFirstFrame(0)
    assert np.array_equal(rows2, rows[2:])
    assert rows[:, :, :3].std() > 5 and (rows[:, :, 3] == 255).all()


def test_cfg3_variance_falls_with_samples(lib, er, synth_dir):
    """minimumSampleRateFinder-style statistic (data/tools/minimumSampleRateFinder.py:36-47): the
    frame-to-frame SD of an ommatidium's output falls ~ 1/sqrt(S).  12-ommatidia 1-steradian eye."""
    from tools import synth
    gltf = os.path.join(synth_dir, "terrain_small.gltf")
    synth.write_eye(os.path.join(synth_dir, "ico.eye"), synth.fibonacci_eye(12))
    synth.write_terrain_gltf(gltf, triangles=20000, eye_file="ico.eye")
    lib.loadGlTFscene(gltf.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    er.setOmmatidiaFromOmmatidiumList(lib, er.getIcoOmmatidia())
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, 12, 1)
    sds = {}
    for S in (16, 256):
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        lib.renderFrame()
        vals = []
        for _ in range(200):
            lib.renderFrame()
            vals.append(er.getOmmatidialData(lib))
        sds[S] = np.stack(vals).std(axis=0).mean()
    ratio = sds[16] / sds[256]
    assert 3.0 < ratio < 5.4, (sds, ratio)                           # expected sqrt(256/16) = 4


def test_cfg5_heterogeneous_eye_pose_batch(lib, er, loader, oracle, ref_data):
    """env_2.gltf + AM_60185-real geometry with log-uniform acceptance angles (numpy rng 5), random
    positions in the 50 mm cube (rng 0): batch rows == oracle frames within the texture tolerance,
    hit ids exact."""
    from tools import synth
    path = os.path.join(ref_data, "sim-environment", "env_2.gltf")
    lib.loadGlTFscene(path.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    sc, sh, ocam = load_oracle_scene(loader, oracle, path, "compound-cam")
    omm = synth.heterogeneous_eye(ocam.ommatidia)
    assert len(omm) == 6374 and omm[:, 6].min() >= 0.02 and omm[:, 6].max() <= 0.35
    er.setOmmatidiaFromArray(lib, omm)
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N, S, P = len(omm), 16, 3
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    pos = np.random.default_rng(0).uniform(-25, 25, (P, 3)).astype(np.float32)
    poses = er.make_poses(pos, x=ocam.x_axis, y=ocam.y_axis, z=ocam.z_axis)
    rows, _ = er.renderPoseBatch(lib, poses)
    eye = oracle.CompoundEyeOracle(sh, omm, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    for p in range(P):
        eye.pose = oracle.make_pose(pos[p], ocam.x_axis, ocam.y_axis, ocam.z_axis)
        eye.render_frame(method="bvh")
        diff = np.abs(rows[p].astype(np.int32) - eye.frame[0].astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.02, (p, diff.max(), (diff > 0).mean())


def test_batched_iterators_match_the_reference_loop(lib, er, ref_data):
    """compound-ray_b200/iterators.py vs the loop of compoundRayIterators.py:84-102 / :121-143
    (setCameraPosition + renderFrame + getFramePointer per item): identical images and positions."""
    import iterators
    scene = os.path.join(ref_data, "sim-environment", "env_2.gltf")
    eye = os.path.join(ref_data, "sim-environment", "eyes", "AM_60185-real.eye")
    S = 24
    # reference-style loop through the plain ABI
    lib.loadGlTFscene(scene.encode())
    lib.gotoCameraByName(b"compound-cam")
    cfg = er.readEyeFile(eye)
    er.setOmmatidiaFromOmmatidiumList(lib, cfg)
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, len(cfg), 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    np.random.seed(123)
    want = []
    for _ in range(7):
        rel = (np.random.random(3) * 2 - 1) * 25.0
        lib.setCameraPosition(*[float(v) for v in rel])
        lib.renderFrame()
        want.append((np.copy(lib.getFramePointer()[:, :, :3]), rel))
    np.random.seed(123)
    it = iter(iterators.RandomCubeIterator(eye, scenePath=scene, samples=S, blockSize=4))
    for k in range(7):
        img, pos = next(it)
        assert img.shape == (1, len(cfg), 3)
        assert np.array_equal(img.numpy(), want[k][0].astype(np.float32)), k
        assert np.allclose(pos.numpy(), want[k][1].astype(np.float32))
    uit = iter(iterators.UniformCubeIterator(eye, scenePath=scene, samples=S, blockSize=3, samplingSize=2, cubeSize=10))
    seen = [next(uit) for _ in range(9)]                       # wraps after 8 lattice points
    assert [tuple(c) for _, _, c in seen[:8]] == [(x, y, z) for z in (0, 1) for y in (0, 1) for x in (0, 1)]
    assert tuple(seen[8][2]) == (0, 0, 0) and seen[0][0].shape == (1, len(cfg))
    gap = 10 / 3
    assert np.allclose(seen[3][1].numpy(), np.array([1, 1, 0]) * gap - gap)


def test_primary_example_call_sequence(lib, er, ref_data, tmp_path):
    """The call sequence of python-examples/primary-example.py:20-98: every camera of the
    test scene is rendered, 'displayed', saved as PPM and read back; compound eyes again at S=100."""
    lib.loadGlTFscene(os.path.join(ref_data, "data", "test-scene", "test-scene.gltf").encode())
    W = H = 200
    er.setRenderSize(lib, W, H)
    n = lib.getCameraCount()
    assert n == 6
    for i in range(n):
        lib.renderFrame()
        lib.displayFrame()                                               # headless no-op
        ppm = tmp_path / f"cam{i}.ppm"
        lib.saveFrameAs(str(ppm).encode())
        frame = np.copy(lib.getFramePointer())
        assert frame.shape == (H, W, 4)
        raw = ppm.read_bytes()
        assert raw.startswith(b"P6\n200 200\n255\n")
        rgb = np.frombuffer(raw[len(b"P6\n200 200\n255\n"):], np.uint8).reshape(H, W, 3)
        assert np.array_equal(rgb, frame[::-1, :, :3])                   # PPM is top-down, the frame bottom-up
        assert frame[:, :, :3].std() > 1                                 # something was drawn
        if lib.isCompoundEyeActive():
            lib.setCurrentEyeSamplesPerOmmatidium(100)
            lib.renderFrame()
            lib.saveFrameAs(str(tmp_path / f"cam{i}_100.ppm").encode())
            smooth = np.copy(lib.getFramePointer())
            assert smooth.shape == (H, W, 4) and (smooth[:, :, 3] == 255).all()
        lib.nextCamera()
    assert lib.getCurrentCameraIndex() == 0
    lib.stop()
    assert lib.getCameraCount() == 0
    lib.loadGlTFscene(os.path.join(ref_data, "data", "test-scene", "test-scene.gltf").encode())   # usable again after stop()
    assert lib.getCameraCount() == 6 and lib.renderFrame() > 0
