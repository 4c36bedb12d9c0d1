"""GPU tests of the round-2 throughput paths of the compound trace kernel, through the C ABI:

  * the per-ommatidium candidate lists (k_buildEntries stage 2 + traceList: the 32 samples a warp holds of one ommatidium
    test a flat list of pre-leaf BVH nodes in lockstep instead of walking the tree per lane) must return the closest
    hit of the per-lane walk BIT FOR BIT -- prim, t, u, v, per-ommatidium RGB, batch rows -- and the oracle's hits;
  * the fused reduction (crSetRenderMode(1, .)) must equal the checker's restatement of its fixed addition order
    (oracle.fused_sum) BIT FOR BIT, and the reference's sequential order within fp32 rounding;
  * the fast-math mode (crSetRenderMode(., 1)) is held to the north_star tolerance against the checker:
    max |dRGB| <= 1/255 and mean |dRGB| <= 1e-4 per ommatidium, 8-bit frames within one step.
"""
import os

import numpy as np
import pytest

from conftest import load_oracle_scene

pytestmark = pytest.mark.gpu

HIT4 = np.dtype([("prim", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])


@pytest.fixture(scope="module")
def terrain(synth_dir):
    from tools import synth
    gltf = os.path.join(synth_dir, "terrain1m_modes.gltf")
    synth.write_eye(os.path.join(synth_dir, "eye10k_modes.eye"), synth.fibonacci_eye(10000))
    synth.write_terrain_gltf(gltf, triangles=1_000_000, eye_file="eye10k_modes.eye")
    return gltf


@pytest.fixture(autouse=True)
def _default_modes(lib):
    yield
    lib.crSetRenderMode(0, 0)
    lib.crDebugSetCandidateLists(2)
    lib.crDebugSetWavefront(0, 24, 0.35)
    lib.crDebugSetNodeLanes(16)
    lib.crDebugSetDynamicChunks(1)
    lib.crDebugSetSmAffine(1, 48)
    lib.crDebugSetZeroCopy(1)
    lib.crDebugSetReadAhead(1, 1.5)
    lib.crDebugSetFrameGroups(1)
    lib.crDebugSetRayDump(False)
    lib.crDebugSetEntryFrontier(1, 2, 0)
    lib.crSetFirstFrame(0)


def _dump(lib, n):
    o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); h = np.zeros(n, HIT4)
    assert lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data) == n
    cnt = np.zeros((n, 2), np.int32)
    assert lib.crDebugCopyLastRayCounts(cnt.ctypes.data) == n
    return o, d, h, cnt


def _same_hits(a, b):
    if not np.array_equal(a["prim"], b["prim"]):
        return False
    hit = a["prim"] >= 0
    return all(np.array_equal(a[k][hit].view(np.uint32), b[k][hit].view(np.uint32)) for k in ("t", "u", "v"))


def test_candidate_lists_equal_per_lane_walk_and_oracle(lib, er, loader, oracle, terrain):
    """Headline geometry (1M-triangle terrain, 10k-ommatidia eye of 0.04 rad acceptance), S=32 so that every warp is
    one ommatidium: candidate lists == per-lane walk from the frontier == per-lane walk from the root == oracle (own
    SAH BVH)."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    N, S = 10000, 32
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, N, 1)
    sc, sh, ocam = load_oracle_scene(loader, oracle, terrain, "compound-cam")
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    eye.render_frame(method="bvh")
    eye.render_frame(method="bvh")                               # frame 1: the cached Box-Muller half is in play
    res = {}
    for lists, frontier in ((0, 1), (1, 1), (0, 0)):
        lib.crDebugSetCandidateLists(2 * lists)
        lib.crDebugSetEntryFrontier(frontier, 2, 0)
        lib.setCurrentEyeSamplesPerOmmatidium(S)                 # restart the streams
        lib.crDebugSetRayDump(True)
        lib.renderFrame(); lib.renderFrame()
        o, d, h, cnt = _dump(lib, N * S)
        res[(lists, frontier)] = (h, er.getOmmatidialData(lib).copy(), er.getFrame(lib, N, 1).copy(), cnt)
        lib.crDebugSetRayDump(False)
        assert np.array_equal(d.view(np.uint32), eye.last["dirs"].view(np.uint32))
        assert _same_hits(h, eye.last["hits"]), f"hits differ from the oracle (lists {lists}, frontier {frontier})"
        assert np.array_equal(res[(lists, frontier)][1].view(np.uint32), eye.last["summed"].view(np.uint32))
        assert np.array_equal(res[(lists, frontier)][2], eye.frame)
    walk, listed = res[(0, 1)], res[(1, 1)]
    assert 0.3 < (walk[0]["prim"] >= 0).mean() < 0.7
    # most ground-looking ommatidia of this eye get a list: their rays then count the list's elements as node fetches and
    # every listed leaf they pass as triangle tests -- the same order of magnitude as the walk, far below a root walk
    n_walk, n_list, n_root = walk[3].sum(axis=0), listed[3].sum(axis=0), res[(0, 0)][3].sum(axis=0)
    assert n_list[0] < 2.5 * n_walk[0] and n_list[1] < 2.5 * n_walk[1], (n_walk, n_list)
    assert n_root[0] > 1.5 * n_walk[0]
    assert not np.array_equal(walk[3], listed[3]), "the candidate lists were not used"
    # structure of the lists of the last frame: header in {-1, 0..15}; elements name internal nodes and a non-empty leaf mask
    lib.crDebugSetCandidateLists(2)
    lib.crDebugSetEntryFrontier(1, 2, 0)
    lib.renderFrame()
    rec = np.zeros((N, 16), np.int32)
    assert lib.crDebugCopyCandidateLists(rec.ctypes.data, N) == N
    hdr = rec[:, 0]
    assert hdr.min() >= -1 and hdr.max() <= 15
    nn = lib.crDebugGetBvhNodeCount()
    for o in np.flatnonzero(hdr > 0)[:2000]:
        el = rec[o, 1:1 + hdr[o]]
        assert ((el >> 2) >= 0).all() and ((el >> 2) < nn).all() and ((el & 3) != 0).all()
        assert len(set((el >> 2).tolist())) == len(el), "a node listed twice"
    sky = np.asarray(ocam.ommatidia)[:, 4] > 0.5                       # steeply upward-looking ommatidia see only sky
    assert (hdr[sky] == 0).all()
    assert (hdr > 0).mean() > 0.25, "fewer than a quarter of the ommatidia got a candidate list"


def _stress_eye(n, seed):
    """Zero / narrow / wide cones, focal offsets (negative too), vertical and non-unit axes, the perp-sum quirk."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    omm = np.zeros((n, 8), np.float32)
    omm[:, 0:3] = d * 0.05 + rng.normal(scale=0.01, size=(n, 3))
    omm[:, 3:6] = d
    omm[:, 6] = np.radians(rng.choice([0.0, 0.05, 1.0, 2.3, 4.0, 8.0, 20.0, 35.0, 120.0], size=n))
    omm[:, 7] = rng.choice([0.0, 0.0, 0.01, 0.5, -0.2, -8.0], size=n)
    omm[0, 3:6] = (0, 1, 0); omm[1, 3:6] = (0, -1, 0)
    omm[2, 3:6] = (0.6, 0.52915026, 0.6)
    omm[3:40, 3:6] *= rng.uniform(0.5, 2.0, size=(37, 1)).astype(np.float32)
    return omm


def test_candidate_lists_stress_eye_and_batches(lib, er, terrain):
    """Mixed eye (listed and walking warps side by side, cones crossing octant boundaries, wide cones, negative focal
    offsets, non-unit axes, origins inside and far from the terrain, rotated poses): hits, RGB and pose-batch rows
    identical with the candidate lists on and off."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    N, S = 2048, 64
    er.setOmmatidiaFromArray(lib, _stress_eye(N, 19))
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, N, 1)
    rng = np.random.default_rng(5)
    poses = []
    for pos in ([0, 12, 0], [3.3, 0.7, -41.0], [0, -5, 0], [400, 50, 0], [-20, 0.05, 7], [0.2, 10.3, 0.1]):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        poses.append(np.concatenate([np.float32(pos), q[:, 0], q[:, 1], q[:, 2]]))
        poses.append(np.concatenate([np.float32(pos), [1, 0, 0, 0, 1, 0, 0, 0, 1]]))
    poses = np.asarray(poses, np.float32)
    res = {}
    for lists in (0, 1):
        lib.crDebugSetCandidateLists(2 * lists)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        lib.crDebugSetRayDump(True)
        frames = []
        for p in poses:
            lib.setCameraPosition(float(p[0]), float(p[1]), float(p[2]))
            er.setCameraLocalSpace(lib, p[3:12].reshape(3, 3).T)
            lib.renderFrame()
            _, d, h, cnt = _dump(lib, N * S)
            frames.append((d.copy(), h.copy(), er.getOmmatidialData(lib).copy(), cnt.copy()))
        lib.crDebugSetRayDump(False)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        rows, _ = er.renderPoseBatch(lib, poses)
        res[lists] = (frames, rows)
    used = 0
    for k, (a, b) in enumerate(zip(res[0][0], res[1][0])):
        assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)), f"directions, pose {k}"
        assert _same_hits(a[1], b[1]), f"hits, pose {k}"
        assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32)), f"RGB, pose {k}"
        used += int(not np.array_equal(a[3], b[3]))
    assert np.array_equal(res[0][1], res[1][1]), "batch rows"
    assert used >= len(poses) // 2, "the candidate lists were hardly used"
    assert sum((f[1]["prim"] >= 0).sum() for f in res[0][0]) > 10000


def test_wavefront_queue_changes_no_bit(lib, er, loader, oracle, terrain):
    """The warp-frames without a candidate list go through the wavefront queue (k_traceQueue with dynamic ray fetch,
    k_shadeQueue) instead of the per-lane walk inside the trace kernel, and both walks switch between their node and leaf
    phases on a lane-count vote (crDebugSetNodeLanes; 1 = classic while-while): per-ommatidium float RGB, 8-bit rows and
    batch rows are bit-identical with the queue off, on, on with other refill and phase thresholds, and with a queue so
    small that most warps find it full and walk inline -- in the ordered and in the fused reduction, per frame and batched,
    with the trace kernel's units handed out by the work counter or by the static grid-stride split; and equal the oracle."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    sc, sh, ocam = load_oracle_scene(loader, oracle, terrain, "compound-cam")
    N, S = 2500, 64
    omm = ocam.ommatidia[:: len(ocam.ommatidia) // N][:N]
    er.setOmmatidiaFromArray(lib, omm)
    er.setRenderSize(lib, N, 1)
    eye = oracle.CompoundEyeOracle(sh, omm, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    eye.render_frame(method="bvh")
    want_seq, want_fused = eye.last["summed"].copy(), oracle.fused_sum(eye.last["compound"], N, S)
    pose0 = np.zeros(12, np.float32); lib.crDebugCopyCameraPose(pose0.ctypes.data)
    poses = np.tile(pose0, (6, 1)); poses[:, 1] += np.float32([0.0, 0.25, 0.5, 1.0, 3.0, 9.0])
    out = {}
    for fused in (0, 1):
        for key, (on, refill, frac, lanes) in {"off": (0, 24, 0.35, 1), "static": (0, 24, 0.35, 16), "off32": (0, 24, 0.35, 32),
                                               "on": (1, 24, 0.35, 16), "refill32": (1, 32, 0.35, 24), "refill1": (1, 1, 0.35, 1),
                                               "tiny": (1, 24, 0.002, 32)}.items():
            lib.crSetRenderMode(fused, 0)
            lib.crDebugSetDynamicChunks(0 if key in ("static", "refill1") else 1)     # static grid-stride split vs the work counter
            lib.crDebugSetNodeLanes(lanes)
            lib.crDebugSetWavefront(on, refill, frac)
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            lib.renderFrame()
            queued = lib.crDebugLastQueuedRays()
            rgb = er.getOmmatidialData(lib).copy()
            row = er.getFrame(lib, N, 1).copy()
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            rows, _ = er.renderPoseBatch(lib, poses)
            queued_batch = lib.crDebugLastQueuedRays()
            out[(fused, key)] = (rgb, row, rows.copy())
            if on:
                assert queued > 0 and queued % 32 == 0 and queued_batch > 0, (key, queued, queued_batch)
                if key == "tiny":
                    assert queued <= 0.002 * N * S + 32
                else:
                    assert 0.02 * N * S < queued < 0.35 * N * S, (key, queued)
            want = want_fused if fused else want_seq
            assert np.array_equal(rgb.view(np.uint32), want.view(np.uint32)), (fused, key)
        for key in ("static", "off32", "on", "refill32", "refill1", "tiny"):
            for a, b in zip(out[(fused, "off")], out[(fused, key)]):
                assert np.array_equal(a, b), (fused, key)
        assert np.array_equal(out[(fused, "off")][1][0], out[(fused, "off")][2][0])
    lib.crSetRenderMode(0, 0)


def test_sm_affine_hand_out_changes_no_bit(lib, er, loader, oracle, terrain):
    """SM-affine hand-out of the trace kernel's units (blocks of 32 units stay on one SM through per-SM ticket counters, the
    block numbers published in an epoch-tagged table): float RGB, 8-bit rows, pose-batch rows and the XORWOW states afterwards
    are bit-identical with the hand-out off, on at its default threshold and forced on for launches of any size -- frames of
    fewer blocks than SMs, a ragged last block, S not a multiple of 32, grouped and ungrouped batches, both reductions, many
    launches in a row (the counters are rearmed by the kernel itself), with the 8-bit row written straight into the mapped host
    frame or copied -- and equal the oracle."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    sc, sh, ocam = load_oracle_scene(loader, oracle, terrain, "compound-cam")
    lib.crDebugSetReadAhead(0, 0.0)
    pose0 = np.zeros(12, np.float32); lib.crDebugCopyCameraPose(pose0.ctypes.data)
    poses = np.tile(pose0, (7, 1)); poses[:, 1] += np.float32([0.0, 0.25, 0.5, 1.0, 3.0, 9.0, 20.0])
    # (N, S): 79 units = 3 blocks, the last one ragged; S = 40: rows straddle warps; 9 280 units = 290 blocks on 148 SMs;
    # 2 500 x 128 = 10 000 units (313 blocks, > 2 per SM)
    for N, S in ((63, 40), (1450, 200), (2320, 128), (2500, 128)):
        omm = ocam.ommatidia[:: len(ocam.ommatidia) // N][:N]
        er.setOmmatidiaFromArray(lib, omm)
        er.setRenderSize(lib, N, 1)
        eye = oracle.CompoundEyeOracle(sh, omm, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
        eye.set_render_size(N, 1)
        eye.render_frame(method="bvh")
        want_seq = eye.last["summed"].copy()
        want_fused = oracle.fused_sum(eye.last["compound"], N, S) if S % 32 == 0 else None
        out = {}
        for fused in (0, 1):
            for key, (on, minb, zc) in {"off": (0, 48, 1), "default": (1, 48, 1), "forced": (1, 0, 1), "two": (1, 2, 1),
                                        "copy": (1, 2, 0)}.items():
                lib.crSetRenderMode(fused, 0)
                lib.crDebugSetZeroCopy(zc)            # 8-bit row straight into the mapped host frame / device-to-host copy
                lib.crDebugSetSmAffine(on, minb)
                lib.setCurrentEyeSamplesPerOmmatidium(S)               # restarts the streams at frame 0
                frames = []
                for _ in range(3):                                      # consecutive launches: counters rearmed, epochs advance
                    lib.renderFrame()
                    frames.append((er.getOmmatidialData(lib).copy(), er.getFrame(lib, N, 1).copy()))
                rows, _ = er.renderPoseBatch(lib, poses)                # starts at frame 3 (odd): ungrouped
                rows2, _ = er.renderPoseBatch(lib, poses[:6])           # starts at frame 10 (even): grouped when the frame is small
                st = np.zeros((N * S, 8), np.uint32)
                lib.crDebugCopyRngStates(st.ctypes.data)
                out[(fused, key)] = (frames, rows.copy(), rows2.copy(), st)
                want = want_fused if (fused and want_fused is not None) else want_seq
                assert np.array_equal(frames[0][0].view(np.uint32), want.view(np.uint32)), (N, S, fused, key)
            ref = out[(fused, "off")]
            for key in ("default", "forced", "two", "copy"):
                got = out[(fused, key)]
                for (a0, a1), (b0, b1) in zip(ref[0], got[0]):
                    assert np.array_equal(a0.view(np.uint32), b0.view(np.uint32)) and np.array_equal(a1, b1), (N, S, fused, key)
                assert np.array_equal(ref[1], got[1]) and np.array_equal(ref[2], got[2]), (N, S, fused, key, "batch rows")
                assert np.array_equal(ref[3], got[3]), (N, S, fused, key, "stream states")
    lib.crSetRenderMode(0, 0)


def test_frame_groups_in_small_batches_change_no_byte(lib, er, ref_data):
    """Batched launches of small frames cut their frames into groups (a work unit = 32 rays x one group; later groups step
    the stream states over the frames before them).  Batches that start at even and at odd frames, of lengths that do and
    do not divide into the groups, in both reductions: rows, float RGB of the last frame and the stream states afterwards
    equal the ungrouped launches and the frame-by-frame loop."""
    path = os.path.join(ref_data, "data", "natural-standin-sky.gltf")
    lib.loadGlTFscene(path.encode())
    er.gotoFirstCompoundEye(lib)
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N = lib.getCurrentEyeOmmatidialCount()
    er.setRenderSize(lib, N, 1)
    lib.crDebugSetReadAhead(0, 0.0)
    pose0 = np.zeros(12, np.float32); lib.crDebugCopyCameraPose(pose0.ctypes.data)
    rng = np.random.default_rng(3)
    lengths = [9, 7, 8, 33, 4, 5]                               # starts at frames 0, 9 (odd: not grouped), 16, 24, 57 (odd), 61 (odd)
    poses = [np.tile(pose0, (n, 1)) for n in lengths]
    for p in poses:
        p[:, :3] += rng.uniform(-0.5, 0.5, size=(len(p), 3)).astype(np.float32)

    def states(S):
        st = np.zeros((N * S, 8), np.uint32)
        lib.crDebugCopyRngStates(st.ctypes.data)
        st[:, 7] = np.where(st[:, 6] == 1, st[:, 7], 0)
        return st

    for fused, S in ((0, 40), (1, 32)):
        out = {}
        for groups in (0, 1):
            lib.crDebugSetFrameGroups(groups)
            lib.crSetRenderMode(fused, 0)
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            rows = [er.renderPoseBatch(lib, p)[0].copy() for p in poses]
            out[groups] = (np.concatenate(rows), er.getOmmatidialData(lib).copy(), states(S))
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        loop = []
        for p in np.concatenate(poses):
            lib.setCameraPosition(*[float(v) for v in p[:3]])
            lib.renderFrame()
            loop.append(er.getFrame(lib, N, 1)[0].copy())
        lib.setCameraPosition(*[float(v) for v in pose0[:3]])
        assert np.array_equal(out[0][0], np.stack(loop)), (fused, "ungrouped batches vs the frame loop")
        for k, name in enumerate(("rows", "float RGB", "stream states")):
            assert np.array_equal(out[0][k].view(np.uint8), out[1][k].view(np.uint8)), (fused, name)
        assert np.array_equal(out[1][2], states(S)), (fused, "states after the frame loop")
    lib.crSetRenderMode(0, 0)


def test_read_ahead_for_a_standing_camera_changes_no_byte(lib, er, ref_data):
    """renderFrame from a standing camera renders several frames per launch and hands them out one per call.  A script of
    calls -- long standing runs, a move in the middle of a run, setOmmatidia with the same count (streams persist),
    setCurrentEyeSamplesPerOmmatidium with the same S (streams restart), a pose batch in between, a mode switch, reading
    the stream states -- gives the same rows, float RGB and XORWOW states with the read-ahead on as with it off."""
    path = os.path.join(ref_data, "data", "natural-standin-sky.gltf")
    lib.loadGlTFscene(path.encode())
    er.gotoFirstCompoundEye(lib)
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    eye = er.readEyeFile(os.path.join(ref_data, "data", "eyes", "1000-equidistant.eye"))
    er.setOmmatidiaFromOmmatidiumList(lib, eye)
    N = lib.getCurrentEyeOmmatidialCount()
    er.setRenderSize(lib, N, 1)
    pose0 = np.zeros(12, np.float32); lib.crDebugCopyCameraPose(pose0.ctypes.data)

    def script(read_ahead):
        lib.crDebugSetReadAhead(read_ahead, 1.5)
        lib.crSetRenderMode(0, 0)
        lib.setCameraPosition(*[float(v) for v in pose0[:3]])
        lib.setCurrentEyeSamplesPerOmmatidium(32)
        out, launches = [], []

        def frames(n):
            for _ in range(n):
                l0 = lib.crGetLaunchCount()
                lib.renderFrame()
                launches.append(lib.crGetLaunchCount() - l0)
                out.append(er.getFrame(lib, N, 1).copy())
                out.append(er.getOmmatidialData(lib).copy())

        def states():
            st = np.zeros((N * lib.getCurrentEyeSamplesPerOmmatidium(), 8), np.uint32)
            lib.crDebugCopyRngStates(st.ctypes.data)
            st[:, 7] = np.where(st[:, 6] == 1, st[:, 7], 0)   # the cached Box-Muller half is dead while its flag is clear
            out.append(st)                                    # (a rewound stream holds 0 there, a run-through stream the stale value)

        frames(45)                                           # long standing run: several read-ahead launches
        states()                                             # ... in the middle of one: the states the caller has reached
        frames(5)
        lib.setCameraPosition(float(pose0[0]), float(pose0[1]) + 0.5, float(pose0[2]))   # move inside a read-ahead run
        frames(9)
        lib.setCurrentEyeSamplesPerOmmatidium(32)            # same S: the streams restart at frame 0
        frames(12)
        er.setOmmatidiaFromOmmatidiumList(lib, eye[::-1])    # same count, other table: the streams persist
        frames(8)
        er.setOmmatidiaFromOmmatidiumList(lib, eye)
        rows, _ = er.renderPoseBatch(lib, np.tile(pose0, (3, 1)))      # a batch in between continues where the caller is
        out.append(rows.copy())
        frames(7)
        lib.crSetRenderMode(1, 0)                            # fused reduction from here on, streams persist
        frames(11)
        lib.setCurrentEyeSamplesPerOmmatidium(64)
        frames(6)
        states()
        lib.crSetRenderMode(0, 0)
        return out, launches

    off, l_off = script(0)
    on, l_on = script(1)
    assert len(off) == len(on)
    for k, (a, b) in enumerate(zip(off, on)):
        assert a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"output {k} differs with the read-ahead on"
    assert sum(1 for v in l_on if v == 0) > len(l_on) // 2, "most frames should have been handed out without a launch"
    assert all(v > 0 for v in l_off)


def test_fused_reduction_fixed_order_and_tolerance(lib, er, loader, oracle, terrain):
    """crSetRenderMode(1, 0): rays, hits and per-sample colours are untouched; the per-ommatidium RGB equals the
    checker's restatement of the fused addition order bit for bit (S = 64 and S = 2048: one and two partials per lane
    of k_sumPartials), the reference's sequential order to <= 2e-5 absolute (mean <= 1e-6), and the 8-bit row within one step."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    sc, sh, ocam = load_oracle_scene(loader, oracle, terrain, "compound-cam")
    for N, S in ((10000, 64), (600, 2048)):
        omm = ocam.ommatidia[:: len(ocam.ommatidia) // N][:N]
        er.setOmmatidiaFromArray(lib, omm)
        er.setRenderSize(lib, N, 1)
        eye = oracle.CompoundEyeOracle(sh, omm, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
        eye.set_render_size(N, 1)
        lib.crSetRenderMode(1, 0)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        for frame in range(2):
            lib.renderFrame()
            eye.render_frame(method="bvh")
            rgb = er.getOmmatidialData(lib)
            want = oracle.fused_sum(eye.last["compound"], N, S)
            assert np.array_equal(rgb.view(np.uint32), want.view(np.uint32)), f"fused order, N={N} S={S} frame {frame}"
            seq = eye.last["summed"]
            # (the sequential fp32 sum of S terms is itself the less accurate of the two: its rounding grows with S)
            assert np.abs(rgb - seq).max() <= 2e-5 and np.abs(rgb - seq).mean() <= 1e-6
            row = er.getFrame(lib, N, 1)
            assert np.array_equal(row, oracle.make_color(want).reshape(1, N, 4))
            diff = np.abs(row.astype(int) - eye.frame.astype(int))
            assert diff.max() <= 1 and (diff > 0).mean() < 2e-3
        # the batched kernel of the fused mode continues the same streams and gives the same rows as per-frame calls
        pose0 = np.zeros(12, np.float32); lib.crDebugCopyCameraPose(pose0.ctypes.data)
        poses = np.tile(pose0, (3, 1)); poses[:, 1] += np.float32([0.0, 0.25, 0.5])
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        rows, _ = er.renderPoseBatch(lib, poses)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        for k in range(3):
            lib.setCameraPosition(*[float(v) for v in poses[k, :3]])
            lib.renderFrame()
            assert np.array_equal(er.getFrame(lib, N, 1)[0], rows[k]), f"fused batch row {k}"
        lib.setCameraPosition(*[float(v) for v in pose0[:3]])
        # S % 32 != 0 falls back to the ordered path: bit-exact against the sequential sum
        lib.setCurrentEyeSamplesPerOmmatidium(40)
        eye.set_samples(40)
        lib.renderFrame(); eye.render_frame(method="bvh")
        assert np.array_equal(er.getOmmatidialData(lib).view(np.uint32), eye.last["summed"].view(np.uint32))
        lib.crSetRenderMode(0, 0)


def test_fused_mode_keeps_raw_sample_projection(lib, er, ref_data, loader, oracle):
    """raw_ommatidial_samples reads the per-sample buffer: with the fused mode on that projection still gets it."""
    path = os.path.join(ref_data, "data", "test-scene", "test-scene.gltf")
    lib.loadGlTFscene(path.encode())
    assert lib.gotoCameraByName(b"insect-cam-1")
    sc, sh, ocam = load_oracle_scene(loader, oracle, path, "insect-cam-1")
    N, S = len(ocam.ommatidia), 32
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), "raw_ommatidial_samples", samples=S)
    eye.set_render_size(N, S)
    lib.crSetRenderMode(1, 0)
    lib.setCurrentEyeShaderName(b"raw_ommatidial_samples")
    er.setRenderSize(lib, N, S)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    lib.renderFrame(); eye.render_frame(method="brute")
    assert np.array_equal(er.getFrame(lib, N, S), eye.frame)
    lib.setCurrentEyeShaderName(ocam.projection.encode())


@pytest.mark.parametrize("scene,cam,S", [("test-scene/test-scene.gltf", "insect-cam-1", 256),
                                         ("natural-standin-sky.gltf", "insect-eye-spherical-projector", 256)])
def test_fast_math_within_tolerance(lib, er, ref_data, loader, oracle, scene, cam, S):
    """crSetRenderMode(0/1, 1): hardware sin/cos/log/pow (what the reference's --use_fast_math build runs).  Against the
    IEEE checker on the reference's own scenes: per-ommatidium RGB max <= 1/255, mean <= 1e-4; primary-hit ids differ
    only where a ray grazes an edge (< 1e-4 of the rays); the RNG integer state is untouched."""
    path = os.path.join(ref_data, "data", scene)
    lib.loadGlTFscene(path.encode())
    assert lib.gotoCameraByName(cam.encode())
    sc, sh, ocam = load_oracle_scene(loader, oracle, path, cam)
    N = len(ocam.ommatidia)
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, N, 1)
    for fused in (0, 1):
        lib.crSetRenderMode(fused, 1)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        eye.set_samples(S)
        for frame in range(2):
            if fused == 0 and frame == 1:
                lib.crDebugSetRayDump(True)
            lib.renderFrame()
            eye.render_frame(method="bvh")
            rgb = er.getOmmatidialData(lib)
            err = np.abs(rgb - eye.last["summed"])
            assert err.max() <= 1.0 / 255.0 and err.mean() <= 1e-4, (fused, frame, float(err.max()), float(err.mean()))
            row = er.getFrame(lib, N, 1).astype(int)
            assert np.abs(row - eye.frame.astype(int)).max() <= 1
        if fused == 0:
            _, d, h, _ = _dump(lib, N * S)
            lib.crDebugSetRayDump(False)
            dd = np.abs(d - eye.last["dirs"]).max()
            assert dd < 2e-5, dd
            assert (h["prim"] != eye.last["hits"]["prim"]).mean() < 1e-4
            st = np.zeros((N * S, 8), np.uint32)
            lib.crDebugCopyRngStates(st.ctypes.data)
            assert np.array_equal(st[:, 0], eye.states["d"]) and np.array_equal(st[:, 1:6], eye.states["v"])
    lib.setCurrentEyeShaderName(ocam.projection.encode())
