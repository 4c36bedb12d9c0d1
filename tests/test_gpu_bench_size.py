"""GPU parity AT the sizes BASELINE.json names (round-2 additions; VERDICT r1 "parity holes at bench size"):

  cfg4  the benched frame itself -- 1M triangles, 10 000 ommatidia, S = 1024 (10.24M rays) -- against the oracle on
        ommatidium ranges (oracle.set_shard reproduces the streams of rows [a, b) of the whole eye), for the per-frame
        kernel, the batched kernel and the dump kernel, in the ordered and the fused mode;
  cfg4  the 10M-triangle point of the speed-test sweep at small S;
  cfg2  natural-standin-sky.gltf seen by a real-insect-scale eye (AM_60185 geometry, 6 374 ommatidia) at S = 64;
  cfg3  the synthetic ofstad arena (JPEG-textured cylinder) with the icosahedral 12-ommatidia eye and the
        minimumSampleRateFinder statistic.
"""
import os

import numpy as np
import pytest

from conftest import load_oracle_scene

pytestmark = pytest.mark.gpu

HIT4 = np.dtype([("prim", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])


@pytest.fixture(autouse=True)
def _default_modes(lib):
    yield
    lib.crSetRenderMode(0, 0)
    lib.crDebugSetRayDump(False)
    lib.crSetFirstFrame(0)


@pytest.fixture(scope="module")
def terrain(synth_dir):
    from tools import synth
    gltf = os.path.join(synth_dir, "terrain1m_bs.gltf")
    synth.write_eye(os.path.join(synth_dir, "eye10k_bs.eye"), synth.fibonacci_eye(10000))
    synth.write_terrain_gltf(gltf, triangles=1_000_000, eye_file="eye10k_bs.eye")
    return gltf


def test_cfg4_benched_frame_against_the_oracle_on_ommatidium_ranges(lib, er, loader, oracle, terrain):
    """The headline frame at full size.  Three ranges of 40 ommatidia (top of the eye = sky, the horizon band, straight
    down): float RGB bit-exact against the oracle in both reduction orders, for renderFrame (k_traceCompound<.,false,..>)
    and for crRenderPoseBatch (k_traceCompound<.,true,..>); hit ids and (t,u,v) of every sample ray of those ranges
    bit-exact through the dump kernel."""
    lib.loadGlTFscene(terrain.encode())
    assert lib.gotoCameraByName(b"compound-cam")
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N, S = 10000, 1024
    er.setRenderSize(lib, N, 1)
    sc, sh, ocam = load_oracle_scene(loader, oracle, terrain, "compound-cam")
    pose = oracle.pose_from_camera(ocam)
    pose12 = np.zeros(12, np.float32)
    lib.crDebugCopyCameraPose(pose12.ctypes.data)
    ranges = [(0, 40), (4980, 5020), (9960, 10000)]
    # oracle: frames 0 and 1 of each range
    want = {}
    for a, b in ranges:
        eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia[a:b], pose, "single_dimension_fast", samples=S)
        eye.set_render_size(b - a, 1)
        eye.set_shard(N, a)
        eye.render_frame(method="bvh")
        eye.render_frame(method="bvh")
        want[(a, b)] = dict(summed=eye.last["summed"].copy(), fused=oracle.fused_sum(eye.last["compound"], b - a, S),
                            hits=eye.last["hits"].copy(), dirs=eye.last["dirs"].copy())
    hit_fracs = [float((want[r]["hits"]["prim"] >= 0).mean()) for r in ranges]
    assert hit_fracs[0] == 0.0 and hit_fracs[2] == 1.0 and 0.05 < hit_fracs[1] < 0.95, hit_fracs
    for fused in (0, 1):
        key = "fused" if fused else "summed"
        lib.crSetRenderMode(fused, 0)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        lib.renderFrame(); lib.renderFrame()
        rgb = er.getOmmatidialData(lib)
        row = er.getFrame(lib, N, 1)[0].copy()
        for a, b in ranges:
            assert np.array_equal(rgb[a:b].view(np.uint32), want[(a, b)][key].view(np.uint32)), (fused, a, "per-frame kernel")
            assert np.array_equal(row[a:b], oracle.make_color(want[(a, b)][key])), (fused, a, "8-bit row")
        lib.setCurrentEyeSamplesPerOmmatidium(S)                     # restart; the same two frames as one batch
        rows, _ = er.renderPoseBatch(lib, np.stack([pose12, pose12]))
        assert np.array_equal(rows[1], row), (fused, "batched kernel row != per-frame kernel row")
    # dump kernel at full size: every sample ray of the three ranges
    lib.crSetRenderMode(0, 0)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    lib.renderFrame()
    lib.crDebugSetRayDump(True)
    lib.renderFrame()
    n = N * S
    o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); h = np.zeros(n, HIT4)
    assert lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data) == n
    lib.crDebugSetRayDump(False)
    d = d.reshape(S, N, 3); h = h.reshape(S, N)                       # stream-id order N*s + o
    for a, b in ranges:
        w = want[(a, b)]
        assert np.array_equal(d[:, a:b].reshape(-1, 3).view(np.uint32), w["dirs"].view(np.uint32)), (a, "directions")
        hh = h[:, a:b].reshape(-1)
        assert np.array_equal(hh["prim"], w["hits"]["prim"]), (a, "hit ids")
        hit = hh["prim"] >= 0
        for k in ("t", "u", "v"):
            assert np.array_equal(hh[k][hit].view(np.uint32), w["hits"][k][hit].view(np.uint32)), (a, k)


def test_cfg4_ten_million_triangles(lib, er, loader, oracle, synth_dir):
    """The 10M-triangle point of BASELINE configs[3]: S = 8 on the 10 000-ommatidia eye (80 000 rays).  Hit ids of every
    ray equal the oracle's instrumented walk of the product's own BVH; 192 of the rays are also settled by brute force
    over all ten million triangles; RGB and the frame equal the oracle's shading of those hits bit for bit."""
    from tools import synth
    gltf = os.path.join(synth_dir, "terrain10m.gltf")
    synth.write_eye(os.path.join(synth_dir, "eye10k_10m.eye"), synth.fibonacci_eye(10000))
    info = synth.write_terrain_gltf(gltf, triangles=10_000_000, eye_file="eye10k_10m.eye")
    assert info["triangles"] >= 10_000_000
    lib.loadGlTFscene(gltf.encode())
    T = lib.crDebugGetTriangleCount()
    assert T == info["triangles"]
    assert lib.gotoCameraByName(b"compound-cam")
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N, S = 10000, 8
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    lib.crDebugSetRayDump(True)
    lib.renderFrame()
    n = N * S
    o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); h = np.zeros(n, HIT4)
    assert lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data) == n
    lib.crDebugSetRayDump(False)
    rgb = er.getOmmatidialData(lib)
    row = er.getFrame(lib, N, 1)
    sc, sh, ocam = load_oracle_scene(loader, oracle, gltf, "compound-cam")
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
    st = np.zeros(n, oracle.STATE_DTYPE)
    orays = oracle.generate_rays(eye.omm, S, eye.pose, st, False)
    assert np.array_equal(d.view(np.uint32), orays[1].view(np.uint32)) and np.array_equal(o.view(np.uint32), orays[0].view(np.uint32))
    nn = lib.crDebugGetBvhNodeCount()
    nodes = np.zeros((nn, 16), np.float32); dtris = np.zeros((T, 12), np.float32)
    lib.crDebugCopyBvh(nodes.ctypes.data, dtris.ctypes.data)
    tm = np.zeros(n, np.float32)
    hits, cnt = oracle.trace_device_bvh(nodes, dtris, o, d, tm)
    assert np.array_equal(hits["prim"], h["prim"])
    hit = h["prim"] >= 0
    assert 0.3 < hit.mean() < 0.7
    for k in ("t", "u", "v"):
        assert np.array_equal(hits[k][hit].view(np.uint32), h[k][hit].view(np.uint32)), k
    pick = np.concatenate([np.flatnonzero(hit)[:: max(1, hit.sum() // 128)][:128], np.flatnonzero(~hit)[:: max(1, (~hit).sum() // 64)][:64]])
    brute = oracle.trace(sh, o[pick], d[pick], tm[pick], method="brute")
    assert np.array_equal(brute["prim"], h["prim"][pick])
    col = oracle.shade(sh, hits, d)
    compound = np.empty((n, 3), np.float32); summed = np.empty((N, 3), np.float32)
    oracle.lib().cro_accumulate(col.ctypes.data, N, S, compound.ctypes.data, summed.ctypes.data)
    assert np.array_equal(rgb.view(np.uint32), summed.view(np.uint32))
    assert np.array_equal(row[0], oracle.make_color(summed))


def test_cfg2_natural_standin_with_a_real_insect_scale_eye(lib, er, loader, oracle, ref_data):
    """BASELINE configs[1]: data/natural-standin-sky.gltf (24 200 triangles, 1024^2 ground texture, simple_sky) through
    its compound camera, with the eye replaced by the AM_60185 geometry (6 374 ommatidia, acceptance 0.045 rad, focal
    offsets 0.2-0.3) at S = 64.  Hit ids exact; per-ommatidium RGB within the textured-scene tolerance (the product samples
    through the hardware texture unit): max <= 1/255, mean <= 1e-4; single_dimension_fast row and the 550x400 spherical
    frame within one 8-bit step."""
    path = os.path.join(ref_data, "data", "natural-standin-sky.gltf")
    lib.loadGlTFscene(path.encode())
    assert lib.gotoCameraByName(b"insect-eye-spherical-projector")
    sc, sh, ocam = load_oracle_scene(loader, oracle, path, "insect-eye-spherical-projector")
    omm = np.asarray([[*o.position, *o.direction, o.acceptanceAngle, o.focalpointOffset]
                      for o in er.readEyeFile(os.path.join(ref_data, "sim-environment", "eyes", "AM_60185-real.eye"))], np.float32)
    N, S = len(omm), 64
    assert N == 6374
    # the toy-experiment eye is in millimetres: bring it to the stand-in scene's scale (eye radius ~0.2, as its own eye)
    omm[:, 0:3] *= np.float32(0.1); omm[:, 7] *= np.float32(0.1)
    er.setOmmatidiaFromArray(lib, omm)
    eye = oracle.CompoundEyeOracle(sh, omm, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
    for fused in (0, 1):
        lib.crSetRenderMode(fused, 0)
        lib.setCurrentEyeShaderName(b"single_dimension_fast")
        er.setRenderSize(lib, N, 1)
        eye.projection = "single_dimension_fast"; eye.set_render_size(N, 1)
        lib.setCurrentEyeSamplesPerOmmatidium(S); eye.set_samples(S)
        lib.crDebugSetRayDump(fused == 0)
        lib.renderFrame(); eye.render_frame(method="bvh")
        lib.renderFrame(); eye.render_frame(method="bvh")
        if fused == 0:
            n = N * S
            o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); h = np.zeros(n, HIT4)
            assert lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data) == n
            lib.crDebugSetRayDump(False)
            assert np.array_equal(d.view(np.uint32), eye.last["dirs"].view(np.uint32))
            assert np.array_equal(h["prim"], eye.last["hits"]["prim"])
            assert 0.2 < (h["prim"] >= 0).mean() < 0.8
        rgb = er.getOmmatidialData(lib)
        ref = oracle.fused_sum(eye.last["compound"], N, S) if fused else eye.last["summed"]
        err = np.abs(rgb - ref)
        assert err.max() <= 1.0 / 255.0 and err.mean() <= 1e-4, (fused, float(err.max()), float(err.mean()))
        assert np.abs(er.getFrame(lib, N, 1).astype(int) - eye.frame.astype(int)).max() <= 1
        lib.setCurrentEyeShaderName(b"spherical_orientationwise")
        er.setRenderSize(lib, 550, 400)
        eye.projection = "spherical_orientationwise"; eye.set_render_size(550, 400)
        lib.renderFrame(); eye.render_frame(method="bvh")
        diff = np.abs(er.getFrame(lib, 550, 400).astype(int) - eye.frame.astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.02


def _variance_image(frames):
    """data/tools/minimumSampleRateFinder.py:36-47: per-ommatidium variance of the RGB vector over consecutive frames."""
    allImages = np.vstack(frames).astype(np.float64)
    diff = allImages - allImages.mean(axis=0)
    mag = np.linalg.norm(diff, axis=2)
    return (mag * mag).sum(axis=0) / (len(frames) - 1)


def test_cfg3_ofstad_arena_icosahedral_eye(lib, er, loader, oracle, synth_dir):
    """BASELINE configs[2]: the synthetic ofstad arena (open cylinder r 12.5 x h 9.19, 32 768 wall triangles, 1024^2
    JPEG pattern, floor disc, default background) with the 12-ommatidia 1-steradian eye of getIcoOmmatidia, as
    minimumSampleRateFinder.py sets it up (:126-137, :267-282).  (i) frames at S = 16 and 256 against the oracle -- hit ids
    exact, RGB within the textured tolerance; (ii) the script's statistic -- max over ommatidia of the frame-to-frame SD of
    the 8-bit single_dimension_fast vector -- computed by the product and by the oracle over the same 12 consecutive
    frames agrees, and falls with S."""
    from tools import synth
    gltf = os.path.join(synth_dir, "arena.gltf")
    synth.write_eye(os.path.join(synth_dir, "ico.eye"), synth.ico_eye())
    info = synth.write_arena_gltf(gltf, eye_file="ico.eye")
    assert info["wall_triangles"] == 32768
    lib.loadGlTFscene(gltf.encode())
    assert lib.crDebugGetTextureCount() == 1 and lib.crDebugGetMissShader() == 0
    assert lib.gotoCameraByName(b"compound-cam")
    sc, sh0, ocam = load_oracle_scene(loader, oracle, gltf, "compound-cam")
    tex = np.zeros((1024, 1024, 4), np.uint8)
    lib.crDebugCopyTexture(0, tex.ctypes.data)                          # the product's decoder is stb_image-exact (CPU KATs);
    d = np.abs(tex.astype(int) - sc.textures[0].astype(int))           # the checker's loader decodes with PIL's libjpeg
    assert d.max() <= 16 and d.mean() < 1.0
    sc.textures[0] = tex
    sh = oracle.SceneHandle(sc)
    er.setOmmatidiaFromOmmatidiumList(lib, er.getIcoOmmatidia())
    omm = np.asarray([[*o.position, *o.direction, o.acceptanceAngle, o.focalpointOffset] for o in er.getIcoOmmatidia()], np.float32)
    assert np.array_equal(omm, synth.ico_eye())
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    er.setRenderSize(lib, 12, 1)
    lib.setCameraPose(3.0, 2.5, -4.0, 0.3, -0.8, 0.1)
    pose12 = np.zeros(12, np.float32)
    lib.crDebugCopyCameraPose(pose12.ctypes.data)
    pose = oracle.make_pose(pose12[0:3], pose12[3:6], pose12[6:9], pose12[9:12])
    sds = {}
    for S in (16, 256):
        eye = oracle.CompoundEyeOracle(sh, omm, pose, "single_dimension_fast", samples=S)
        eye.set_render_size(12, 1)
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        lib.crDebugSetRayDump(True)
        mine, theirs = [], []
        for k in range(12):
            lib.renderFrame(); eye.render_frame(method="bvh")
            if k == 0:
                n = 12 * S
                o = np.zeros((n, 3), np.float32); dd = np.zeros((n, 3), np.float32); h = np.zeros(n, HIT4)
                assert lib.crDebugCopyLastRays(o.ctypes.data, dd.ctypes.data, h.ctypes.data) == n
                assert np.array_equal(dd.view(np.uint32), eye.last["dirs"].view(np.uint32))
                assert np.array_equal(h["prim"], eye.last["hits"]["prim"])
                assert (h["prim"] >= 0).mean() > 0.5
            err = np.abs(er.getOmmatidialData(lib) - eye.last["summed"])
            assert err.max() <= 1.0 / 255.0 and err.mean() <= 1e-4, (S, k, float(err.max()))
            mine.append(er.getFrame(lib, 12, 1)[:, :, :3].copy()); theirs.append(eye.frame[:, :, :3].copy())
        lib.crDebugSetRayDump(False)
        sd_mine, sd_theirs = np.sqrt(_variance_image(mine).max()), np.sqrt(_variance_image(theirs).max())
        assert abs(sd_mine - sd_theirs) <= 0.5, (S, sd_mine, sd_theirs)
        sds[S] = sd_mine
    assert sds[16] > 2.0 * sds[256] > 0.0, sds
