"""GPU parity tests: the CUDA product (through its C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star / DESIGN.md):
  * integer work -- XORWOW states, primary-hit primitive ids, pixel->ommatidium maps, id frames:
    BIT-EXACT;
  * rays (origin, direction), hit parameters (t,u,v), per-ommatidium float RGB on scenes without
    textures, and the uchar4 frames derived from them: BIT/BYTE-EXACT as well, because both sides
    evaluate the same specified binary32 arithmetic (cr_math);
  * textured scenes: the product samples through the hardware texture unit; tolerance
    max |dRGB| <= 1/255 and mean <= 1e-4 on per-ommatidium float RGB (stated in the test).
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_oracle_scene

pytestmark = pytest.mark.gpu

HIT4 = np.dtype([("prim", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])


def _frame(er, L, w, h):
    return er.getFrame(L, w, h)


def _product_rays(L, n):
    o = np.zeros((n, 3), np.float32)
    d = np.zeros((n, 3), np.float32)
    h = np.zeros(n, HIT4)
    got = L.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data)
    assert got == n
    return o, d, h


def _product_states(L, n):
    st = np.zeros((n, 8), np.uint32)
    L.crDebugCopyRngStates(st.ctypes.data)
    return st


def _oracle_states(eye):
    s = eye.states
    out = np.zeros((len(s), 8), np.uint32)
    out[:, 0] = s["d"]
    out[:, 1:6] = s["v"]
    out[:, 6] = s["flag"].astype(np.uint32)
    out[:, 7] = s["extra"].view(np.uint32)
    return out


def test_math_bit_exact(lib, oracle):
    """cr_math on the device == the oracle's cr_math.h, bit for bit, on dense samples."""
    rng = np.random.default_rng(123)
    OL = oracle.lib()
    cases = [
        (0, OL.cro_sinf, rng.uniform(-20, 20, 20000)), (1, OL.cro_cosf, rng.uniform(-20, 20, 20000)),
        (2, OL.cro_logf, rng.uniform(1e-12, 1.0, 20000)), (3, OL.cro_expf, rng.uniform(-30, 10, 20000)),
        (5, OL.cro_asinf, rng.uniform(-1.01, 1.01, 20000)), (6, OL.cro_acosf, rng.uniform(-1.01, 1.01, 20000)),
    ]
    for fn, ofn, xs in cases:
        xs = xs.astype(np.float32)
        out = np.zeros_like(xs)
        lib.crDebugEvalMath(fn, xs.ctypes.data, None, out.ctypes.data, len(xs))
        ref = np.array([ofn(float(x)) for x in xs], dtype=np.float32)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), f"math fn {fn} differs"
    xs = rng.uniform(0, 1, 20000).astype(np.float32)
    for ex in (np.float32(2.2), np.float32(1.0 / np.float64(np.float32(2.2)))):
        ys = np.full_like(xs, ex)
        out = np.zeros_like(xs)
        lib.crDebugEvalMath(4, xs.ctypes.data, ys.ctypes.data, out.ctypes.data, len(xs))
        ref = np.array([OL.cro_powf(float(x), float(ex)) for x in xs], dtype=np.float32)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    a = rng.uniform(-1, 1, 20000).astype(np.float32)
    b = rng.uniform(-1, 1, 20000).astype(np.float32)
    out = np.zeros_like(a)
    lib.crDebugEvalMath(7, a.ctypes.data, b.ctypes.data, out.ctypes.data, len(a))
    ref = np.array([OL.cro_atan2f(float(y), float(x)) for y, x in zip(a, b)], dtype=np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def _setup_test_scene(lib, er, ref_data, loader, oracle, scene="test-scene.gltf", cam="insect-cam-1", S=32, size=(400, 400)):
    path = os.path.join(ref_data, "data", "test-scene", scene)
    lib.loadGlTFscene(path.encode())
    assert lib.gotoCameraByName(cam.encode())
    er.setRenderSize(lib, *size)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    sc, sh, ocam = load_oracle_scene(loader, oracle, path, cam)
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), ocam.projection, samples=S)
    eye.set_render_size(*size)
    return sc, sh, ocam, eye


def test_cfg1_test_scene_frames_0_and_1(lib, er, ref_data, loader, oracle):
    """BASELINE config 1: test-scene.gltf + test.eye, spherical_orientationwise, S=32, frames 0 and 1."""
    sc, sh, ocam, eye = _setup_test_scene(lib, er, ref_data, loader, oracle)
    N, S = len(ocam.ommatidia), 32
    lib.crDebugSetRayDump(True)
    for frame in range(2):
        ms = lib.renderFrame()
        assert ms > 0
        eye.render_frame(method="brute")
        # RNG: integer-exact state after the frame (d, v[5], flag) and the cached normal's bits
        assert np.array_equal(_product_states(lib, N * S), _oracle_states(eye)), f"RNG state, frame {frame}"
        o, d, h = _product_rays(lib, N * S)
        assert np.array_equal(o.view(np.uint32), eye.last["origins"].view(np.uint32)), "ray origins"
        assert np.array_equal(d.view(np.uint32), eye.last["dirs"].view(np.uint32)), "ray directions"
        oh = eye.last["hits"]
        assert np.array_equal(h["prim"], oh["prim"]), "primary-hit primitive ids"
        hit = h["prim"] >= 0
        for k in ("t", "u", "v"):
            assert np.array_equal(h[k][hit].view(np.uint32), oh[k][hit].view(np.uint32)), k
        rgb = er.getOmmatidialData(lib)
        assert np.array_equal(rgb.view(np.uint32), eye.last["summed"].view(np.uint32)), "per-ommatidium RGB"
        fr = _frame(er, lib, 400, 400)
        assert np.array_equal(fr, eye.frame), "uchar4 frame"
    lib.crDebugSetRayDump(False)
    assert (h["prim"] >= 0).sum() > 100


def test_bvh_equals_bruteforce_on_random_rays(lib, er, ref_data, loader, oracle):
    """Traversal exactness: device BVH result == oracle brute force on identical arbitrary rays,
    including axis-parallel and zero-component directions and rays starting inside geometry."""
    path = os.path.join(ref_data, "data", "test-scene", "test-scene.gltf")
    lib.loadGlTFscene(path.encode())
    sc, sh, _ = load_oracle_scene(loader, oracle, path)
    rng = np.random.default_rng(7)
    n = 60000
    o = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:2000, 0] = 0.0
    d[2000:4000, 1] = 0.0
    d[4000:6000, 2] = 0.0
    d[6000:6500] = np.array([1, 0, 0], np.float32)
    d[6500:7000] = np.array([0, -1, 0], np.float32)
    d[7000:7500] = np.array([0, 0, 1], np.float32)
    o[7500:9000] = rng.uniform(-0.9, 0.9, (1500, 3)).astype(np.float32)     # inside the cube
    # rays aimed exactly at vertices / edge midpoints of the cube (ties and edge cases)
    verts = sc.verts.reshape(-1, 3)[:36]
    tgt = np.concatenate([verts, 0.5 * (verts + np.roll(verts, 1, axis=0))]).astype(np.float32)
    k = len(tgt)
    o[9000:9000 + k] = np.array([3.0, 2.5, 4.0], np.float32)
    d[9000:9000 + k] = tgt - o[9000:9000 + k]
    tmin = np.zeros(n, np.float32)
    tmin[::3] = 0.01
    hits8 = np.zeros((n, 8), np.int32)
    lib.crDebugTraceRays(o.ctypes.data, d.ctypes.data, tmin.ctypes.data, n, hits8.ctypes.data)
    oh = oracle.trace(sh, o, d, tmin, method="brute")
    assert np.array_equal(hits8[:, 0], oh["prim"])
    hit = oh["prim"] >= 0
    assert hit.sum() > 2000
    assert np.array_equal(hits8[hit, 1].view(np.float32).view(np.uint32), oh["t"][hit].view(np.uint32))
    assert np.array_equal(hits8[hit, 2].view(np.float32).view(np.uint32), oh["u"][hit].view(np.uint32))


def test_bvh_structure(lib, ref_data):
    """Every triangle sits in exactly one reachable leaf; child boxes contain their triangles."""
    path = os.path.join(ref_data, "data", "natural-standin-sky.gltf")
    lib.loadGlTFscene(path.encode())
    T = lib.crDebugGetTriangleCount()
    nn = lib.crDebugGetBvhNodeCount()
    nodes = np.zeros((nn, 16), np.float32)
    tris = np.zeros((T, 12), np.float32)
    lib.crDebugCopyBvh(nodes.ctypes.data, tris.ctypes.data)
    prim = tris[:, 3].view(np.int32)
    assert np.array_equal(np.sort(prim), np.arange(T)), "sorted triangle array is a permutation"
    refs = nodes[:, 12:14].view(np.int32)
    seen = np.zeros(T, np.int32)
    visited = np.zeros(nn, bool)
    stack = [0]
    while stack:
        i = stack.pop()
        assert not visited[i]
        visited[i] = True
        n = nodes[i]
        boxes = [((n[0], n[2], n[8]), (n[1], n[3], n[9])), ((n[4], n[6], n[10]), (n[5], n[7], n[11]))]
        for c in range(2):
            r = int(refs[i, c])
            if r >= 0:
                stack.append(r)
                continue
            x = ~r
            first, cnt = x >> 3, (x & 7) + 1
            seen[first:first + cnt] += 1
            v0 = tris[first:first + cnt, 0:3]
            pts = np.concatenate([v0, v0 + tris[first:first + cnt, 4:7], v0 + tris[first:first + cnt, 8:11]])
            assert (pts >= np.array(boxes[c][0]) - 0).all() and (pts <= np.array(boxes[c][1]) + 0).all()
    assert (seen == 1).all()


def test_state_reset_rules_and_shader_switch(lib, er, ref_data, loader, oracle):
    """RNG streams persist across frames / pose / ommatidia VALUE changes and reset on S or COUNT
    changes (cameras/CompoundEye.cpp:30-62,98-183); vector and raw projections byte-exact."""
    sc, sh, ocam, eye = _setup_test_scene(lib, er, ref_data, loader, oracle, cam="insect-cam-2", S=7, size=(100, 7))
    N = len(ocam.ommatidia)
    lib.setCurrentEyeShaderName(b"raw_ommatidial_samples")
    eye.projection = "raw_ommatidial_samples"
    lib.renderFrame(); eye.render_frame(method="brute")
    assert np.array_equal(_frame(er, lib, 100, 7), eye.frame)
    # pose change keeps the streams
    lib.setCameraPose(-3.5, 0.4, 4.0, 0.2, -0.7, 0.1)
    eye.pose = oracle.set_camera_pose(-3.5, 0.4, 4.0, 0.2, -0.7, 0.1)
    lib.setCurrentEyeShaderName(b"single_dimension")
    eye.projection = "single_dimension"
    er.setRenderSize(lib, 37, 3); eye.set_render_size(37, 3)
    lib.renderFrame(); eye.render_frame(method="brute")
    assert np.array_equal(_product_states(lib, N * 7), _oracle_states(eye))
    assert np.array_equal(_frame(er, lib, 37, 3), eye.frame)
    # same-count ommatidia change keeps the streams; different count resets them
    omm2 = ocam.ommatidia.copy(); omm2[:, 6] = 0.3
    er.setOmmatidiaFromArray(lib, omm2); eye.set_ommatidia(omm2)
    lib.setCurrentEyeShaderName(b"single_dimension_fast"); eye.projection = "single_dimension_fast"
    er.setRenderSize(lib, N, 1); eye.set_render_size(N, 1)
    lib.renderFrame(); eye.render_frame(method="brute")
    assert np.array_equal(_frame(er, lib, N, 1), eye.frame)
    omm3 = omm2[:60]
    er.setOmmatidiaFromArray(lib, omm3); eye.set_ommatidia(omm3)
    assert lib.getCurrentEyeOmmatidialCount() == 60
    lib.renderFrame(); eye.render_frame(method="brute")
    assert np.array_equal(_product_states(lib, 60 * 7), _oracle_states(eye))
    # S change resets; S <= 0 clamps to 1
    lib.setCurrentEyeSamplesPerOmmatidium(0); eye.set_samples(0)
    assert lib.getCurrentEyeSamplesPerOmmatidium() == 1
    lib.renderFrame(); eye.render_frame(method="brute")
    assert np.array_equal(_product_states(lib, 60), _oracle_states(eye))
    assert np.array_equal(er.getOmmatidialData(lib).view(np.uint32), eye.last["summed"].view(np.uint32))


@pytest.mark.parametrize("mode", ["spherical_positionwise", "spherical_orientationwise", "spherical_split_orientationwise",
                                  "spherical_orientationwise_ids", "spherical_positionwise_ids"])
def test_projection_maps_bit_exact(lib, er, ref_data, loader, oracle, mode):
    sc, sh, ocam, eye = _setup_test_scene(lib, er, ref_data, loader, oracle, scene="test-scene-sky.gltf", S=4, size=(160, 90))
    lib.setCurrentEyeShaderName(mode.encode()); eye.projection = mode
    lib.renderFrame(); eye.render_frame(method="brute")
    m = np.zeros((90, 160), np.uint32)
    lib.crDebugCopyProjectionMap(m.ctypes.data)
    assert np.array_equal(m, oracle.projection_map(ocam.ommatidia, mode, 160, 90))
    assert np.array_equal(_frame(er, lib, 160, 90), eye.frame)


def test_large_sample_counts_chunked_sum(lib, er, ref_data, loader, oracle):
    """S above the per-CTA chunk (512) exercises the carried sequential sum; still bit-exact."""
    sc, sh, ocam, eye = _setup_test_scene(lib, er, ref_data, loader, oracle, cam="insect-cam-2", S=1300, size=(100, 1))
    lib.setCurrentEyeShaderName(b"single_dimension_fast"); eye.projection = "single_dimension_fast"
    lib.renderFrame(); eye.render_frame(method="bvh")
    assert np.array_equal(er.getOmmatidialData(lib).view(np.uint32), eye.last["summed"].view(np.uint32))
    assert np.array_equal(_frame(er, lib, 100, 1), eye.frame)


def test_pose_batch_equals_sequential_frames(lib, er, ref_data, loader, oracle):
    sc, sh, ocam, eye = _setup_test_scene(lib, er, ref_data, loader, oracle, cam="insect-cam-2", S=16, size=(100, 1))
    lib.setCurrentEyeShaderName(b"single_dimension_fast"); eye.projection = "single_dimension_fast"
    rng = np.random.default_rng(0)
    pos = rng.uniform(-2, 2, (5, 3)).astype(np.float32) + np.array([-4, 0.2, 4.5], np.float32)
    poses = er.make_poses(pos, x=ocam.x_axis, y=ocam.y_axis, z=ocam.z_axis)
    out, ms = er.renderPoseBatch(lib, poses)
    assert ms > 0
    for p in range(5):
        eye.pose = oracle.make_pose(pos[p], ocam.x_axis, ocam.y_axis, ocam.z_axis)
        eye.render_frame(method="brute")
        assert np.array_equal(out[p], eye.frame[0]), f"pose {p}"
    # restart in the middle: a fresh stream set positioned at frame 3 reproduces poses 3 and 4
    lib.crSetFirstFrame(3)
    out2, _ = er.renderPoseBatch(lib, poses[3:])
    assert np.array_equal(out2, out[3:])
    lib.crSetFirstFrame(0)


def test_textured_scenes_within_tolerance(lib, er, ref_data, loader, oracle):
    """cfg2-style: natural-standin-sky (1024^2 ground texture, hardware bilinear) and env_2."""
    for rel, cam in (("data/natural-standin-sky.gltf", "insect-eye-spherical-projector"), ("sim-environment/env_2.gltf", "compound-cam")):
        path = os.path.join(ref_data, rel)
        lib.loadGlTFscene(path.encode())
        assert lib.gotoCameraByName(cam.encode())
        S = 64
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        lib.setCurrentEyeShaderName(b"single_dimension_fast")
        sc, sh, ocam = load_oracle_scene(loader, oracle, path, cam)
        N = len(ocam.ommatidia)
        er.setRenderSize(lib, N, 1)
        eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
        eye.set_render_size(N, 1)
        lib.crDebugSetRayDump(True)
        lib.renderFrame(); eye.render_frame(method="bvh")
        o, d, h = _product_rays(lib, N * S)
        lib.crDebugSetRayDump(False)
        assert np.array_equal(h["prim"], eye.last["hits"]["prim"]), "hit ids stay bit-exact on textured scenes"
        rgb = er.getOmmatidialData(lib)
        diff = np.abs(rgb - eye.last["summed"])
        assert diff.max() <= 1.0 / 255.0 and diff.mean() <= 1e-4, (rel, diff.max(), diff.mean())
        fr = _frame(er, lib, N, 1).astype(np.int32)
        assert np.abs(fr - eye.frame.astype(np.int32)).max() <= 1


def test_ordinary_cameras(lib, er, ref_data, loader, oracle):
    """pinhole / orthographic / panoramic share the traversal: frames byte-exact."""
    path = os.path.join(ref_data, "data", "test-scene", "test-scene.gltf")
    lib.loadGlTFscene(path.encode())
    sc = loader.load_scene(path)
    sh = oracle.SceneHandle(sc)
    er.setRenderSize(lib, 96, 64)
    kinds = {"perspective": 0, "panoramic": 1, "orthographic": 2}
    for i, cam in enumerate(sc.cameras):
        if cam.kind == "compound":
            continue
        lib.gotoCamera(i)
        lib.renderFrame()
        fr = _frame(er, lib, 96, 64)
        o = np.zeros((96 * 64, 3), np.float32); d = np.zeros_like(o); tm = np.zeros(96 * 64, np.float32)
        pose = oracle.pose_from_camera(cam)
        scale = np.ascontiguousarray(cam.scale, np.float32)
        oracle.lib().cro_camera_rays(kinds[cam.kind], C.byref(pose), scale.ctypes.data, 96, 64, o.ctypes.data, d.ctypes.data, tm.ctypes.data)
        hits = oracle.trace(sh, o, d, tm, method="brute")
        rgb = oracle.shade(sh, hits, d)
        ref = oracle.make_color(rgb).reshape(64, 96, 4)
        assert np.array_equal(fr, ref), cam.name


def _write_tri_scene(path, n_tris, background="simple_sky"):
    """n_tris random triangles around the origin (0 = empty scene) + a compound camera at the origin."""
    import base64
    import json
    rng = np.random.default_rng(n_tris + 11)
    nodes = [{"camera": 0, "name": "cam"}]
    gltf = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0], "extras": {"background-shader": background}}], "nodes": nodes,
            "cameras": [{"name": "cam", "type": "perspective", "perspective": {"yfov": 0.5, "znear": 0.1},
                         "extras": {"compound-eye": True, "compound-projection": "single_dimension_fast", "compound-structure": "e.eye"}}]}
    if n_tris:
        c = rng.uniform(-3, 3, (n_tris, 1, 3))
        v = (c + rng.uniform(-1.5, 1.5, (n_tris, 3, 3))).astype(np.float32).reshape(-1, 3)
        col = rng.uniform(0, 1, (len(v), 4)).astype(np.float32)
        blob = v.tobytes() + col.tobytes()
        gltf["scenes"][0]["nodes"].append(1)
        nodes.append({"mesh": 0, "name": "soup"})
        gltf["meshes"] = [{"name": "soup", "primitives": [{"attributes": {"POSITION": 0, "COLOR_0": 1}}]}]      # non-indexed
        gltf["accessors"] = [{"bufferView": 0, "componentType": 5126, "count": len(v), "type": "VEC3",
                              "min": v.min(0).tolist(), "max": v.max(0).tolist()},
                             {"bufferView": 1, "componentType": 5126, "count": len(v), "type": "VEC4"}]
        gltf["bufferViews"] = [{"buffer": 0, "byteOffset": 0, "byteLength": v.nbytes},
                               {"buffer": 0, "byteOffset": v.nbytes, "byteLength": col.nbytes}]
        gltf["buffers"] = [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]
    with open(path, "w") as f:
        json.dump(gltf, f)


@pytest.mark.parametrize("n_tris", [0, 1, 2, 3, 4, 5, 9])
def test_tiny_and_empty_scenes(lib, er, loader, oracle, tmp_path, n_tris):
    """Builder edge cases: empty scene (every ray misses), a single triangle, scenes that fit one
    leaf (n <= 4), the first real hierarchies; non-indexed geometry with float COLOR_0."""
    from tools import synth
    gltf = str(tmp_path / f"t{n_tris}.gltf")
    synth.write_eye(str(tmp_path / "e.eye"), synth.fibonacci_eye(300, radius=0.05, acceptance=0.3))
    _write_tri_scene(gltf, n_tris, background="simple_sky" if n_tris % 2 == 0 else "default_background")
    lib.loadGlTFscene(gltf.encode())
    assert lib.crDebugGetTriangleCount() == n_tris
    assert lib.gotoCameraByName(b"cam")
    N, S = 300, 5
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    sc, sh, ocam = load_oracle_scene(loader, oracle, gltf, "cam")
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    lib.crDebugSetRayDump(True)
    lib.renderFrame(); eye.render_frame(method="brute")
    o, d, h = _product_rays(lib, N * S)
    lib.crDebugSetRayDump(False)
    assert np.array_equal(h["prim"], eye.last["hits"]["prim"])
    if n_tris == 0:
        assert (h["prim"] == -1).all()
    else:
        assert (h["prim"] >= 0).any()
    assert np.array_equal(er.getOmmatidialData(lib).view(np.uint32), eye.last["summed"].view(np.uint32))
    assert np.array_equal(_frame(er, lib, N, 1), eye.frame)


def test_texture_unit_vs_oracle_emulation(lib, oracle, loader, ref_data):
    """The product samples textures through the hardware unit (wrap + bilinear, as the reference does);
    the oracle emulates it with 8-bit weights.  Measured gap on a real texture: max 3.8e-4, mean 5e-6
    per channel (compound-ray_b200/tools/tex_probe.py) -- two orders below 1/255.  Tolerance: 1e-3."""
    import ctypes as C
    path = os.path.join(ref_data, "data", "natural-standin-sky.gltf")
    lib.loadGlTFscene(path.encode())
    sc = loader.load_scene(path)
    sh = oracle.SceneHandle(sc)
    rng = np.random.default_rng(4)
    n = 50000
    uv = rng.uniform(-1.5, 2.5, (n, 2)).astype(np.float32)
    got = np.zeros((n, 4), np.float32)
    lib.crDebugSampleTexture(0, uv.ctypes.data, n, got.ctypes.data)
    # oracle path: a hit on a triangle whose interpolated UV is exactly (u, v): use the shading entry directly
    tex = sc.textures[0].astype(np.float64) / 255.0
    H, W = tex.shape[:2]
    fu = uv[:, 0] - np.floor(uv[:, 0]); fv = uv[:, 1] - np.floor(uv[:, 1])
    xb = fu * np.float32(W) - np.float32(0.5); yb = fv * np.float32(H) - np.float32(0.5)
    xf = np.floor(xb); yf = np.floor(yb)
    a = np.floor((xb - xf) * 256 + 0.5) / 256; b = np.floor((yb - yf) * 256 + 0.5) / 256
    x0 = xf.astype(np.int64) % W; x1 = (x0 + 1) % W; y0 = yf.astype(np.int64) % H; y1 = (y0 + 1) % H
    want = ((1 - a) * (1 - b))[:, None] * tex[y0, x0, :3] + (a * (1 - b))[:, None] * tex[y0, x1, :3] + \
           ((1 - a) * b)[:, None] * tex[y1, x0, :3] + (a * b)[:, None] * tex[y1, x1, :3]
    d = np.abs(got[:, :3] - want)
    assert d.max() < 1e-3 and d.mean() < 2e-5, (d.max(), d.mean())
    assert (got[:, 3] == 1.0).all()


def test_coincident_and_degenerate_triangles(lib, er, loader, oracle, tmp_path):
    """Tie policy and builder robustness: 40 copies of one triangle (identical Morton codes, equal t on
    every hit -> the LOWEST primitive index must win regardless of traversal order), zero-area
    triangles (never hit), and two coplanar overlapping triangles."""
    import base64
    import json
    from tools import synth
    tri = np.array([[-2, -1.5, 3], [2, -1.5, 3], [0, 2.5, 3]], np.float32)
    v = [tri] * 40
    v += [np.array([[0, 0, 2], [0, 0, 2], [1, 1, 2]], np.float32)] * 3            # zero area, in front of the stack
    v += [np.array([[-3, -3, 5], [3, -3, 5], [0, 3, 5]], np.float32), np.array([[-3, -3, 5], [0, 3, 5], [3, -3, 5]], np.float32)]
    v = np.concatenate(v).astype(np.float32)
    blob = v.tobytes()
    gltf = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0, 1]}],
            "nodes": [{"camera": 0, "name": "cam"}, {"mesh": 0, "name": "stack"}],
            "cameras": [{"name": "cam", "type": "perspective", "perspective": {"yfov": 0.5, "znear": 0.1},
                         "extras": {"compound-eye": "TRUE", "compound-projection": "single_dimension_fast", "compound-structure": "e.eye"}}],
            "meshes": [{"name": "stack", "primitives": [{"attributes": {"POSITION": 0}}]}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": len(v), "type": "VEC3", "min": v.min(0).tolist(), "max": v.max(0).tolist()}],
            "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": len(blob)}],
            "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    path = str(tmp_path / "dups.gltf")
    with open(path, "w") as f:
        json.dump(gltf, f)
    omm = synth.fibonacci_eye(200, radius=0.01, acceptance=0.4)
    omm[:, 3:6] = omm[:, 3:6] * 0.3 + np.array([0, 0, 1], np.float32)             # mostly towards +z (glTF camera looks down -z: z axis flips)
    omm[:, 3:6] /= np.linalg.norm(omm[:, 3:6], axis=1, keepdims=True)
    synth.write_eye(str(tmp_path / "e.eye"), omm)
    lib.loadGlTFscene(path.encode())
    assert lib.crDebugGetTriangleCount() == 45 and lib.gotoCameraByName(b"cam")
    lib.setCameraLocalSpace(1, 0, 0, 0, 1, 0, 0, 0, 1)                            # look along +z
    N, S = 200, 8
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    sc, sh, ocam = load_oracle_scene(loader, oracle, path, "cam")
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.make_pose(ocam.position), "single_dimension_fast", samples=S)
    eye.set_render_size(N, 1)
    lib.crDebugSetRayDump(True)
    lib.renderFrame(); eye.render_frame(method="brute")
    o, d, h = _product_rays(lib, N * S)
    lib.crDebugSetRayDump(False)
    assert np.array_equal(h["prim"], eye.last["hits"]["prim"])
    hit = h["prim"] >= 0
    assert hit.sum() > 200
    assert set(np.unique(h["prim"][hit]).tolist()) <= {0, 43, 44}, "coincident copies resolve to primitive 0; degenerate ones never hit"
    assert (h["prim"][hit] == 0).sum() > 100
    assert np.array_equal(_frame(er, lib, N, 1), eye.frame)
