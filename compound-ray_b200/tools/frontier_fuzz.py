"""Randomised A/B campaign for the entry frontier (k_buildEntries): random triangle soups (sizes 1..20k, extents
1e-3..1e4, offsets up to 1e5, clustered / sheet-like / uniform), random eyes (acceptance 0..2 rad, focal offsets,
slightly non-unit axes), random poses (inside / outside / far, rotated).  For every configuration the same RNG
streams are traced with the frontier off and on; hits (prim, t, u, v) and per-ommatidium RGB must be bit-identical.
  python compound-ray_b200/tools/frontier_fuzz.py [--configs 200] [--seed 0]"""
import argparse, base64, json, os, sys, tempfile
import numpy as np
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(_ROOT, "compound-ray_b200")); sys.path.insert(0, _ROOT)
import eye_renderer as er

HIT4 = np.dtype([("prim", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])


def write_soup(path, rng):
    n = int(rng.choice([1, 2, 3, 7, 50, 400, 3000, 20000]))
    extent = float(10.0 ** rng.uniform(-3, 4))
    offset = rng.uniform(-1, 1, 3) * float(10.0 ** rng.uniform(-2, 5)) * float(rng.choice([0, 1]))
    kind = rng.choice(["uniform", "clustered", "sheet", "slivers"])
    if kind == "uniform":
        c = rng.uniform(-1, 1, (n, 1, 3))
    elif kind == "clustered":
        c = rng.normal(size=(n, 1, 3)) * 0.05 + rng.uniform(-1, 1, (1, 1, 3))
    elif kind == "sheet":
        c = rng.uniform(-1, 1, (n, 1, 3)); c[:, :, 1] *= 0.01
    else:
        c = rng.uniform(-1, 1, (n, 1, 3))
    size = float(10.0 ** rng.uniform(-3, 0))
    tri = c + rng.uniform(-1, 1, (n, 3, 3)) * size
    if kind == "slivers":
        tri[:, 2] = tri[:, 0] + (tri[:, 1] - tri[:, 0]) * rng.uniform(0.4, 0.6, (n, 1)) + rng.normal(size=(n, 3)) * 1e-6
    v = (tri * extent + offset).astype(np.float32).reshape(-1, 3)
    col = rng.uniform(0, 1, (len(v), 4)).astype(np.float32)
    blob = v.tobytes() + col.tobytes()
    gltf = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0, 1], "extras": {"background-shader": "simple_sky"}}],
            "nodes": [{"camera": 0, "name": "cam"}, {"mesh": 0, "name": "soup"}],
            "cameras": [{"name": "cam", "type": "perspective", "perspective": {"yfov": 0.5, "znear": 0.1},
                         "extras": {"compound-eye": True, "compound-projection": "single_dimension_fast", "compound-structure": "e.eye"}}],
            "meshes": [{"name": "soup", "primitives": [{"attributes": {"POSITION": 0, "COLOR_0": 1}}]}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": len(v), "type": "VEC3", "min": v.min(0).tolist(), "max": v.max(0).tolist()},
                          {"bufferView": 1, "componentType": 5126, "count": len(v), "type": "VEC4"}],
            "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": v.nbytes}, {"buffer": 0, "byteOffset": v.nbytes, "byteLength": col.nbytes}],
            "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    with open(path, "w") as f:
        json.dump(gltf, f)
    return v.reshape(-1, 3, 3), extent, offset, kind


def random_eye(rng, n, scale):
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    omm = np.zeros((n, 8), np.float32)
    omm[:, 0:3] = d * 0.01 * scale * rng.uniform(0, 1)
    omm[:, 3:6] = d * (1.0 + rng.choice([0.0, 0.0, 3e-5, -6e-5, 1e-3], size=(n, 1)))
    omm[:, 6] = rng.choice([0.0, 1e-4, 0.01, 0.05, 0.2, 0.6, 1.2, 2.0], size=n)
    omm[:, 7] = rng.choice([0.0, 0.0, 0.0, 1e-3, 0.3], size=n) * scale
    return omm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    lib = er.load_library(device=0); lib.setVerbosity(False)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "e.eye"), "w") as f:
        f.write("0 0 0 0 0 1 0.1 0\n")
    N, S = 256, 32
    bad = rays = hits_total = 0
    for cfg in range(args.configs):
        tris, extent, offset, kind = write_soup(os.path.join(tmp, "s.gltf"), rng)
        lib.loadGlTFscene(os.path.join(tmp, "s.gltf").encode())
        assert lib.gotoCameraByName(b"cam")
        lib.setCurrentEyeShaderName(b"single_dimension_fast"); er.setRenderSize(lib, N, 1)
        er.setOmmatidiaFromArray(lib, random_eye(rng, N, extent))
        for pose_i in range(3):
            where = rng.choice(["inside", "inside", "near", "far", "vertex", "vertex"])
            if where == "inside": pos = offset + rng.uniform(-1, 1, 3) * extent
            elif where == "near": pos = offset + rng.normal(size=3) * extent * 2.5
            elif where == "far": pos = offset + rng.normal(size=3) * extent * 300
            else: pos = tris[rng.integers(len(tris)), rng.integers(3)].astype(np.float64) + rng.normal(size=3) * extent * float(rng.choice([0.0, 1e-6, 1e-3, 0.05]))
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            out = []
            for mode in (0, 1):
                lib.crDebugSetEntryFrontier(mode, 2, 0)
                lib.setCurrentEyeSamplesPerOmmatidium(S)
                lib.setCameraPosition(float(pos[0]), float(pos[1]), float(pos[2]))
                er.setCameraLocalSpace(lib, q)
                lib.crDebugSetRayDump(True); lib.renderFrame(); lib.renderFrame(); lib.crDebugSetRayDump(False)
                o = np.zeros((N * S, 3), np.float32); d = np.zeros((N * S, 3), np.float32); h = np.zeros(N * S, HIT4)
                lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data)
                out.append((h.copy(), er.getOmmatidialData(lib).copy()))
            (h0, c0), (h1, c1) = out
            ok = np.array_equal(h0["prim"], h1["prim"]) and np.array_equal(c0.view(np.uint32), c1.view(np.uint32))
            hit = h0["prim"] >= 0
            ok = ok and all(np.array_equal(h0[k][hit].view(np.uint32), h1[k][hit].view(np.uint32)) for k in ("t", "u", "v"))
            rays += N * S; hits_total += int(hit.sum())
            if not ok:
                bad += 1
                print("MISMATCH cfg", cfg, "pose", pose_i, kind, "tris", len(tris), "extent", extent, "offset", offset, "where", where,
                      "differing rays", int((h0["prim"] != h1["prim"]).sum()), file=sys.stderr)
    print(json.dumps({"configs": args.configs, "poses": 3 * args.configs, "rays": rays, "hits": hits_total, "mismatching_frames": bad}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
