"""Randomised exactness campaign for the device BVH (LBVH build + conservative slab test + tie rule): random triangle
soups (uniform / clustered / sheet / slivers / duplicated triangles, extents 1e-3..1e4, offsets up to 1e5) traced
with arbitrary rays (random, axis-parallel, zero components, origins on vertices, aimed at vertices and edge
midpoints, tmin > 0) by the library (crDebugTraceRays) and by the CPU oracle's BRUTE-FORCE loop over the same
flattened triangles.  Hit primitive and (t, u, v) must be bit-identical.  The oracle is the checker here.
The one admitted difference: a NUMERICALLY DEGENERATE ray/triangle pair (|det| < 1e-5 |d||e1||e2|: a sliver met
edge-on, e.g. a ray aimed exactly at a vertex of a needle triangle), where the binary32 Moller-Trumbore test itself
returns a t that is wrong by per cent, so the "hit point" lies outside the triangle's bounding box; the BVH, which also
requires the ray to pass that box within (tmin, best t], rejects it, a loop over all triangles does not.  Such rays
are counted separately (`degenerate_differences`); any other difference fails the run.
  python compound-ray_b200/tools/bvh_fuzz.py [--configs 100] [--seed 0]"""
import argparse, json, os, sys, tempfile
import numpy as np
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(_ROOT, "compound-ray_b200")); sys.path.insert(0, _ROOT)
import eye_renderer as er
from oracle import oracle as O
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import frontier_fuzz as FF


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--rays", type=int, default=20000)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    lib = er.load_library(device=0); lib.setVerbosity(False)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "e.eye"), "w") as f:
        f.write("0 0 0 0 0 1 0.1 0\n")
    bad = rays = hits_total = degenerate = 0
    n = args.rays
    for cfg in range(args.configs):
        while True:
            tris, extent, offset, kind = FF.write_soup(os.path.join(tmp, "s.gltf"), rng)
            if len(tris) <= 3000:
                break
        lib.loadGlTFscene(os.path.join(tmp, "s.gltf").encode())
        T = lib.crDebugGetTriangleCount()
        flat = np.zeros((T, 9), np.float32)
        lib.crDebugCopyTriangles(flat.ctypes.data)                  # v0, e1, e2 per flattened primitive, as the kernels see them
        verts = tris.reshape(-1, 3).astype(np.float32)
        o = (offset + rng.uniform(-1.5, 1.5, (n, 3)) * extent).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        q = n // 10
        d[0:q // 2, 0] = 0.0; d[q // 2:q, 1] = 0.0
        d[q:q + q // 3] = np.array([1, 0, 0], np.float32); d[q + q // 3:q + 2 * (q // 3)] = np.array([0, -1, 0], np.float32)
        d[q + 2 * (q // 3):2 * q] = np.array([0, 0, 1], np.float32)
        o[2 * q:3 * q] = verts[rng.integers(len(verts), size=q)]                      # origins exactly on vertices
        tgt = verts[rng.integers(len(verts), size=2 * q)]
        tgt[q:] = 0.5 * (tgt[q:] + verts[rng.integers(len(verts), size=q)])            # vertices and (pseudo) edge midpoints
        d[3 * q:5 * q] = tgt - o[3 * q:5 * q]
        d[5 * q:6 * q] *= np.float32(10.0 ** rng.uniform(-6, 6))                        # un-normalised directions
        tmin = np.zeros(n, np.float32)
        tmin[::3] = np.float32(0.01 * extent)
        hits8 = np.zeros((n, 8), np.int32)
        lib.crDebugTraceRays(o.ctypes.data, d.ctypes.data, tmin.ctypes.data, n, hits8.ctypes.data)
        oh = np.empty(n, dtype=O.HIT_DTYPE)
        O.lib().cro_trace_bruteforce(flat.ctypes.data_as(O.C.c_void_p), T, o.ctypes.data_as(O.C.c_void_p), d.ctypes.data_as(O.C.c_void_p),
                                     tmin.ctypes.data_as(O.C.c_void_p), n, np.float32(1e16), oh.ctypes.data_as(O.C.c_void_p))
        hit = oh["prim"] >= 0
        ok = np.array_equal(hits8[:, 0], oh["prim"])
        ok = ok and all(np.array_equal(hits8[hit, 1 + i].view(np.uint32), oh[k][hit].view(np.uint32)) for i, k in enumerate(("t", "u", "v")))
        rays += n; hits_total += int(hit.sum())
        if not ok:
            same = (hits8[:, 0] == oh["prim"])
            for i, k in enumerate(("t", "u", "v")):
                same &= ~hit | (hits8[:, 1 + i].view(np.uint32) == oh[k].view(np.uint32))
            unexplained = 0
            for r in np.where(~same)[0]:
                cond = []
                for prim in {int(hits8[r, 0]), int(oh["prim"][r])} - {-1}:
                    D = d[r].astype(np.float64); e1 = flat[prim, 3:6].astype(np.float64); e2 = flat[prim, 6:9].astype(np.float64)
                    cond.append(abs(e1 @ np.cross(D, e2)) / (np.linalg.norm(D) * np.linalg.norm(e1) * np.linalg.norm(e2) + 1e-300))
                if cond and min(cond) < 1e-5:
                    degenerate += 1
                else:
                    unexplained += 1
            if unexplained:
                bad += 1
                print("MISMATCH cfg", cfg, kind, "tris", T, "extent", extent, "offset", offset, "unexplained rays", unexplained, file=sys.stderr)
    print(json.dumps({"configs": args.configs, "rays": rays, "hits": hits_total, "degenerate_differences": degenerate, "mismatching_configs": bad}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
