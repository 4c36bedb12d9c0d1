"""Experiment: how many node visits per ray remain if each ommatidium's rays start from a small
frontier of subtree roots found by a conservative cone/box test."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, '/root/repo/compound-ray_b200'); sys.path.insert(0, '/root/repo')
import eye_renderer as er
import bench
from oracle import oracle as O
gltf, _ = bench.make_workload(1_000_000, 10_000)
lib = er.load_library(device=0); lib.setVerbosity(False)
lib.loadGlTFscene(gltf.encode()); lib.gotoCameraByName(b"compound-cam")
lib.setCurrentEyeShaderName(b"single_dimension_fast"); er.setRenderSize(lib, 10000, 1)
N = 10000; S = 32
lib.setCurrentEyeSamplesPerOmmatidium(S)
lib.crDebugSetRayDump(True); lib.renderFrame(); lib.crDebugSetRayDump(False)
n = N * S
o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); h = np.zeros((n, 4), np.int32)
lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data)
T = lib.crDebugGetTriangleCount(); nn = lib.crDebugGetBvhNodeCount()
nodes = np.zeros((nn, 16), np.float32); tris = np.zeros((T, 12), np.float32)
lib.crDebugCopyBvh(nodes.ctypes.data, tris.ctypes.data)
o_os = np.ascontiguousarray(o.reshape(S, N, 3).transpose(1, 0, 2)); d_os = np.ascontiguousarray(d.reshape(S, N, 3).transpose(1, 0, 2))
omm = np.zeros((N, 8), np.float32)
from oracle import gltf_loader
sc = gltf_loader.load_scene(gltf); cam = [c for c in sc.cameras if c.kind == "compound"][0]
acc = np.asarray(cam.ommatidia, np.float32).reshape(-1, 8)[:, 6]
print("acceptance (deg) min/mean/max", np.degrees(acc.min()), np.degrees(acc.mean()), np.degrees(acc.max()), file=sys.stderr)
hits, cnt = O.trace_device_bvh(nodes, tris, o, d, np.zeros(n, np.float32))
print("baseline nodes/ray %.2f tris/ray %.2f" % (cnt[0] / n, cnt[1] / n), file=sys.stderr)
L = O.lib()
for nsig, mode in ((4.0, 0), (4.0, 1)):
    halfs = (acc / 2.35482 * nsig).astype(np.float32)
    for K in (1, 2, 3, 4, 6, 8):
        c = np.zeros(4, np.int64)
        L.cro_sim_frontier(nodes.ctypes.data_as(C.c_void_p), C.c_int64(nn), tris.ctypes.data_as(C.c_void_p),
                           o_os.ctypes.data_as(C.c_void_p), d_os.ctypes.data_as(C.c_void_p), C.c_int64(N), C.c_int64(S),
                           halfs.ctypes.data_as(C.c_void_p), C.c_int(K), C.c_int(mode), c.ctypes.data_as(C.c_void_p))
        print("mode", mode, "nsig %.0f K %2d: nodes/ray %.2f tris/ray %.2f frontier/omm %.2f outside %.5f" % (nsig, K, c[0] / n, c[1] / n, c[2] / N, c[3] / n), file=sys.stderr)
