"""Per-warp traversal-length statistics on the headline workload: how much of the SIMD loss is inherent
variance of node counts between the 32 samples of one ommatidium."""
import os, sys
import numpy as np
sys.path.insert(0, '/root/repo/compound-ray_b200'); sys.path.insert(0, '/root/repo')
import eye_renderer as er
import bench
gltf, _ = bench.make_workload(1_000_000, 10_000)
lib = er.load_library(device=0); lib.setVerbosity(False)
lib.loadGlTFscene(gltf.encode()); lib.gotoCameraByName(b"compound-cam")
lib.setCurrentEyeShaderName(b"single_dimension_fast"); er.setRenderSize(lib, 10000, 1)
S = 32
lib.setCurrentEyeSamplesPerOmmatidium(S)
lib.crDebugSetRayDump(True); lib.renderFrame(); lib.crDebugSetRayDump(False)
n = 10000 * S
o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); h = np.zeros((n, 4), np.int32)
lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data)
hits8 = np.zeros((n, 8), np.int32)
tm = np.zeros(n, np.float32)
lib.crDebugTraceRays(o.ctypes.data, d.ctypes.data, tm.ctypes.data, n, hits8.ctypes.data)
nc = hits8[:, 4].astype(np.float64); tc = hits8[:, 5].astype(np.float64)
# dump order is stream id N*s+o -> regroup as [o][s] (the kernel's warp = 32 samples of one ommatidium)
nc_os = nc.reshape(S, 10000).T; tc_os = tc.reshape(S, 10000).T
print("nodes/ray mean %.2f  tris/ray mean %.2f" % (nc.mean(), tc.mean()))
eff = nc_os.mean(axis=1) / np.maximum(nc_os.max(axis=1), 1)
print("node-loop SIMD efficiency if a warp waits for its longest ray: %.3f (weighted %.3f)" % (eff.mean(), nc_os.sum() / (nc_os.max(axis=1).sum() * S)))
hit = (h[:, 0] >= 0).reshape(S, 10000).T
mixed = (hit.any(axis=1) & ~hit.all(axis=1)).mean()
print("warps with mixed hit/miss rays: %.3f" % mixed)
srt = np.sort(nc_os.reshape(-1)); 
print("percentiles of nodes/ray:", [float(np.percentile(nc, p)) for p in (5, 25, 50, 75, 95, 99)])
# hypothetical: sort each ommatidium block of 1024 (32 warps) by node count before forming warps -- upper bound of regrouping
lib.setCurrentEyeSamplesPerOmmatidium(1024)
