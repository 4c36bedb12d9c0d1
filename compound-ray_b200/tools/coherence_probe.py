"""Experiment: would grouping the samples of one ommatidium into warps by direction (instead of by
sample index) even out the per-lane traversal lengths?  Uses the device-side per-ray counters."""
import os, sys
import numpy as np
sys.path.insert(0, '/root/repo/compound-ray_b200'); sys.path.insert(0, '/root/repo')
import eye_renderer as er
import bench
gltf, _ = bench.make_workload(1_000_000, 10_000)
lib = er.load_library(device=0); lib.setVerbosity(False)
lib.loadGlTFscene(gltf.encode()); lib.gotoCameraByName(b"compound-cam")
lib.setCurrentEyeShaderName(b"single_dimension_fast")
from oracle import gltf_loader
sc = gltf_loader.load_scene(gltf); cam = [c for c in sc.cameras if c.kind == "compound"][0]
omm = np.asarray(cam.ommatidia, np.float32).reshape(-1, 8)[::10].copy()      # 1000 evenly strided ommatidia
N = len(omm); S = 1024
er.setOmmatidiaFromArray(lib, omm); er.setRenderSize(lib, N, 1)
lib.setCurrentEyeSamplesPerOmmatidium(S)
lib.crDebugSetRayDump(True); lib.renderFrame(); lib.crDebugSetRayDump(False)
n = N * S
o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); h = np.zeros((n, 4), np.int32); c = np.zeros((n, 2), np.int32)
lib.crDebugCopyLastRays(o.ctypes.data, d.ctypes.data, h.ctypes.data); lib.crDebugCopyLastRayCounts(c.ctypes.data)
d = d.reshape(S, N, 3).transpose(1, 0, 2); c = c.reshape(S, N, 2).transpose(1, 0, 2).astype(np.float64)
cost = 42.0 * c[..., 0] + 45.0 * c[..., 1]                       # thread instructions, roughly
print("nodes/ray %.2f tris/ray %.2f" % (c[..., 0].mean(), c[..., 1].mean()), file=sys.stderr)
def util(cost_os):                                               # [N][S] -> sum / (32 * sum of warp maxima)
    w = cost_os.reshape(N, S // 32, 32)
    return cost_os.sum() / (32.0 * w.max(axis=2).sum())
print("lane utilisation, warps by sample index: %.3f" % util(cost), file=sys.stderr)
axis = omm[:, 3:6] / np.linalg.norm(omm[:, 3:6], axis=1, keepdims=True)
keys = np.empty((N, S)); rad = np.empty((N, S))
for i in range(N):
    a = axis[i]; t = np.array([1.0, 0, 0]) if abs(a[0]) < 0.57 else np.array([0, 1.0, 0])
    u = np.cross(a, t); u /= np.linalg.norm(u); v = np.cross(a, u)
    dd = d[i] / np.linalg.norm(d[i], axis=1, keepdims=True)
    keys[i] = np.arctan2(dd @ v, dd @ u); rad[i] = np.hypot(dd @ v, dd @ u)
os.makedirs('/root/repo/gpurun_out', exist_ok=True)
np.savez_compressed('/root/repo/gpurun_out/coherence_dump.npz', d=d.astype(np.float32), cost=cost.astype(np.float32), keys=keys.astype(np.float32), rad=rad.astype(np.float32), nodes=c[..., 0].astype(np.int16), tris=c[..., 1].astype(np.int16), hit=(h[:, 0] >= 0).reshape(S, N).T)
