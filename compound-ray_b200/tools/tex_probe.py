"""Probe the CUDA texture unit's bilinear filter (wrap, normalised coords, unorm8 -> float) against
candidate CPU emulations; prints the mismatch statistics of each candidate."""
import os, sys
import numpy as np
sys.path.insert(0, '/root/repo/compound-ray_b200'); sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/benchmarks')
import eye_renderer as er, speed_test
from oracle import gltf_loader
data = speed_test.fixtures()
path = os.path.join(data, 'data', 'natural-standin-sky.gltf')
lib = er.load_library(device=0); lib.setVerbosity(False)
lib.loadGlTFscene(path.encode())
sc = gltf_loader.load_scene(path)
rng = np.random.default_rng(0)
n = 200000
uv = rng.uniform(-1.5, 2.5, (n, 2)).astype(np.float32)
uv[:1000] = (rng.integers(0, 1024, (1000, 2)) + rng.choice([0.0, 0.5, 0.25, 1/256, 255/256], (1000, 2))).astype(np.float32) / 1024
for ti, tex in enumerate(sc.textures):
    out = np.zeros((n, 4), np.float32)
    lib.crDebugSampleTexture(ti, uv.ctypes.data, n, out.ctypes.data)
    H, W = tex.shape[:2]
    t = tex.astype(np.int64)
    u = uv[:, 0].astype(np.float32); v = uv[:, 1].astype(np.float32)
    def emulate(frac_mode, coord_mode):
        fu = (u - np.floor(u)).astype(np.float32); fv = (v - np.floor(v)).astype(np.float32)
        if coord_mode == 'f32':
            xb = (fu * np.float32(W) - np.float32(0.5)).astype(np.float32); yb = (fv * np.float32(H) - np.float32(0.5)).astype(np.float32)
            xf = np.floor(xb); yf = np.floor(yb); a = (xb - xf).astype(np.float64); b = (yb - yf).astype(np.float64)
        else:   # fixed point coordinates with 8 fractional bits computed from float64
            xb = fu.astype(np.float64) * W - 0.5; yb = fv.astype(np.float64) * H - 0.5
            xf = np.floor(xb); yf = np.floor(yb); a = xb - xf; b = yb - yf
        if frac_mode == 'round': wa = np.floor(a * 256 + 0.5); wb = np.floor(b * 256 + 0.5)
        elif frac_mode == 'trunc': wa = np.floor(a * 256); wb = np.floor(b * 256)
        else: wa = a * 256; wb = b * 256
        x0 = xf.astype(np.int64) % W; x1 = (x0 + 1) % W; y0 = yf.astype(np.int64) % H; y1 = (y0 + 1) % H
        res = np.zeros((n, 3))
        for ch in range(3):
            t00 = t[y0, x0, ch]; t10 = t[y0, x1, ch]; t01 = t[y1, x0, ch]; t11 = t[y1, x1, ch]
            res[:, ch] = ((256 - wa) * (256 - wb) * t00 + wa * (256 - wb) * t10 + (256 - wa) * wb * t01 + wa * wb * t11) / (65536.0 * 255.0)
        return res
    for fm in ('round', 'trunc', 'none'):
        for cm in ('f32', 'f64'):
            e = emulate(fm, cm)
            d = np.abs(e - out[:, :3].astype(np.float64))
            print(f"tex{ti} {W}x{H} frac={fm:5s} coord={cm}: max {d.max():.3e} mean {d.mean():.3e} exact-float32 {(e.astype(np.float32) == out[:, :3]).mean():.4f} >1e-6: {(d > 1e-6).mean():.5f}")
