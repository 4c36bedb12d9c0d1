#!/usr/bin/env python
"""benchmarks/speed_probe.py -- per-frame time series of the speed-test protocol at chosen S (diagnostics for outliers)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "compound-ray_b200"))
import numpy as np
import eye_renderer as er
from benchmarks.speed_test import fixtures

def main():
    Ss = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8,32").split(",")]
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    data = fixtures()
    lib = er.load_library(device=0)
    lib.setVerbosity(False)
    lib.loadGlTFscene(os.path.join(data, "data", "natural-standin-sky.gltf").encode())
    er.gotoFirstCompoundEye(lib)
    er.setOmmatidiaFromOmmatidiumList(lib, er.readEyeFile(os.path.join(data, "data", "eyes", "1000-equidistant.eye")))
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N = lib.getCurrentEyeOmmatidialCount()
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(1)
    t0 = time.time()
    while time.time() - t0 <= 1.0:
        lib.renderFrame()
    for S in Ss:
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        lib.renderFrame(); lib.renderFrame()
        ms = np.array([lib.renderFrame() for _ in range(frames)])
        top = np.argsort(ms)[::-1][:8]
        print(json.dumps({"S": S, "mean_ms": float(ms.mean()), "mean_without_top3_ms": float(np.sort(ms)[:-3].mean()), "median_ms": float(np.median(ms)), "p90_ms": float(np.percentile(ms, 90)),
                          "over_1ms": int((ms > 1.0).sum()), "over_100us": int((ms > 0.1).sum()),
                          "top": [(int(i), round(float(ms[i]), 3)) for i in top]}))
    lib.stop()

if __name__ == "__main__":
    main()
