#!/usr/bin/env python
"""benchmarks/speed_test.py -- the reference's OWN speed-test protocol against this library.

Method of python-examples/speed-test/speedTest.py:76-129: first compound eye of the scene,
ommatidia := 1000-equidistant.eye (N = 1000), projection single_dimension_fast, render size N x 1,
warm-up, then for each S: setCurrentEyeSamplesPerOmmatidium(S), two throw-away frames, mean of the
value RETURNED by renderFrame() over `--frames` frames.  fps = 1000 / mean ms; rays/s = fps*N*S.

Scenes: the reference's "rothamstead"/"ofstad" speed-test scenes are not in its checkout
(.MISSING_LARGE_BLOBS); the runs here use data/natural-standin-sky.gltf (the stand-in the reference
ships for the natural scene; 24 200 triangles + 1024^2 ground texture) from the fixture archive.
Published reference numbers (other hardware) are printed alongside for context only.
"""
import argparse
import json
import os
import sys
import tarfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "compound-ray_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# python-examples/speed-test/NVIDIA_GeForce_RTX_2080_Ti-rothamstead-...-average-FPSs-(1-3200-rays,500-samples).txt
# and ...-ofstad-... (line number = S), as tabulated in BASELINE.md
PUBLISHED_2080TI_FPS = {"natural": {64: 4427.4, 1000: 1469.6, 3200: 572.0},
                        "ofstad": {1: 5417.6, 32: 5251.6, 64: 4785.7, 1000: 1566.1, 3200: 612.3}}


def fixtures():
    data = os.path.join(ROOT, "tests", "_data")
    if not os.path.exists(os.path.join(data, "data", "natural-standin-sky.gltf")):
        os.makedirs(data, exist_ok=True)
        with tarfile.open(os.path.join(ROOT, "tests", "golden", "reference_data.tar.gz")) as tar:
            tar.extractall(data, filter="data")
    return data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", default="1,2,4,8,16,32,64,128,256,512,1000,2000,3200")
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--warmup-seconds", type=float, default=2.0)
    ap.add_argument("--gltf", default=None)
    ap.add_argument("--eye", default=None)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    import eye_renderer as er
    data = fixtures()
    gltf = args.gltf or os.path.join(data, "data", "natural-standin-sky.gltf")
    eye = args.eye or os.path.join(data, "data", "eyes", "1000-equidistant.eye")
    lib = er.load_library(device=args.device)
    lib.setVerbosity(False)
    lib.loadGlTFscene(gltf.encode())
    er.gotoFirstCompoundEye(lib)
    er.setOmmatidiaFromOmmatidiumList(lib, er.readEyeFile(eye))
    lib.setCurrentEyeShaderName(b"single_dimension_fast")
    N = lib.getCurrentEyeOmmatidialCount()
    er.setRenderSize(lib, N, 1)
    lib.setCurrentEyeSamplesPerOmmatidium(1)
    t0 = time.time()
    while time.time() - t0 <= args.warmup_seconds:
        lib.renderFrame()
    rows = []
    for S in [int(s) for s in args.samples.split(",")]:
        lib.setCurrentEyeSamplesPerOmmatidium(S)
        lib.renderFrame()
        lib.renderFrame()
        total = 0.0
        wall0 = time.perf_counter()
        for _ in range(args.frames):
            total += lib.renderFrame()
        wall = time.perf_counter() - wall0
        ms = total / args.frames
        row = {"S": S, "ms_per_frame": ms, "fps": 1000.0 / ms, "rays_per_sec": 1000.0 / ms * N * S,
               "wall_fps_incl_python": args.frames / wall,
               "published_2080ti_fps_natural": PUBLISHED_2080TI_FPS["natural"].get(S),
               "published_2080ti_fps_ofstad": PUBLISHED_2080TI_FPS["ofstad"].get(S)}
        rows.append(row)
        print("S=%5d  %8.4f ms/frame  %9.1f fps  %8.3f Grays/s   (published RTX 2080 Ti: natural %s, ofstad %s fps)" % (
            S, ms, row["fps"], row["rays_per_sec"] / 1e9, row["published_2080ti_fps_natural"], row["published_2080ti_fps_ofstad"]),
            file=sys.stderr)
    print(json.dumps({"benchmark": "reference speed-test protocol (speedTest.py:76-129)", "scene": os.path.basename(gltf),
                      "eye": os.path.basename(eye), "ommatidia": int(N), "frames_per_point": args.frames, "rows": rows}))
    lib.stop()


if __name__ == "__main__":
    main()
