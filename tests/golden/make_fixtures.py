"""tests/golden/make_fixtures.py -- regenerates the committed fixtures (run in the dev container,
where /root/reference and the CUDA toolkit exist; the GPU box has neither the reference tree).

  reference_data.tar.gz  INPUT DATA only (scenes + .eye tables) packed from the reference's data
                         directories, so the parity tests can run on the reference's own scenes on
                         a box without /root/reference.  No reference source code is included.
  xorwow_kat.json        cuRAND's own host XORWOW implementation (oracle/_ref/curand_kat)
  sutil_kat.json         the reference's sutil math headers evaluated on fixed inputs (oracle/_ref/sutil_kat)
  jpeg/*.jpg             small JPEG test streams written with PIL (4:4:4 / 4:2:2 / 4:2:0 / grey / progressive)
  stb_jpeg_kat.{json,npz} those streams decoded by the reference's vendored stb_image.h (oracle/_ref/stb_kat)
"""
import os
import subprocess
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
TOY = "python-examples/position-estimation-toy-experiment/sim-environment"

FILES = {
    # archive name                                  : path under /root/reference
    "data/test-scene/test-scene.gltf": "data/test-scene/test-scene.gltf",
    "data/test-scene/test-scene-sky.gltf": "data/test-scene/test-scene-sky.gltf",
    "data/test-scene/test.eye": "data/test-scene/test.eye",
    "data/test-scene/test100.eye": "data/test-scene/test100.eye",
    "data/natural-standin-sky.gltf": "data/natural-standin-sky.gltf",
    "data/eyes/1000-equidistant.eye": "data/eyes/1000-equidistant.eye",
    "data/eyes/1000-horizontallyAcute-variableDegree.eye": "data/eyes/1000-horizontallyAcute-variableDegree.eye",
    "sim-environment/env_2.gltf": TOY + "/env_2.gltf",
    "sim-environment/eyes/AM_60185-real.eye": TOY + "/eyes/AM_60185-real.eye",
}


def make_jpeg_kat():
    import json
    import sys
    import tempfile
    import numpy as np
    from PIL import Image
    sys.path.insert(0, ROOT)
    from oracle import gltf_loader
    tex = gltf_loader.load_scene(os.path.join(REF, "data", "natural-standin-sky.gltf")).textures[0][:, :, :3]
    rng = np.random.default_rng(0)
    out = os.path.join(HERE, "jpeg")
    os.makedirs(out, exist_ok=True)

    def save(name, arr, **kw):
        Image.fromarray(arr).save(os.path.join(out, name), format="JPEG", **kw)

    crop = np.ascontiguousarray(tex[100:145, 200:267])
    save("tex_444_q90.jpg", crop, quality=90, subsampling=0)
    save("tex_422_q75.jpg", crop, quality=75, subsampling=1)
    save("tex_420_q60.jpg", crop, quality=60, subsampling=2)
    save("tex_420_128x96.jpg", np.ascontiguousarray(tex[0:96, 0:128]), quality=85, subsampling=2)
    noise = rng.integers(0, 256, (33, 17, 3), dtype=np.uint8)
    save("noise_420_q95.jpg", noise, quality=95, subsampling=2)
    save("noise_444_q100.jpg", noise, quality=100, subsampling=0)
    yy, xx = np.mgrid[0:40, 0:50]
    Image.fromarray(((xx * 5 + yy * 3) % 256).astype(np.uint8), "L").save(os.path.join(out, "grey_q80.jpg"), format="JPEG", quality=80)
    save("pixel_420.jpg", np.full((1, 1, 3), (200, 30, 90), np.uint8), quality=90, subsampling=2)
    save("tex_420_optimized.jpg", crop, quality=70, subsampling=2, optimize=True)
    save("tex_progressive.jpg", crop, quality=70, progressive=True)
    # progressive streams (SOF2): spectral selection + successive approximation as written by libjpeg(-turbo)
    save("prog_444_q90.jpg", crop, quality=90, subsampling=0, progressive=True)
    save("prog_422_q75.jpg", crop, quality=75, subsampling=1, progressive=True)
    save("prog_420_odd_q85.jpg", np.ascontiguousarray(tex[3:80, 5:136]), quality=85, subsampling=2, progressive=True)
    save("prog_noise_420_q95.jpg", noise, quality=95, subsampling=2, progressive=True, optimize=True)
    save("prog_noise_444_q30.jpg", noise, quality=30, subsampling=0, progressive=True)
    save("prog_pixel_420.jpg", np.full((1, 1, 3), (200, 30, 90), np.uint8), quality=90, subsampling=2, progressive=True)
    save("prog_420_restart.jpg", crop, quality=80, subsampling=2, progressive=True, restart_marker_blocks=3)
    save("base_420_restart.jpg", crop, quality=80, subsampling=2, restart_marker_rows=1)
    Image.fromarray(((xx * 5 + yy * 3) % 256).astype(np.uint8), "L").save(os.path.join(out, "prog_grey_q80.jpg"), format="JPEG", quality=80,
                                                                           progressive=True)
    files = sorted(os.path.join(out, f) for f in os.listdir(out) if f.endswith(".jpg"))
    with tempfile.TemporaryDirectory() as tmp:
        idx = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "stb_kat"), tmp] + files, text=True)
        with open(os.path.join(HERE, "stb_jpeg_kat.json"), "w") as f:
            f.write(idx)
        arrs = {}
        for name, e in json.loads(idx).items():
            if "error" not in e:
                arrs[name] = np.fromfile(os.path.join(tmp, name + ".rgba"), dtype=np.uint8).reshape(e["height"], e["width"], 4)
        np.savez_compressed(os.path.join(HERE, "stb_jpeg_kat.npz"), **arrs)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    with open(os.path.join(HERE, "xorwow_kat.json"), "w") as f:
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "curand_kat")], stdout=f)
    with open(os.path.join(HERE, "sutil_kat.json"), "w") as f:
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "sutil_kat")], stdout=f)
    make_jpeg_kat()
    out = os.path.join(HERE, "reference_data.tar.gz")
    with tarfile.open(out, "w:gz", compresslevel=9) as tar:
        for arc, src in sorted(FILES.items()):
            info = tar.gettarinfo(os.path.join(REF, src), arcname=arc)
            info.mtime = 0
            info.uid = info.gid = 0
            info.uname = info.gname = ""
            with open(os.path.join(REF, src), "rb") as fh:
                tar.addfile(info, fh)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
