"""tests/golden/make_fixtures.py -- regenerates the committed fixtures (run in the dev container,
where /root/reference and the CUDA toolkit exist; the GPU box has neither the reference tree).

  reference_data.tar.gz  INPUT DATA only (scenes + .eye tables) packed from the reference's data
                         directories, so the parity tests can run on the reference's own scenes on
                         a box without /root/reference.  No reference source code is included.
  xorwow_kat.json        cuRAND's own host XORWOW implementation (oracle/_ref/curand_kat)
  sutil_kat.json         the reference's sutil math headers evaluated on fixed inputs (oracle/_ref/sutil_kat)
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


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    with open(os.path.join(HERE, "xorwow_kat.json"), "w") as f:
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "curand_kat")], stdout=f)
    with open(os.path.join(HERE, "sutil_kat.json"), "w") as f:
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "sutil_kat")], stdout=f)
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
