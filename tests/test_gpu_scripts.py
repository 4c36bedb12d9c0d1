"""The reference's own Python, UNMODIFIED, against this library on a GPU (INTEGRATION.md's drop-in claim).

tests/golden/reference_data.tar.gz carries, byte for byte, python-examples/eyeRendererHelperFunctions.py,
python-examples/primary-example.py and position-estimation-toy-experiment/compoundRayIterators.py.  They are laid out
the way the reference's checkout lays them out, the library is put where they look for it
(`../build/make/lib/libEyeRenderer3.so`, primary-example.py:20; eyeRendererPaths.EYE_RENDERER_LIB_PATH,
compoundRayIterators.py:32) and they are run as separate processes.  Nothing in the scripts is edited; the only
interpreter-level accommodation is that time.sleep and PIL's Image.show are no-ops (a headless box has no viewer and the
example sleeps 5 s between cameras).  What they write is then compared with the same call sequence issued in-process."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_oracle_scene

pytestmark = pytest.mark.gpu

RUNNER = ("import sys, time, runpy; time.sleep = lambda s: None\n"
          "import PIL.Image as I; I.Image.show = lambda self, *a, **k: None\n"
          "runpy.run_path(sys.argv[1], run_name='__main__')\n")


def _checkout(tmp_path, ref_data, er):
    """A directory that looks like the reference's checkout: python-examples/, data/, build/make/lib/."""
    root = tmp_path / "compound-ray"
    (root / "build" / "make" / "lib").mkdir(parents=True)
    os.symlink(os.environ.get("CR_LIB_PATH") or er.LIB_PATH, root / "build" / "make" / "lib" / "libEyeRenderer3.so")
    shutil.copytree(os.path.join(ref_data, "python-examples"), root / "python-examples")
    shutil.copytree(os.path.join(ref_data, "data"), root / "data")
    return root


def test_unmodified_primary_example_runs_and_renders(lib, er, ref_data, loader, oracle, tmp_path):
    """python-examples/primary-example.py as shipped.  Its scene (data/ofstad-arena/ofstad-acceptance-angle.gltf) is one
    of the reference's missing large blobs; the synthetic arena of SURVEY 8(d)3 (JPEG-textured cylinder + floor, a
    compound and a panoramic camera) is written to that path.  The script's nine PPMs must exist and equal, byte for
    byte, what the same call sequence gives in this process; its 100-sample compound frame is also held against the
    oracle (textured scene: every pixel within one 8-bit step)."""
    from tools import synth
    root = _checkout(tmp_path, ref_data, er)
    arena = root / "data" / "ofstad-arena"
    arena.mkdir()
    shutil.copy(root / "data" / "eyes" / "1000-equidistant.eye", arena / "1000-equidistant.eye")
    gltf = str(arena / "ofstad-acceptance-angle.gltf")
    synth.write_arena_gltf(gltf, eye_file="1000-equidistant.eye", projection="spherical_orientationwise")
    r = subprocess.run([sys.executable, "-c", RUNNER, "primary-example.py"], cwd=root / "python-examples", capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Successfully loaded" in r.stdout and "Traceback" not in r.stderr
    out = root / "python-examples" / "test-images"
    names = [f"test-image-{i}.ppm" for i in range(5)] + [f"test-image-{i}-100samples.ppm" for i in (0, 2, 4)]
    for n in names:
        assert (out / n).exists(), (n, r.stdout[-1500:])
    assert r.stdout.count("This one's a compound eye") == 3 and r.stdout.count("rendered in") == 5

    # the same calls, in this process (primary-example.py:26-95)
    mine = tmp_path / "mine"
    mine.mkdir()
    lib.loadGlTFscene(gltf.encode())
    lib.setRenderSize(200, 200)
    for i in range(5):
        assert lib.renderFrame() > 0
        lib.displayFrame()
        lib.saveFrameAs(str(mine / f"test-image-{i}.ppm").encode())
        if lib.isCompoundEyeActive():
            lib.setCurrentEyeSamplesPerOmmatidium(100)
            lib.renderFrame()
            lib.saveFrameAs(str(mine / f"test-image-{i}-100samples.ppm").encode())
            if i == 0:
                frame100 = er.getFrame(lib, 200, 200)
        lib.nextCamera()
    for n in names:
        assert (out / n).read_bytes() == (mine / n).read_bytes(), n
    hdr = b"P6\n200 200\n255\n"
    raw = (out / "test-image-0-100samples.ppm").read_bytes()
    assert raw.startswith(hdr)
    assert np.array_equal(np.frombuffer(raw[len(hdr):], np.uint8).reshape(200, 200, 3), frame100[::-1, :, :3])

    # ... and the checker on that frame: frame 0 at S = 1, then the streams restart at S = 100
    sc, sh, ocam = load_oracle_scene(loader, oracle, gltf, "compound-cam")
    tex = np.zeros((1024, 1024, 4), np.uint8)
    lib.crDebugCopyTexture(0, tex.ctypes.data)
    assert np.abs(tex.astype(int) - sc.textures[0].astype(int)).max() <= 6        # stb_image-exact decoder vs PIL's libjpeg
    sc.textures[0] = tex                                                           # same texels on both sides
    sh = oracle.SceneHandle(sc)
    eye = oracle.CompoundEyeOracle(sh, ocam.ommatidia, oracle.pose_from_camera(ocam), ocam.projection, samples=100)
    eye.set_render_size(200, 200)
    eye.render_frame(method="bvh")
    diff = np.abs(eye.frame.astype(int) - frame100.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.02, (int(diff.max()), float((diff > 0).mean()))
    assert frame100[:, :, :3].std() > 10


def test_unmodified_compound_ray_iterators(lib, er, ref_data, tmp_path):
    """position-estimation-toy-experiment/compoundRayIterators.py as shipped (RandomCubeIterator: env_2.gltf, AM_60185
    eye, single_dimension_fast, S = 1000, setCameraPosition + renderFrame + getFramePointer per item,
    :27-102), configured only through the eyeRendererPaths module it imports.  Its first items equal, byte for byte,
    the batched iterator of compound-ray_b200/iterators.py under the same numpy seed."""
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    (cfg / "eyeRendererPaths.py").write_text(
        f"PYTHON_EXAMPLES_PATH = {os.path.join(ref_data, 'python-examples')!r}\n"
        f"EYE_RENDERER_LIB_PATH = {(os.environ.get('CR_LIB_PATH') or er.LIB_PATH)!r}\n")
    out = tmp_path / "items.npz"
    code = ("import sys, numpy as np\n"
            f"sys.path.insert(0, {str(cfg)!r}); sys.path.insert(0, {os.path.join(ref_data, 'python-examples', 'position-estimation-toy-experiment')!r})\n"
            "import compoundRayIterators as C\n"
            "np.random.seed(77)\n"
            "it = iter(C.RandomCubeIterator('sim-environment/eyes/AM_60185-real.eye'))\n"
            "items = [next(it) for _ in range(3)]\n"
            f"np.savez({str(out)!r}, imgs=np.stack([i.numpy() for i, _ in items]), pos=np.stack([p.numpy() for _, p in items]))\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ref_data, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(out)
    assert got["imgs"].shape == (3, 1, 6374, 3) and got["imgs"].std() > 1
    import iterators
    np.random.seed(77)
    it = iter(iterators.RandomCubeIterator(os.path.join(ref_data, "sim-environment", "eyes", "AM_60185-real.eye"),
                                           scenePath=os.path.join(ref_data, "sim-environment", "env_2.gltf"), samples=1000, blockSize=3))
    for k in range(3):
        img, pos = next(it)
        assert np.array_equal(img.numpy(), got["imgs"][k]), k
        assert np.allclose(pos.numpy(), got["pos"][k])
