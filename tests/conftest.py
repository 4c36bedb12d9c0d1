"""Shared fixtures.  `-m "not gpu"` runs everywhere; `-m gpu` needs a B200 and calls the product
only through the C ABI of compound-ray_b200/lib/libEyeRenderer3.so.  Nothing here reads
/root/reference at run time: the reference's scenes/eyes come from tests/golden/reference_data.tar.gz."""
import os
import sys
import tarfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "compound-ray_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

DATA_DIR = os.path.join(ROOT, "tests", "_data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def ref_data():
    """Directory holding the reference's input scenes and eye files (unpacked fixture archive)."""
    stamp = os.path.join(DATA_DIR, ".unpacked")
    arc = os.path.join(ROOT, "tests", "golden", "reference_data.tar.gz")
    if not os.path.exists(stamp) or os.path.getmtime(stamp) < os.path.getmtime(arc):
        os.makedirs(DATA_DIR, exist_ok=True)
        with tarfile.open(arc) as tar:
            tar.extractall(DATA_DIR, filter="data")
        open(stamp, "w").close()
    return DATA_DIR


@pytest.fixture(scope="session")
def ref_outputs():
    """Directory holding frames the reference itself rendered (tests/golden/reference_outputs.tar.gz)."""
    out = os.path.join(DATA_DIR, "reference-outputs")
    stamp = os.path.join(out, ".unpacked")
    arc = os.path.join(ROOT, "tests", "golden", "reference_outputs.tar.gz")
    if not os.path.exists(stamp) or os.path.getmtime(stamp) < os.path.getmtime(arc):
        os.makedirs(out, exist_ok=True)
        with tarfile.open(arc) as tar:
            tar.extractall(out, filter="data")
        open(stamp, "w").close()
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def loader():
    from oracle import gltf_loader
    return gltf_loader


@pytest.fixture(scope="session")
def er():
    import eye_renderer
    return eye_renderer


@pytest.fixture(scope="session")
def lib(er):
    """The product library (loads without a GPU; rendering calls need one)."""
    L = er.load_library()
    L.setVerbosity(False)
    # The parity tests use small frames; by default those skip the entry-frontier pass (it only pays above
    # ~0.5M rays per frame).  Force it on for S >= 2 so every oracle comparison also exercises that path.
    L.crDebugSetEntryFrontier(1, 2, 0)
    # ... and the per-ommatidium candidate lists, which by default are only built in batches of >= 4 frames
    L.crDebugSetCandidateLists(2)
    return L


@pytest.fixture(scope="session")
def synth_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("synth"))


def load_oracle_scene(loader, oracle, path, camera_name=None):
    sc = loader.load_scene(path)
    sh = oracle.SceneHandle(sc)
    cam = None
    for c in sc.cameras:
        if (camera_name is None and c.kind == "compound") or c.name == camera_name:
            cam = c
            break
    return sc, sh, cam
