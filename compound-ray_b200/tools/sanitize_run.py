"""Tiny tour of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
  compute-sanitizer --tool memcheck python compound-ray_b200/tools/sanitize_run.py"""
import os, sys, tarfile
import numpy as np
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(_ROOT, "compound-ray_b200")); sys.path.insert(0, _ROOT)
import eye_renderer as er
data = os.path.join(_ROOT, "tests", "_data")
if not os.path.exists(os.path.join(data, "data", "test-scene", "test-scene.gltf")):
    os.makedirs(data, exist_ok=True)
    with tarfile.open(os.path.join(_ROOT, "tests", "golden", "reference_data.tar.gz")) as tar:
        tar.extractall(data, filter="data")
lib = er.load_library(device=0); lib.setVerbosity(False)
if os.environ.get("CR_SANITIZE_SECTION", "") == "affine":
    # last session's paths only: SM-affine hand-out forced on for launches of any size (per-SM tickets, epoch-tagged block table,
    # ragged last blocks, grouped and ungrouped batches, consecutive launches), the second form of the frontier pass, programmatic
    # dependent launch of trace and reduction kernels, read-ahead buffers sized for the ramp's longest batch
    lib.loadGlTFscene(os.path.join(data, "data/natural-standin-sky.gltf").encode())
    er.gotoFirstCompoundEye(lib)
    N = lib.getCurrentEyeOmmatidialCount()
    lib.setCurrentEyeShaderName(b"single_dimension_fast"); er.setRenderSize(lib, N, 1)
    lib.crDebugSetEntryFrontier(1, 2, 0)
    for on, minb in ((1, 0), (1, 2), (0, 48)):
        lib.crDebugSetSmAffine(on, minb)
        for fused in (1, 0):
            lib.crSetRenderMode(fused, 0)
            for S in (32, 40, 96, 7):
                lib.setCurrentEyeSamplesPerOmmatidium(S)
                for k in range(3):
                    lib.setCameraPosition(0.1 * k, 0.2, 0.05 * k); lib.renderFrame(); lib.getFramePointer()
                for k in range(12):
                    lib.renderFrame(); lib.getFramePointer()
                lib.setCameraPosition(0.3, 0.1, 0.0); lib.renderFrame(); lib.getFramePointer()
                for n in (9, 6):
                    er.renderPoseBatch(lib, er.make_poses(np.random.default_rng(n).uniform(-1, 1, (n, 3))))
    lib.stop()
    print("sanitize tour (affine section) done")
    sys.exit(0)
for scene, cam in (("data/test-scene/test-scene.gltf", b"insect-cam-2"), ("data/natural-standin-sky.gltf", None)):
    lib.loadGlTFscene(os.path.join(data, scene).encode())
    if cam: assert lib.gotoCameraByName(cam)
    else: er.gotoFirstCompoundEye(lib)
    N = lib.getCurrentEyeOmmatidialCount()
    for frontier in (0, 1):
        lib.crDebugSetEntryFrontier(frontier, 2, 0)
        for S in (1, 4, 5, 33, 64, 130, 132):                       # partial warps, partial K1b tiles, TMA (S % 4 == 0) and cp.async forms
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            for shader, size in ((b"spherical_orientationwise", (64, 48)), (b"single_dimension_fast", (N, 1)), (b"raw_ommatidial_samples", (N, S)),
                                 (b"single_dimension", (77, 3)), (b"spherical_positionwise_ids", (40, 30))):
                lib.setCurrentEyeShaderName(shader); er.setRenderSize(lib, *size)
                lib.renderFrame(); lib.getFramePointer()
            lib.setCurrentEyeShaderName(b"single_dimension_fast"); er.setRenderSize(lib, N, 1)
            poses = er.make_poses(np.random.default_rng(S).uniform(-1, 1, (7, 3)))
            er.renderPoseBatch(lib, poses)
    # round 2: fused / fast modes, candidate lists, wavefront queue, static split, frame groups, read-ahead, zero-copy rows
    lib.setCurrentEyeShaderName(b"single_dimension_fast"); er.setRenderSize(lib, N, 1)
    for fused, fast, lists, wavefront, dynamic in ((1, 0, 2, 0, 1), (1, 1, 2, 1, 1), (0, 0, 2, 1, 0), (1, 0, 0, 0, 0)):
        lib.crSetRenderMode(fused, fast); lib.crDebugSetCandidateLists(lists)
        lib.crDebugSetWavefront(wavefront, 7, 0.05); lib.crDebugSetDynamicChunks(dynamic)
        for S in (32, 96, 40):
            lib.setCurrentEyeSamplesPerOmmatidium(S)
            for k in range(4):                                       # moving camera: zero-copy rows (fused), per-frame kernels
                lib.setCameraPosition(0.1 * k, 0.2, 0.05 * k); lib.renderFrame(); lib.getFramePointer()
            for k in range(9):                                       # standing camera: read-ahead launches, then a move drops the rest
                lib.renderFrame(); lib.getFramePointer()
            lib.setCameraPosition(0.3, 0.1, 0.0); lib.renderFrame(); lib.getFramePointer()
            for n in (9, 5, 12):                                     # batches at even and odd starts: frame groups on and off
                er.renderPoseBatch(lib, er.make_poses(np.random.default_rng(n).uniform(-1, 1, (n, 3))))
            st = np.zeros((N * S, 8), np.uint32); lib.crDebugCopyRngStates(st.ctypes.data)
    lib.crSetRenderMode(0, 0); lib.crDebugSetCandidateLists(0); lib.crDebugSetWavefront(0, 24, 0.35); lib.crDebugSetDynamicChunks(1)
    lib.crSetFirstFrame(3); lib.renderFrame(); lib.crSetFirstFrame(0)
    lib.crSetOmmatidialShard(5 * N, N); lib.renderFrame(); lib.crSetOmmatidialShard(0, 0)
    lib.crDebugSetRayDump(True); lib.renderFrame(); lib.crDebugSetRayDump(False)
    for i in range(lib.getCameraCount()):                            # ordinary cameras too
        lib.gotoCamera(i); er.setRenderSize(lib, 50, 40); lib.renderFrame(); lib.getFramePointer()
lib.stop()
print("sanitize tour done")
