"""Python host side of the B200-native libEyeRenderer3.

Mirrors the reference's ctypes helper module (python-examples/eyeRendererHelperFunctions.py:
configureFunctions :40-71, setRenderSize :80-83, setOmmatidiaFrom* :85-101, gotoFirst* :103-129,
readEyeFile/saveEyeFile :131-151, decodeProjectionMapID :153-160, getProjectionImageUsingMap
:162-169, getIcoOmmatidia :171-194) with the same function names and argument meaning, so code
written against the reference helper runs against this module unchanged.  The reference helper
itself also works unmodified with this library; this module exists because the reference tree is
not shipped with the product, and to expose the additive cr* entry points.

The library is the product: there is no Python or CPU fallback behind these calls.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
from numpy.ctypeslib import ndpointer

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libEyeRenderer3.so")


class c_float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def toNumpy(self):
        return np.asarray([self.x, self.y, self.z])


class c_ommatidiumPacket(C.Structure):
    _fields_ = [("posX", C.c_float), ("posY", C.c_float), ("posZ", C.c_float),
                ("dirX", C.c_float), ("dirY", C.c_float), ("dirZ", C.c_float),
                ("acceptanceAngle", C.c_float), ("focalpointOffset", C.c_float)]


class Ommatidium:
    def __init__(self, position, direction, acceptanceAngle, focalpointOffset):
        self.position = position
        self.direction = direction
        self.acceptanceAngle = acceptanceAngle
        self.focalpointOffset = focalpointOffset

    def getSolidAngle(self):
        """Solid angle (steradians) of the acceptance cone."""
        return 2.0 * math.pi * (1.0 - math.cos(self.acceptanceAngle / 2.0))

    def copy(self):
        return Ommatidium(self.position.copy(), self.direction.copy(), self.acceptanceAngle, self.focalpointOffset)


def load_library(path: str | None = None, device: int | None = None):
    """dlopen the renderer, declare the signatures and optionally pin the CUDA device."""
    path = path or os.environ.get("CR_LIB_PATH") or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: build it with `make -C {_HERE}` (nvcc, sm_100a)")
    lib = C.CDLL(path)
    configureFunctions(lib)
    if device is not None:
        lib.crSetDevice(int(device))
    return lib


def configureFunctions(eyeRenderer):
    """argtypes / restypes of every entry point except setOmmatidia (sized per call)."""
    f, vp = C.c_float, C.c_void_p
    r = eyeRenderer
    r.setVerbosity.argtypes = [C.c_bool]
    r.loadGlTFscene.argtypes = [C.c_char_p]
    r.setRenderSize.argtypes = [C.c_int, C.c_int]
    r.renderFrame.restype = C.c_double
    r.saveFrameAs.argtypes = [C.c_char_p]
    r.getFramePointer.restype = vp
    r.getCameraCount.restype = C.c_size_t
    r.getCurrentCameraIndex.restype = C.c_size_t
    r.getCurrentCameraName.restype = C.c_char_p
    r.gotoCamera.argtypes = [C.c_int]
    r.gotoCameraByName.argtypes = [C.c_char_p]
    r.gotoCameraByName.restype = C.c_bool
    r.setCameraPosition.argtypes = [f] * 3
    r.getCameraPosition.argtypes = [C.POINTER(f)] * 3
    r.setCameraLocalSpace.argtypes = [f] * 9
    r.rotateCameraAround.argtypes = [f] * 4
    r.rotateCameraLocallyAround.argtypes = [f] * 4
    r.translateCamera.argtypes = [f] * 3
    r.translateCameraLocally.argtypes = [f] * 3
    r.setCameraPose.argtypes = [f] * 6
    r.isCompoundEyeActive.restype = C.c_bool
    r.setCurrentEyeSamplesPerOmmatidium.argtypes = [C.c_int]
    r.getCurrentEyeSamplesPerOmmatidium.restype = C.c_int
    r.changeCurrentEyeSamplesPerOmmatidiumBy.argtypes = [C.c_int]
    r.getCurrentEyeOmmatidialCount.restype = C.c_size_t
    r.getCurrentEyeDataPath.restype = C.c_char_p
    r.setCurrentEyeShaderName.argtypes = [C.c_char_p]
    r.isInsideHitGeometry.argtypes = [f, f, f, C.c_char_p]
    r.isInsideHitGeometry.restype = C.c_bool
    r.getGeometryMaxBounds.argtypes = [C.c_char_p]
    r.getGeometryMaxBounds.restype = c_float3
    r.getGeometryMinBounds.argtypes = [C.c_char_p]
    r.getGeometryMinBounds.restype = c_float3
    # additions
    r.crSetDevice.argtypes = [C.c_int]
    r.crGetOmmatidialData.argtypes = [vp]
    r.crRenderPoseBatch.argtypes = [vp, C.c_size_t, vp, vp]
    r.crRenderPoseBatch.restype = C.c_double
    r.crSetFirstFrame.argtypes = [C.c_uint64]
    r.crSetOmmatidialShard.argtypes = [C.c_uint64, C.c_uint64]
    r.crCommGetUniqueId.argtypes = [vp]
    r.crCommInit.argtypes = [vp, C.c_int, C.c_int]
    r.crAllGatherRows.argtypes = [vp, vp, C.c_size_t]
    r.crRenderPoseBatchSharded.argtypes = [vp, C.c_size_t, vp, vp, C.c_size_t, C.c_uint64]
    r.crRenderPoseBatchSharded.restype = C.c_double
    r.crSetRenderMode.argtypes = [C.c_int, C.c_int]
    r.crGetRenderMode.restype = C.c_int
    r.crDebugSetCandidateLists.argtypes = [C.c_int]
    r.crDebugSetWavefront.argtypes = [C.c_int, C.c_int, C.c_double]
    r.crDebugSetNodeLanes.argtypes = [C.c_int]
    r.crDebugSetFrameGroups.argtypes = [C.c_int]
    r.crDebugSetReadAhead.argtypes = [C.c_int, C.c_double]
    r.crDebugSetFrameProfile.argtypes = [C.c_int]
    r.crDebugFrameBreakdown.argtypes = [vp]
    r.crDebugSetDynamicChunks.argtypes = [C.c_int]
    r.crDebugSetSmAffine.argtypes = [C.c_int, C.c_int]
    r.crDebugSetZeroCopy.argtypes = [C.c_int]
    r.crDebugLastQueuedRays.restype = C.c_ulonglong
    r.crDebugCopyCandidateLists.argtypes = [vp, C.c_size_t]
    r.crDebugCopyCandidateLists.restype = C.c_size_t
    r.crGetLastBatchFrames.restype = C.c_int
    r.crGetLastTraceMs.restype = C.c_double
    r.crGetLaunchCount.restype = C.c_ulonglong
    r.crGetBvhBuildMs.restype = C.c_double
    r.crDebugDecodeImageFile.argtypes = [C.c_char_p, vp, vp]
    r.crDebugDecodeImageFile.restype = C.c_bool
    r.crDebugCopyDecodedImage.argtypes = [vp]
    r.crDebugGetTextureSize.argtypes = [C.c_int, vp, vp]
    r.crDebugCopyTexture.argtypes = [C.c_int, vp]
    for name in ("crDebugGetTriangleCount", "crDebugGetVertexCount", "crDebugGetMeshCount", "crDebugGetBvhNodeCount",
                 "crDebugGetTextureCount"):
        getattr(r, name).restype = C.c_size_t
    r.crDebugCopyTriangles.argtypes = [vp]
    r.crDebugCopyTriangleMesh.argtypes = [vp]
    r.crDebugCopyMeshInfo.argtypes = [vp, vp]
    r.crDebugCopyCornerAttributes.argtypes = [vp, vp]
    r.crDebugCopyCameraPose.argtypes = [vp]
    r.crDebugCopyCameraScale.argtypes = [vp]
    r.crDebugCopyOmmatidia.argtypes = [vp]
    r.crDebugCopyBvh.argtypes = [vp, vp]
    r.crDebugSetRayDump.argtypes = [C.c_bool]
    r.crDebugSetEntryFrontier.argtypes = [C.c_int, C.c_int, C.c_longlong]
    r.crDebugXorwowInit.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, vp]
    r.crDebugCopyLastRayCounts.argtypes = [vp]
    r.crDebugCopyLastRayCounts.restype = C.c_size_t
    r.crDebugCopyLastRays.argtypes = [vp, vp, vp]
    r.crDebugCopyLastRays.restype = C.c_size_t
    r.crDebugCopyRngStates.argtypes = [vp]
    r.crDebugTraceRays.argtypes = [vp, vp, vp, C.c_int, vp]
    r.crDebugCopyProjectionMap.argtypes = [vp]
    r.crDebugSampleTexture.argtypes = [C.c_int, vp, C.c_int, vp]
    r.crDebugEvalMath.argtypes = [C.c_int, vp, vp, vp, C.c_int]


def setCameraLocalSpace(eyeRenderer, npMatrix):
    """Columns of a 3x3 matrix are the camera's x, y and z axes."""
    m = np.asarray(npMatrix, dtype=np.float32)
    eyeRenderer.setCameraLocalSpace(*[float(v) for v in m[:, 0]], *[float(v) for v in m[:, 1]], *[float(v) for v in m[:, 2]])


def setRenderSize(eyeRenderer, width, height):
    """Resizes the frame and retypes getFramePointer as uint8[height][width][4] (row 0 = bottom)."""
    eyeRenderer.setRenderSize(int(width), int(height))
    eyeRenderer.getFramePointer.restype = ndpointer(dtype=C.c_ubyte, shape=(int(height), int(width), 4))


def setOmmatidiaFromPacketList(eyeRenderer, packetList):
    count = len(packetList)
    array_type = c_ommatidiumPacket * count
    eyeRenderer.setOmmatidia.argtypes = [array_type, C.c_size_t]
    eyeRenderer.setOmmatidia(array_type(*packetList), C.c_size_t(count))


def setOmmatidiaFromOmmatidiumList(eyeRenderer, ommList):
    packets = [c_ommatidiumPacket(*[float(n) for n in o.position], *[float(n) for n in o.direction],
                                  float(o.acceptanceAngle), float(o.focalpointOffset)) for o in ommList]
    setOmmatidiaFromPacketList(eyeRenderer, packets)


def setOmmatidiaFromArray(eyeRenderer, omm):
    """Addition: float32[N][8] rows (position, direction, acceptance angle, focal offset)."""
    omm = np.ascontiguousarray(omm, dtype=np.float32).reshape(-1, 8)
    eyeRenderer.setOmmatidia.argtypes = [C.c_void_p, C.c_size_t]
    eyeRenderer.setOmmatidia(omm.ctypes.data_as(C.c_void_p), C.c_size_t(len(omm)))


def _gotoFirst(eyeRenderer, wantCompound, what):
    for i in range(eyeRenderer.getCameraCount()):
        eyeRenderer.gotoCamera(int(i))
        if bool(eyeRenderer.isCompoundEyeActive()) == wantCompound:
            return
    raise Exception(f"Error: Could not find {what} in provided GlTF scene.")


def gotoFirstCompoundEye(eyeRenderer):
    _gotoFirst(eyeRenderer, True, "compound eye")


def gotoFirstRegularCamera(eyeRenderer):
    _gotoFirst(eyeRenderer, False, "regular camera")


def _getEyeFeatures(line):
    v = [float(n) for n in line.split(" ") if n.strip() != ""]
    return Ommatidium(np.asarray(v[0:3]), np.asarray(v[3:6]), v[6], v[7])


def readEyeFile(path):
    with open(path) as f:
        return [_getEyeFeatures(line) for line in f if line.strip() != ""]


def saveEyeFile(path, omms):
    with open(path, "w") as f:
        for o in omms:
            vals = [*o.position[:3], *o.direction[:3], o.acceptanceAngle, o.focalpointOffset]
            f.write(" ".join("{:0.10f}".format(float(v)) for v in vals) + "\n")


def decodeProjectionMapID(RGBAquadlet):
    """Index written by the "_ids" projections: big-endian bytes in R, G, B, A."""
    return (int(RGBAquadlet[0]) << 24) | (int(RGBAquadlet[1]) << 16) | (int(RGBAquadlet[2]) << 8) | int(RGBAquadlet[3])


def getProjectionImageUsingMap(vector, idMap, pjWidth, pjHeight):
    ids = (idMap[..., 0].astype(np.uint32) << 24) | (idMap[..., 1].astype(np.uint32) << 16) | \
          (idMap[..., 2].astype(np.uint32) << 8) | idMap[..., 3].astype(np.uint32)
    return np.asarray(vector)[ids[:pjHeight, :pjWidth]].astype(np.uint8)


def getIcoOmmatidia():
    """Twelve equidistant ommatidia on the icosahedron's vertices, 1 steradian each."""
    lat = math.atan(0.5)
    pts = [[0.0, 1.0, 0.0]]
    for ring, sign in ((0.0, 1.0), (0.2 * math.pi, -1.0)):
        for i in range(5):
            a = 0.4 * math.pi * i + ring
            pts.append([math.cos(a) * math.cos(lat), sign * math.sin(lat), math.sin(a) * math.cos(lat)])
    pts.append([0.0, -1.0, 0.0])
    acceptance = math.acos(-(1 / (2 * math.pi) - 1)) * 2
    return [Ommatidium(np.zeros(3), np.asarray(p), acceptance, 0.0) for p in pts]


# ----------------------------------------------------------------------------- additions
def getFrame(eyeRenderer, width, height):
    """Copy of the current frame as uint8[height][width][4]."""
    setattr(eyeRenderer.getFramePointer, "restype", ndpointer(dtype=C.c_ubyte, shape=(int(height), int(width), 4)))
    return np.copy(eyeRenderer.getFramePointer())


def getOmmatidialData(eyeRenderer):
    """float32[N][3]: per-ommatidium linear RGB of the last compound frame."""
    n = eyeRenderer.getCurrentEyeOmmatidialCount()
    out = np.zeros((n, 3), dtype=np.float32)
    eyeRenderer.crGetOmmatidialData(out.ctypes.data_as(C.c_void_p))
    return out


def setRenderMode(eyeRenderer, fused=None, fast_math=None):
    """Addition: crSetRenderMode.  fused=True: in-kernel reduction (fixed order, RGB equal to fp32 rounding);
    fast_math=True: hardware sin/cos/log/pow as in the reference's --use_fast_math build.  None keeps a switch."""
    eyeRenderer.crSetRenderMode(-1 if fused is None else int(bool(fused)), -1 if fast_math is None else int(bool(fast_math)))


def make_poses(positions, x=(1, 0, 0), y=(0, 1, 0), z=(0, 0, 1)):
    """float32[P][12] pose rows from positions and a fixed orientation."""
    positions = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
    axes = np.concatenate([np.asarray(x, np.float32), np.asarray(y, np.float32), np.asarray(z, np.float32)])
    return np.ascontiguousarray(np.concatenate([positions, np.tile(axes, (len(positions), 1))], axis=1), dtype=np.float32)


def renderPoseBatch(eyeRenderer, poses, out_device_ptr=None):
    """Render one frame per pose row; returns (uint8[P][N][4] or None, milliseconds)."""
    poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 12)
    n = eyeRenderer.getCurrentEyeOmmatidialCount()
    if out_device_ptr is not None:
        ms = eyeRenderer.crRenderPoseBatch(poses.ctypes.data_as(C.c_void_p), len(poses), None, C.c_void_p(out_device_ptr))
        return None, ms
    out = np.zeros((len(poses), n, 4), dtype=np.uint8)
    ms = eyeRenderer.crRenderPoseBatch(poses.ctypes.data_as(C.c_void_p), len(poses), out.ctypes.data_as(C.c_void_p), None)
    return out, ms
