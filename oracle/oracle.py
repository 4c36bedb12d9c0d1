"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU oracle (oracle/_build/libcr_oracle.so, built by oracle/Makefile from
cr_oracle.c) plus a small stateful renderer that mirrors the reference's frame loop
(libEyeRenderer3/libEyeRenderer.cpp:152-195) so tests can be written the way the reference's
python-examples drive the library.  The product never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcr_oracle.so")

PROJECTIONS = {
    "raw_ommatidial_samples": 0, "single_dimension": 1, "single_dimension_fast": 2,
    "spherical_positionwise": 3, "spherical_orientationwise": 4,
    "spherical_split_orientationwise": 5, "spherical_orientationwise_ids": 6,
    "spherical_positionwise_ids": 7,
}
MISS_SHADERS = {"default_background": 0, "simple_sky": 1}


class XwState(C.Structure):
    _fields_ = [("d", C.c_uint32), ("v", C.c_uint32 * 5), ("flag", C.c_int32), ("extra", C.c_float)]


class Pose(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("ax", C.c_float * 3), ("ay", C.c_float * 3), ("az", C.c_float * 3)]


class MeshInfo(C.Structure):
    _fields_ = [("color_type", C.c_int32), ("has_uv", C.c_int32), ("tex", C.c_int32), ("base_color", C.c_float * 4)]


class Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgba", C.c_void_p)]


class SceneC(C.Structure):
    _fields_ = [("T", C.c_int64), ("tris", C.c_void_p), ("tri_mesh", C.c_void_p), ("corner_uv", C.c_void_p),
                ("corner_col", C.c_void_p), ("meshes", C.c_void_p), ("n_meshes", C.c_int32),
                ("textures", C.c_void_p), ("n_textures", C.c_int32), ("miss_shader", C.c_int32),
                ("tex_frac_bits", C.c_int32)]


HIT_DTYPE = np.dtype([("prim", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])
STATE_DTYPE = np.dtype([("d", np.uint32), ("v", np.uint32, (5,)), ("flag", np.int32), ("extra", np.float32)])


def build(force=False):
    """Compile the oracle (gcc, a second or two)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("cr_oracle.c", "cr_math.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/libcr_oracle.so"])
    return _LIB_PATH


def build_native(out_dir):
    """The SAME source compiled for the machine this runs on (`gcc -O3 -march=native`, still without contraction or
    fast-math, so the bits do not change): the CPU-baseline timing build of bench.py.  It is compiled where
    it is timed -- never shipped -- because -march=native code built on one host may not run on another.  Returns the
    path, or None when there is no compiler."""
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libcr_oracle_native.so")
    cmd = ["gcc", "-O3", "-march=native", "-std=c11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
           "-fvisibility=hidden", "-shared", "-o", out, os.path.join(_HERE, "cr_oracle.c"), "-lm"]
    try:
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        return None
    return out


_lib = None


def lib():
    """The checker library.  CR_ORACLE_SO (set by bench.py's CPU-baseline subprocess) selects another build of it."""
    global _lib
    if _lib is None:
        path = os.environ.get("CR_ORACLE_SO") or _LIB_PATH
        if path == _LIB_PATH and not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(path)
        vp, i64, f32 = C.c_void_p, C.c_int64, C.c_float
        L.cro_xorwow_init.argtypes = [C.POINTER(XwState), C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]
        L.cro_xorwow_skipahead.argtypes = [C.POINTER(XwState), C.c_ulonglong]
        L.cro_xorwow_next.argtypes = [C.POINTER(XwState)]
        L.cro_xorwow_next.restype = C.c_uint32
        L.cro_xorwow_normal.argtypes = [C.POINTER(XwState)]
        L.cro_xorwow_normal.restype = f32
        L.cro_xorwow_uniform.argtypes = [C.POINTER(XwState)]
        L.cro_xorwow_uniform.restype = f32
        L.cro_draws_before_frame.argtypes = [C.c_ulonglong]
        L.cro_draws_before_frame.restype = C.c_ulonglong
        L.cro_position_streams.argtypes = [vp, i64, i64, C.c_ulonglong]
        L.cro_position_streams_shard.argtypes = [vp, i64, i64, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]
        L.cro_generate_rays.argtypes = [vp, i64, i64, C.POINTER(Pose), vp, C.c_int, vp, vp, vp]
        L.cro_trace_bruteforce.argtypes = [vp, i64, vp, vp, vp, i64, f32, vp]
        L.cro_bvh_build.argtypes = [vp, i64]
        L.cro_bvh_build.restype = vp
        L.cro_bvh_free.argtypes = [vp]
        L.cro_trace_bvh.argtypes = [vp, vp, vp, vp, i64, f32, vp, vp]
        L.cro_trace_device_bvh.argtypes = [vp, i64, vp, i64, vp, vp, vp, i64, f32, vp, vp]
        L.cro_shade.argtypes = [C.POINTER(SceneC), vp, vp, i64, vp]
        L.cro_accumulate.argtypes = [vp, i64, i64, vp, vp]
        L.cro_make_color.argtypes = [vp, vp]
        L.cro_projection_map.argtypes = [vp, i64, C.c_int, C.c_int, C.c_int, vp]
        L.cro_project.argtypes = [vp, i64, i64, C.c_int, C.c_int, C.c_int, vp, vp]
        L.cro_camera_rays.argtypes = [C.c_int, C.POINTER(Pose), vp, C.c_int, C.c_int, vp, vp, vp]
        L.cro_pose_rotate_around.argtypes = [C.POINTER(Pose), f32, vp]
        for name in ("sinf", "cosf", "logf", "expf", "asinf", "acosf"):
            fn = getattr(L, "cro_" + name)
            fn.argtypes = [f32]
            fn.restype = f32
        L.cro_powf.argtypes = [f32, f32]
        L.cro_powf.restype = f32
        L.cro_atan2f.argtypes = [f32, f32]
        L.cro_atan2f.restype = f32
        L.cro_num_threads.restype = C.c_int
        L.cro_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_pose(position=(0, 0, 0), x=(1, 0, 0), y=(0, 1, 0), z=(0, 0, 1)):
    p = Pose()
    p.pos[:] = [float(v) for v in position]
    p.ax[:] = [float(v) for v in x]
    p.ay[:] = [float(v) for v in y]
    p.az[:] = [float(v) for v in z]
    return p


def pose_from_camera(cam):
    return make_pose(cam.position, cam.x_axis, cam.y_axis, cam.z_axis)


def set_camera_pose(px, py, pz, rx, ry, rz):
    """libEyeRenderer.cpp:380-388: reset; rotate about world X, Y, Z; translate."""
    p = make_pose()
    for ang, axis in ((rx, (1, 0, 0)), (ry, (0, 1, 0)), (rz, (0, 0, 1))):
        a = np.array(axis, dtype=np.float32)
        lib().cro_pose_rotate_around(C.byref(p), np.float32(ang), _p(a))
    p.pos[:] = [np.float32(px), np.float32(py), np.float32(pz)]
    return p


def make_color(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((len(rgb), 4), dtype=np.uint8)
    for i in range(len(rgb)):
        lib().cro_make_color(_p(rgb[i:i + 1]), _p(out[i:i + 1]))
    return out


class SceneHandle:
    """Keeps the numpy arrays alive behind the C scene struct."""

    def __init__(self, scene, tex_frac_bits=8):
        self.scene = scene
        self.tris = np.ascontiguousarray(scene.tris, dtype=np.float32)
        self.tri_mesh = np.ascontiguousarray(scene.tri_mesh, dtype=np.int32)
        self.corner_uv = np.ascontiguousarray(scene.corner_uv, dtype=np.float32)
        self.corner_col = np.ascontiguousarray(scene.corner_col, dtype=np.float32)
        self.meshes = (MeshInfo * max(1, len(scene.mesh_info)))()
        for i, mi in enumerate(scene.mesh_info):
            self.meshes[i].color_type = mi["color_type"]
            self.meshes[i].has_uv = mi["has_uv"]
            self.meshes[i].tex = mi["tex"]
            self.meshes[i].base_color[:] = [float(v) for v in mi["base_color"]]
        self.tex_arrays = [np.ascontiguousarray(t, dtype=np.uint8) for t in scene.textures]
        self.textures = (Texture * max(1, len(self.tex_arrays)))()
        for i, t in enumerate(self.tex_arrays):
            self.textures[i].width = t.shape[1]
            self.textures[i].height = t.shape[0]
            self.textures[i].rgba = t.ctypes.data
        self.c = SceneC()
        self.c.T = len(self.tris)
        self.c.tris = self.tris.ctypes.data
        self.c.tri_mesh = self.tri_mesh.ctypes.data
        self.c.corner_uv = self.corner_uv.ctypes.data
        self.c.corner_col = self.corner_col.ctypes.data
        self.c.meshes = C.addressof(self.meshes)
        self.c.n_meshes = len(scene.mesh_info)
        self.c.textures = C.addressof(self.textures)
        self.c.n_textures = len(self.tex_arrays)
        self.c.miss_shader = MISS_SHADERS.get(scene.miss_shader, 0)
        self.c.tex_frac_bits = tex_frac_bits
        self._bvh = None

    @property
    def T(self):
        return len(self.tris)

    def bvh(self):
        if self._bvh is None:
            self._bvh = lib().cro_bvh_build(_p(self.tris), len(self.tris))
        return self._bvh

    def __del__(self):
        if getattr(self, "_bvh", None):
            lib().cro_bvh_free(self._bvh)
            self._bvh = None


def generate_rays(omm, S, pose, states, configured):
    omm = np.ascontiguousarray(omm, dtype=np.float32)
    N = len(omm)
    R = N * S
    origins = np.empty((R, 3), dtype=np.float32)
    dirs = np.empty((R, 3), dtype=np.float32)
    tmins = np.empty(R, dtype=np.float32)
    lib().cro_generate_rays(_p(omm), N, S, C.byref(pose), _p(states), int(configured), _p(origins), _p(dirs), _p(tmins))
    return origins, dirs, tmins


def trace(sh: SceneHandle, origins, dirs, tmins, tmax=1e16, method="auto", counters=None):
    R = len(origins)
    hits = np.empty(R, dtype=HIT_DTYPE)
    if method == "auto":
        method = "brute" if sh.T * R <= 2_000_000_000 and sh.T <= 4096 else "bvh"
    if method == "brute":
        lib().cro_trace_bruteforce(_p(sh.tris), sh.T, _p(origins), _p(dirs), _p(tmins), R, np.float32(tmax), _p(hits))
    else:
        cnt = np.zeros(2, dtype=np.int64)
        lib().cro_trace_bvh(sh.bvh(), _p(origins), _p(dirs), _p(tmins), R, np.float32(tmax), _p(hits), _p(cnt))
        if counters is not None:
            counters[:] = cnt
    return hits


def trace_device_bvh(nodes, dtris, origins, dirs, tmins, tmax=1e16):
    """Instrumented traversal of the product's BVH (arrays downloaded through crDebug*)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1, 16)
    dtris = np.ascontiguousarray(dtris, dtype=np.float32).reshape(-1, 12)
    R = len(origins)
    hits = np.empty(R, dtype=HIT_DTYPE)
    cnt = np.zeros(2, dtype=np.int64)
    lib().cro_trace_device_bvh(_p(nodes), len(nodes), _p(dtris), len(dtris), _p(origins), _p(dirs), _p(tmins), R,
                               np.float32(tmax), _p(hits), _p(cnt))
    return hits, cnt


def shade(sh: SceneHandle, hits, dirs):
    R = len(hits)
    rgb = np.empty((R, 3), dtype=np.float32)
    lib().cro_shade(C.byref(sh.c), _p(hits), _p(dirs), R, _p(rgb))
    return rgb


def projection_map(omm, mode, W, H):
    omm = np.ascontiguousarray(omm, dtype=np.float32)
    out = np.empty((H, W), dtype=np.uint32)
    lib().cro_projection_map(_p(omm), len(omm), PROJECTIONS[mode], W, H, _p(out))
    return out


def fused_sum(compound, N, S):
    """Per-ommatidium RGB in the FIXED addition order of the product's fused reduction (crSetRenderMode(1, .);
    k_traceCompound<FUSED> + k_sumPartials in csrc/cr_kernels.cu), restated on the checker's per-sample colours
    (`compound`: [S*N][3] in stream-id order N*s+o, already divided by S, shaders.cu:730):
      1. samples 32b .. 32b+31 of an ommatidium are combined by a butterfly -- level d in (16, 8, 4, 2, 1) adds
         element i and element i+d of the surviving 2d elements;
      2. "lane" l adds the block sums l, l+32, l+64, ... in ascending order, starting from 0;
      3. the same butterfly combines the 32 lanes.
    The reference's own order is the plain sequential sum (shaders.cu:341-347) -- cro_accumulate; the two differ by
    fp32 rounding only."""
    assert S % 32 == 0
    c = np.ascontiguousarray(compound, dtype=np.float32).reshape(S, N, 3).transpose(1, 0, 2).reshape(N, S // 32, 32, 3)

    def butterfly(a):                                   # [..., 32, 3] -> [..., 3]
        for d in (16, 8, 4, 2, 1):
            a = a[..., :d, :] + a[..., d:2 * d, :]
        return a[..., 0, :]

    blocks = butterfly(c)                               # [N][S/32][3]
    lanes = np.zeros((N, 32, 3), dtype=np.float32)
    for k in range(0, S // 32, 32):
        part = blocks[:, k:k + 32]
        lanes[:, :part.shape[1]] = lanes[:, :part.shape[1]] + part
    return butterfly(lanes).astype(np.float32)


class CompoundEyeOracle:
    """Stateful mirror of one CompoundEye camera + the frame loop.

    State rules follow cameras/CompoundEye.cpp:30-62,98-183: RNG streams are reset
    ("randomsConfigured = false") when S or the ommatidial COUNT changes, and persist otherwise.
    """

    def __init__(self, sh: SceneHandle, ommatidia, pose, projection="spherical_orientationwise", samples=1):
        self.sh = sh
        self.omm = np.ascontiguousarray(ommatidia, dtype=np.float32).reshape(-1, 8)
        self.pose = pose
        self.projection = projection
        self.S = 1
        self.configured = False
        self.states = np.zeros(len(self.omm) * self.S, dtype=STATE_DTYPE)
        self.W, self.H = 400, 400                                  # libEyeRenderer.cpp:85-86
        self.frame = np.zeros((self.H, self.W, 4), dtype=np.uint8)
        self.last = {}
        if samples != 1:
            self.set_samples(samples)

    def set_samples(self, s):
        self.S = max(1, int(s))
        self.states = np.zeros(len(self.omm) * self.S, dtype=STATE_DTYPE)
        self.configured = False

    def set_ommatidia(self, omm):
        omm = np.ascontiguousarray(omm, dtype=np.float32).reshape(-1, 8)
        if len(omm) != len(self.omm):
            self.states = np.zeros(len(omm) * self.S, dtype=STATE_DTYPE)
            self.configured = False
        self.omm = omm

    def set_first_frame(self, k):
        """Position every stream as if k frames had been rendered (pose sharding / restart)."""
        self.states = np.zeros(len(self.omm) * self.S, dtype=STATE_DTYPE)
        lib().cro_position_streams(_p(self.states), len(self.omm), self.S, int(k))
        self.configured = True

    def set_shard(self, n_global, o_first, first_frame=0):
        """The table holds rows [o_first, o_first+N) of an eye of n_global ommatidia: streams keep their
        global ids (== crSetOmmatidialShard in the product)."""
        self.states = np.zeros(len(self.omm) * self.S, dtype=STATE_DTYPE)
        lib().cro_position_streams_shard(_p(self.states), len(self.omm), self.S, C.c_ulonglong(int(first_frame)),
                                         C.c_ulonglong(int(n_global)), C.c_ulonglong(int(o_first)))
        self.configured = True

    def set_render_size(self, w, h):
        self.W, self.H = int(w), int(h)
        self.frame = np.zeros((self.H, self.W, 4), dtype=np.uint8)

    def render_frame(self, method="auto", project=True):
        N, S = len(self.omm), self.S
        origins, dirs, tmins = generate_rays(self.omm, S, self.pose, self.states, self.configured)
        self.configured = True
        cnt = np.zeros(2, dtype=np.int64)
        hits = trace(self.sh, origins, dirs, tmins, method=method, counters=cnt)
        rgb = shade(self.sh, hits, dirs)
        compound = np.empty((N * S, 3), dtype=np.float32)
        summed = np.empty((N, 3), dtype=np.float32)
        lib().cro_accumulate(_p(rgb), N, S, _p(compound), _p(summed))
        self.last = dict(origins=origins, dirs=dirs, tmins=tmins, hits=hits, rgb=rgb, compound=compound,
                         summed=summed, counters=cnt)
        if project:
            lib().cro_project(_p(self.omm), N, S, PROJECTIONS[self.projection], self.W, self.H, _p(compound),
                              _p(self.frame))
        return self.frame
