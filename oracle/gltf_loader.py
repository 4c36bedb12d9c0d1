"""oracle/gltf_loader.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

numpy restatement of the reference's scene loading for the compound-eye render path:
``loadScene`` / ``processGLTFNode`` (libEyeRenderer3/MulticamScene.cpp:165-526, 531-736), the
``.eye`` format (data/eyes/eye-specification.txt, MulticamScene.cpp:290-299) and the sutil math
conventions it relies on (sutil/Matrix.h:344-359,472-490,677-700; sutil/Quaternion.h:239-269;
sutil/Aabb.h:351-371).  All arithmetic is IEEE binary32 with one rounding per written operation,
so results are bit-comparable with the C++ product loader.

Difference from the reference by design (documented in DESIGN.md): the reference keeps vertices in
object space and lets OptiX apply the instance transform to the ray (MulticamScene.cpp:1356-1371);
here -- as in the product -- static instances are flattened into world space once at load:
``world = ((m0*x + m1*y) + m2*z) + m3`` per row of the node transform.
"""
from __future__ import annotations

import base64
import io
import json
import os
from dataclasses import dataclass, field

import numpy as np

F = np.float32

_COMPONENT_DTYPE = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16,
                    5125: np.uint32, 5126: np.float32}
_TYPE_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


# ----------------------------------------------------------------------------- float32 matrix math
def mat_identity():
    return np.eye(4, dtype=F)


def mat_mul(a, b):
    """sutil/Matrix.h:344-359: sum = 0; sum += a[i][k]*b[k][j] for k = 0..3 (float32)."""
    out = np.zeros((4, 4), dtype=F)
    for i in range(4):
        for j in range(4):
            s = F(0.0)
            for k in range(4):
                s = F(s + F(a[i, k] * b[k, j]))
            out[i, j] = s
    return out


def mat_vec4(m, v):
    """sutil/Matrix.h:472-490: m0*x + m1*y + m2*z + m3*w, left to right."""
    out = np.zeros(4, dtype=F)
    for i in range(4):
        out[i] = F(F(F(F(m[i, 0] * v[0]) + F(m[i, 1] * v[1])) + F(m[i, 2] * v[2])) + F(m[i, 3] * v[3]))
    return out


def quat_rotation_matrix(w, x, y, z):
    """sutil/Quaternion.h:239-269 (ctor order w,x,y,z: MulticamScene.cpp:183-188)."""
    qw, qx, qy, qz = F(w), F(x), F(y), F(z)
    two = F(2.0)
    one = F(1.0)
    m = np.zeros((4, 4), dtype=F)
    m[0, 0] = F(F(one - F(F(two * qy) * qy)) - F(F(two * qz) * qz))
    m[0, 1] = F(F(F(two * qx) * qy) - F(F(two * qz) * qw))
    m[0, 2] = F(F(F(two * qx) * qz) + F(F(two * qy) * qw))
    m[1, 0] = F(F(F(two * qx) * qy) + F(F(two * qz) * qw))
    m[1, 1] = F(F(one - F(F(two * qx) * qx)) - F(F(two * qz) * qz))
    m[1, 2] = F(F(F(two * qy) * qz) - F(F(two * qx) * qw))
    m[2, 0] = F(F(F(two * qx) * qz) - F(F(two * qy) * qw))
    m[2, 1] = F(F(F(two * qy) * qz) + F(F(two * qx) * qw))
    m[2, 2] = F(F(one - F(F(two * qx) * qx)) - F(F(two * qy) * qy))
    m[3, 3] = one
    return m


def node_local_matrices(node):
    """MulticamScene.cpp:173-205.  Returns (matrix, translation, rotation, scale)."""
    t = mat_identity()
    if node.get("translation"):
        tr = node["translation"]
        t[0, 3], t[1, 3], t[2, 3] = F(tr[0]), F(tr[1]), F(tr[2])
    r = mat_identity()
    if node.get("rotation"):
        q = node["rotation"]
        r = quat_rotation_matrix(q[3], q[0], q[1], q[2])
    s = mat_identity()
    if node.get("scale"):
        sc = node["scale"]
        s[0, 0], s[1, 1], s[2, 2] = F(sc[0]), F(sc[1]), F(sc[2])
    m = mat_identity()
    if node.get("matrix"):
        m = np.array([F(x) for x in node["matrix"]], dtype=F).reshape(4, 4).T.copy()
    return m, t, r, s


def node_xform(parent, node):
    m, t, r, s = node_local_matrices(node)
    return mat_mul(mat_mul(mat_mul(mat_mul(parent, m), t), r), s)


def transform_points(m, pts):
    """world = ((m0*x + m1*y) + m2*z) + m3 per row, float32, vectorised."""
    pts = np.asarray(pts, dtype=F)
    out = np.empty_like(pts)
    for i in range(3):
        acc = (m[i, 0] * pts[:, 0]).astype(F)
        acc = (acc + (m[i, 1] * pts[:, 1]).astype(F)).astype(F)
        acc = (acc + (m[i, 2] * pts[:, 2]).astype(F)).astype(F)
        acc = (acc + m[i, 3]).astype(F)
        out[:, i] = acc
    return out


def aabb_transform(m, bmin, bmax):
    """sutil/Aabb.h:351-371: include the 8 transformed corners."""
    corners = []
    for cx in (bmin[0], bmax[0]):
        for cy in (bmin[1], bmax[1]):
            for cz in (bmin[2], bmax[2]):
                corners.append(mat_vec4(m, np.array([cx, cy, cz, 1.0], dtype=F))[:3])
    c = np.array(corners, dtype=F)
    return c.min(axis=0), c.max(axis=0)


# ----------------------------------------------------------------------------- data classes
@dataclass
class Camera:
    name: str
    kind: str                       # "perspective" | "panoramic" | "orthographic" | "compound"
    position: np.ndarray            # float32[3]
    x_axis: np.ndarray              # local space (right, up, forward): MulticamScene.cpp:215-217,
    y_axis: np.ndarray              #   setLocalSpace(rightAxis, upAxis, forwardAxis)
    z_axis: np.ndarray
    scale: np.ndarray = field(default_factory=lambda: np.zeros(3, dtype=F))
    projection: str = ""
    eye_path: str = ""
    ommatidia: np.ndarray | None = None     # float32[N][8]


@dataclass
class MeshGroup:
    name: str
    first_tri: int
    n_tris: int
    color_type: int                 # -1 or glTF componentType
    has_uv: bool
    material: int
    world_min: np.ndarray
    world_max: np.ndarray


@dataclass
class Scene:
    cameras: list
    meshes: list
    hitboxes: list
    tris: np.ndarray                # float32[T][9]  v0, e1, e2 (world)
    verts: np.ndarray               # float32[T][3][3] world-space corners
    tri_mesh: np.ndarray            # int32[T]
    corner_uv: np.ndarray | None    # float32[T][3][2]
    corner_col: np.ndarray | None   # float32[T][3][4]
    mesh_info: list                 # dicts: color_type, has_uv, tex, base_color
    textures: list                  # uint8[h][w][4] per glTF *texture* index
    miss_shader: str


def parse_eye_lines(lines):
    """MulticamScene.cpp:290-299: split on single spaces (empty pieces dropped), 8 std::stof."""
    rows = []
    for line in lines:
        parts = [p for p in line.rstrip("\n").rstrip("\r").split(" ") if p != ""]
        if not parts:
            raise ValueError("blank line in .eye file (std::stof would throw in the reference)")
        rows.append([F(float(p)) for p in parts[:8]])
    return np.array(rows, dtype=F).reshape(-1, 8)


def read_eye_file(path):
    with open(path) as f:
        return parse_eye_lines(f.readlines())


def _truthy(extras, key):
    """MulticamScene.cpp:131-146."""
    if not isinstance(extras, dict) or key not in extras:
        return False
    v = extras[key]
    if isinstance(v, bool):
        return v
    if isinstance(v, str):
        return v.lower() == "true"
    return False


def _load_uri(uri, base_dir):
    if uri.startswith("data:"):
        return base64.b64decode(uri.split(",", 1)[1])
    with open(os.path.join(base_dir, uri), "rb") as f:
        return f.read()


class _Model:
    def __init__(self, path):
        self.path = path
        # MulticamScene.cpp:548-551: directory part including the trailing slash
        slash = max(path.rfind("/"), path.rfind("\\")) + 1
        self.dir = path[:slash]
        with open(path) as f:
            self.j = json.load(f)
        self.buffers = [_load_uri(b["uri"], self.dir) for b in self.j.get("buffers", [])]

    def accessor(self, idx):
        acc = self.j["accessors"][idx]
        bv = self.j["bufferViews"][acc["bufferView"]]
        dt = np.dtype(_COMPONENT_DTYPE[acc["componentType"]])
        nc = _TYPE_NCOMP[acc["type"]]
        off = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * nc
        buf = self.buffers[bv["buffer"]]
        count = acc["count"]
        arr = np.ndarray((count, nc), dtype=dt, buffer=buf, offset=off, strides=(stride, dt.itemsize))
        return np.array(arr), acc

    def image_rgba(self, idx):
        from PIL import Image
        img = self.j["images"][idx]
        if "bufferView" in img:
            bv = self.j["bufferViews"][img["bufferView"]]
            off = bv.get("byteOffset", 0)
            data = self.buffers[bv["buffer"]][off:off + bv["byteLength"]]
        else:
            data = _load_uri(img["uri"], self.dir)
        im = Image.open(io.BytesIO(data))
        return np.array(im.convert("RGBA"), dtype=np.uint8)       # stb req_comp = 4


def load_scene(path) -> Scene:
    md = _Model(path)
    j = md.j

    miss = "default_background"                                    # MulticamScene.h:187
    for sc in j.get("scenes", []):                                 # MulticamScene.cpp:555-565
        bg = sc.get("extras", {}).get("background-shader", "") if isinstance(sc.get("extras"), dict) else ""
        if isinstance(bg, str) and bg != "":
            miss = bg

    images = [md.image_rgba(i) for i in range(len(j.get("images", [])))]
    textures = [images[t["source"]] for t in j.get("textures", [])]   # sampler: always wrap+linear (:801-834)

    materials = []
    for m in j.get("materials", []):                               # :629-718
        pbr = m.get("pbrMetallicRoughness", {})
        bc = pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0])
        tex = pbr.get("baseColorTexture", {}).get("index", -1)
        materials.append({"base_color": np.array(bc, dtype=F), "tex": tex})

    cameras, meshes, hitboxes, mesh_info = [], [], [], []
    tri_chunks, vert_chunks, mesh_ids, uv_chunks, col_chunks = [], [], [], [], []
    state = {"T": 0}

    def process(node, parent):
        xf = node_xform(parent, node)
        if "camera" in node:                                       # :207-328
            cam = j["cameras"][node["camera"]]
            up = mat_vec4(xf, np.array([0, 1, 0, 0], dtype=F))[:3]
            fwd = mat_vec4(xf, np.array([0, 0, -1, 0], dtype=F))[:3]
            right = mat_vec4(xf, np.array([1, 0, 0, 0], dtype=F))[:3]
            eye = mat_vec4(xf, np.array([0, 0, 0, 1], dtype=F))[:3]
            name = cam.get("name", "")
            extras = cam.get("extras", {})
            if cam.get("type") == "orthographic":
                o = cam["orthographic"]
                cameras.append(Camera(name, "orthographic", eye, right, up, fwd,
                                      scale=np.array([o["xmag"], o["ymag"], 0], dtype=F)))
                return
            if _truthy(extras, "panoramic"):
                cameras.append(Camera(name, "panoramic", eye, right, up, fwd))
                return
            if _truthy(extras, "compound-eye"):
                eye_path = extras.get("compound-structure", "")
                proj = extras.get("compound-projection", "")
                if not isinstance(eye_path, str) or eye_path == "" or not isinstance(proj, str) or proj == "":
                    return
                used = None
                if os.path.isfile(eye_path):
                    used = eye_path
                elif os.path.isfile(md.dir + eye_path):
                    used = md.dir + eye_path
                if used is None:
                    return                                          # camera silently skipped (:276-279)
                omm = read_eye_file(used)
                if len(omm) == 0:
                    return
                cameras.append(Camera(name, "compound", eye, right, up, fwd, projection=proj,
                                      eye_path=used, ommatidia=omm))
                return
            yfov = F(F(F(cam.get("perspective", {}).get("yfov", 0.0)) * F(180.0)) / F(np.pi))
            # PerspectiveCamera ctor scale (10,10,1) then setYFOV (cameras/PerspectiveCamera.cpp:3-23)
            rad = F(F(yfov / F(180)) * F(3.14159265358979323846))
            sy = F(F(np.tan(np.float64(F(rad / F(2.0))))) * F(1.0))
            cameras.append(Camera(name, "perspective", eye, right, up, fwd,
                                  scale=np.array([F(sy * F(1.0)), sy, 1.0], dtype=F)))
            return
        if "mesh" in node:
            gm = j["meshes"][node["mesh"]]
            is_hitbox = _truthy(gm.get("extras", {}), "hitbox")
            for prim in gm.get("primitives", []):
                if prim.get("mode", 4) != 4:
                    continue
                pos, pacc = md.accessor(prim["attributes"]["POSITION"])
                if "indices" in prim:
                    idx, _ = md.accessor(prim["indices"])
                    idx = idx.reshape(-1).astype(np.int64)
                else:
                    idx = np.arange(len(pos), dtype=np.int64)
                ntri = len(idx) // 3
                idx = idx[:ntri * 3].reshape(ntri, 3)
                if is_hitbox:
                    hitboxes.append({"name": gm.get("name", ""), "xform": xf,
                                     "tris": pos.astype(F)[idx]})
                    continue
                wpos = transform_points(xf, pos.astype(F))
                corners = wpos[idx]                                 # [ntri][3][3]
                v0 = corners[:, 0, :]
                e1 = (corners[:, 1, :] - v0).astype(F)
                e2 = (corners[:, 2, :] - v0).astype(F)
                tri_chunks.append(np.concatenate([v0, e1, e2], axis=1).astype(F))
                vert_chunks.append(corners.astype(F))
                mesh_ids.append(np.full(ntri, len(meshes), dtype=np.int32))
                has_uv = "TEXCOORD_0" in prim["attributes"]
                if has_uv:
                    uv, _ = md.accessor(prim["attributes"]["TEXCOORD_0"])
                    uv_chunks.append(uv.astype(F)[idx])
                else:
                    uv_chunks.append(np.zeros((ntri, 3, 2), dtype=F))
                color_type = -1
                if "COLOR_0" in prim["attributes"]:
                    col, cacc = md.accessor(prim["attributes"]["COLOR_0"])
                    if cacc["type"] == "VEC4" and cacc["componentType"] in (5126, 5123, 5121):
                        color_type = cacc["componentType"]
                        if color_type == 5126:
                            colf = col.astype(F)
                        elif color_type == 5123:                    # LocalGeometry.h:126-137: *= 1/65535
                            colf = (col.astype(F) * F(F(1.0) / F(65535.0))).astype(F)
                        else:                                       # :113-125
                            colf = (col.astype(F) * F(F(1.0) / F(255.0))).astype(F)
                        col_chunks.append(colf[idx])
                if color_type == -1:
                    col_chunks.append(np.zeros((ntri, 3, 4), dtype=F))
                mat = prim.get("material", -1)
                bmin = np.array(pacc.get("min", pos.min(axis=0)), dtype=F)
                bmax = np.array(pacc.get("max", pos.max(axis=0)), dtype=F)
                wmin, wmax = aabb_transform(xf, bmin, bmax)
                meshes.append(MeshGroup(gm.get("name", ""), state["T"], ntri, color_type, has_uv, mat, wmin, wmax))
                mi = {"color_type": color_type, "has_uv": int(has_uv), "tex": -1,
                      "base_color": np.array([1, 1, 1, 1], dtype=F)}
                if mat >= 0:
                    mi["tex"] = materials[mat]["tex"]
                    mi["base_color"] = materials[mat]["base_color"]
                mesh_info.append(mi)
                state["T"] += ntri
            return
        for c in node.get("children", []):
            process(j["nodes"][c], xf)

    nodes = j.get("nodes", [])
    is_root = [True] * len(nodes)
    for n in nodes:
        for c in n.get("children", []):
            is_root[c] = False
    for i, n in enumerate(nodes):
        if is_root[i]:
            process(n, mat_identity())

    T = state["T"]
    if T:
        tris = np.concatenate(tri_chunks).astype(F)
        verts = np.concatenate(vert_chunks).astype(F)
        tri_mesh = np.concatenate(mesh_ids).astype(np.int32)
        corner_uv = np.ascontiguousarray(np.concatenate(uv_chunks).astype(F))
        corner_col = np.ascontiguousarray(np.concatenate(col_chunks).astype(F))
    else:
        tris = np.zeros((0, 9), dtype=F)
        verts = np.zeros((0, 3, 3), dtype=F)
        tri_mesh = np.zeros(0, dtype=np.int32)
        corner_uv = np.zeros((0, 3, 2), dtype=F)
        corner_col = np.zeros((0, 3, 4), dtype=F)
    return Scene(cameras, meshes, hitboxes, np.ascontiguousarray(tris), verts, tri_mesh, corner_uv, corner_col,
                 mesh_info, textures, miss)
