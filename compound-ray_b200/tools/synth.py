"""Synthetic scenes and eyes for the BASELINE.json configs whose original assets are not in the
reference checkout (.MISSING_LARGE_BLOBS) -- SURVEY.md section 8(d).

  write_terrain_gltf   config 4 "speed-test": heightfield terrain of ~T triangles with per-vertex
                       u16 colours and the simple_sky background, a compound camera and a panoramic
                       camera.  Geometry goes to an external .bin next to the .gltf.
  fibonacci_eye        N ommatidia evenly spread on a sphere (seed-free), acceptance angle set so the
                       cones tile the sphere.
  heterogeneous_eye    config 5: positions/directions of a given eye, acceptance angles log-uniform.
  write_eye            .eye writer (data/eyes/eye-specification.txt: 8 floats per line).

All generators are deterministic (fixed seeds) and use numpy only.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np


def _value_noise(x, z, seed, octaves=3):
    """3-octave value noise on a lattice hashed from `seed` (deterministic, vectorised)."""
    rng = np.random.default_rng(seed)
    table = rng.random((256, 256)).astype(np.float64)
    out = np.zeros_like(x, dtype=np.float64)
    amp, freq = 1.0, 1.0
    for _ in range(octaves):
        xs, zs = x * freq, z * freq
        x0, z0 = np.floor(xs).astype(np.int64), np.floor(zs).astype(np.int64)
        fx, fz = xs - x0, zs - z0
        sx, sz = fx * fx * (3 - 2 * fx), fz * fz * (3 - 2 * fz)
        v00 = table[x0 & 255, z0 & 255]
        v10 = table[(x0 + 1) & 255, z0 & 255]
        v01 = table[x0 & 255, (z0 + 1) & 255]
        v11 = table[(x0 + 1) & 255, (z0 + 1) & 255]
        out += amp * ((v00 * (1 - sx) + v10 * sx) * (1 - sz) + (v01 * (1 - sx) + v11 * sx) * sz)
        amp *= 0.5
        freq *= 2.0
    return out / 1.75


def terrain_height(x, z, extent=1000.0, amplitude=15.0, seed=7):
    return amplitude * _value_noise((x / extent + 0.5) * 8.0, (z / extent + 0.5) * 8.0, seed)


def write_eye(path, omm):
    omm = np.asarray(omm, dtype=np.float64).reshape(-1, 8)
    with open(path, "w") as f:
        for row in omm:
            f.write(" ".join("{:0.10f}".format(v) for v in row) + "\n")


def fibonacci_eye(n=10000, radius=0.1, focal=0.0, acceptance=None):
    """float32[n][8]: positions radius*dir, directions on the Fibonacci sphere."""
    i = np.arange(n, dtype=np.float64) + 0.5
    y = 1.0 - 2.0 * i / n
    r = np.sqrt(np.maximum(0.0, 1.0 - y * y))
    phi = i * math.pi * (3.0 - math.sqrt(5.0))
    d = np.stack([np.cos(phi) * r, y, np.sin(phi) * r], axis=1)
    if acceptance is None:
        acceptance = 2.0 * math.sqrt(4.0 * math.pi / n / math.pi)
    omm = np.zeros((n, 8), dtype=np.float64)
    omm[:, 0:3] = d * radius
    omm[:, 3:6] = d
    omm[:, 6] = acceptance
    omm[:, 7] = focal
    return omm.astype(np.float32)


def heterogeneous_eye(base_omm, lo=0.02, hi=0.35, seed=5):
    """Same geometry, acceptance angles drawn log-uniform in [lo, hi] (numpy default_rng(seed))."""
    omm = np.array(base_omm, dtype=np.float32).reshape(-1, 8).copy()
    rng = np.random.default_rng(seed)
    omm[:, 6] = np.exp(rng.uniform(math.log(lo), math.log(hi), len(omm))).astype(np.float32)
    return omm


def write_terrain_gltf(path, triangles=1_000_000, extent=1000.0, amplitude=15.0, seed=7, eye_file="eye.eye",
                       projection="single_dimension_fast", camera_height=2.0):
    """Writes <path> (.gltf) + <path>.bin.  Returns dict(triangles, vertices, camera_position)."""
    n = int(math.ceil(math.sqrt(triangles / 2.0))) + 1
    lin = (np.arange(n, dtype=np.float64) / (n - 1) - 0.5) * extent
    gx, gz = np.meshgrid(lin, lin, indexing="xy")
    gy = terrain_height(gx, gz, extent, amplitude, seed)
    pos = np.stack([gx, gy, gz], axis=-1).reshape(-1, 3).astype(np.float32)
    # colours: height/slope tinted greens and browns, u16 RGBA
    h = (gy - gy.min()) / max(1e-9, (gy.max() - gy.min()))
    tint = _value_noise((gx / extent + 0.5) * 64.0, (gz / extent + 0.5) * 64.0, seed + 1, octaves=2)
    rgb = np.stack([0.25 + 0.5 * h + 0.15 * tint, 0.35 + 0.4 * (1 - h) + 0.2 * tint, 0.15 + 0.2 * tint], axis=-1)
    col = np.concatenate([np.clip(rgb, 0, 1), np.ones(rgb.shape[:2] + (1,))], axis=-1).reshape(-1, 4)
    col16 = np.round(col * 65535.0).astype(np.uint16)
    # two triangles per cell, alternating diagonal
    ii, jj = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="xy")
    v00 = (jj * n + ii).reshape(-1).astype(np.uint32)
    v10, v01, v11 = v00 + 1, v00 + n, v00 + n + 1
    flip = ((ii + jj) & 1).reshape(-1).astype(bool)
    t1 = np.where(flip[:, None], np.stack([v00, v01, v10], 1), np.stack([v00, v01, v11], 1))
    t2 = np.where(flip[:, None], np.stack([v10, v01, v11], 1), np.stack([v00, v11, v10], 1))
    idx = np.stack([t1, t2], axis=1).reshape(-1, 3).astype(np.uint32)

    blob = bytearray()
    views = []

    def add(arr, target=None):
        while len(blob) % 4:
            blob.append(0)
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": arr.nbytes})
        if target:
            views[-1]["target"] = target
        blob.extend(arr.tobytes())
        return len(views) - 1

    v_pos = add(pos, 34962)
    v_col = add(col16, 34962)
    v_idx = add(idx.reshape(-1), 34963)
    bin_name = os.path.basename(path) + ".bin"
    cam_y = float(terrain_height(np.array([0.0]), np.array([0.0]), extent, amplitude, seed)[0] + camera_height)
    rot_up = [0.7071067690849304, 0, 0, 0.7071067690849304]       # Blender-style camera rig as in the
    rot_dn = [-0.7071067690849304, 0, 0, 0.7071067690849304]      # reference scenes (parent + _Orientation)
    gltf = {
        "asset": {"version": "2.0", "generator": "compound-ray_b200 tools/synth.py"},
        "scene": 0,
        "scenes": [{"name": "Scene", "nodes": [0, 2, 4], "extras": {"background-shader": "simple_sky"}}],
        "nodes": [
            {"mesh": 0, "name": "Terrain"},
            {"camera": 0, "name": "compound-cam_Orientation", "rotation": rot_dn},
            {"children": [1], "name": "compound-cam", "rotation": rot_up, "translation": [0.0, cam_y, 0.0]},
            {"camera": 1, "name": "pano-cam_Orientation", "rotation": rot_dn},
            {"children": [3], "name": "pano-cam", "rotation": rot_up, "translation": [0.0, cam_y, 0.0]},
        ],
        "cameras": [
            {"name": "compound-cam", "type": "perspective",
             "perspective": {"yfov": 0.4, "znear": 0.1, "zfar": 1000},
             "extras": {"compound-eye": "true", "compound-projection": projection, "compound-structure": eye_file}},
            {"name": "pano-cam", "type": "perspective", "perspective": {"yfov": 0.4, "znear": 0.1, "zfar": 1000},
             "extras": {"panoramic": "true"}},
        ],
        "meshes": [{"name": "Terrain", "primitives": [{"attributes": {"POSITION": 0, "COLOR_0": 1}, "indices": 2}]}],
        "accessors": [
            {"bufferView": v_pos, "componentType": 5126, "count": int(len(pos)), "type": "VEC3",
             "min": [float(v) for v in pos.min(axis=0)], "max": [float(v) for v in pos.max(axis=0)]},
            {"bufferView": v_col, "componentType": 5123, "count": int(len(col16)), "type": "VEC4", "normalized": True},
            {"bufferView": v_idx, "componentType": 5125, "count": int(idx.size), "type": "SCALAR"},
        ],
        "bufferViews": views,
        "buffers": [{"byteLength": len(blob), "uri": bin_name}],
    }
    with open(path, "w") as f:
        json.dump(gltf, f)
    with open(os.path.join(os.path.dirname(path), bin_name), "wb") as f:
        f.write(bytes(blob))
    return {"triangles": int(len(idx)), "vertices": int(len(pos)), "camera_position": [0.0, cam_y, 0.0]}


# ------------------------------------------------------------------------------------------------
# BASELINE config 3: the "ofstad arena" of data/tools/minimumSampleRateFinder.py.  The reference's own
# data/ofstad-arena/ofstad-arena.gltf and ofstad_patterning.jpg are not in the checkout; SURVEY 8(d)3 fixes the
# stand-in: an open cylinder of radius 12.5 and height 9.19 (the bounds the script samples inside,
# minimum-samples-calculation/readme.txt:15), 256 segments x 64 rings = 32 768 triangles, a 1024^2 black/white
# rectangle pattern on its inside -- written as a JPEG so that the product's own decoder is on the path --, a floor
# disc, default background, a compound camera and a panoramic camera at the centre.
# ------------------------------------------------------------------------------------------------
ARENA_RADIUS = 12.5
ARENA_HEIGHT = 9.19


def arena_pattern(size=1024, seed=3):
    """uint8[size][size][3]: white wall with black rectangles of assorted widths/heights (a stand-in for the bars,
    boxes and gratings of the Ofstad et al. arena)."""
    rng = np.random.default_rng(seed)
    img = np.full((size, size, 3), 255, np.uint8)
    for k in range(48):
        w = int(rng.integers(size // 64, size // 6))
        h = int(rng.integers(size // 32, size // 2))
        x = int(rng.integers(0, size - w))
        y = int(rng.integers(0, size - h))
        img[y:y + h, x:x + w] = 0 if k % 5 else 96
    # one grating sector: 16 vertical bars
    gx0 = size // 8
    for b in range(16):
        if b % 2 == 0:
            img[size // 2:size // 2 + size // 4, gx0 + b * (size // 64):gx0 + (b + 1) * (size // 64)] = 0
    return img


def write_arena_gltf(path, segments=256, rings=64, radius=ARENA_RADIUS, height=ARENA_HEIGHT, eye_file="eye.eye",
                     projection="single_dimension_fast", camera_height=None, texture="jpeg", jpeg_quality=92):
    """Writes <path> (.gltf) + <path>.bin + the pattern image next to it.  Returns dict(triangles, ...).
    texture = "jpeg" | "png" (needs PIL to write the image) | "none" (white wall, base colour only)."""
    th = np.arange(segments + 1, dtype=np.float64) / segments * 2.0 * math.pi
    ys = np.arange(rings + 1, dtype=np.float64) / rings * height
    T, Y = np.meshgrid(th, ys, indexing="xy")                        # [rings+1][segments+1]
    wall = np.stack([radius * np.cos(T), Y, radius * np.sin(T)], axis=-1).reshape(-1, 3).astype(np.float32)
    wall_uv = np.stack([T / (2.0 * math.pi), 1.0 - Y / height], axis=-1).reshape(-1, 2).astype(np.float32)
    n1 = segments + 1
    ii, jj = np.meshgrid(np.arange(segments), np.arange(rings), indexing="xy")
    v00 = (jj * n1 + ii).reshape(-1).astype(np.uint32)
    v10, v01, v11 = v00 + 1, v00 + n1, v00 + n1 + 1
    wall_idx = np.stack([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=1).reshape(-1, 3).astype(np.uint32)
    # floor disc: fan of `segments` triangles, plain grey material (the base-colour path of the shader)
    floor = np.concatenate([np.zeros((1, 3)), np.stack([radius * np.cos(th[:-1]), np.zeros(segments), radius * np.sin(th[:-1])], 1)]).astype(np.float32)
    k = np.arange(segments, dtype=np.uint32)
    floor_idx = np.stack([np.zeros(segments, np.uint32), 1 + k, 1 + (k + 1) % segments], axis=1).astype(np.uint32)

    blob = bytearray()
    views = []

    def add(arr, target=None):
        while len(blob) % 4:
            blob.append(0)
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": arr.nbytes})
        if target:
            views[-1]["target"] = target
        blob.extend(arr.tobytes())
        return len(views) - 1

    v_wp, v_wuv, v_wi = add(wall, 34962), add(wall_uv, 34962), add(wall_idx.reshape(-1), 34963)
    v_fp, v_fi = add(floor, 34962), add(floor_idx.reshape(-1), 34963)
    bin_name = os.path.basename(path) + ".bin"
    base = os.path.dirname(path)
    cam_y = float(height * 0.25 if camera_height is None else camera_height)
    rot_up = [0.7071067690849304, 0, 0, 0.7071067690849304]
    rot_dn = [-0.7071067690849304, 0, 0, 0.7071067690849304]
    wall_prim = {"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 0}
    materials = [{"name": "wall", "pbrMetallicRoughness": {"baseColorFactor": [1.0, 1.0, 1.0, 1.0]}},
                 {"name": "floor", "pbrMetallicRoughness": {"baseColorFactor": [0.35, 0.35, 0.35, 1.0]}}]
    gltf = {
        "asset": {"version": "2.0", "generator": "compound-ray_b200 tools/synth.py"},
        "scene": 0,
        "scenes": [{"name": "Scene", "nodes": [0, 1, 3, 5]}],               # no background-shader extra: default_background
        "nodes": [
            {"mesh": 0, "name": "ArenaWall"},
            {"mesh": 1, "name": "ArenaFloor"},
            {"camera": 0, "name": "compound-cam_Orientation", "rotation": rot_dn},
            {"children": [2], "name": "compound-cam", "rotation": rot_up, "translation": [0.0, cam_y, 0.0]},
            {"camera": 1, "name": "pano-cam_Orientation", "rotation": rot_dn},
            {"children": [4], "name": "pano-cam", "rotation": rot_up, "translation": [0.0, cam_y, 0.0]},
        ],
        "cameras": [
            {"name": "compound-cam", "type": "perspective", "perspective": {"yfov": 0.4, "znear": 0.1, "zfar": 1000},
             "extras": {"compound-eye": "true", "compound-projection": projection, "compound-structure": eye_file}},
            {"name": "pano-cam", "type": "perspective", "perspective": {"yfov": 0.4, "znear": 0.1, "zfar": 1000},
             "extras": {"panoramic": "true"}},
        ],
        "meshes": [{"name": "ArenaWall", "primitives": [wall_prim]},
                   {"name": "ArenaFloor", "primitives": [{"attributes": {"POSITION": 3}, "indices": 4, "material": 1}]}],
        "materials": materials,
        "accessors": [
            {"bufferView": v_wp, "componentType": 5126, "count": int(len(wall)), "type": "VEC3",
             "min": [float(v) for v in wall.min(axis=0)], "max": [float(v) for v in wall.max(axis=0)]},
            {"bufferView": v_wuv, "componentType": 5126, "count": int(len(wall_uv)), "type": "VEC2"},
            {"bufferView": v_wi, "componentType": 5125, "count": int(wall_idx.size), "type": "SCALAR"},
            {"bufferView": v_fp, "componentType": 5126, "count": int(len(floor)), "type": "VEC3",
             "min": [float(v) for v in floor.min(axis=0)], "max": [float(v) for v in floor.max(axis=0)]},
            {"bufferView": v_fi, "componentType": 5125, "count": int(floor_idx.size), "type": "SCALAR"},
        ],
        "bufferViews": views,
        "buffers": [{"byteLength": len(blob), "uri": bin_name}],
    }
    image_name = None
    if texture in ("jpeg", "png"):
        from PIL import Image
        image_name = os.path.basename(path) + (".pattern.jpg" if texture == "jpeg" else ".pattern.png")
        im = Image.fromarray(arena_pattern())
        if texture == "jpeg":
            im.save(os.path.join(base, image_name), format="JPEG", quality=jpeg_quality, subsampling=2)
        else:
            im.save(os.path.join(base, image_name), format="PNG")
        gltf["images"] = [{"uri": image_name}]
        gltf["samplers"] = [{"magFilter": 9729, "minFilter": 9987, "wrapS": 10497, "wrapT": 10497}]
        gltf["textures"] = [{"source": 0, "sampler": 0}]
        materials[0]["pbrMetallicRoughness"]["baseColorTexture"] = {"index": 0}
    with open(path, "w") as f:
        json.dump(gltf, f)
    with open(os.path.join(base, bin_name), "wb") as f:
        f.write(bytes(blob))
    return {"triangles": int(len(wall_idx) + len(floor_idx)), "wall_triangles": int(len(wall_idx)), "vertices": int(len(wall) + len(floor)),
            "camera_position": [0.0, cam_y, 0.0], "image": image_name}


def ico_eye():
    """float32[12][8]: the icosahedral 12-ommatidia eye of eyeRendererHelperFunctions.getIcoOmmatidia (:171-194): vertices
    of an icosahedron, 1 steradian each (acceptance = 2*acos(1 - 1/(2*pi)))."""
    lat = math.atan(0.5)
    pts = [[0.0, 1.0, 0.0]]
    for ring, sign in ((0.0, 1.0), (0.2 * math.pi, -1.0)):
        for i in range(5):
            a = 0.4 * math.pi * i + ring
            pts.append([math.cos(a) * math.cos(lat), sign * math.sin(lat), math.sin(a) * math.cos(lat)])
    pts.append([0.0, -1.0, 0.0])
    omm = np.zeros((12, 8), np.float32)
    omm[:, 3:6] = np.asarray(pts, np.float32)
    omm[:, 6] = math.acos(-(1 / (2 * math.pi) - 1)) * 2
    return omm
