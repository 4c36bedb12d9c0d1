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
