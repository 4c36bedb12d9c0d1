"""tests/golden/make_fixtures.py -- regenerates the committed fixtures (run in the dev container,
where /root/reference and the CUDA toolkit exist; the GPU box has neither the reference tree).

  reference_data.tar.gz  INPUT DATA (scenes + .eye tables) packed from the reference's data directories, so the
                         parity tests can run on the reference's own scenes on a box without /root/reference --
                         plus, as acceptance fixtures, the reference's ctypes helper and two example scripts, which
                         tests/test_gpu_scripts.py runs unmodified against this library.
  xorwow_kat.json        cuRAND's own host XORWOW implementation (oracle/_ref/curand_kat)
  sutil_kat.json         the reference's sutil math headers evaluated on fixed inputs (oracle/_ref/sutil_kat)
  hitscan_kat.json       the reference's sutil/hitscanprocessing.cpp (point-in-hitbox, bounds) on three meshes x 400 points
  helper_kat.json        the reference's eyeRendererHelperFunctions.py (pure-Python parts) evaluated on fixed inputs
  tinygltf_kat.json      the reference's vendored tinygltf + stb_image + sutil on its own scenes: cameras in insertion order
                         with pose bits, per-primitive facts and digests of world-space triangles / UVs / colours, texture digests
  jpeg/*.jpg             small JPEG test streams written with PIL (4:4:4 / 4:2:2 / 4:2:0 / grey / progressive)
  stb_jpeg_kat.{json,npz} those streams decoded by the reference's vendored stb_image.h (oracle/_ref/stb_kat)
  png/*.png              small PNG streams of every colour type / bit depth / tRNS form, plain and Adam7 (own writer)
  stb_png_kat.{json,npz}  those streams decoded by the reference's vendored stb_image.h
  reference_outputs.tar.gz  OUTPUT frames the reference itself rendered and its authors committed next to the
                         scripts that made them (python-examples/alias-demonstration/viewpoint-experiment.py,
                         heterogeneous-demonstration/demonstration.py, overview-images/overviewImages.py), plus the
                         two .eye tables those scripts read.  They were rendered in the authors' (unpublished)
                         natural environment, of which data/natural-standin-sky.gltf keeps the camera, the eye and
                         the simple_sky background: every ommatidium that sees only sky is a known answer
                         (tests/test_reference_outputs.py).  Also docs/images/standin-sky-render.png, a screenshot
                         of the reference's viewer on the stand-in scene itself (ground texture and terrain included).
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
    # The reference's own ctypes helper and two of its example scripts, as ACCEPTANCE FIXTURES: tests/test_gpu_scripts.py
    # runs them UNMODIFIED (byte for byte as packed here) against this library -- the drop-in claim of INTEGRATION.md.
    # They are test inputs, never imported by the product.
    "python-examples/eyeRendererHelperFunctions.py": "python-examples/eyeRendererHelperFunctions.py",
    "python-examples/primary-example.py": "python-examples/primary-example.py",
    "python-examples/position-estimation-toy-experiment/compoundRayIterators.py":
        "python-examples/position-estimation-toy-experiment/compoundRayIterators.py",
}


PY = "python-examples"
OUTPUTS = {
    # archive name                                   : path under /root/reference
    "alias-demonstration/spherical-image-0-samples.ppm": PY + "/alias-demonstration/output/view-images/spherical-image-0-samples.ppm",
    "alias-demonstration/spherical-image-700-samples.ppm": PY + "/alias-demonstration/output/view-images/spherical-image-700-samples.ppm",
    # viewpoint-experiment.py:51-66 -- column k of this image is column k of the frame rendered with k+1 samples
    "alias-demonstration/combinedImage-700segs.png": PY + "/alias-demonstration/output/view-images/combinedImage-700segs.png",
    "heterogeneous-demonstration/1000-extreme-horizontallyAcute-variableDegree.eye":
        PY + "/heterogeneous-demonstration/1000-extreme-horizontallyAcute-variableDegree.eye",
    "heterogeneous-demonstration/heterogeneous-omms-4.ppm": PY + "/heterogeneous-demonstration/heterogeneous-omms-4.ppm",
    "heterogeneous-demonstration/homogeneous-omms-small-4.ppm": PY + "/heterogeneous-demonstration/homogeneous-omms-small-4.ppm",
    "overview-images/uniform-omms.ppm": PY + "/overview-images/uniform-omms.ppm",
    "overview-images/acute-omms.ppm": PY + "/overview-images/acute-omms.ppm",
    # screenshot of the reference's GUI showing data/natural-standin-sky.gltf through its first camera (README figure)
    "docs/images/standin-sky-render.png": "docs/images/standin-sky-render.png",
    # ... and showing data/test-scene/test-scene.gltf through insect-cam-1 (found to be S = 41, frame 8248)
    "docs/images/test-scene-running.png": "docs/images/test-scene-running.png",
    # quantified-experiment.py: per-ommatidium variance of the 8-bit eye vector over 1000 / 100 consecutive frames;
    # the number in the name is the loop index, i.e. samples per ommatidium minus one
    "alias-demonstration/vector-data/variance-0-samples.txt": PY + "/alias-demonstration/output/vector-data/variance-0-samples.txt",
    "alias-demonstration/vector-data-100samples/variance-0-samples.txt": PY + "/alias-demonstration/output/vector-data-100samples/variance-0-samples.txt",
    "alias-demonstration/vector-data-100samples/variance-3-samples.txt": PY + "/alias-demonstration/output/vector-data-100samples/variance-3-samples.txt",
    "alias-demonstration/vector-data-100samples/variance-15-samples.txt": PY + "/alias-demonstration/output/vector-data-100samples/variance-15-samples.txt",
}


def pack(out, files):
    with tarfile.open(out, "w:gz", compresslevel=9) as tar:
        for arc, src in sorted(files.items()):
            info = tar.gettarinfo(os.path.join(REF, src), arcname=arc)
            info.mtime = 0
            info.uid = info.gid = 0
            info.uname = info.gname = ""
            with open(os.path.join(REF, src), "rb") as fh:
                tar.addfile(info, fh)
    print(out, os.path.getsize(out), "bytes")


def make_helper_kat():
    """helper_kat.json: the reference's own python-examples/eyeRendererHelperFunctions.py evaluated here (pure Python,
    no renderer needed): getIcoOmmatidia, getSolidAngle, saveEyeFile's text, readEyeFile, decodeProjectionMapID."""
    import json
    import sys
    import tempfile
    sys.path.insert(0, os.path.join(REF, "python-examples"))
    import eyeRendererHelperFunctions as R
    ico = R.getIcoOmmatidia()
    with tempfile.TemporaryDirectory() as tmp:
        R.saveEyeFile(os.path.join(tmp, "ico.eye"), ico)
        text = open(os.path.join(tmp, "ico.eye")).read()
    eye = R.readEyeFile(os.path.join(REF, "data", "test-scene", "test100.eye"))
    out = {"ico": [[float(v).hex() for v in (*o.position, *o.direction, o.acceptanceAngle, o.focalpointOffset)] for o in ico],
           "ico_solid_angles": [o.getSolidAngle().hex() for o in ico],
           "ico_eye_file": text,
           "test100_rows": [[float(v).hex() for v in (*o.position, *o.direction, o.acceptanceAngle, o.focalpointOffset)] for o in eye[:5] + eye[-2:]],
           "test100_solid_angles": [o.getSolidAngle().hex() for o in eye[:5]],
           "decode_ids": [[q, R.decodeProjectionMapID(q)] for q in ([0, 0, 0, 0], [1, 2, 3, 4], [0, 0, 3, 231], [255, 255, 255, 255])]}
    with open(os.path.join(HERE, "helper_kat.json"), "w") as f:
        json.dump(out, f, indent=1)


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


def _png_chunk(tag, body):
    import struct
    import zlib
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)


def write_png(path, samples, depth, ctype, interlace=False, plte=None, trns=None):
    """Minimal PNG writer for fixtures PIL/OpenCV cannot produce (Adam7, colour keys at every depth).
    samples: uint16 array [H][W][channels] of raw sample values; filter type cycles 0..4 over the rows."""
    import struct
    import zlib
    import numpy as np
    H, W, ch = samples.shape

    def pack_rows(img):
        h, w, _ = img.shape
        out = bytearray()
        prev = None
        for y in range(h):
            if depth == 16:
                row = img[y].astype(">u2").tobytes()
            elif depth == 8:
                row = img[y].astype(np.uint8).tobytes()
            else:
                bits = np.unpackbits(img[y].astype(np.uint8).reshape(-1, 1), axis=1)[:, 8 - depth:].reshape(-1)
                row = np.packbits(bits).tobytes()
            bpp = max(1, ch * depth // 8)
            cur = np.frombuffer(row, np.uint8).astype(np.int32)
            up = np.zeros_like(cur) if prev is None else prev
            left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
            ul = np.concatenate([np.zeros(bpp, np.int32), up[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
            ft = y % 5
            if ft == 0: f = cur
            elif ft == 1: f = cur - left
            elif ft == 2: f = cur - up
            elif ft == 3: f = cur - ((left + up) >> 1)
            else:
                pp = left + up - ul
                pa, pb, pc = np.abs(pp - left), np.abs(pp - up), np.abs(pp - ul)
                pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
                f = cur - pred
            out += bytes([ft]) + (f & 255).astype(np.uint8).tobytes()
            prev = cur
        return bytes(out)

    if interlace:
        ox, oy, sx, sy = (0, 4, 0, 2, 0, 1, 0), (0, 0, 4, 0, 2, 0, 1), (8, 8, 4, 4, 2, 2, 1), (8, 8, 8, 4, 4, 2, 2)
        raw = b"".join(pack_rows(samples[oy[p]::sy[p], ox[p]::sx[p]]) for p in range(7)
                       if samples[oy[p]::sy[p], ox[p]::sx[p]].size)
    else:
        raw = pack_rows(samples)
    body = b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, ctype, 0, 0, 1 if interlace else 0))
    if plte is not None:
        body += _png_chunk(b"PLTE", bytes(plte))
    if trns is not None:
        body += _png_chunk(b"tRNS", bytes(trns))
    half = len(raw) // 2 or 1
    z = zlib.compress(raw, 9)
    body += _png_chunk(b"IDAT", z[:len(z) // 2]) + _png_chunk(b"IDAT", z[len(z) // 2:]) + _png_chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(body)


def make_png_kat():
    """png/*.png: every colour type / bit depth / tRNS form / Adam7, decoded by the reference's stb_image -> stb_png_kat.*"""
    import json
    import struct
    import tempfile
    import numpy as np
    out = os.path.join(HERE, "png")
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(3)
    W, H = 13, 11

    def smooth(ch, maxv):
        yy, xx = np.mgrid[0:H, 0:W]
        base = np.stack([(xx * (k + 2) + yy * (3 - k)) for k in range(ch)], axis=2).astype(np.float64)
        return np.clip(base / base.max() * maxv + rng.integers(0, max(1, maxv // 8) + 1, (H, W, ch)), 0, maxv).astype(np.uint16)

    pal16 = rng.integers(0, 256, 16 * 3, dtype=np.uint8)
    pal256 = rng.integers(0, 256, 200 * 3, dtype=np.uint8)
    for il in (False, True):
        tagi = "_adam7" if il else ""
        for depth in (1, 2, 4, 8, 16):
            g = smooth(1, (1 << depth) - 1)
            write_png(os.path.join(out, f"grey{depth}{tagi}.png"), g, depth, 0, il)
            key = int(g[2, 3, 0])
            write_png(os.path.join(out, f"grey{depth}_key{tagi}.png"), g, depth, 0, il, trns=struct.pack(">H", key))
        for depth in (8, 16):
            m = (1 << depth) - 1
            rgb = smooth(3, m)
            write_png(os.path.join(out, f"rgb{depth}{tagi}.png"), rgb, depth, 2, il)
            rgb[4:6, 2:9] = rgb[0, 0]                              # a block of the key colour
            write_png(os.path.join(out, f"rgb{depth}_key{tagi}.png"), rgb, depth, 2, il, trns=struct.pack(">HHH", *[int(v) for v in rgb[0, 0]]))
            write_png(os.path.join(out, f"greyalpha{depth}{tagi}.png"), smooth(2, m), depth, 4, il)
            write_png(os.path.join(out, f"rgba{depth}{tagi}.png"), smooth(4, m), depth, 6, il)
        for depth, pal in ((1, pal16[:6]), (2, pal16[:12]), (4, pal16), (8, pal256)):
            n = len(pal) // 3
            idx = rng.integers(0, n, (H, W, 1)).astype(np.uint16)
            write_png(os.path.join(out, f"pal{depth}{tagi}.png"), idx, depth, 3, il, plte=pal)
            write_png(os.path.join(out, f"pal{depth}_trns{tagi}.png"), idx, depth, 3, il, plte=pal,
                      trns=rng.integers(0, 256, max(1, n // 2), dtype=np.uint8))
    for w, h in ((1, 1), (2, 3), (5, 1), (1, 9), (9, 8)):        # Adam7 passes that come out empty
        write_png(os.path.join(out, f"tiny_{w}x{h}_adam7.png"), rng.integers(0, 256, (h, w, 3)).astype(np.uint16), 8, 2, True)
    files = sorted(os.path.join(out, f) for f in os.listdir(out) if f.endswith(".png"))
    with tempfile.TemporaryDirectory() as tmp:
        idx = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "stb_kat"), tmp] + files, text=True)
        with open(os.path.join(HERE, "stb_png_kat.json"), "w") as f:
            f.write(idx)
        arrs = {}
        for name, e in json.loads(idx).items():
            if "error" not in e:
                arrs[name] = np.fromfile(os.path.join(tmp, name + ".rgba"), dtype=np.uint8).reshape(e["height"], e["width"], 4)
        np.savez_compressed(os.path.join(HERE, "stb_png_kat.npz"), **arrs)
    return len(files), len(arrs)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    with open(os.path.join(HERE, "xorwow_kat.json"), "w") as f:
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "curand_kat")], stdout=f)
    with open(os.path.join(HERE, "sutil_kat.json"), "w") as f:
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "sutil_kat")], stdout=f)
    with open(os.path.join(HERE, "hitscan_kat.json"), "w") as f:
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "hitscan_kat")], stdout=f)
    with open(os.path.join(HERE, "tinygltf_kat.json"), "w") as f:      # keys = archive names of reference_data.tar.gz
        args = []
        for arc, src in sorted(FILES.items()):
            if arc.endswith(".gltf"):
                args += [arc, os.path.join(REF, src)]
        # + a hand-made scene covering what the reference's scenes do not: a matrix node above TRS nodes, interleaved
        # attributes (byteStride), u32 indices, accessor byteOffset, u16 and float COLOR_0, a non-indexed primitive,
        # baseColorFactor, an external-file and a data-URI image, an orthographic camera under a rotated parent
        args += ["synthetic/features.gltf", os.path.join(HERE, "synthetic", "features.gltf")]
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "tinygltf_kat")] + args, stdout=f, cwd="/")
    make_helper_kat()
    make_jpeg_kat()
    make_png_kat()
    pack(os.path.join(HERE, "reference_data.tar.gz"), FILES)
    pack(os.path.join(HERE, "reference_outputs.tar.gz"), OUTPUTS)


if __name__ == "__main__":
    main()
