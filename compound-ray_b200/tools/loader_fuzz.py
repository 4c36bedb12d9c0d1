#!/usr/bin/env python
"""tools/loader_fuzz.py -- mutation fuzzing of the host-side parsers (glTF/JSON/base64 loader, .eye reader, PNG and
JPEG decoders) under AddressSanitizer + UndefinedBehaviourSanitizer.

The parsers are plain C++ (csrc/cr_scene.cpp and its headers), so they are compiled without CUDA into
tools/loader_fuzz_harness.cpp and fed mutated copies of real scenes and images: byte flips, numeric-token swaps
("count": 4294967295, -1, 1e30 ...), truncations, deleted and duplicated slices.  Every input must either load or
be rejected with an exception; any sanitizer report fails the run.

  python compound-ray_b200/tools/loader_fuzz.py --gltf 12000 --images 25000 [--seed 1] [--work DIR]

Campaign of record (round 1): 12 000 scenes + 25 000 images, no report -- after it had found and fixed a heap overflow
behind wrapped accessor range checks, unbounded recursion on cyclic node graphs, a Huffman-table overflow on
over-subscribed JPEG code lengths, header-driven giant allocations and signed overflow in the IDCT.
"""
from __future__ import annotations

import argparse
import glob
import os
import random
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
_NUM = re.compile(r'(?<=[:\[, ])-?\d+(\.\d+)?(?=[,\]\} \n])')
_SWAPS = ['-1', '0', '1', '2', '3', '7', '12', '4294967295', '99999999', '1e30', '-0.0', '65536', '2147483648', '1e-30',
          '5121', '5123', '5125', '5126', '100000', '-2147483649', '0.5', '1e300']


def build_harness(out_dir: str) -> str:
    exe = os.path.join(out_dir, "loader_fuzz_harness")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-fno-omit-frame-pointer", "-I" + os.path.join(PKG, "csrc"), os.path.join(HERE, "loader_fuzz_harness.cpp"),
                           os.path.join(PKG, "csrc", "cr_scene.cpp"), "-lz", "-o", exe])
    return exe


def mutate_gltf(src: bytes, rnd: random.Random) -> bytes:
    b = bytearray(src)
    head = min(len(b), 14000)                 # the JSON structure precedes the base64 payload: aim most edits there
    kind = rnd.random()
    if kind < 0.3:
        for _ in range(rnd.randint(1, 6)):
            p = rnd.randrange(head) if rnd.random() < 0.8 else rnd.randrange(len(b))
            b[p] = rnd.randrange(256)
    elif kind < 0.75:
        s = b.decode("latin1")
        text = s[:head]
        for _ in range(rnd.randint(1, 4)):
            toks = list(_NUM.finditer(text))
            if not toks:
                break
            m = rnd.choice(toks)
            text = text[:m.start()] + rnd.choice(_SWAPS) + text[m.end():]
        b = bytearray((text + s[head:]).encode("latin1"))
    elif kind < 0.85:
        b = b[:rnd.randrange(len(b))]
    else:
        a = rnd.randrange(head)
        c = min(len(b), a + rnd.randint(1, 300))
        if rnd.random() < 0.5:
            del b[a:c]
        else:
            b[a:a] = b[a:c]
    return bytes(b)


def mutate_image(src: bytes, rnd: random.Random) -> bytes:
    b = bytearray(src)
    kind = rnd.random()
    if kind < 0.6:
        for _ in range(rnd.randint(1, 5)):
            p = rnd.randrange(len(b))
            b[p] = rnd.choice([0, 255, rnd.randrange(256), b[p] ^ (1 << rnd.randrange(8))])
    elif kind < 0.8:
        b = b[:rnd.randrange(len(b))]
    else:
        a = rnd.randrange(len(b))
        c = min(len(b), a + rnd.randint(1, 64))
        if rnd.random() < 0.5:
            del b[a:c]
        else:
            b[a:a] = b[a:c]
    return bytes(b)


def run(work: str, scenes: list[str], images: list[str], n_gltf: int, n_images: int, seed: int = 1, exe: str | None = None):
    """Returns (accepted, rejected, sanitizer report text).  `scenes` must sit next to the .eye files they name."""
    os.makedirs(work, exist_ok=True)
    exe = exe or build_harness(work)
    rnd = random.Random(seed)
    files = []
    scene_bytes = [(k, os.path.dirname(p), open(p, "rb").read()) for k, p in enumerate(scenes)]
    for i in range(n_gltf):
        k, d, src = rnd.choice(scene_bytes)
        sub = os.path.join(work, "g%d" % k)
        if not os.path.isdir(sub):
            os.makedirs(sub)
            for entry in os.listdir(d):                                   # eye tables / eyes/ directories referenced by the scene
                if entry.endswith(".eye") or os.path.isdir(os.path.join(d, entry)):
                    os.symlink(os.path.join(d, entry), os.path.join(sub, entry))
        path = os.path.join(sub, "f%06d.gltf" % i)
        with open(path, "wb") as f:
            f.write(mutate_gltf(src, rnd))
        files.append(path)
    image_bytes = [open(p, "rb").read() for p in images]
    for i in range(n_images):
        path = os.path.join(work, "i%06d.bin" % i)
        with open(path, "wb") as f:
            f.write(mutate_image(rnd.choice(image_bytes), rnd))
        files.append(path)
    accepted = rejected = 0
    report = ""
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=1", UBSAN_OPTIONS="print_stacktrace=1")
    for k in range(0, len(files), 400):
        r = subprocess.run([exe] + files[k:k + 400], capture_output=True, text=True, env=env)
        for line in r.stdout.splitlines():
            if line.startswith("accepted "):
                parts = line.split()
                accepted += int(parts[1]); rejected += int(parts[3])
        if r.returncode != 0 or "Sanitizer" in r.stderr or "runtime error" in r.stderr:
            report += r.stderr[-4000:]
    return accepted, rejected, report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gltf", type=int, default=2000)
    ap.add_argument("--images", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--work", default="/tmp/cr_loader_fuzz")
    args = ap.parse_args()
    data = os.path.join(ROOT, "tests", "_data")
    scenes = [os.path.join(data, "data", "test-scene", "test-scene.gltf"), os.path.join(data, "data", "test-scene", "test-scene-sky.gltf"),
              os.path.join(data, "sim-environment", "env_2.gltf")]
    images = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "png", "*.png")) + glob.glob(os.path.join(ROOT, "tests", "golden", "jpeg", "*.jpg")))
    acc, rej, report = run(args.work, scenes, images, args.gltf, args.images, args.seed)
    print(f"accepted {acc}, rejected {rej}, sanitizer reports: {'none' if not report else 'YES'}")
    if report:
        print(report)
        sys.exit(1)


if __name__ == "__main__":
    main()
