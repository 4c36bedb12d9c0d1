#!/usr/bin/env python
"""profiles/region_breakdown.py <cuda,sass source csv> <rays per launch> -- instruction share per code region of
cr_kernels.cu for the first kernel in an `ncu --page source --csv --print-source cuda,sass` export."""
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
rays = float(sys.argv[2])
hdr_idx = [i for i, r in enumerate(rows) if len(r) > 5 and 'Instructions Executed' in r]
h = rows[hdr_idx[0]]
i_s, i_i, i_t = h.index('# Samples'), h.index('Instructions Executed'), h.index('Thread Instructions Executed')


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


fname, lines = '', []
for r in rows:
    if r and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    if r and r[0].isdigit() and len(r) > i_t and r[i_i].isdigit():
        lines.append((fname, int(r[0]), I(r[i_s]), I(r[i_i]), I(r[i_t])))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'compound-ray_b200', 'csrc', 'cr_kernels.cu')).read().split('\n')


def find(s):
    return [i + 1 for i, l in enumerate(src) if s in l][0]


marks = [('vec helpers', '// small vector helpers'), ('rng', '// cuRAND XORWOW on a 32-byte'), ('raygen', '// Ommatidial sample ray'),
         ('box setup', '// Closest hit (stands in'), ('triTest', '__device__ __forceinline__ bool triTest'),
         ('stack', '// Traversal stack:'), ('trav loop', 'template <bool COUNT>'), ('shading', '// Shading: closest hit'),
         ('k1 body', '// K1.  Persistent warps'), ('rest', '// K1b:')]
bounds = [(n, find(m)) for n, m in marks] + [('end', 10 ** 9)]
ti = sum(x[3] for x in lines); ts = sum(x[2] for x in lines); tt = sum(x[4] for x in lines)
print(f'warp-inst/ray {ti / rays:.1f}  thread-inst/ray {tt / rays:.1f}  SIMD lanes {tt / ti:.1f}/32')
acc = {}
for f, ln, s, i, t in lines:
    key = 'cr_math.h' if f == 'cr_math.h' else ('other:' + f)
    if f == 'cr_kernels.cu':
        for k in range(len(bounds) - 1):
            if bounds[k][1] <= ln < bounds[k + 1][1]:
                key = bounds[k][0]
    a = acc.setdefault(key, [0, 0, 0]); a[0] += i; a[1] += s; a[2] += t
for k, (i, s, t) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    if i:
        print(f'{k:14s} warp-inst {100 * i / ti:5.1f}%  stall samples {100 * s / ts:5.1f}%  lanes {t / max(i, 1):5.1f}  thread-inst/ray {t / rays:7.1f}')
