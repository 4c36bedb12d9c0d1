#!/usr/bin/env python
"""profiles/make_sass_evidence.py -> profiles/r02_sass_evidence.txt: static SASS mnemonic counts and excerpts of the
shipped kernels (`cuobjdump -sass compound-ray_b200/lib/libEyeRenderer3.so`, sm_100a) plus the ptxas register / spill
lines of the build (compound-ray_b200/build/cr_kernels.ptxas.log).  Needs no GPU."""
import collections
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "compound-ray_b200", "lib", "libEyeRenderer3.so")
PAT = {"LDG.E.*256 (256-bit node loads, new with sm_100)": r"LDG\.E\S*\.256", "FMNMX3 (three-input min/max slab test)": r"FMNMX3",
       "UBLKCP (bulk TMA copy)": r"UBLKCP", "SYNCS (mbarrier)": r"SYNCS", "SHFL.BFLY (fused reduction butterfly)": r"SHFL\.BFLY",
       "MUFU (hardware sin/cos/lg2/ex2/rcp/rsq)": r"MUFU\.", "VOTE / ballot": r"VOTE", "WARPSYNC": r"WARPSYNC", "LDS": r"\bLDS", "STS": r"\bSTS",
       "LDL (local-memory stack spill levels)": r"\bLDL", "STL": r"\bSTL", "FFMA": r"FFMA",
       "ACQBULK (griddepcontrol.wait: programmatic dependent launch)": r"ACQBULK", "PREEXIT (griddepcontrol.launch_dependents)": r"PREEXIT",
       "S2R SR_VIRTUALSMID (%smid: SM-affine hand-out)": r"SR_VIRTUALSMID", "ATOMG (work counter, per-SM tickets)": r"ATOMG",
       "HMMA/tensor ops (none expected: no contraction on this path)": r"HMMA|UTCHMMA|TCGEN|QGMMA"}
WANT = ["k_traceCompound", "k_sumSamplesTma", "k_sumPartials", "k_buildEntries", "k_camera"]


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    for p in re.split(r"\n\s*Function : ", txt)[1:]:
        funcs[p.split("\n", 1)[0].strip()] = p
    out = ["SASS evidence, round 2 -- `cuobjdump -sass compound-ray_b200/lib/libEyeRenderer3.so` (sm_100a): static mnemonic counts per",
           "kernel, excerpts, and the build's register/spill lines (profiles/make_sass_evidence.py regenerates this file).", ""]
    for name, body in funcs.items():
        if not any(w in name for w in WANT):
            continue
        lines = [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        out.append(f"=== {demangle(name)}\n    {len(lines)} SASS instructions")
        for k, rx in PAT.items():
            c = sum(1 for l in lines if re.search(rx, l))
            if c or "none expected" in k:
                out.append(f"    {c:5d}  {k}")
        mufu = collections.Counter(m.group(0) for l in lines for m in [re.search(r"MUFU\.\w+", l)] if m)
        if mufu:
            out.append("           " + ", ".join(f"{k} x{v}" for k, v in sorted(mufu.items())))
        out.append("")

    def excerpt(part, rx, n, title):
        for name, body in funcs.items():
            if part in name:
                out.append(f"--- {title} ({demangle(name)[:110]})")
                out.extend("    " + l.strip() for l in [l for l in body.split("\n") if re.search(rx, l)][:n])
                out.append("")
                return
    excerpt("k_traceCompoundILb0ELb1ELb1ELb0", r"LDG\.E\S*\.256", 4, "256-bit BVH node loads in the batched fused trace kernel")
    excerpt("k_traceCompoundILb0ELb1ELb1ELb0", r"FMNMX3", 4, "three-input min/max of the slab test")
    excerpt("k_traceCompoundILb0ELb1ELb1ELb0", r"SHFL\.BFLY", 3, "shuffle butterfly of the fused reduction")
    excerpt("k_traceCompoundILb0ELb1ELb1ELb1", r"MUFU\.(SIN|COS|LG2|EX2)", 6, "hardware elementary functions of the fast-math variant")
    excerpt("k_sumSamplesTma", r"UBLKCP|SYNCS", 8, "bulk TMA copy + mbarrier of the ordered-sum kernel")
    log = os.path.join(ROOT, "compound-ray_b200", "build", "cr_kernels.ptxas.log")
    if os.path.exists(log):
        out.append("--- ptxas -v (compound-ray_b200/build/cr_kernels.ptxas.log): registers, stack frame, spills, shared memory per entry")
        cur = None
        for l in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", l)
            if m:
                cur = demangle(m.group(1))
            elif cur and ("spill" in l or "Used" in l) and any(w in cur for w in WANT):
                out.append(f"    {cur[:100]:100s} | {l.strip().replace('ptxas info    : ', '')}")
    open(os.path.join(HERE, "r02_sass_evidence.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[-40:]))


if __name__ == "__main__":
    main()
