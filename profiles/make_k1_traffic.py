#!/usr/bin/env python
"""profiles/make_k1_traffic.py <mode>:<frames>:<ncu-csv-or-rep> ... -> profiles/k1_traffic.json

Builds the table bench.py reconstructs `roofline.traffic` / `roofline.issue` from: per render mode, the ncu figures of
ONE launch of the batched trace kernel on the headline workload at two or more frames-per-launch values
(`ncu --clock-control none -k regex:k_traceCompound -s 1 -c 1 ... python bench.py --mode M --steps F --repeats 1`).
Inputs are `.ncu-rep` files (read with `ncu -i ... --page raw --csv`) or that CSV saved to a file."""
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = {"dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "smsp__inst_executed.sum": "inst_executed", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "thread_inst_per_warp_inst", "gpu__time_duration.sum": "duration"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "inst": 1.0, "%": 1.0, "": 1.0}


def rows_of(path):
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        text = open(path).read()
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.reader(lines))


def capture(path):
    rows = rows_of(path)
    hdr, units = rows[0], rows[1]
    r = [x for x in rows[2:] if "k_traceCompound" in x[hdr.index("Kernel Name")]][0]
    out = {"kernel": r[hdr.index("Kernel Name")]}
    for i, h in enumerate(hdr):
        if h in KEYS:
            out[KEYS[h]] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    out["dram_bytes"] = out.pop("dram_bytes_read") + out.pop("dram_bytes_write")
    out["duration_ms_under_ncu"] = out.pop("duration")
    return out


def main():
    modes = {}
    for arg in sys.argv[1:]:
        mode, frames, path = arg.split(":", 2)
        c = capture(path)
        c["frames_per_launch"] = int(frames)
        c["source"] = os.path.basename(path)
        modes.setdefault(mode, {"captures": []})["captures"].append(c)
    for m in modes.values():
        m["captures"].sort(key=lambda c: c["frames_per_launch"])
        m["source"] = ", ".join(c["source"] for c in m["captures"])
    out = {"workload": "speed-test terrain 1M triangles, 10 000-ommatidia Fibonacci eye, S=1024 (bench.py defaults); one launch of the "
                       "batched trace kernel = frames_per_launch x 10.24M rays",
           "how": "ncu --clock-control none -k regex:k_traceCompound -s 1 -c 1 python bench.py --mode <mode> --steps <F> --repeats 1 "
                  "--no-cpu-baseline --no-modes; profiles/make_k1_traffic.py",
           "modes": modes}
    json.dump(out, open(os.path.join(HERE, "k1_traffic.json"), "w"), indent=1)
    for name, m in modes.items():
        for c in m["captures"]:
            rays = c["frames_per_launch"] * 10.24e6
            print(f"{name:12s} F={c['frames_per_launch']:3d}  {c['dram_bytes'] / rays:6.2f} B/ray DRAM  {c['inst_executed'] / rays:6.2f} warp-inst/ray  "
                  f"issue {c['issue_active_pct']:.1f}%  lanes {c['thread_inst_per_warp_inst']:.2f}")


if __name__ == "__main__":
    main()
