#!/usr/bin/env python
"""profiles/summarize_ncu.py <report.ncu-rep> [kernel-substring] -- text summary of an `ncu --set full`
capture: headline metrics per launch and the hottest source lines (needs -lineinfo)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__cycles_active.avg"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(run(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if len(sys.argv) > 2 and sys.argv[2] not in name:
            continue
        print(f"=== launch {r[hdr.index('ID')]}: {name[:90]}")
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"  {h} [{units[i]}] = {r[i]}")
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]]
        print("  stall cycles per issued instruction: " +
              ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}={v:.2f}"
                        for v, h in sorted(stalls, reverse=True)[:9]))
    src = list(csv.reader(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]).splitlines()))
    hdr_idx = [i for i, r in enumerate(src) if len(r) > 5 and "Instructions Executed" in r]
    if not hdr_idx:
        return
    h = src[hdr_idx[0]]
    i_s, i_i, i_t = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    lines, fname = [], ""
    first_end = hdr_idx[2] if len(hdr_idx) > 2 else len(src)      # sections of the first launch only
    for r in src[:first_end]:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if r and r[0].isdigit() and len(r) > i_t and r[i_i].isdigit():
            lines.append((fname, int(r[0]), r[1], int(r[i_s] or 0), int(r[i_i]), int(r[i_t] or 0)))
    ti, ts = sum(x[4] for x in lines) or 1, sum(x[3] for x in lines) or 1
    print(f"--- hottest source lines of the first captured launch (share of warp instructions / stall samples; threads per instruction)")
    for f, ln, s, smp, ins, thr in sorted(lines, key=lambda x: -x[4])[:40]:
        print(f"  {f}:{ln:<4d} inst {100 * ins / ti:5.1f}%  samples {100 * smp / ts:5.1f}%  thr/inst {thr / max(ins, 1):5.1f} | {s.strip()[:80]}")


if __name__ == "__main__":
    main()
