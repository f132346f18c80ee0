#!/usr/bin/env python
"""Summarises `ncu --set full` reports (gpurun_out/<tag>/full_*.ncu-rep) into a markdown table
for profiles/. Runs on the CPU box: `ncu -i <rep> --page raw --csv` needs no GPU.

    python scripts/ncu_summary.py gpurun_out/r1 > profiles/r1_ncu_full.md
"""
import csv
import glob
import io
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def main():
    d = sys.argv[1]
    print(f"# ncu --set full summaries ({d})\n")
    print("One row per captured launch; `dram rd/wr` are per launch (traffic for bench.py's roofline object); "
          "durations are under the profiler (cold caches, serialised) and are not bench values.\n")
    print("| report | kernel | grid x block | " + " | ".join(k for _, k in KEYS) + " | top stalls (pc samples) |")
    print("|---|---|---|" + "---|" * (len(KEYS) + 1))
    reps = sorted(glob.glob(os.path.join(d, "full_*.ncu-rep"))) or sorted(glob.glob(os.path.join(d, "full_*.csv")))
    for rep in reps:
        if rep.endswith(".csv"):      # raw page exported on the GPU box (the reports are too large to travel)
            out = open(rep).read()
        else:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            row = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            cells = []
            for k, _ in KEYS:
                v = row.get(k, "")
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u.get(k, '')}".strip())
            stalls = []
            for k in hdr:
                if k.startswith(STALL) and "not_issued" not in k:
                    try:
                        stalls.append((float(row[k]), k[len(STALL):]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            tot = sum(s for s, _ in stalls) or 1
            top = ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in stalls[:4])
            name = row["Kernel Name"].replace("|", "\\|")
            if len(name) > 90:
                name = name[:87] + "..."
            print(f"| {os.path.basename(rep)} | `{name}` | {row['Grid Size']} x {row['Block Size']} | " + " | ".join(cells) + f" | {top} |")


if __name__ == "__main__":
    main()
