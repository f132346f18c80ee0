#!/usr/bin/env python
"""Per-kernel share of the step from an ncu launch list (gpu__time_duration.sum csv).
    python scripts/launch_shares.py profiles/r1_launches_bench.csv [bench.json]
Compares with the CUDA-event per-primitive times of the bench line when given."""
import csv
import json
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("djb::", "")
    if name.startswith(("at::", "native::", "at_cuda_detail::", "cuda::")) or "fill_fmix32" in name:
        continue        # torch kernels of bench.py's verification pass / input generation
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[iv]) / 1e3
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | us / launch | share of step (ncu) |\n|---|---|---|---|---|")
for k, (n, t) in agg.items():
    print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f} % |")
if len(sys.argv) > 2:
    b = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ms = {k: v["ms"] for k, v in b["primitives"].items()}
    t = sum(ms.values())
    print("\n| primitive (CUDA events, bench.py) | ms | share of step |\n|---|---|---|")
    for k, v in ms.items():
        print(f"| {k} | {v} | {100 * v / t:.1f} % |")
