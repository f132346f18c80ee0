#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an ncu report.

Joins `ncu --page source --csv` (per-SASS-instruction metrics) with `nvdisasm -g` line info of
the cubin extracted from the shared library (both work without a GPU).

    python scripts/ncu_lines.py <report.ncu-rep> <kernel-substring> [cubin-name-substring] [launch-index]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    cubin_pat = sys.argv[3] if len(sys.argv) > 3 else ""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "drjit_b200", "lib", "libdrjit_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)

    # ---- line table from nvdisasm: function -> offset -> (file, line) -------------------
    table = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin") or cubin_pat not in f:
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        func, cur = None, ("?", 0)
        for line in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", line)
            if m:
                func = m.group(1); table[func] = {}; continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m and func:
                table[func][int(m.group(1), 16)] = cur

    if rep.endswith(".csv"):      # source page exported on the GPU box (`ncu -i rep --page source --csv`)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the csv holds one block per launch: "Kernel Name", name / header / rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}; blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(row)
    want = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    blocks = [b for b in blocks if pat in b["name"]]
    b = blocks[want]
    hdr = b["rows"][0]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    # mangled name lookup: pick the function whose instruction count matches best
    rows = b["rows"][1:]
    base = int(rows[0][ia], 16)
    n = len(rows)
    cands = [f for f, t in table.items() if len(t) == n] or [f for f, t in table.items() if abs(len(t) - n) < 8]
    func = cands[0] if cands else None
    print(f"# {b['name']}\n# {n} SASS instructions; line table: {func}")
    per = {}
    tot_i = tot_s = 0
    for r in rows:
        off = int(r[ia], 16) - base
        key = table.get(func, {}).get(off, ("?", 0))
        inst, samp = int(r[ii] or 0), int(r[isamp] or 0)
        p = per.setdefault(key, [0, 0]); p[0] += inst; p[1] += samp
        tot_i += inst; tot_s += samp
    src_cache = {}
    print(f"# total warp instructions {tot_i}, samples {tot_s}\n# inst%  samp%  file:line  source")
    for key, (inst, samp) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        if inst < tot_i * 0.004 and samp < tot_s * 0.004:
            continue
        fn, ln = key
        text = ""
        for d in ("drjit_b200/csrc", "scripts"):
            path = os.path.join(root, d, fn)
            if os.path.exists(path):
                src_cache.setdefault(path, open(path).read().splitlines())
                if 0 < ln <= len(src_cache[path]):
                    text = src_cache[path][ln - 1].strip()
        print(f"{100 * inst / tot_i:5.1f}  {100 * samp / max(1, tot_s):5.1f}  {fn}:{ln}  {text[:110]}")


if __name__ == "__main__":
    main()
