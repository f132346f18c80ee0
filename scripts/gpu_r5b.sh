#!/bin/bash
# Round-2 visit r5b (one GPU): source-level ncu captures of the segmented scan (f32, block 1000) and the
# 64-bit scan -- one launch each, reports small enough to travel (read here with scripts/ncu_lines.py).
TAG=${1:-r5b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp prims; timeout 300 python scripts/time_prims.py scan scan64 scanseg > $OUT/prims.txt 2>&1; cat $OUT/prims.txt
for P in scanseg scan64; do
  stamp "ncu-full $P"
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:prefix_reduce" -s 1 -c 1 -f -o $OUT/full_$P \
      python scripts/time_prims.py $P --reps 1 --warm 1 > $OUT/ncu_full_$P.log 2>&1; echo "ncu full $P rc=$?"
  ls -la $OUT/full_$P.ncu-rep
done
stamp done
