#!/bin/bash
# Round-2 visit r5c (one GPU): group-per-block segmented scan (parity + timing over block sizes), and
# source-level ncu of the 64-bit scan exported as CSV on the box (the reports are too large to travel).
TAG=${1:-r5c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "prefix" --maxfail=10 > $OUT/pytest_prefix.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_prefix.log | head -20
stamp prims; timeout 300 python scripts/time_prims.py scan64 scanseg > $OUT/prims.txt 2>&1; cat $OUT/prims.txt
stamp segsweep; timeout 300 python scripts/time_scanseg.py > $OUT/scanseg_sweep.txt 2>&1; cat $OUT/scanseg_sweep.txt
for P in scanseg scan64; do
  stamp "ncu-full $P"
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:prefix_" -s 1 -c 1 -f -o /tmp/full_$P \
      python scripts/time_prims.py $P --reps 1 --warm 1 > $OUT/ncu_full_$P.log 2>&1; echo "ncu full $P rc=$?"
  ncu -i /tmp/full_$P.ncu-rep --page raw --csv > $OUT/full_$P.csv 2>/dev/null
  ncu -i /tmp/full_$P.ncu-rep --page source --csv > $OUT/source_$P.csv 2>/dev/null
  ls -la $OUT/full_$P.csv $OUT/source_$P.csv
done
stamp done
