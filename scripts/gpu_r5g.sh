#!/bin/bash
# Round-2 visit r5g (one GPU): source-level ncu of the compaction at 1 % and 10 % mask density.
TAG=${1:-r5g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp prims; timeout 300 python scripts/time_prims.py compress compress01 compress99 > $OUT/prims.txt 2>&1; cat $OUT/prims.txt
for P in compress01; do
  stamp "ncu-full $P"
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:compress_kernel" -s 1 -c 1 -f -o /tmp/full_$P \
      python scripts/time_prims.py $P --reps 1 --warm 1 > $OUT/ncu_full_$P.log 2>&1; echo "ncu full $P rc=$?"
  ncu -i /tmp/full_$P.ncu-rep --page raw --csv > $OUT/full_$P.csv 2>/dev/null
  ncu -i /tmp/full_$P.ncu-rep --page source --csv > $OUT/source_$P.csv 2>/dev/null
done
stamp done
