#!/bin/bash
# Round-2 visit r5j (one GPU): evict-first stores A/B for the 64-bit scan and for smaller u32 scans.
TAG=${1:-r5j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
for V in shipped ab_NO_EVICT_FIRST shipped ab_NO_EVICT_FIRST; do
  if [ $V = shipped ]; then unset DRJIT_B200_LIB; else export DRJIT_B200_LIB=$PWD/build/$V/libdrjit_b200.so; fi
  stamp "prims $V"
  timeout 300 python scripts/time_prims.py scan64 --reps 20 >> $OUT/prims_$V.txt 2>&1
  timeout 300 python scripts/time_prims.py scan64 --reps 20 --log2 26 >> $OUT/prims_$V.txt 2>&1
  timeout 300 python scripts/time_prims.py scan --reps 20 --log2 28 >> $OUT/prims_$V.txt 2>&1
  timeout 300 python scripts/time_prims.py scan --reps 20 --log2 27 >> $OUT/prims_$V.txt 2>&1
  tail -4 $OUT/prims_$V.txt
done
stamp done
