#!/bin/bash
# Round-2 visit r5d (one GPU): optimised group-per-block segmented scan: parity, sweep over block sizes
# with the group kernel on and off (experiments build), ncu raw + source pages as CSV.
TAG=${1:-r5d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "prefix" --maxfail=10 > $OUT/pytest_prefix.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_prefix.log | head -20
stamp segsweep; timeout 300 python scripts/time_scanseg.py > $OUT/scanseg_sweep.txt 2>&1; cat $OUT/scanseg_sweep.txt
stamp segsweep-general; DRJIT_B200_LIB=$PWD/build/exp/libdrjit_b200.so DRJIT_B200_SCAN_NO_GROUP=1 timeout 300 python scripts/time_scanseg.py > $OUT/scanseg_sweep_general_kernel.txt 2>&1; cat $OUT/scanseg_sweep_general_kernel.txt
for P in scanseg; do
  stamp "ncu-full $P"
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:prefix_" -s 1 -c 1 -f -o /tmp/full_$P \
      python scripts/time_prims.py $P --reps 1 --warm 1 > $OUT/ncu_full_$P.log 2>&1; echo "ncu full $P rc=$?"
  ncu -i /tmp/full_$P.ncu-rep --page raw --csv > $OUT/full_$P.csv 2>/dev/null
  ncu -i /tmp/full_$P.ncu-rep --page source --csv > $OUT/source_$P.csv 2>/dev/null
done
stamp done
