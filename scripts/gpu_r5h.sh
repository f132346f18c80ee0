#!/bin/bash
# Round-2 visit r5h (one GPU): carry-window polling with back-off against the 20 ns polls (A/B build),
# compaction at 1 / 10 / 50 / 99 % density, compress parity tests.
TAG=${1:-r5h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
for V in shipped ab_COMPRESS_POLL20 shipped; do
  if [ $V = shipped ]; then unset DRJIT_B200_LIB; else export DRJIT_B200_LIB=$PWD/build/$V/libdrjit_b200.so; fi
  stamp "prims $V"; timeout 300 python scripts/time_prims.py compress compress01 compress10 compress99 --reps 20 >> $OUT/prims_$V.txt 2>&1; tail -4 $OUT/prims_$V.txt
done
unset DRJIT_B200_LIB
stamp pytest; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "compress" --maxfail=10 > $OUT/pytest_compress.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_compress.log | head
stamp done
