#!/bin/bash
TAG=${1:-r1l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:prefix_reduce" -s 1 -c 1 -f -o $OUT/full_scanseg \
    python scripts/time_prims.py scanseg --reps 1 --warm 1 > $OUT/ncu_full_scanseg.log 2>&1; echo "ncu rc=$?"
