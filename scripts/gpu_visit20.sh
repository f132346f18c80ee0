#!/bin/bash
TAG=${1:-r1w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:mkperm_tile_scatter_stable" -s 1 -c 1 -f -o $OUT/full_mkperm_stable \
    python scripts/time_prims.py mkperm256 --reps 1 --warm 1 > $OUT/ncu_full_mkperm_stable.log 2>&1; echo "ncu rc=$?"
timeout 300 python scripts/incumbent.py --buckets 256 > $OUT/incumbent_256.txt 2>&1; grep "mkperm" $OUT/incumbent_256.txt
