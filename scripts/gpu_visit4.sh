#!/bin/bash
OUT=gpurun_out/${1:-r1h}; mkdir -p $OUT
echo "== sweep scan 2^30"; timeout 300 build/sweep_scan 30 2>&1 | tee $OUT/sweep_scan30.log
echo "== sweep scan 2^24"; timeout 300 build/sweep_scan 24 2>&1 | tee $OUT/sweep_scan24.log
echo "== sweep scan 2^30 no look-back"; timeout 300 build/sweep_scan 30 "" 1 2>&1 | tee $OUT/sweep_scan30_nolb.log
