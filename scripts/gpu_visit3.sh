#!/bin/bash
OUT=gpurun_out/${1:-r1d}; mkdir -p $OUT
echo "== sweep scan 2^30"; timeout 300 build/sweep_scan 30 2>&1 | tee $OUT/sweep_scan30.log
echo "== sweep scan 2^24"; timeout 120 build/sweep_scan 24 2>&1 | tee $OUT/sweep_scan24.log
echo "== sweep compress 2^30 50%"; timeout 300 build/sweep_compress 30 128 2>&1 | tee $OUT/sweep_compress30.log
echo "== sweep compress 2^30 99%"; timeout 300 build/sweep_compress 30 253 2>&1 | tee $OUT/sweep_compress30_99.log
