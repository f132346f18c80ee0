#!/bin/bash
# Round-2 visit B (one GPU): the in-situ tests (reference test programs through the patched
# libdrjit-core.so), compress copy-out sweep (LSU vs bulk shared->global), racecheck of the bulk variant.
TAG=${1:-r4b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp insitu; timeout 900 python -m pytest tests/test_insitu_gpu.py -q -m gpu -x > $OUT/pytest_insitu.log 2>&1; echo "insitu rc=$?"; tail -30 $OUT/pytest_insitu.log
mkdir -p /tmp/ins/out_reductions /tmp/ins/out_vcall
( cd /tmp/ins && timeout 600 $OLDPWD/oracle/_ref_b200/test_reductions -c > $OLDPWD/$OUT/ref_test_reductions.txt 2>&1; echo "test_reductions rc=$?"; tail -4 $OLDPWD/$OUT/ref_test_reductions.txt )
( cd /tmp/ins && timeout 600 $OLDPWD/oracle/_ref_b200/test_vcall -c > $OLDPWD/$OUT/ref_test_vcall.txt 2>&1; echo "test_vcall rc=$?"; tail -4 $OLDPWD/$OUT/ref_test_vcall.txt )
stamp sweep-compress
for T in 128 3 253; do timeout 120 build/sweep_compress 30 $T "ROWS=8" > $OUT/sweep_compress_t$T.txt 2>&1; echo "rc=$?"; cat $OUT/sweep_compress_t$T.txt; done
timeout 120 build/sweep_compress 30 128 "bulk" > $OUT/sweep_compress_bulk_all.txt 2>&1; cat $OUT/sweep_compress_bulk_all.txt
stamp racecheck
timeout 300 compute-sanitizer --tool racecheck build/sweep_compress 22 128 "bulk ROWS=8 S=1 min3" > $OUT/racecheck_compress_bulk.txt 2>&1; tail -5 $OUT/racecheck_compress_bulk.txt
timeout 300 compute-sanitizer --tool memcheck build/sweep_compress 24 77 "bulk ROWS=8 S=1 min3" > $OUT/memcheck_compress_bulk.txt 2>&1; tail -5 $OUT/memcheck_compress_bulk.txt
stamp done
