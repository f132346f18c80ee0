#!/bin/bash
# Round-2 visit E (one GPU): sort + call_reduce tests, compress paired staging after the fix.
TAG=${1:-r4e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests/test_sort_gpu.py tests/test_call_reduce_gpu.py -q -m gpu > $OUT/pytest_sort_call.log 2>&1; echo "rc=$?"; tail -25 $OUT/pytest_sort_call.log
stamp prims; timeout 300 python scripts/time_prims.py call_reduce mkperm --reps 10 > $OUT/prims_call_reduce.txt 2>&1; cat $OUT/prims_call_reduce.txt
stamp sweep-compress
for T in 128 3 253 26; do timeout 120 build/sweep_compress 30 $T "S=1 min" > $OUT/sweep_compress_t$T.txt 2>&1; echo "rc=$?"; grep -v "vec\|bulk" $OUT/sweep_compress_t$T.txt; done
timeout 120 build/sweep_compress 30 128 "pairs" > $OUT/sweep_compress_pairs_all.txt 2>&1; cat $OUT/sweep_compress_pairs_all.txt
for LG in 20 24 27; do timeout 60 build/sweep_compress $LG 77 "pairs ROWS=8 S=1 min3" | tail -1; done
stamp sanitizers
timeout 300 compute-sanitizer --tool racecheck build/sweep_compress 22 128 "pairs ROWS=8 S=1 min3" > $OUT/racecheck_compress_pairs.txt 2>&1; tail -4 $OUT/racecheck_compress_pairs.txt
timeout 300 compute-sanitizer --tool memcheck build/sweep_compress 24 200 "pairs ROWS=8 S=1 min3" > $OUT/memcheck_compress_pairs.txt 2>&1; tail -4 $OUT/memcheck_compress_pairs.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_sort_gpu.py tests/test_call_reduce_gpu.py -q -m gpu -k "(bit_exact and 8193) or (70_001 or 70001)" > $OUT/memcheck_sort_call.txt 2>&1; tail -6 $OUT/memcheck_sort_call.txt
stamp done
