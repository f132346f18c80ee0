#!/bin/bash
# Round-2 visit D (one GPU): sort tests + timings, compress staging sweep (paired entries).
TAG=${1:-r4d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest-sort; timeout 900 python -m pytest tests/test_sort_gpu.py -q -m gpu -x > $OUT/pytest_sort.log 2>&1; echo "sort rc=$?"; tail -25 $OUT/pytest_sort.log
stamp prims-sort; timeout 300 python scripts/time_prims.py sort sortkeys sort_composed torch_sort --reps 10 > $OUT/prims_sort.txt 2>&1; cat $OUT/prims_sort.txt
stamp sweep-compress
for T in 128 3 253 26; do timeout 120 build/sweep_compress 30 $T "S=1 min" > $OUT/sweep_compress_t$T.txt 2>&1; echo "rc=$?"; grep -v "vec\|bulk" $OUT/sweep_compress_t$T.txt; done
timeout 120 build/sweep_compress 30 128 "pairs" > $OUT/sweep_compress_pairs_all.txt 2>&1; cat $OUT/sweep_compress_pairs_all.txt
stamp sanitizers
timeout 300 compute-sanitizer --tool racecheck build/sweep_compress 22 128 "pairs ROWS=8 S=1 min3" > $OUT/racecheck_compress_pairs.txt 2>&1; tail -5 $OUT/racecheck_compress_pairs.txt
timeout 300 compute-sanitizer --tool memcheck build/sweep_compress 24 200 "pairs ROWS=8 S=1 min3" > $OUT/memcheck_compress_pairs.txt 2>&1; tail -5 $OUT/memcheck_compress_pairs.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_sort_gpu.py -q -m gpu -k "bit_exact and 8193" > $OUT/memcheck_sort.txt 2>&1; tail -6 $OUT/memcheck_sort.txt
stamp done
