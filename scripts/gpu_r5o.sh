#!/bin/bash
# Round-2 visit r5o (one GPU): racecheck after the single-CTA barrier fix (small / unaligned scans, the
# whole prefix test group), memcheck of the small-array kernel, small-size table.
TAG=${1:-r5o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp racecheck; timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "small_arrays or int_ops or literals" > $OUT/racecheck.log 2>&1; tail -3 $OUT/racecheck.log
stamp memcheck; timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "small_arrays" > $OUT/memcheck.log 2>&1; tail -3 $OUT/memcheck.log
stamp pytest; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "prefix or scan" --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head
stamp small; timeout 600 python scripts/small_sizes.py 2>/dev/null > $OUT/small_sizes.txt; grep -E "primitive|prefix" $OUT/small_sizes.txt
stamp done
