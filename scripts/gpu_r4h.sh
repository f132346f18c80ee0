#!/bin/bash
# Round-2 visit H (one GPU): full GPU suite after the single-CTA scan path / payload-in-flag exchange,
# small-size table, scatter partition bound.
TAG=${1:-r4h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp small; timeout 600 python scripts/small_sizes.py > $OUT/small_sizes.txt 2>&1; head -14 $OUT/small_sizes.txt
stamp partition; timeout 120 build/microbench_partition > $OUT/microbench_partition.txt 2>&1; cat $OUT/microbench_partition.txt
stamp done
