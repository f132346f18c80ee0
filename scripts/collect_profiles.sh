#!/bin/bash
# Copies the text evidence of one GPU visit (gpurun_out/<tag>) into profiles/ under the name
# <name>_*, deriving the summaries from the .ncu-rep files with the scripts next to this one.
#   scripts/collect_profiles.sh r1c r1c
TAG=$1; NAME=${2:-$1}
SRC=gpurun_out/$TAG; DST=profiles
set -e
cp $SRC/bench.json $DST/${NAME}_bench_n1.json
cp $SRC/bench_ref.json $DST/${NAME}_bench_reference_cpu.json
cp $SRC/prims.log $DST/${NAME}_prims.txt
cp $SRC/gpu.txt $DST/${NAME}_gpu.txt
cp $SRC/launches_bench.csv $DST/${NAME}_launches_bench.csv
python scripts/launch_shares.py $SRC/launches_bench.csv $SRC/bench.json > $DST/${NAME}_launch_shares.md
python scripts/ncu_summary.py $SRC > $DST/${NAME}_ncu_full.md
for f in sweep_scan_d0 sweep_scan_d1 sweep_compress_t128 sweep_compress_t3 sweep_compress_t253 microbench; do
  [ -f $SRC/$f.txt ] && cp $SRC/$f.txt $DST/${NAME}_$f.txt
done
grep -E "passed|failed" $SRC/pytest.log > $DST/${NAME}_pytest.txt || true
tail -1 $SRC/smoke.log >> $DST/${NAME}_pytest.txt || true
echo "collected $SRC -> $DST/${NAME}_*"
