#!/bin/bash
# Round-2 visit I (one GPU): suite after the small-array scan path, small-size table, scatter partition
# bound, evict-first scan stores A/B inside the bench step, ncu evidence for the new compress kernel.
TAG=${1:-r4i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -20
stamp small; timeout 600 python scripts/small_sizes.py 2>/dev/null > $OUT/small_sizes.txt; head -12 $OUT/small_sizes.txt
stamp partition; timeout 120 build/microbench_partition > $OUT/microbench_partition.txt 2>&1; cat $OUT/microbench_partition.txt
stamp scan-cs
export DRJIT_B200_LIB=$PWD/build/exp/libdrjit_b200.so
for D in 0 4; do
  DRJIT_B200_SCAN_DEBUG=$D timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/bench_scan_debug$D.json 2>/dev/null
  python -c "
import json; d=json.loads([l for l in open('$OUT/bench_scan_debug$D.json') if l.startswith('{')][-1]); print('scan debug $D:', d['ms_per_step'], {k:v['ms'] for k,v in d['primitives'].items()})"
done
unset DRJIT_B200_LIB
stamp ncu-launches; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
K='regex:reduce|compress|mkperm|scatter|sort'
for P in compress sort; do
  stamp "ncu-full $P"
  timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s 1 -c 4 -f -o /tmp/full_$P \
      python scripts/time_prims.py $P --reps 1 --warm 1 > $OUT/ncu_full_$P.log 2>&1; echo "ncu full $P rc=$?"
  ncu -i /tmp/full_$P.ncu-rep --page raw --csv > $OUT/full_$P.csv 2>/dev/null
done
stamp done
