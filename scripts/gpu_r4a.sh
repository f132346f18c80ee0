#!/bin/bash
# Round-2 visit (one GPU): smoke, the GPU test-suite, both bench arms with the driver's parameters,
# the ncu launch list of the bench command and one `ncu --set full` capture per primitive (raw and
# source pages exported to csv on the box: the .ncu-rep files are too large to travel back).
TAG=${1:-r4a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; free -g | head -2 >> $OUT/nproc.txt
stamp smoke; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp prims; timeout 300 python scripts/time_prims.py all > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp bench; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 6000 $OUT/bench.json; tail -5 $OUT/bench.err
stamp bench-ref; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; tail -c 1500 $OUT/bench_ref.json; tail -5 $OUT/bench_ref.err
stamp ncu-launches; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
K='regex:reduce|compress|mkperm|scatter'
for P in scan compress sum block_reduce dot mkperm scatter; do
  stamp "ncu-full $P"
  timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s 1 -c 4 -f -o /tmp/full_$P \
      python scripts/time_prims.py $P --reps 1 --warm 1 > $OUT/ncu_full_$P.log 2>&1; echo "ncu full $P rc=$?"
  ncu -i /tmp/full_$P.ncu-rep --page raw --csv > $OUT/full_$P.csv 2>/dev/null
  if [ "$P" = compress ] || [ "$P" = mkperm ] || [ "$P" = scatter ]; then
    ncu -i /tmp/full_$P.ncu-rep --page source --csv > $OUT/source_$P.csv 2>/dev/null
  fi
done
stamp done
ls -la $OUT
