#!/bin/bash
# Round-1 evidence visit: smoke, parity tests, per-primitive timings, both bench arms, ncu launch
# list of the bench command, and one `ncu --set full` capture per hot kernel.
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
stamp smoke; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
stamp prims; timeout 600 python scripts/time_prims.py all > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp pytest; timeout 1200 python -m pytest tests -q -m gpu --maxfail=30 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -40
stamp bench; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 4000 $OUT/bench.json; tail -5 $OUT/bench.err
stamp bench-ref; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; tail -c 1500 $OUT/bench_ref.json; tail -5 $OUT/bench_ref.err
stamp ncu-launches; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
K='regex:reduce|compress|mkperm|scatter'
for P in scan compress sum block_reduce dot mkperm scatter; do
  stamp "ncu-full $P"
  timeout 420 ncu --set full --clock-control none --import-source on -k "$K" -s 1 -c 4 -f -o $OUT/full_$P \
      python scripts/time_prims.py $P --reps 1 --warm 1 > $OUT/ncu_full_$P.log 2>&1; echo "ncu full $P rc=$?"
done
stamp done
ls -la $OUT
