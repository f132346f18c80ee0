#!/bin/bash
# One GPU-box visit: smoke, parity tests, per-primitive timings, bench, ncu launch list.
# Logs -> gpurun_out/<tag>/. Usage (repo root on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=30 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -40
echo "== prims"; timeout 600 python scripts/time_prims.py all > $OUT/prims.log 2>&1; cat $OUT/prims.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; tail -c 1500 $OUT/bench_ref.json; tail -5 $OUT/bench_ref.err
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/time_prims.py all --reps 1 > $OUT/ncu_prims.log 2>&1; echo "ncu rc=$?"
