#!/bin/bash
# Evidence visit without the `ncu --set full` captures (those reports are ~60 MB: see gpu_evidence.sh):
# per-primitive timings, both bench arms, ncu launch list of the bench command, and one
# `ncu --set full` capture per hot kernel.
TAG=${1:-r1c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
stamp sweep-scan
for D in 0 1; do timeout 60 build/sweep_scan 30 "" $D > $OUT/sweep_scan_d$D.txt 2>&1; echo "rc=$?"; done
cat $OUT/sweep_scan_d0.txt $OUT/sweep_scan_d1.txt
stamp sweep-compress
for T in 128 3 253; do timeout 60 build/sweep_compress 30 $T > $OUT/sweep_compress_t$T.txt 2>&1; echo "rc=$?"; cat $OUT/sweep_compress_t$T.txt; done
stamp smoke; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp prims; timeout 300 python scripts/time_prims.py all > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp bench; timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 4500 $OUT/bench.json; tail -5 $OUT/bench.err
stamp bench-ref; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; tail -c 600 $OUT/bench_ref.json; tail -5 $OUT/bench_ref.err
stamp ncu-launches; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
stamp done
ls -la $OUT
