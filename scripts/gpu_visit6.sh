#!/bin/bash
# Sweep visit: scan / compress geometry and variant sweeps, atomic-rate microbenchmarks,
# GPU parity tests of the current library, per-primitive timings, ncu source capture of mkperm.
TAG=${1:-r1b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
stamp sweep-scan
for D in 0 1 2; do timeout 300 build/sweep_scan 30 "" $D > $OUT/sweep_scan_d$D.txt 2>&1; done
cat $OUT/sweep_scan_d0.txt
stamp sweep-compress
for T in 128 3 253; do timeout 300 build/sweep_compress 30 $T > $OUT/sweep_compress_t$T.txt 2>&1; done
cat $OUT/sweep_compress_t128.txt; grep -E "v2|early ROWS=8 S=2 min3" $OUT/sweep_compress_t3.txt $OUT/sweep_compress_t253.txt
stamp microbench; timeout 300 build/microbench > $OUT/microbench.txt 2>&1; tail -40 $OUT/microbench.txt
stamp pytest; timeout 1200 python -m pytest tests -q -m gpu --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp prims; timeout 600 python scripts/time_prims.py all > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp "ncu mkperm"
timeout 420 ncu --set full --clock-control none --import-source on -k "regex:mkperm_tile" -s 2 -c 2 -f -o $OUT/full_mkperm \
    python scripts/time_prims.py mkperm --reps 1 --warm 1 > $OUT/ncu_full_mkperm.log 2>&1; echo "ncu rc=$?"
stamp "ncu compress"
timeout 420 ncu --set full --clock-control none --import-source on -k "regex:compress" -s 1 -c 1 -f -o $OUT/full_compress \
    python scripts/time_prims.py compress --reps 1 --warm 1 > $OUT/ncu_full_compress.log 2>&1; echo "ncu rc=$?"
stamp done
ls -la $OUT
