#!/bin/bash
TAG=${1:-r1e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp sweep-compress
for D in 0 1; do timeout 60 build/sweep_compress 30 128 "" $D > $OUT/sweep_compress_t128_d$D.txt 2>&1; echo "rc=$?"; cat $OUT/sweep_compress_t128_d$D.txt; done
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp prims; timeout 300 python scripts/time_prims.py mkperm mkperm256 > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp prims32; DRJIT_B200_MKPERM_TILE_KEYS=32 timeout 300 python scripts/time_prims.py mkperm mkperm256 > $OUT/prims32.log 2>&1; cat $OUT/prims32.log
stamp bench; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 1500 $OUT/bench.json; tail -5 $OUT/bench.err
stamp done
