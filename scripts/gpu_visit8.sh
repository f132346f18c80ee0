#!/bin/bash
# Quick check visit: compress decoded two tiles ahead, new mkperm tile path (16 Ki / 32 Ki keys).
TAG=${1:-r1d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp sweep-compress
for T in 128 3 253; do timeout 60 build/sweep_compress 30 $T > $OUT/sweep_compress_t$T.txt 2>&1; echo "rc=$?"; cat $OUT/sweep_compress_t$T.txt; done
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp pytest32; DRJIT_B200_MKPERM_TILE_KEYS=32 timeout 900 python -m pytest tests -q -m gpu -k "mkperm or smoke" --maxfail=20 > $OUT/pytest32.log 2>&1; echo "pytest32 rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest32.log | head -30
stamp prims; timeout 300 python scripts/time_prims.py compress compress01 compress99 mkperm mkperm256 scan scan64 > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp prims32; DRJIT_B200_MKPERM_TILE_KEYS=32 timeout 300 python scripts/time_prims.py mkperm mkperm256 > $OUT/prims32.log 2>&1; cat $OUT/prims32.log
stamp "ncu mkperm16"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:mkperm_tile" -s 2 -c 2 -f -o $OUT/full_mkperm16 \
    python scripts/time_prims.py mkperm --reps 1 --warm 1 > $OUT/ncu_full_mkperm16.log 2>&1; echo "ncu rc=$?"
stamp "ncu mkperm32"
DRJIT_B200_MKPERM_TILE_KEYS=32 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:mkperm_tile" -s 2 -c 2 -f -o $OUT/full_mkperm32 \
    python scripts/time_prims.py mkperm --reps 1 --warm 1 > $OUT/ncu_full_mkperm32.log 2>&1; echo "ncu rc=$?"
stamp "ncu compress"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:compress" -s 1 -c 1 -f -o $OUT/full_compress \
    python scripts/time_prims.py compress --reps 1 --warm 1 > $OUT/ncu_full_compress.log 2>&1; echo "ncu rc=$?"
stamp done
