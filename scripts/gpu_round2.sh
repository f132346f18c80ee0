#!/bin/bash
TAG=${1:-r1b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== pytest"; timeout 2400 python -m pytest tests -q -m gpu --maxfail=30 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -40
echo "== prims (default)"; timeout 600 python scripts/time_prims.py all > $OUT/prims.log 2>&1; cat $OUT/prims.log
echo "== scan small tiles"; DRJIT_B200_SCAN_BIG=0 timeout 300 python scripts/time_prims.py scan scan64 2>&1 | tee $OUT/scan_small.log
for mode in warp cta global; do echo "== mkperm mode=$mode"; DRJIT_B200_MKPERM_MODE=$mode timeout 300 python scripts/time_prims.py mkperm mkperm256 2>&1 | tee $OUT/mkperm_$mode.log; done
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/time_prims.py all --reps 1 > $OUT/ncu_prims.log 2>&1; echo "ncu rc=$?"
